"""Summarise an .ncu-rep (read here on the CPU box): key roofline counters per captured launch."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "launch__grid_size", "launch__cluster", "launch__registers_per_thread", "gpu__time_duration.sum",
        "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct", "sm__warps_active.avg.pct", "lts__t_sectors_srcunit_tex", "smsp__warp_issue_stalled", "lts__t_sectors_op_read.sum", "lts__t_sectors.sum "]
for i, h in enumerate(hdr):
    if any(h.startswith(w) or w in h for w in want) and "smsp__warp_issue_stalled" not in h:
        print("%-80s %-10s %s" % (h, units[i], " | ".join(r[i] for r in rows[2:])))
