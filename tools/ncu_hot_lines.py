"""Hottest source lines of one kernel in an .ncu-rep (warp-stall samples per CUDA line, with the dominant stall reasons).
Usage: python tools/ncu_hot_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(raw.splitlines()))
cur_file, hdr, lines = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        lines.append((cur_file, r))
i_s = hdr.index("# Samples")
i_ex = hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
total = sum(int(r[i_s] or 0) for _, r in lines)
print("total samples", total)
for f, r in sorted(lines, key=lambda fr: -int(fr[1][i_s] or 0))[:top_n]:
    n = int(r[i_s] or 0)
    top = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stalls), reverse=True)[:3]
    print("%5.1f%% %-18s:%-5s inst=%-9s %-34s %s" % (100.0 * n / max(total, 1), f, r[0], r[i_ex],
          " ".join("%s=%d" % (k, v) for v, k in top if v), r[1].strip()[:90]))
