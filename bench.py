#!/usr/bin/env python
"""bench.py — queries/sec of the batched exact search (BASELINE.json configs[1]) on N B200s.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload exact|hnsw]

One "step" = one pass of the hot path over one 10,000-query batch against the 1,000,000 x 768 cosine corpus
(k = 10, recall@10 = 1.0 by construction: the scan is exact). N > 1 (launched under torchrun, one rank per GPU)
shards the corpus by vector id (global = local * N + rank, src/index/sharded_rwlock.rs:54-66), every rank scans
its shard for the whole batch, the per-shard top-k lists are all-gathered over NCCL and merged on the device
("scaling": "strong": the corpus is fixed, per-GPU work shrinks with N).

`--impl reference` times the reference algorithm's CPU form of the same path (the oracle's brute-force scan —
the Rust crate cannot be built in this image) on the host cores, on a bounded query sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ROWS = 1_000_000
DIM = 768
N_QUERIES = 10_000
K = 10
SEED = 0x5EED
QUERY_BATCHES = 4          # distinct query batches rotated through the steps


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def quiet_nccl_stdout():
    """stdout carries ONE JSON line: NCCL's log goes to stderr, and its bare version banner (printed to stdout at
    NCCL_DEBUG=VERSION) is dropped; an explicit NCCL_DEBUG=INFO/TRACE from the caller is kept."""
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"


def measured_traffic(kernel, **match):
    """DRAM bytes per launch of `kernel` from the committed ncu capture, if it was taken on this configuration."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[kernel]
    except Exception:
        return None
    return t["dram_bytes_per_launch"] if all(t.get(k) == v for k, v in match.items()) else None


def make_corpus(kind, n_rows):
    from chronomind_b200.datasets import ChunkedCorpus
    return ChunkedCorpus(kind, n_rows, DIM, SEED)


def make_data(n_rows, n_queries_total, kind="emb", rank=0, world=1):
    """(this rank's shard of the corpus, all queries, the corpus generator). Chunked generation: a rank never holds
    more than its shard, a checker can stream the whole corpus."""
    t0 = time.time()
    gen = make_corpus(kind, n_rows)
    shard = gen.shard(rank, world)
    queries = gen.queries(n_queries_total)
    log("[bench] synthetic %s data: shard %d/%d %s of %d rows, queries %s in %.1fs"
        % ("embedding-like (16-d subspace)" if kind == "emb" else "uniform", rank, world, shard.shape, n_rows, queries.shape, time.time() - t0))
    return shard, queries, gen


def total_keys(dist, ids):
    """(TotalF32, u32) tuples as sortable u64 keys (src/index/mod.rs:37-52; lockfree_hnsw.rs:149-150)."""
    b = np.ascontiguousarray(dist, np.float32).view(np.uint32).astype(np.uint64)
    k = np.where(b & np.uint64(0x80000000), ~b & np.uint64(0xFFFFFFFF), b | np.uint64(0x80000000))
    return (k << np.uint64(32)) | ids.astype(np.uint64)


def cpu_scan_streaming(gen, q, k, threads, cache=None):
    """The reference algorithm's CPU brute force (oracle, tests/property_test.rs:201-206 arithmetic, queries split
    statically over `threads` like benches/comparison.rs:125-139) over the WHOLE corpus, streamed chunk by chunk so
    the corpus never has to fit in host memory (`cache`: keep the prepared chunks between calls when it does).
    Returns (ids, dist, seconds spent scanning)."""
    import oracle
    q = np.ascontiguousarray(q, np.float32)
    best = None
    scan_s = 0.0
    for ci in range(gen.n_chunks):
        if cache is not None and ci in cache:
            prepared = cache[ci]
        else:
            prepared = oracle.preprocess_rows(gen.chunk(ci), threads=threads)   # the index stores preprocess()ed rows
            if cache is not None:
                cache[ci] = prepared
        t0 = time.perf_counter()
        ids, d = oracle.brute_force(prepared, q, k, mode="prepared", threads=threads)
        scan_s += time.perf_counter() - t0
        keys = total_keys(d, ids.astype(np.uint64) + np.uint64(ci * gen.CHUNK))
        best = keys if best is None else np.sort(np.concatenate([best, keys], axis=1), axis=1)[:, :k]
    best = np.sort(best, axis=1)[:, :k]
    out_ids = (best & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    kb = (best >> np.uint64(32)).astype(np.uint32)
    bits = np.where(kb & np.uint32(0x80000000), kb & np.uint32(0x7FFFFFFF), ~kb)
    return out_ids, bits.astype(np.uint32).view(np.float32), scan_s


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self._stop_evt = gpu_index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "power_w_max": max(float(r[2]) for r in self.rows), "samples": len(self.rows)}


def run_reference(args, rank, world):
    """`--impl reference`: the reference's own CPU form of the path. The Rust crate cannot be built in this image, so
    this times the oracle port of its brute-force scan on all host threads, a bounded query sample per step."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_rows, nq = args.rows, args.queries
    gen = make_corpus(args.data, n_rows)
    queries = gen.queries(nq)
    sample = min(args.cpu_sample or 2 * threads, nq)
    cache = {} if n_rows * DIM * 4 <= 8e9 else None      # hold the prepared corpus when it fits (the index does)
    for _ in range(min(args.warmup, 1)):
        cpu_scan_streaming(gen, queries[:sample], K, threads, cache)
    total = 0.0
    for s in range(args.steps):
        q = queries[(s * sample) % max(1, nq - sample):][:sample]
        total += cpu_scan_streaming(gen, q, K, threads, cache)[2]
    qps = sample * args.steps / total
    line = {
        "impl": "reference", "metric": "queries/sec at recall@10>=0.95, 1Mx768 cosine, batch 10k", "value": qps,
        "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * total / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "exact brute-force kNN %dx%d cosine (%s), %d-query batch, k=%d (BASELINE configs[1])" % (n_rows, DIM, args.data, nq, K),
                   "recall_at_10": 1.0, "engine": "CPU scan of the reference algorithm (oracle port, AVX2+FMA)"},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": "%d queries per step x %d steps against the full %d x %d corpus (scan time only)" % (sample, args.steps, n_rows, DIM)},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _preprocess_mt(corpus, threads):
    import oracle
    return oracle.preprocess_rows(corpus, threads=threads)


def run_hnsw(args, rank, world, local_rank):
    """BASELINE configs[2]: batched HNSW search (one small CTA per query, rows gathered by the copy engine) + temporal
    rerank, w = 0.3, on the graph the host builder constructs over the synthetic corpus (the snapshot carries no graph,
    SURVEY.md F2).
      value = queries/s of `VectorIndex::search(q, ef)` + take(10) at --ef, device resident, with recall@10 against the
              exact scan reported beside it. N > 1: every rank searches ITS shard's graph (independent HNSW sub-graph per
              shard, rows by id mod N), lists all-gathered over NCCL and merged by (distance, global handle).
      e2e   = `ChronoMind::search` semantics with HOST buffers: H2D of the queries, HNSW pool of ef candidates,
              (N > 1: all-gather + merge to the GLOBAL top-ef,) temporal rerank, D2H of [nq][10] handles and scores."""
    import torch
    import torch.distributed as dist
    from chronomind_b200 import datasets, engine, sharding
    from chronomind_b200.engine import ChronoMindB200, default_config
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        quiet_nccl_stdout()
        dist.init_process_group("nccl", device_id=dev)
    n_rows, nq = args.rows, args.queries
    efs = sorted({int(e) for e in args.efs.split(",")} | {args.ef})
    shard, queries_all, gen = make_data(n_rows, nq * 2, args.data, rank, world)
    now = (1_790_000_000, 0)
    W, BASE = 0.3, 0.1
    ts_s, ts_n, decay = datasets.temporal_attrs(n_rows, 1, now[0])          # indexed by GLOBAL handle
    cfg = default_config(dimensions=DIM, ef_construction=args.efc)
    threads = max(1, (os.cpu_count() or 1) // world)
    st = ChronoMindB200.from_matrix(shard, cfg, sharding.shard_attrs(ts_s, rank, world), sharding.shard_attrs(ts_n, rank, world),
                                    sharding.shard_attrs(decay, rank, world)).set_shard(rank, world)
    t0 = time.time()
    if args.build == "gpu":
        st.build_graph_gpu(SEED + rank, local_rank, threads)
    else:
        st.build_graph(SEED + rank, threads)
    build_s = time.time() - t0
    log("[bench] rank %d %s graph build: %d rows, M=16 efC=%d, %d threads, %.1fs (%.0f inserts/s), invariants=%d"
        % (rank, args.build, shard.shape[0], args.efc, threads, build_s, shard.shape[0] / build_s, st.check_invariants()))
    st.upload(local_rank)
    g_ts_s = torch.from_numpy(ts_s.view(np.int64)).to(dev)
    g_ts_n = torch.from_numpy(ts_n.view(np.int32)).to(dev)
    g_decay = torch.from_numpy(decay).to(dev)
    q_host = [torch.from_numpy(queries_all[i * nq:(i + 1) * nq]).pin_memory() for i in range(2)]
    q_dev = [q.to(dev) for q in q_host]
    ids_l = torch.empty((nq, K), dtype=torch.int32, device=dev)
    dist_l = torch.empty((nq, K), dtype=torch.float32, device=dev)
    out_i = torch.empty((nq, K), dtype=torch.int32).pin_memory()
    out_s = torch.empty((nq, K), dtype=torch.float32).pin_memory()

    rows_per_rank = (nq + world - 1) // world
    q_gather = torch.empty((world * rows_per_rank, DIM), dtype=torch.float32, device=dev) if world > 1 else None
    q_slice = torch.empty((rows_per_rank, DIM), dtype=torch.float32, device=dev) if world > 1 else None

    def merged(ids, d, length):
        if world == 1:
            return ids, d
        ig, dg = sharding.allgather_lists(ids, d)
        return engine.merge_topk_cuda(ig, dg, length)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    truth = [merged(*st.search_exact_cuda(q, K), K)[0].cpu().numpy().view(np.uint32) for q in q_dev]   # exact ground truth
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs") or 6650.0

    def recall(got, want):
        return float(np.mean([len(set(a) & set(b)) / K for a, b in zip(got.tolist(), want.tolist())]))

    def step_device(i, ef):
        st.search_index_cuda(q_dev[i % 2], ef, K, ids_l, dist_l)
        return merged(ids_l, dist_l, K)

    def step_e2e(i, ef):
        """ChronoMind::search through host buffers."""
        if world == 1:
            return st.search(q_host[i % 2].numpy(), K, now=now, ef_search=ef, temporal_weight=W, base_decay_rate=BASE)
        qh = q_host[i % 2]                                               # 1/N of the batch per rank over PCIe,
        lo = min(nq, rank * rows_per_rank)                               # the slices all-gathered over NVLink
        hi = min(nq, lo + rows_per_rank)
        q_slice[:hi - lo].copy_(qh[lo:hi], non_blocking=True)
        dist.all_gather_into_tensor(q_gather, q_slice)
        qd = q_gather[:nq]
        pool = max(ef, 3 * K)                                            # OVERSAMPLE, src/store.rs:40,323
        pi, pd = st.search_index_cuda(qd, pool, pool)
        mi, md = merged(pi, pd, pool)                                    # global top-ef first ...
        ri, rs, _ = engine.rerank_cuda(mi, md, K, g_ts_s, g_ts_n, g_decay, W, BASE, now)   # ... then the temporal blend
        out_i.copy_(ri, non_blocking=True)
        out_s.copy_(rs, non_blocking=True)
        torch.cuda.synchronize()
        return out_i, out_s

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    per_ef, head = {}, None
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    for ef in efs:
        for i in range(args.warmup):
            step_device(i, ef)
        barrier()
        engine.profile_enable(True)
        engine.profile_read()
        launches0 = engine.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for i in range(args.steps):
            res = step_device(args.warmup + i, ef)
        ev1.record()
        barrier()
        ms = max_over_ranks(ev0.elapsed_time(ev1))
        launches = engine.launch_count() - launches0
        kernel_ms, kernel_n = engine.profile_read()
        engine.profile_enable(False)
        last = (args.warmup + args.steps - 1) % 2
        got = res[0].cpu().numpy().view(np.uint32)
        for i in range(min(args.warmup, 2)):
            step_e2e(i, ef)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            step_e2e(args.warmup + i, ef)
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        # work counters of this rank's shard (== the CPU oracle's at equal ef on the same graph: tests/test_gpu_parity.py)
        st.search_index(q_host[last].numpy(), ef, K)
        c = engine.last_counters()
        bytes_per_batch = c["dist_evals"] * DIM * 4 + c["list_entries"] * 4 + nq * DIM * 4
        kern_avg = kernel_ms / max(kernel_n, 1)
        r = {"ef": ef, "qps": nq * args.steps / (ms * 1e-3), "ms_per_step": ms / args.steps, "recall_at_10": recall(got, truth[last]),
             "dist_evals_per_query": c["dist_evals"] / nq, "expansions_per_query": c["expansions"] / nq,
             "algorithmic_bytes_per_query": bytes_per_batch / nq, "kernel_ms_avg": kern_avg,
             "achieved_gbs": bytes_per_batch / (kern_avg * 1e-3) / 1e9 if kern_avg > 0 else None,
             "hbm_frac": bytes_per_batch / (kern_avg * 1e-3) / 1e9 / hbm_peak if kern_avg > 0 else None,
             "temporal_e2e_qps": nq * args.steps / e2e_s, "gpu_launches": launches, "visited_table_restarts": c["fallback_queries"]}
        per_ef[str(ef)] = r
        if rank == 0:
            log("[bench] hnsw ef=%d: %s" % (ef, json.dumps(r)))
        if ef == args.ef:
            head = r
    clocks = sampler.stop() if sampler else None

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        import oracle
        import tempfile
        t0 = time.time()
        sample = args.cpu_sample or 64 * threads
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "graph.cmg")
            st.save_graph(path)
            g = oracle.read_graph_sidecar(path)
        orc = oracle.OracleHnsw(16, args.efc, 50, SEED)
        orc.import_graph(_preprocess_mt(shard, threads), g)
        qs = queries_all[:sample]
        orc.search_batch(qs, args.ef, K, threads=threads)
        best = None
        for _ in range(3):                                   # best_qps (benches/external.rs:118-124)
            t1 = time.perf_counter()
            oi, od, _, octr = orc.search_batch(qs, args.ef, K, threads=threads)
            dt = time.perf_counter() - t1
            best = dt if best is None else min(best, dt)
        gi, gd = st.search_index(qs, args.ef, K)
        parity = bool(np.array_equal(gi, oi) and np.array_equal(gd.view(np.uint32), od.view(np.uint32)))
        cpu = {"value": sample / best, "unit": "queries/s", "cores": threads, "kind": "port",
               "sample": "%d queries, ef=%d, same graph (sidecar import), best of 3 passes (%.2fs each)" % (sample, args.ef, best),
               "parity_on_sample": parity, "dist_evals_per_query": float(octr[0]) / sample}
        log("[bench] cpu baseline in %.1fs: %s" % (time.time() - t0, cpu))
    if rank == 0:
        line = {
            "metric": "queries/sec at recall@10>=0.95, 1Mx768 cosine, batch 10k", "value": head["qps"], "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "HNSW batched search %dx%d cosine (%s), %d-query batch, k=%d, ef=%d, M=16 efC=%d host-built graph%s "
                                   "(BASELINE configs[2]); temporal rerank w=0.3 in the e2e leg"
                                   % (n_rows, DIM, args.data, nq, K, args.ef, args.efc, "" if world == 1 else " per shard, %d shards by id" % world),
                       "recall_at_10": head["recall_at_10"], "per_ef": per_ef, "graph_build": args.build, "graph_build_s": build_s,
                       "graph_build_inserts_per_s": shard.shape[0] / build_s, "build_threads": threads,
                       "parallelism": "1 GPU" if world == 1 else "independent HNSW sub-graph per shard on %d GPUs, NCCL allgather + device merge (+ rerank of the global top-ef in e2e)" % world,
                       "l2": "%.2f GB fp32 rows per GPU gathered at random: larger than L2; 2 query batches rotated" % (n_rows / world * DIM * 4 / 1e9)},
            "e2e": {"value": head["temporal_e2e_qps"], "unit": "queries/s", "h2d_bytes_per_step": nq * DIM * 4, "d2h_bytes_per_step": nq * K * 8},
            "gpu_launches": head["gpu_launches"],
            "roofline": {"bound": "hbm", "achieved": head["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": head["hbm_frac"],
                         "traffic": measured_traffic("k_hnsw_search", rows=n_rows, queries=nq, n_gpus=world, data=args.data, ef=args.ef),
                         "traffic_unit": "DRAM bytes per launch (ncu) vs algorithmic %d" % int(head["algorithmic_bytes_per_query"] * nq),
                         "kernel": "k_hnsw_search (rank 0's shard)", "kernel_ms_avg": head["kernel_ms_avg"],
                         "algorithmic_bytes": "dist_evals*dim*4 + list_entries*4 + dim*4 per query, counters == CPU oracle's at equal ef",
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback of B200_PROFILING.md"},
            "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    global DIM, K
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=N_ROWS, help="corpus rows (default: the metric's 1,000,000)")
    ap.add_argument("--queries", type=int, default=N_QUERIES)
    ap.add_argument("--engine", type=int, default=0, help="0 auto, 1 fp32 CUDA-core scan, 2 tcgen05 scan + rescore")
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the cpu_baseline sample (0: 2 x cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="exact", choices=["exact", "hnsw"],
                    help="exact: BASELINE configs[1] (the metric's workload); hnsw: configs[2], batched HNSW + temporal rerank")
    ap.add_argument("--ef", type=int, default=128, help="hnsw: the ef the headline value is quoted at")
    ap.add_argument("--efs", default="64,128,200", help="hnsw: every ef measured (reported under config.per_ef)")
    ap.add_argument("--efc", type=int, default=100, help="hnsw: ef_construction of the host-built graph")
    ap.add_argument("--data", default="emb", choices=["emb", "uniform"])
    ap.add_argument("--dim", type=int, default=DIM, help="vector dimension (default 768; 32 with --k 100 is BASELINE configs[4])")
    ap.add_argument("--k", type=int, default=K, help="neighbors per query (default 10)")
    ap.add_argument("--build", default="host", choices=["host", "gpu"],
                    help="hnsw: graph construction — host (LockFreeHnsw::insert restatement, all cores) or gpu-assisted "
                         "(all-pairs kNN candidates on the device, select/link on the host)")
    args = ap.parse_args()

    DIM, K = args.dim, args.k            # the workload shape is module state for the helpers above

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload == "hnsw":
        run_hnsw(args, rank, world, local_rank)
        return

    import torch
    import torch.distributed as dist
    from chronomind_b200 import engine, sharding
    from chronomind_b200.engine import ChronoMindB200, default_config

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        quiet_nccl_stdout()
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"

    n_rows, nq = args.rows, args.queries
    shard, queries_all, gen = make_data(n_rows, nq * QUERY_BATCHES, args.data, rank, world)
    t0 = time.time()
    st = ChronoMindB200.from_matrix(shard, default_config(dimensions=DIM)).set_shard(rank, world).upload(local_rank)
    st.set_exact_engine(args.engine)
    log("[bench] rank %d: index of %d rows resident on cuda:%d in %.1fs" % (rank, len(st), local_rank, time.time() - t0))

    q_host = [torch.from_numpy(queries_all[i * nq:(i + 1) * nq]).pin_memory() for i in range(QUERY_BATCHES)]
    q_dev = [q.to(dev) for q in q_host]
    ids_l = torch.empty((nq, K), dtype=torch.int32, device=dev)
    dist_l = torch.empty((nq, K), dtype=torch.float32, device=dev)
    out_ids_host = torch.empty((nq, K), dtype=torch.int32).pin_memory()
    out_dist_host = torch.empty((nq, K), dtype=torch.float32).pin_memory()

    def step_device(i):
        """Inputs resident in HBM; result stays on the device."""
        st.search_exact_cuda(q_dev[i % QUERY_BATCHES], K, ids_l, dist_l)
        if world > 1:
            ids_g, dist_g = sharding.allgather_lists(ids_l, dist_l)
            return engine.merge_topk_cuda(ids_g, dist_g, K)
        return ids_l, dist_l

    # N > 1: every rank needs the whole batch on its GPU. Each rank copies 1/N of it from pinned host memory over
    # ITS PCIe link and the slices are all-gathered over NVLink (30.7 MB at 10k x 768) — the node's aggregate host
    # bandwidth instead of N full copies.
    rows_per_rank = (nq + world - 1) // world
    q_gather = torch.empty((world * rows_per_rank, DIM), dtype=torch.float32, device=dev) if world > 1 else None
    q_slice = torch.empty((rows_per_rank, DIM), dtype=torch.float32, device=dev) if world > 1 else None

    def step_e2e(i):
        """Public API with HOST buffers: H2D of the query batch and D2H of the result inside the step."""
        if world == 1:
            return st.search_exact(q_host[i % QUERY_BATCHES].numpy(), K)
        qh = q_host[i % QUERY_BATCHES]
        lo = min(nq, rank * rows_per_rank)
        hi = min(nq, lo + rows_per_rank)
        q_slice[:hi - lo].copy_(qh[lo:hi], non_blocking=True)
        dist.all_gather_into_tensor(q_gather, q_slice)
        st.search_exact_cuda(q_gather[:nq], K, ids_l, dist_l)
        ids_g, dist_g = sharding.allgather_lists(ids_l, dist_l)
        mi, md = engine.merge_topk_cuda(ids_g, dist_g, K)
        if rank == 0:                                   # the caller reads the merged result once
            out_ids_host.copy_(mi, non_blocking=True)
            out_dist_host.copy_(md, non_blocking=True)
        torch.cuda.synchronize()
        return out_ids_host, out_dist_host

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------------------------
    for i in range(args.warmup):
        step_device(i)
    barrier()
    engine.profile_enable(True)
    engine.profile_read()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = engine.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        res = step_device(args.warmup + i)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    last_batch = (args.warmup + args.steps - 1) % QUERY_BATCHES
    dev_ids = res[0].cpu().numpy().view(np.uint32).copy()      # what the timed device path returned last
    dev_dist = res[1].cpu().numpy().copy()
    launches = engine.launch_count() - launches0
    kernel_ms, kernel_n = engine.profile_read()
    engine.profile_enable(False)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())

    # ---- end-to-end through the host-buffer API -----------------------------------------------------------------
    for i in range(min(args.warmup, 2)):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        out = step_e2e(args.warmup + i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        counters = engine.last_counters()                      # of the last host-buffer call (the e2e leg)
        tc = counters.get("exact_engine") == 2
        unresolved = int((dev_ids[:, 0] == 0xFFFFFFFF).sum() + np.isnan(dev_dist[:, 0]).sum())
        log("[bench] counters of the last e2e call: %s; unresolved rows in the timed device result: %d" % (counters, unresolved))
        flops_per_launch = 2.0 * nq * (n_rows / world) * DIM * args.steps / max(kernel_n, 1)
        kern_avg_ms = kernel_ms / max(kernel_n, 1)
        achieved = flops_per_launch / (kern_avg_ms * 1e-3) / 1e12 if kern_avg_ms > 0 else None
        peak = peaks.get("bf16_tflops_sustained") or 1400.0
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": (achieved / peak) if achieved else None,
                    "traffic": measured_traffic("scan_tc_kernel", rows=n_rows, queries=nq, n_gpus=world, data=args.data) if tc else None,
                    "traffic_unit": "DRAM bytes per launch (ncu); algorithmic minimum = one pass over the bf16 corpus + queries = %d" % (n_rows // world * DIM * 2 + nq * DIM * 2),
                    "kernel": "scan_tc (tcgen05 bf16)" if tc else "k_exact_dist (fp32 CUDA cores, bit-exact chain)",
                    "kernel_ms_avg": kern_avg_ms, "kernel_launches_per_step": kernel_n / args.steps,
                    "scores_per_s_per_gpu": (nq * (n_rows / world) / (kern_avg_ms * 1e-3)) if kern_avg_ms > 0 else None,
                    "note": None if DIM >= 256 else "short rows: one or two UMMA k-steps per tile, the TMEM-epilogue select bounds "
                                                    "this shape (SURVEY.md H6); read scores_per_s_per_gpu, not frac",
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
                    if peaks else "fallback of B200_PROFILING.md"}
        cpu = None
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sample = args.cpu_sample or 2 * threads
            t0 = time.time()
            # the oracle's scan of the sample: the CPU number AND the parity check of the TIMED device result
            # (last step's batch)
            qs = queries_all[last_batch * nq:last_batch * nq + sample]
            wi, wd, secs = cpu_scan_streaming(gen, qs, K, threads)
            qps = sample / secs
            gi, gd = dev_ids[:sample], dev_dist[:sample]
            parity = bool(np.array_equal(gi, wi) and np.array_equal(gd.view(np.uint32), wd.view(np.uint32)))
            recall = float(np.mean([len(set(a) & set(b)) / K for a, b in zip(gi.tolist(), wi.tolist())]))
            if DIM != 768:
                log("[bench] note: unit vectors, so the L2 ranking of BASELINE configs[4] is this cosine ranking "
                    "(|a-b|^2 = 2(1 - a.b)); the reference has no L2 metric (SURVEY.md F3)")
            cpu = {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                   "sample": "%d of the %d queries against the full %dx%d corpus, one pass (%.1fs of scan time)"
                             % (sample, nq, n_rows, DIM, secs), "parity_on_sample": parity,
                   "recall_at_10_on_sample": recall}
            log("[bench] cpu baseline in %.1fs" % (time.time() - t0))
        cfg_no = "1" if (n_rows <= 1_000_000 and DIM == 768) else ("3" if DIM == 768 else "4")
        line = {
            "metric": "queries/sec at recall@10>=0.95, 1Mx768 cosine, batch 10k", "value": nq * args.steps / (ms * 1e-3),
            "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16" if tc else "f32", "data": "synthetic",
            "config": {"workload": "exact brute-force kNN %dx%d cosine (%s data), %d-query batch, k=%d (BASELINE configs[%s])"
                                   % (n_rows, DIM, args.data, nq, K, cfg_no),
                       "recall_at_10": (cpu or {}).get("recall_at_10_on_sample", "exact by construction (unchecked: --no-cpu-baseline)"),
                       "unresolved_queries": unresolved, "fallback_queries_e2e": counters.get("fallback_queries"),
                       "engine": "tcgen05 bf16 scan + fp32 rescore" if tc else "fp32 CUDA-core scan",
                       "parallelism": "corpus sharded by id over %d GPU(s), NCCL allgather + device merge" % world if world > 1 else "1 GPU",
                       "l2": "inputs larger than L2: %.2f GB bf16 corpus per GPU streamed per step; %d query batches rotated" % (n_rows / world * DIM * 2 / 1e9, QUERY_BATCHES)},
            "e2e": {"value": nq * args.steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": nq * DIM * 4,
                    "d2h_bytes_per_step": nq * K * 8},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
