#!/usr/bin/env python
"""bench.py — queries/sec at recall@10 >= 0.95 of the batched search path (BASELINE.json) on N B200s.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload exact|hnsw]

One "step" = one pass of the hot path over one 10,000-query batch against the 1,000,000 x 768 cosine corpus (k = 10).
The JSON line's headline (`value`, `e2e`, `roofline`) is the exact scan of BASELINE configs[1] (tcgen05 bf16 scan + fp32
rescore, recall@10 = 1.0 by construction). The same line carries an `hnsw` block for configs[2]: the GPU-assisted
graph of the same corpus, the batched traversal swept over ef, recall@10 against the exact scan, achieved HBM
bandwidth on algorithmic bytes, and the FRONTIER point — the smallest ef whose recall@10 reaches 0.95.

N > 1 (launched under torchrun, one rank per GPU) shards the corpus by vector id (global = local * N + rank,
src/index/sharded_rwlock.rs:54-66); every rank searches its shard for the whole batch, the per-shard top-k lists are
all-gathered over NCCL and merged on the device ("scaling": "strong": the corpus is fixed).

`--impl reference` times the reference's PRODUCTION read path on the host cores: `LockFreeHnsw::search`
(src/index/lockfree_hnsw.rs:469-495; the oracle port — the Rust crate cannot be built in this image) over the same
graph, at the smallest ef that reaches recall@10 >= 0.95, queries split statically over all host threads
(benches/comparison.rs:125-139). Its brute-force scan (tests/recall_test.rs:70-79) is reported beside it.
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
METRIC = "queries/sec at recall@10>=0.95, 1Mx768 cosine, batch 10k"
DEFAULT_EFS = "16,20,24,28,32,40,48,64,128,200"
NOW = (1_790_000_000, 0)
TEMPORAL_W, BASE_DECAY = 0.3, 0.1


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def workload_string(n_rows, dim, data, nq, k):
    """The TASK both arms run (identical in the `ours` and `reference` lines); the engine that serves it is named in
    `config.engine`."""
    cfg_no = "1 and 2" if (n_rows <= 1_000_000 and dim == 768) else ("3" if dim == 768 else "4")
    return ("kNN search %dx%d cosine (%s data), %d-query batch, k=%d, recall@%d>=0.95 (BASELINE configs[%s])"
            % (n_rows, dim, data, nq, k, k, cfg_no))


def quiet_nccl_stdout():
    """stdout carries ONE JSON line: NCCL's log goes to stderr, and its bare version banner (printed to stdout at
    NCCL_DEBUG=VERSION) is dropped; an explicit NCCL_DEBUG=INFO/TRACE from the caller is kept."""
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


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


def recall_at_k(got, want):
    return float(np.mean([len(set(a) & set(b)) / want.shape[1] for a, b in zip(got.tolist(), want.tolist())]))


def order_mismatches_beyond_ties(got_ids, want_ids, got_scores, want_scores, rel=1e-5, atol=1e-7):
    """north_star's parity rule for scored lists: scores within 1e-5 relative, ids equal wherever the oracle's neighbouring
    scores differ by more than that. Returns the number of positions that BREAK the rule (0 = parity)."""
    gs, ws = np.asarray(got_scores, np.float64), np.asarray(want_scores, np.float64)
    fin = np.isfinite(ws)
    bad = int((fin != np.isfinite(gs)).sum())
    tol = rel * np.abs(np.where(fin, ws, 0.0)) + atol
    bad += int((np.abs(np.where(fin, gs - ws, 0.0)) > tol).sum())
    k = want_ids.shape[1]
    for r, j in zip(*np.nonzero(np.asarray(got_ids) != np.asarray(want_ids))):
        near = j == k - 1
        if j > 0:
            near |= abs(ws[r, j] - ws[r, j - 1]) <= tol[r, j]
        if j + 1 < k:
            near |= abs(ws[r, j] - ws[r, j + 1]) <= tol[r, j]
        bad += 0 if near else 1
    return bad


def frontier_ef(per_ef, gate=0.95):
    """Smallest ef whose recall@10 reaches the gate (the metric's operating point); None when none does."""
    ok = sorted(int(e) for e, r in per_ef.items() if r["recall_at_10"] >= gate)
    return ok[0] if ok else None


class ClockSampler(threading.Thread):
    """SM clock / power / throttle reasons sampled DURING the timed regions, in-process through NVML every 5 ms
    (a 0.25 s region yields ~50 samples); falls back to polling nvidia-smi when NVML cannot be loaded."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index, uuid=None):
        super().__init__(daemon=True)
        self.gpu, self.uuid = gpu_index, uuid
        self.rows, self._stop_evt, self._active = [], threading.Event(), threading.Event()
        self.h = self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            if uuid:
                try:
                    self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
                except Exception:
                    self.h = None
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:                                   # noqa: BLE001
            log("[bench] NVML unavailable (%r): sampling nvidia-smi instead" % (e,))
            self.nvml = None

    def region(self, on):
        """Only samples taken while a timed region is open are kept."""
        (self._active.set if on else self._active.clear)()

    def _sample(self):
        if self.nvml is not None:
            n = self.nvml
            sm = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
            try:
                mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            return sm, self.max_sm, float(n.nvmlDeviceGetPowerUsage(self.h)) / 1000.0, mask
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout
        p = [x.strip() for x in out.strip().split(",")]
        mask = sum(bit for (name, bit), v in zip(self.REASONS, p[3:7]) if v.lower().startswith("active"))
        return float(p[0]), float(p[1]), float(p[2]), mask

    def run(self):
        while not self._stop_evt.is_set():
            if self._active.is_set():
                try:
                    self.rows.append(self._sample())
                except Exception:
                    pass
                self._stop_evt.wait(0.005 if self.nvml is not None else 0.05)
            else:
                self._stop_evt.wait(0.002)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"], "samples": 0}
        sm = sorted(r[0] for r in self.rows)
        mask = 0
        for r in self.rows:
            mask |= r[3]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.rows[0][1],
                "reasons": [name for name, bit in self.REASONS if mask & bit],
                "power_w_max": max(r[2] for r in self.rows), "samples": len(self.rows),
                "source": "nvml (in-process, 5 ms period, timed regions only)" if self.nvml is not None else "nvidia-smi"}


def graph_cache_path(n_rows, dim, data, efc, rank, world):
    """Both arms search the SAME graph. The GPU-assisted build is deterministic, so either arm can rebuild it; whichever
    runs first on a box also leaves the sidecar here so the other one can skip the build."""
    return "/tmp/chronomind_bench_%dx%d_%s_efc%d_seed%x_shard%dof%d.cmg" % (n_rows, dim, data, efc, SEED, rank, world)


def build_or_load_graph(st, args, path, seed, device, threads, have_cuda):
    """(how, seconds). GPU-assisted build when a device is there, the host replay of `LockFreeHnsw::insert` otherwise."""
    from chronomind_b200.engine import CmError
    t0 = time.time()
    if not args.no_graph_cache and os.path.exists(path):
        try:
            st.load_graph(path)
            return "sidecar %s (written by the other arm's build on this box)" % path, time.time() - t0
        except CmError as e:
            log("[bench] stale graph sidecar ignored: %s" % e)
    if have_cuda and args.build == "gpu":
        st.build_graph_gpu(seed, device, threads)
        how = "gpu-assisted (predecessor kNN, select_diverse and reverse-edge re-selection on the device; reachability check on the host)"
    else:
        st.build_graph(seed, threads)
        how = "host (LockFreeHnsw::insert restatement, %d threads)" % threads
    secs = time.time() - t0
    if not args.no_graph_cache:
        try:
            st.save_graph(path + ".tmp%d" % os.getpid())
            os.replace(path + ".tmp%d" % os.getpid(), path)
        except Exception as e:                                   # noqa: BLE001
            log("[bench] could not cache the graph sidecar: %r" % (e,))
    return how, secs


def oracle_on_same_graph(st, shard, efc, threads):
    """The CPU oracle over exactly the graph the engine searches (sidecar export -> import)."""
    import oracle
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "graph.cmg")
        st.save_graph(path)
        g = oracle.read_graph_sidecar(path)
    orc = oracle.OracleHnsw(16, efc, 50, SEED)
    orc.import_graph(oracle.preprocess_rows(shard, threads=threads), g)
    return orc


def cpu_hnsw_best_qps(orc, qs, ef, k, threads, passes=3):
    """`best_qps` (benches/external.rs:118-124): best of `passes` timed passes after one untimed."""
    orc.search_batch(qs[:max(64, len(qs) // 8)], ef, k, threads=threads)
    best, res = None, None
    for _ in range(passes):
        t1 = time.perf_counter()
        res = orc.search_batch(qs, ef, k, threads=threads)
        dt = time.perf_counter() - t1
        best = dt if best is None else min(best, dt)
    return len(qs) / best, best, res


# =====================================================================================================================
# --impl reference
# =====================================================================================================================
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path on this box's host cores. Production read path = HNSW
    (`LockFreeHnsw::search` under `ChronoMind::search`), so THAT is `value`; the brute-force scan the reference uses as
    ground truth is reported beside it."""
    if rank != 0:
        return
    import oracle
    from chronomind_b200.engine import ChronoMindB200, default_config
    threads = os.cpu_count() or 1
    n_rows, nq = args.rows, args.queries
    t_start = time.time()
    try:
        import torch
        have_cuda = torch.cuda.is_available()
    except Exception:
        have_cuda = False
    shard, queries_all, gen = make_data(n_rows, nq * QUERY_BATCHES, args.data, 0, 1)
    cfg = default_config(dimensions=DIM, ef_construction=args.efc)
    st = ChronoMindB200.from_matrix(shard, cfg)
    how, build_s = build_or_load_graph(st, args, graph_cache_path(n_rows, DIM, args.data, args.efc, 0, 1), SEED, 0, threads, have_cuda)
    log("[bench] reference arm graph: %s in %.1fs, invariants=%d" % (how, build_s, st.check_invariants()))
    orc = oracle_on_same_graph(st, shard, args.efc, threads)
    prepared_cache = {}
    # operating point: smallest ef with recall@10 >= 0.95 against the reference's own brute force, on a bounded sample
    rs = min(args.recall_sample, nq)
    t0 = time.time()
    truth, _, bf_secs = cpu_scan_streaming(gen, queries_all[:rs], K, threads, prepared_cache if n_rows * DIM * 4 <= 8e9 else None)
    brute_qps = rs / bf_secs
    log("[bench] brute force (ground truth) on %d queries: %.1f q/s (%.1fs)" % (rs, brute_qps, time.time() - t0))
    # The traversal is deterministic and identical on CPU and GPU (same graph, same arithmetic), so recall at a given ef
    # is ONE number for both arms; this arm can only afford a small truth sample, so the gate is applied with the
    # sample's two-sigma allowance — in the reference's favour (a smaller ef is faster).
    sweep = {}
    ef_star = None
    for ef in ([args.ef] if args.ef else sorted({int(e) for e in args.efs.split(",")})):
        if ef < K:
            continue
        ids, _, _, _ = orc.search_batch(queries_all[:rs], ef, K, threads=threads)
        per_q = np.array([len(set(a) & set(b)) / K for a, b in zip(ids.tolist(), truth.tolist())])
        sigma = float(per_q.std(ddof=1) / np.sqrt(len(per_q))) if len(per_q) > 1 else 0.0
        sweep[str(ef)] = {"recall_at_10": float(per_q.mean()), "stderr": sigma}
        if ef_star is None and per_q.mean() + 2.0 * sigma >= 0.95:
            ef_star = ef
            break
    if ef_star is None:
        ef_star = max(int(e) for e in sweep)
    del prepared_cache
    sample = min(args.cpu_sample or 2000, nq)                     # SEARCH_OPS-sized slices (benches/comparison.rs:35)
    for s in range(args.warmup):
        orc.search_batch(queries_all[(s * sample) % max(1, nq - sample):][:sample], ef_star, K, threads=threads)
    total = 0.0
    for s in range(args.steps):
        q = queries_all[(s * sample) % max(1, len(queries_all) - sample):][:sample]
        t1 = time.perf_counter()
        orc.search_batch(q, ef_star, K, threads=threads)
        total += time.perf_counter() - t1
    qps = sample * args.steps / total
    rec = sweep.get(str(ef_star), {}).get("recall_at_10")
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * total / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(n_rows, DIM, args.data, nq, K),
                   "engine": "CPU LockFreeHnsw::search (oracle port, AVX2+FMA), ef=%d = smallest ef whose recall@10 reaches 0.95 within "
                             "two standard errors of the %d-query truth sample, %d host threads" % (ef_star, rs, threads),
                   "recall_at_10": rec, "ef": ef_star, "ef_sweep_recall": sweep,
                   "graph": how, "graph_build_s": build_s, "brute_force_qps": brute_qps,
                   "brute_force_note": "tests/recall_test.rs:70-79 linear scan, the reference's ground-truth helper (not its serving path)",
                   "setup_s": time.time() - t_start},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": "%d-query slices of the %d-query batch per step x %d steps, ef=%d, same graph as the GPU arm"
                                   % (sample, nq, args.steps, ef_star)},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# =====================================================================================================================
# --impl ours
# =====================================================================================================================
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from chronomind_b200 import datasets, engine, sharding
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
    peaks = load_peaks()
    threads_all = os.cpu_count() or 1
    threads = max(1, threads_all // world)
    n_rows, nq = args.rows, args.queries
    want_hnsw = not args.no_hnsw and DIM == 768
    if args.workload == "hnsw" and not want_hnsw:
        raise SystemExit("--workload hnsw needs the 768-d corpus and the HNSW block")
    want_exact = args.workload == "exact" or want_hnsw          # the exact scan is also the ground truth of the HNSW block

    shard, queries_all, gen = make_data(n_rows, nq * QUERY_BATCHES, args.data, rank, world)
    ts_s, ts_n, decay = datasets.temporal_attrs(n_rows, 1, NOW[0])              # indexed by GLOBAL handle
    t0 = time.time()
    cfg = default_config(dimensions=DIM, ef_construction=args.efc)
    use_snapshot = world == 1 and (args.source == "snapshot" or (args.source == "auto" and n_rows * DIM * 4 <= (4 << 30)))
    source, ctx_rows = "matrix handed to cm_index_from_matrix", None
    if use_snapshot:
        # BASELINE configs[2]: "from reference snapshot". The synthetic store goes through the reference's own file format:
        # save_snapshot's container (src/persistence.rs:41-64) written by datasets.write_snapshot, read back by the
        # engine's load_snapshot restatement (cm_index_from_snapshot: header, CRC-32, bincode body, validate per record).
        ctx_rows = np.random.default_rng(SEED ^ 0xC7).integers(0, max(1, args.contexts), size=n_rows).astype(np.uint32)
        snap_path = os.path.join(os.environ.get("TMPDIR", "/tmp"), "chronomind_bench_%d_%dx%d.chrono" % (os.getpid(), n_rows, DIM))
        st = None
        try:
            tw = time.time()
            info = datasets.write_snapshot(snap_path, shard, ts_s, ts_n, decay, ctx_rows, max(1, args.contexts),
                                           config={"ef_construction": args.efc, "max_memories": max(n_rows, 100_000)})
            tw = time.time() - tw
            tl = time.time()
            st = ChronoMindB200.from_snapshot(snap_path)
            tl = time.time() - tl
        except OSError as e:                      # no room for a 3 GB scratch file: the same rows go in as a matrix
            if args.source == "snapshot":
                raise
            log("[bench] snapshot file not written (%s): falling back to cm_index_from_matrix" % e)
            st, use_snapshot, ctx_rows = None, False, None
            source = "matrix handed to cm_index_from_matrix (the scratch snapshot file could not be written: %s)" % e
        finally:
            if os.path.exists(snap_path):
                os.unlink(snap_path)
    if use_snapshot:
        assert len(st) == n_rows and st.config.dimensions == DIM and st.external_id(n_rows - 1) == "m%09d" % (n_rows - 1)
        source = ("CHRONO1 v2 snapshot file, %.2f GB, %d records over %d contexts: written in %.1fs, loaded by cm_index_from_snapshot "
                  "in %.1fs (%.2f GB/s: read + CRC-32 + bincode decode + validate + preprocess)"
                  % (info["bytes"] / 1e9, n_rows, max(1, args.contexts), tw, tl, info["bytes"] / 1e9 / tl))
        log("[bench] source: " + source)
    else:
        st = ChronoMindB200.from_matrix(shard, cfg, sharding.shard_attrs(ts_s, rank, world), sharding.shard_attrs(ts_n, rank, world),
                                        sharding.shard_attrs(decay, rank, world)).set_shard(rank, world)
    graph_how = graph_s = None
    if want_hnsw:
        graph_how, graph_s = build_or_load_graph(st, args, graph_cache_path(n_rows, DIM, args.data, args.efc, rank, world),
                                                 SEED + rank, local_rank, threads, True)
        log("[bench] rank %d graph: %s, %d rows, M=16 efC=%d, %.1fs, invariants=%d unreachable=%d"
            % (rank, graph_how, shard.shape[0], args.efc, graph_s, st.check_invariants(), st.unreachable()))
    st.upload(local_rank)
    st.set_exact_engine(args.engine)
    log("[bench] rank %d: index of %d rows resident on cuda:%d in %.1fs" % (rank, len(st), local_rank, time.time() - t0))

    q_host = [torch.from_numpy(queries_all[i * nq:(i + 1) * nq]).pin_memory() for i in range(QUERY_BATCHES)]
    q_dev = [q.to(dev) for q in q_host]
    n_streams = max(1, args.streams)
    # one notch above the default (= lowest) priority: the library runs each batch's tail (rescore, fallback) on a
    # default-priority side stream, so the next batch's scan gets the SMs first and the tail fills in underneath
    streams = [torch.cuda.Stream(dev, priority=-1) for _ in range(n_streams)] if n_streams > 1 else []
    ids_s = [torch.empty((nq, K), dtype=torch.int32, device=dev) for _ in range(n_streams)]      # one result buffer per
    dist_s = [torch.empty((nq, K), dtype=torch.float32, device=dev) for _ in range(n_streams)]   # batch in flight
    ids_l, dist_l = ids_s[0], dist_s[0]
    out_ids_host = torch.empty((nq, K), dtype=torch.int32).pin_memory()
    out_dist_host = torch.empty((nq, K), dtype=torch.float32).pin_memory()
    # N > 1: every rank needs the whole batch on its GPU. Each rank copies 1/N of it from pinned host memory over
    # ITS PCIe link and the slices are all-gathered over NVLink (30.7 MB at 10k x 768) — the node's aggregate host
    # bandwidth instead of N full copies.
    rows_per_rank = (nq + world - 1) // world
    q_gather = torch.empty((world * rows_per_rank, DIM), dtype=torch.float32, device=dev) if world > 1 else None
    q_slice = torch.empty((rows_per_rank, DIM), dtype=torch.float32, device=dev) if world > 1 else None

    def merged(ids, d, length):
        if world == 1:
            return ids, d
        # one packed (distance, local handle) array per rank over NVLink, merged by (distance, global handle)
        return engine.merge_topk_keys_cuda(sharding.allgather_keys(engine.pack_topk_cuda(ids, d)), length)

    def gather_queries(i):
        qh = q_host[i % QUERY_BATCHES]
        lo = min(nq, rank * rows_per_rank)
        hi = min(nq, lo + rows_per_rank)
        q_slice[:hi - lo].copy_(qh[lo:hi], non_blocking=True)
        dist.all_gather_into_tensor(q_gather, q_slice)
        return q_gather[:nq]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    uuid = None
    try:
        uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        pass
    sampler = ClockSampler(local_rank, uuid) if rank == 0 else None
    if sampler:
        sampler.start()

    def timed_device(step_fn, warmup, steps, pipelined=False):
        """W untimed + K timed steps: barrier + synchronize on both sides, CUDA events on the launching stream, max over
        ranks. `pipelined`: consecutive batches go to alternating CUDA streams (a serving loop with two batches in
        flight: the rescore / exchange / merge of batch i runs under the scan of batch i + 1); the events still bracket
        ALL the work — the side streams fork from the timing stream after ev0 and join it before ev1.
        Returns (ms, result of the last step, kernels launched, (dominant-kernel ms, launches))."""
        res = None
        cur = torch.cuda.current_stream(dev)
        use = streams if pipelined else []

        def run(i):
            if use:
                with torch.cuda.stream(use[i % len(use)]):
                    return step_fn(i)
            return step_fn(i)
        for i in range(warmup):
            run(i)
        barrier()
        engine.profile_enable(True)
        engine.profile_read()
        launches0 = engine.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if sampler:
            sampler.region(True)
        ev0.record()
        for sm in use:
            sm.wait_stream(cur)
        for i in range(steps):
            res = run(warmup + i)
        for sm in use:
            cur.wait_stream(sm)
        ev1.record()
        barrier()
        if sampler:
            sampler.region(False)
        ms = max_over_ranks(ev0.elapsed_time(ev1))
        launches = engine.launch_count() - launches0
        prof = engine.profile_read()
        engine.profile_enable(False)
        return ms, res, launches, prof

    def timed_host(step_fn, warmup, steps, callers=1):
        """K timed calls of a blocking host-buffer entry point. `callers` > 1: that many caller threads share the index
        (`&self` + Send + Sync; the reference's own benches time search from T threads, benches/comparison.rs:125-139) —
        one thread's H2D / D2H then overlaps another's kernels. Every call still moves its own batch in and its own
        result out inside the timed region. Returns (seconds, result of the last call)."""
        out = [None]

        def run(first, count):
            if callers <= 1:
                for i in range(first, first + count):
                    out[0] = step_fn(i)
                return
            errors = []

            def worker(t):
                try:
                    for i in range(first + t, first + count, callers):
                        r = step_fn(i)
                        if i == first + count - 1:
                            out[0] = r
                except BaseException as e:                       # noqa: BLE001 — re-raised on the timing thread
                    errors.append(e)
            ts = [threading.Thread(target=worker, args=(t,)) for t in range(callers)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
            if errors:
                raise errors[0]
        # warm up at the SAME concurrency: every caller thread's workspace (hundreds of MB of scratch, allocated on first
        # use) must exist before the timed region, not be cudaMalloc'ed inside it
        run(0, max(min(warmup, 3), 2 * callers))
        barrier()
        if sampler:
            sampler.region(True)
        t0 = time.perf_counter()
        run(warmup, steps)
        torch.cuda.synchronize()
        secs = time.perf_counter() - t0
        if sampler:
            sampler.region(False)
        return max_over_ranks(secs), out[0]

    e2e_callers = max(1, args.callers) if world == 1 else 1      # N > 1: the e2e leg drives torch collectives from one thread

    # ---- exact scan (configs[1]; ground truth of everything else) ---------------------------------------------------
    exact = None
    truth = None
    if want_exact:
        def step_device(i):
            """Inputs resident in HBM; result stays on the device."""
            b = i % n_streams
            st.search_exact_cuda(q_dev[i % QUERY_BATCHES], K, ids_s[b], dist_s[b])
            return merged(ids_s[b], dist_s[b], K)

        def step_e2e(i):
            """Public API with HOST buffers: H2D of the query batch and D2H of the result inside the step."""
            if world == 1:
                return st.search_exact(q_host[i % QUERY_BATCHES].numpy(), K)
            st.search_exact_cuda(gather_queries(i), K, ids_l, dist_l)
            mi, md = merged(ids_l, dist_l, K)
            if rank == 0:                                   # the caller reads the merged result once
                out_ids_host.copy_(mi, non_blocking=True)
                out_dist_host.copy_(md, non_blocking=True)
            torch.cuda.synchronize()
            return out_ids_host, out_dist_host

        # Pass 1, one batch in flight: the dominant kernel's own duration (CUDA events around each launch) for the roofline.
        ms_seq, res, launches, (kernel_ms, kernel_n) = timed_device(step_device, args.warmup, args.steps, pipelined=False)
        ms, kernel_ms_pipe, kernel_n_pipe, ms_pipe, in_flight = ms_seq, None, None, None, 1
        if n_streams > 1:
            # Pass 2, two batches in flight: same K steps over alternating streams. A kernel's event bracket now also
            # contains whatever co-runs with it, so it is reported but not used for the roofline. The headline quotes the
            # faster of the two serving modes and says which (overlap pays when the per-batch tail is a visible share of
            # the step — small shards; with long scans on many ranks the desynchronised collectives cost more than it hides).
            ms_pipe, res_p, launches_p, (kernel_ms_pipe, kernel_n_pipe) = timed_device(step_device, args.warmup, args.steps, pipelined=True)
            if ms_pipe < ms_seq:
                ms, res, launches, in_flight = ms_pipe, res_p, launches_p, n_streams
        last_batch = (args.warmup + args.steps - 1) % QUERY_BATCHES
        dev_ids = res[0].cpu().numpy().view(np.uint32).copy()      # what the timed device path returned last
        dev_dist = res[1].cpu().numpy().copy()
        unresolved_status = st.dev_status()                        # queries the async path could not resolve (0 normally)
        dev_ctr = st.dev_counters()                                # of every device-resident step so far (warm-up included)
        e2e_s, _ = timed_host(step_e2e, args.warmup, args.steps, callers=1)
        e2e_modes = {"1": nq * args.steps / e2e_s}
        if e2e_callers > 1:                                    # both serving modes are timed; the line quotes the faster
            e2e_s2, _ = timed_host(step_e2e, args.warmup, args.steps, callers=e2e_callers)
            e2e_modes[str(e2e_callers)] = nq * args.steps / e2e_s2
            e2e_s = min(e2e_s, e2e_s2)
        if world == 1:
            step_e2e(0)                                            # once more on THIS thread: its counters are thread-local
        counters = engine.last_counters() if world == 1 else {}    # of the last host-buffer call (the e2e leg)
        truth = [None] * QUERY_BATCHES
        truth[last_batch] = dev_ids
        tc = (counters.get("exact_engine") == 2) if world == 1 else (args.engine != 1 and shard.shape[0] >= 8192 and K <= 128)
        unresolved = int((dev_ids[:, 0] == 0xFFFFFFFF).sum() + np.isnan(dev_dist[:, 0]).sum())
        flops_per_launch = 2.0 * nq * shard.shape[0] * DIM * args.steps / max(kernel_n, 1)
        kern_avg_ms = kernel_ms / max(kernel_n, 1)
        achieved = flops_per_launch / (kern_avg_ms * 1e-3) / 1e12 if kern_avg_ms > 0 else None
        burst = peaks.get("bf16_tflops") or 1650.0
        sustained = peaks.get("bf16_tflops_sustained") or 1400.0
        exact = {
            "ms": ms, "ms_seq": ms_seq, "ms_pipe": ms_pipe, "in_flight": in_flight, "e2e_modes": e2e_modes, "qps": nq * args.steps / (ms * 1e-3), "e2e_qps": nq * args.steps / e2e_s, "launches": launches,
            "dev_ids": dev_ids, "dev_dist": dev_dist, "last_batch": last_batch, "tc": tc, "unresolved": unresolved + unresolved_status,
            "counters": counters, "dev_counters": dev_ctr, "dev_steps": args.warmup + args.steps,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": burst, "unit": "TFLOP/s",
                         "frac": (achieved / burst) if achieved else None,
                         "frac_burst": (achieved / burst) if achieved else None,
                         "frac_sustained": (achieved / sustained) if achieved else None, "peak_sustained": sustained,
                         "traffic": measured_traffic("scan_tc_kernel", rows=n_rows, queries=nq, n_gpus=world, data=args.data) if tc else None,
                         "traffic_unit": "DRAM bytes per launch (ncu); algorithmic minimum = one pass over the bf16 corpus + queries = %d"
                                         % (shard.shape[0] * DIM * 2 + nq * DIM * 2),
                         "kernel": "scan_tc (tcgen05 bf16)" if tc else "k_exact_dist (fp32 CUDA cores, bit-exact chain)",
                         "kernel_ms_avg": kern_avg_ms, "kernel_launches_per_step": kernel_n / args.steps,
                         "timed_in": "the one-batch-in-flight pass of K steps (%.3f ms/step); in the two-in-flight pass the kernel's "
                                     "event bracket also holds the previous batch's co-running tail: %s ms"
                                     % (ms_seq / args.steps, ("%.3f" % (kernel_ms_pipe / max(kernel_n_pipe, 1))) if kernel_ms_pipe else "n/a")
                                     if n_streams > 1 else "the timed region",
                         "algorithmic_flop_per_launch": flops_per_launch,
                         "scores_per_s_per_gpu": (nq * shard.shape[0] / (kern_avg_ms * 1e-3)) if kern_avg_ms > 0 else None,
                         "note": None if DIM >= 256 else "short rows: one or two UMMA k-steps per tile, the TMEM-epilogue select bounds "
                                                         "this shape (SURVEY.md H6); read scores_per_s_per_gpu, not frac",
                         "peak_source": ("MEASURED_PEAKS.json bf16_tflops (burst: the timed region is a fraction of a second); "
                                         "frac_sustained uses bf16_tflops_sustained") if peaks else "fallback of B200_PROFILING.md"}}
        if rank == 0:
            log("[bench] exact: %.0f q/s device, %.0f q/s e2e, K1 %.0f TFLOP/s; counters of the last e2e call: %s; unresolved: %d"
                % (exact["qps"], exact["e2e_qps"], achieved or 0, counters, exact["unresolved"]))

    # ---- HNSW block (configs[2]) -------------------------------------------------------------------------------------
    hnsw = None
    if want_hnsw:
        hbm_peak = peaks.get("hbm_gbs") or 6650.0
        g_ts_s = torch.from_numpy(ts_s.view(np.int64)).to(dev)
        g_ts_n = torch.from_numpy(ts_n.view(np.int32)).to(dev)
        g_decay = torch.from_numpy(decay).to(dev)
        efs = sorted({int(e) for e in args.efs.split(",")} | ({args.ef} if args.ef else set()))
        efs = [e for e in efs if e >= K]
        nb = 2                                                       # query batches rotated by the HNSW legs
        for b in range(nb):                                          # exact ground truth for the batches used
            if truth is None or truth[b] is None:
                t_ids = merged(*st.search_exact_cuda(q_dev[b], K), K)[0].cpu().numpy().view(np.uint32)
                if truth is None:
                    truth = [None] * QUERY_BATCHES
                truth[b] = t_ids
        h_steps, h_warm = max(3, min(args.steps, args.hnsw_steps)), min(args.warmup, 3)
        per_ef = {}
        for ef in efs:
            def step_device(i, ef=ef):
                st.search_index_cuda(q_dev[i % nb], ef, K, ids_l, dist_l)
                return merged(ids_l, dist_l, K)

            def step_e2e_index(i, ef=ef):
                """`VectorIndex::search(q, ef)` + take(10) through host buffers — the call the reference arm times."""
                return st.search_index(q_host[i % nb].numpy(), ef, K)

            def step_e2e_temporal(i, ef=ef):
                """`ChronoMind::search` through host buffers: pool of max(ef, 3k), (N > 1: merge to the global pool,)
                fused temporal rerank w = 0.3."""
                if world == 1:
                    return st.search(q_host[i % nb].numpy(), K, now=NOW, ef_search=ef, temporal_weight=TEMPORAL_W, base_decay_rate=BASE_DECAY)
                pool = max(ef, 3 * K)                                    # OVERSAMPLE, src/store.rs:40,323
                pi, pd = st.search_index_cuda(gather_queries(i), pool, pool)
                mi, md = merged(pi, pd, pool)                            # global top-ef first ...
                ri, rs_, _ = engine.rerank_cuda(mi, md, K, g_ts_s, g_ts_n, g_decay, TEMPORAL_W, BASE_DECAY, NOW)   # ... then the blend
                out_ids_host.copy_(ri, non_blocking=True)
                out_dist_host.copy_(rs_, non_blocking=True)
                torch.cuda.synchronize()
                return out_ids_host, out_dist_host

            ms, res, launches, (kernel_ms, kernel_n) = timed_device(step_device, h_warm, h_steps)
            last = (h_warm + h_steps - 1) % nb
            got = res[0].cpu().numpy().view(np.uint32)
            e2e_steps = 4
            e2e_t, _ = timed_host(step_e2e_temporal, 1, e2e_steps, callers=e2e_callers)
            e2e_i = None
            if world == 1:
                e2e_i, _ = timed_host(step_e2e_index, 1, e2e_steps, callers=e2e_callers)
                step_e2e_index(last)
                c = engine.last_counters()                               # of a host-buffer index search on this thread == one batch
            else:
                st.search_index(q_host[last].numpy(), ef, K)
                c = engine.last_counters()
            bytes_per_batch = c["dist_evals"] * DIM * 4 + c["list_entries"] * 4 + nq * DIM * 4
            kern_avg = kernel_ms / max(kernel_n, 1)
            per_ef[str(ef)] = {
                "ef": ef, "qps": nq * h_steps / (ms * 1e-3), "ms_per_step": ms / h_steps, "recall_at_10": recall_at_k(got, truth[last]),
                "dist_evals_per_query": c["dist_evals"] / nq, "expansions_per_query": c["expansions"] / nq,
                "algorithmic_bytes_per_query": bytes_per_batch / nq, "kernel_ms_avg": kern_avg,
                "achieved_gbs": bytes_per_batch / (kern_avg * 1e-3) / 1e9 if kern_avg > 0 else None,
                "hbm_frac": bytes_per_batch / (kern_avg * 1e-3) / 1e9 / hbm_peak if kern_avg > 0 else None,
                "e2e_index_qps": (nq * e2e_steps / e2e_i) if e2e_i else None, "e2e_temporal_qps": nq * e2e_steps / e2e_t,
                "gpu_launches_per_step": launches / h_steps, "visited_table_restarts": c["fallback_queries"],
                "tie_replays": c.get("tie_replays", 0), "failures": c.get("failures", 0)}
            if rank == 0:
                log("[bench] hnsw ef=%d: %s" % (ef, json.dumps(per_ef[str(ef)])))
        ef_star = args.ef or frontier_ef(per_ef) or max(efs)
        head = per_ef[str(ef_star)]
        hnsw = {
            "workload": "HNSW batched search %dx%d cosine (%s), %d-query batch, k=%d, M=16 efC=%d graph%s (BASELINE configs[2]); "
                        "temporal rerank w=%.1f in e2e_temporal" % (n_rows, DIM, args.data, nq, K, args.efc,
                                                                      "" if world == 1 else " per shard, %d shards by id" % world, TEMPORAL_W),
            "frontier": {"ef": ef_star, "rule": "smallest measured ef with recall@10 >= 0.95 against the exact scan", "qps": head["qps"],
                         "recall_at_10": head["recall_at_10"], "e2e_index_qps": head["e2e_index_qps"],
                         "e2e_temporal_qps": head["e2e_temporal_qps"], "hbm_frac": head["hbm_frac"]},
            "per_ef": per_ef, "steps_per_ef": h_steps, "warmup_per_ef": h_warm,
            "graph_build": graph_how, "graph_build_s": graph_s, "graph_build_inserts_per_s": (shard.shape[0] / graph_s) if graph_s else None,
            "build_threads": threads,
            "roofline": {"bound": "hbm", "achieved": head["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": head["hbm_frac"],
                         "ef": ef_star,
                         "traffic": measured_traffic("k_hnsw_search", rows=n_rows, queries=nq, n_gpus=world, data=args.data, ef=ef_star),
                         "traffic_unit": "DRAM bytes per launch (ncu) vs algorithmic %d" % int(head["algorithmic_bytes_per_query"] * nq),
                         "kernel": "k_hnsw_search (rank 0's shard)", "kernel_ms_avg": head["kernel_ms_avg"],
                         "algorithmic_bytes": "dist_evals*dim*4 + list_entries*4 + dim*4 per query, counters == CPU oracle's at equal ef",
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback of B200_PROFILING.md"},
            "l2": "%.2f GB fp32 rows per GPU gathered at random: larger than L2; %d query batches rotated" % (shard.shape[0] * DIM * 4 / 1e9, nb)}
    # ---- search_in_context block (only a snapshot-loaded store carries contexts) -------------------------
    context = None
    if ctx_rows is not None and args.contexts > 1 and K <= 128:
        c_steps = max(3, min(args.steps, 10))
        rngc = np.random.default_rng(SEED ^ 0xC8)
        want_labels = [rngc.integers(0, args.contexts, size=nq) for _ in range(2)]
        label_to_id = np.array([st.context_id("ctx%05d" % c) for c in range(args.contexts)], dtype=np.uint32)   # ids = order of first appearance
        want_ids = [label_to_id[w] for w in want_labels]

        def step_ctx(i):
            return st.search_in_context(want_ids[i % 2], q_host[i % 2].numpy(), K, now=NOW, temporal_weight=TEMPORAL_W,
                                        base_decay_rate=BASE_DECAY)
        step_ctx(0)
        engine.profile_enable(True)
        engine.profile_read()
        c_secs, c_res = timed_host(step_ctx, 2, c_steps, callers=1)
        c_kms, c_kn = engine.profile_read()
        engine.profile_enable(False)
        c_res = step_ctx(1)
        cc = engine.last_counters()
        context = {"workload": "ChronoMind::search_in_context, %d-query batch over %d contexts of the %dx%d store, k=%d, w=%.1f, host buffers"
                               % (nq, args.contexts, n_rows, DIM, K, TEMPORAL_W),
                   "e2e_qps": nq * c_steps / c_secs, "ms_per_batch": 1000.0 * c_secs / c_steps, "engine": cc["exact_engine"],
                   "fallback_queries": cc["fallback_queries"], "rescored_per_query": cc["rescored"] / nq,
                   "scan_kernel_ms_avg": c_kms / max(c_kn, 1),
                   "layout": "rows sorted by context on the device, queries grouped by context per call, per-pair tile ranges",
                   "_last": (want_labels[1], c_res)}
        log("[bench] search_in_context: %.0f q/s e2e (%.2f ms per batch, biased scan %.2f ms)"
            % (context["e2e_qps"], context["ms_per_batch"], context["scan_kernel_ms_avg"]))
    clocks = sampler.stop() if sampler else None

    # ---- CPU baseline + parity of the timed results (rank 0, N = 1) --------------------------------------------------
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu = {"unit": "queries/s", "cores": threads_all, "kind": "port"}
        t0 = time.time()
        if exact is not None:
            # the reference's brute force on a sample of the TIMED batch: parity check of the timed device result
            sample = min(args.cpu_sample or 256, nq)
            lb = exact["last_batch"]
            qs = queries_all[lb * nq:lb * nq + sample]
            wi, wd, secs = cpu_scan_streaming(gen, qs, K, threads_all)
            gi, gd = exact["dev_ids"][:sample], exact["dev_dist"][:sample]
            cpu.update({"brute_force_qps": sample / secs,
                        "brute_force_sample": "%d of the %d queries of the last timed batch against the full %dx%d corpus (%.1fs of scan time)"
                                              % (sample, nq, n_rows, DIM, secs),
                        "parity_on_sample": bool(np.array_equal(gi, wi) and np.array_equal(gd.view(np.uint32), wd.view(np.uint32))),
                        "recall_at_10_on_sample": recall_at_k(gi, wi)})
            if DIM != 768:
                log("[bench] note: unit vectors, so the L2 ranking of BASELINE configs[4] is this cosine ranking "
                    "(|a-b|^2 = 2(1 - a.b)); the reference has no L2 metric (SURVEY.md F3)")
        if hnsw is not None and world == 1:
            # the reference's production path on the same graph, at the frontier ef
            orc = oracle_on_same_graph(st, shard, args.efc, threads_all)
            ef_star = hnsw["frontier"]["ef"]
            hs = min(args.cpu_hnsw_sample, nq)
            qs = queries_all[:hs]
            qps, best, (oi, od, _, octr) = cpu_hnsw_best_qps(orc, qs, ef_star, K, threads_all)
            gi, gd = st.search_index(qs, ef_star, K)
            gc = engine.last_counters()
            par = {str(ef_star): bool(np.array_equal(gi, oi) and np.array_equal(gd.view(np.uint32), od.view(np.uint32))
                                      and gc["dist_evals"] == int(octr[0]) and gc["expansions"] == int(octr[1]))}
            for ef in (64, 200):
                oi2, od2, _, octr2 = orc.search_batch(qs[:512], ef, K, threads=threads_all)
                gi2, gd2 = st.search_index(qs[:512], ef, K)
                gc2 = engine.last_counters()
                par[str(ef)] = bool(np.array_equal(gi2, oi2) and np.array_equal(gd2.view(np.uint32), od2.view(np.uint32))
                                    and gc2["dist_evals"] == int(octr2[0]) and gc2["expansions"] == int(octr2[1]))
            cpu.update({"value": qps, "engine": "LockFreeHnsw::search (oracle port) at the frontier ef=%d on the same graph" % ef_star,
                        "sample": "%d queries, ef=%d, same graph (sidecar export -> oracle import), best of 3 passes (%.3fs each)"
                                  % (hs, ef_star, best),
                        "hnsw_parity_ids_dists_counters": par, "dist_evals_per_query": float(octr[0]) / hs})
            hnsw["cpu_baseline"] = {k: cpu[k] for k in ("value", "sample", "hnsw_parity_ids_dists_counters")}
        elif exact is not None:
            cpu.update({"value": cpu["brute_force_qps"], "sample": cpu["brute_force_sample"], "engine": "brute force (oracle port)"})
        if context is not None:
            import oracle
            want_l, (gi, gs) = context.pop("_last")
            ns = 64
            t1 = time.time()
            rc, wi, ws = oracle.search_in_context(shard, ctx_rows, queries_all[nq:nq + ns], want_l[:ns].astype(np.uint32), K, TEMPORAL_W,
                                                  BASE_DECAY, ts_s, ts_n, decay, NOW[0], NOW[1], threads=threads_all)
            bad = order_mismatches_beyond_ties(gi[:ns], wi, gs[:ns], ws) if rc == 0 else -1
            context["parity_on_sample"] = {"queries": ns, "oracle_rc": rc, "order_mismatches_beyond_1e-5_ties": bad,
                                           "cpu_qps": ns / (time.time() - t1)}
        log("[bench] cpu baseline + parity in %.1fs: %s" % (time.time() - t0, cpu))

    if rank == 0:
        wl = workload_string(n_rows, DIM, args.data, nq, K)
        common = {"metric": METRIC, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                  "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "data": "synthetic"}
        if args.workload == "exact":
            line = dict(common, value=exact["qps"], ms_per_step=exact["ms"] / args.steps, dtype="bf16" if exact["tc"] else "f32",
                        config={"workload": wl, "engine": ("tcgen05 bf16 scan + fp32 rescore (exact: recall@10 = 1.0)" if exact["tc"]
                                                           else "fp32 CUDA-core scan (exact)"),
                                "recall_at_10": (cpu or {}).get("recall_at_10_on_sample", "exact by construction (unchecked: --no-cpu-baseline)"),
                                "unresolved_queries": exact["unresolved"], "fallback_queries_e2e": exact["counters"].get("fallback_queries"),
                                "rescored_per_query_e2e": (exact["counters"].get("rescored", 0) / nq) if exact["counters"] else None,
                                "fallback_queries_device_steps": exact["dev_counters"]["fallback_queries"],
                                "rescored_per_query_per_shard_device_steps": exact["dev_counters"]["rescored"] / (nq * exact["dev_steps"]),
                                "parallelism": "corpus sharded by id over %d GPU(s), NCCL allgather + device merge" % world if world > 1 else "1 GPU",
                                "batches_in_flight": {"device": exact["in_flight"], "e2e_qps_by_caller_threads": exact["e2e_modes"],
                                                      "device_qps_one_in_flight": nq * args.steps / (exact["ms_seq"] * 1e-3),
                                                      "ms_per_step_one_in_flight": exact["ms_seq"] / args.steps,
                                                      "device_qps_two_in_flight": (nq * args.steps / (exact["ms_pipe"] * 1e-3)) if exact["ms_pipe"] else None,
                                                      "headline": "the faster of the two passes (both K timed steps after W warm-up steps)",
                                                      "note": "consecutive 10k-query batches alternate over CUDA streams (device) / caller threads "
                                                              "(e2e): the rescore, exchange and copies of one batch overlap the scan of the next; "
                                                              "--streams 1 --callers 1 measures one batch at a time"},
                                "l2": "inputs larger than L2: %.2f GB bf16 corpus per GPU streamed per step; %d query batches rotated"
                                      % (shard.shape[0] * DIM * 2 / 1e9, QUERY_BATCHES)},
                        e2e={"value": exact["e2e_qps"], "unit": "queries/s", "h2d_bytes_per_step": nq * DIM * 4, "d2h_bytes_per_step": nq * K * 8},
                        gpu_launches=exact["launches"], roofline=exact["roofline"])
        else:
            f = hnsw["frontier"]
            head = hnsw["per_ef"][str(f["ef"])]
            line = dict(common, value=head["qps"], ms_per_step=head["ms_per_step"], dtype="f32", steps=hnsw["steps_per_ef"],
                        warmup=hnsw["warmup_per_ef"],
                        config={"workload": wl, "engine": "batched HNSW traversal at ef=%d (frontier)" % f["ef"], "recall_at_10": head["recall_at_10"],
                                "parallelism": "1 GPU" if world == 1 else "independent HNSW sub-graph per shard on %d GPUs, NCCL allgather + device merge" % world,
                                "l2": hnsw["l2"]},
                        e2e={"value": head["e2e_index_qps"] or head["e2e_temporal_qps"], "unit": "queries/s", "h2d_bytes_per_step": nq * DIM * 4,
                             "d2h_bytes_per_step": nq * K * 8},
                        gpu_launches=int(head["gpu_launches_per_step"] * hnsw["steps_per_ef"]), roofline=hnsw["roofline"])
        line["cpu_baseline"] = cpu
        line["clocks"] = clocks
        if hnsw is not None:
            line["hnsw"] = hnsw
        if context is not None:
            context.pop("_last", None)
            line["search_in_context"] = context
        line["config"]["source"] = source
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# =====================================================================================================================
# --abi-group: the sharded path as ONE object behind the C ABI (cm_group_*), one process driving N devices
# =====================================================================================================================
def run_group(args):
    """`python bench.py --gpus N --abi-group`: no torchrun, no torch.distributed — `cm_group_create(devices)` loads NCCL
    inside the library, `cm_group_search_*` takes HOST buffers and returns merged global handles. Every number here is
    end to end (H2D of the batch, sharded search, NCCL all-gather + merge, D2H inside the timed region)."""
    import torch
    from chronomind_b200 import datasets
    from chronomind_b200.engine import ShardedChronoMindB200, default_config
    G = args.gpus
    quiet_nccl_stdout()                       # the library loads NCCL: keep its banner off the JSON line
    if torch.cuda.device_count() < G:
        raise SystemExit("--abi-group --gpus %d needs %d visible CUDA devices" % (G, G))
    n_rows, nq = args.rows, args.queries
    threads_all = os.cpu_count() or 1
    shard, queries_all, gen = make_data(n_rows, nq * QUERY_BATCHES, args.data, 0, 1)
    ts_s, ts_n, decay = datasets.temporal_attrs(n_rows, 1, NOW[0])
    cfg = default_config(dimensions=DIM, ef_construction=args.efc)
    t0 = time.time()
    grp = ShardedChronoMindB200.from_matrix(list(range(G)), shard, cfg, ts_s, ts_n, decay)
    want_hnsw = not args.no_hnsw and DIM == 768
    build_s = None
    if want_hnsw:
        t1 = time.time()
        grp.build_graphs(SEED, threads_all, gpu=args.build == "gpu")
        build_s = time.time() - t1
    grp.upload()
    log("[bench] cm_group over %d device(s): %d rows resident in %.1fs (graphs %.1fs)" % (G, n_rows, time.time() - t0, build_s or 0.0))
    q_host = [torch.from_numpy(queries_all[i * nq:(i + 1) * nq]).pin_memory().numpy() for i in range(QUERY_BATCHES)]

    def timed(fn, warmup, steps):
        for i in range(warmup):
            fn(i)
        t1 = time.perf_counter()
        out = None
        for i in range(steps):
            out = fn(warmup + i)
        return time.perf_counter() - t1, out

    secs, out = timed(lambda i: grp.search_exact(q_host[i % QUERY_BATCHES], K), args.warmup, args.steps)
    last = (args.warmup + args.steps - 1) % QUERY_BATCHES
    exact_qps = nq * args.steps / secs
    parity = None
    if not args.no_cpu_baseline:
        sample = min(args.cpu_sample or 64, nq)
        wi, wd, _ = cpu_scan_streaming(gen, queries_all[last * nq:last * nq + sample], K, threads_all)
        parity = bool(np.array_equal(out[0][:sample], wi) and np.array_equal(out[1][:sample].view(np.uint32), wd.view(np.uint32)))
    hn = None
    if want_hnsw:
        truth = out[0]
        per_ef = {}
        for ef in sorted({int(e) for e in args.efs.split(",")}):
            if ef < K:
                continue
            s1, o1 = timed(lambda i, ef=ef: grp.search_index(q_host[last], ef, K), 1, 3)
            s2, _ = timed(lambda i, ef=ef: grp.search(q_host[last], K, now=NOW, ef_search=ef, temporal_weight=TEMPORAL_W,
                                                      base_decay_rate=BASE_DECAY), 1, 3)
            per_ef[str(ef)] = {"ef": ef, "e2e_index_qps": nq * 3 / s1, "e2e_temporal_qps": nq * 3 / s2,
                               "recall_at_10": recall_at_k(o1[0], truth)}
            log("[bench] group hnsw ef=%d: %s" % (ef, json.dumps(per_ef[str(ef)])))
        ef_star = frontier_ef(per_ef) or max(int(e) for e in per_ef)
        hn = {"frontier": dict(per_ef[str(ef_star)], rule="smallest measured ef with recall@10 >= 0.95 against the exact scan"),
              "per_ef": per_ef, "graph_build_s": build_s, "graph_build": args.build}
    line = {"metric": METRIC, "value": exact_qps, "unit": "queries/s", "n_gpus": G, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * secs / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_string(n_rows, DIM, args.data, nq, K),
                       "engine": "cm_group_search_exact: tcgen05 bf16 scan + fp32 rescore per shard, NCCL all-gather + device merge inside the library",
                       "parallelism": "cm_group: ONE process, %d device(s), one host thread + stream per device, ncclCommInitAll inside libchronomind_b200.so" % G,
                       "parity_on_sample": parity, "timing": "host wall clock around the blocking C-ABI call (it synchronises before returning)"},
            "e2e": {"value": exact_qps, "unit": "queries/s", "h2d_bytes_per_step": nq * DIM * 4, "d2h_bytes_per_step": nq * K * 8},
            "gpu_launches": None, "hnsw": hn}
    print(json.dumps(line), flush=True)
    grp.close()


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
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the brute-force parity sample (0: 256) / per reference step (0: 2000)")
    ap.add_argument("--cpu-hnsw-sample", type=int, default=2048, help="queries timed through the CPU HNSW oracle in the ours arm")
    ap.add_argument("--recall-sample", type=int, default=256, help="reference arm: queries whose brute-force truth picks the operating ef")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hnsw", action="store_true", help="skip the HNSW block (graph build + ef sweep)")
    ap.add_argument("--source", default="auto", choices=["auto", "snapshot", "matrix"],
                    help="how the corpus reaches the engine at N = 1: 'snapshot' writes it as a CHRONO1 v2 file (save_snapshot's "
                         "container) and loads that through cm_index_from_snapshot, like a ChronoMind deployment's data; 'matrix' "
                         "hands the rows over directly; auto = snapshot up to 4 GB of rows")
    ap.add_argument("--contexts", type=int, default=64, help="contexts the snapshot's records are spread over (search_in_context block)")
    ap.add_argument("--no-graph-cache", action="store_true", help="neither read nor write the /tmp graph sidecar")
    ap.add_argument("--abi-group", action="store_true",
                    help="drive the N GPUs from ONE process through cm_group_* (NCCL inside the library) instead of torchrun ranks")
    ap.add_argument("--streams", type=int, default=2, help="device-resident leg: CUDA streams the batches alternate over (1: one batch at a time)")
    ap.add_argument("--callers", type=int, default=2, help="e2e leg (N = 1): caller threads sharing the index (1: one blocking call at a time)")
    ap.add_argument("--workload", default="exact", choices=["exact", "hnsw"],
                    help="which leg is the line's headline: exact (BASELINE configs[1], the metric's workload) or hnsw (configs[2])")
    ap.add_argument("--ef", type=int, default=0, help="hnsw: quote the headline at this ef instead of the recall frontier")
    ap.add_argument("--efs", default=DEFAULT_EFS, help="hnsw: every ef measured (reported under hnsw.per_ef)")
    ap.add_argument("--hnsw-steps", type=int, default=10, help="timed steps per ef of the sweep")
    ap.add_argument("--efc", type=int, default=100, help="hnsw: ef_construction of the graph")
    ap.add_argument("--data", default="emb", choices=["emb", "uniform"])
    ap.add_argument("--dim", type=int, default=DIM, help="vector dimension (default 768; 32 with --k 100 is BASELINE configs[4])")
    ap.add_argument("--k", type=int, default=K, help="neighbors per query (default 10)")
    ap.add_argument("--build", default="gpu", choices=["host", "gpu"],
                    help="hnsw: graph construction — gpu-assisted (predecessor kNN + select/link on the device, reachability on the "
                         "host) or host (LockFreeHnsw::insert restatement, all cores)")
    args = ap.parse_args()

    DIM, K = args.dim, args.k            # the workload shape is module state for the helpers above

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.abi_group:
        run_group(args)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
