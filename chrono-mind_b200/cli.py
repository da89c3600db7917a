"""`query` front-end of the reference CLI (`chronomind query`, src/main.rs:47-69,195-230) on the GPU engine.

    python -m chronomind_b200 query --file store.chrono --vector "[0.1, 0.2, 0.3, 0.4]" [--limit 10]
                                    [--context LABEL] [--normalize] [--device 0] [--seed S] [--threads T] [--build gpu|host]

Loads the snapshot (`load_snapshot`, src/persistence.rs:73-123), rebuilds the HNSW graph — GPU-assisted by default (predecessor kNN, select and link on the device),
`--build host` replays the reference's one-by-one insertion — or reads the `<file>.graph` sidecar when present,
uploads to the GPU and prints the reference
CLI's lines. There is no CPU search path: without a CUDA device the command fails.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np


def parse_vector(text: str) -> np.ndarray:
    """`parse_vector` (src/main.rs:259-272): a JSON array or comma-separated floats."""
    t = text.strip()
    if t.startswith("["):
        return np.asarray(json.loads(t), dtype=np.float32)
    try:
        return np.asarray([float(p.strip()) for p in t.split(",")], dtype=np.float32)
    except ValueError as e:
        raise SystemExit("Error: invalid vector data: bad component: %s" % e)


def normalize_vector(v: np.ndarray) -> np.ndarray:
    """`normalize_vector` (src/main.rs:274-281)."""
    n = np.sqrt(np.sum(v * v, dtype=np.float32), dtype=np.float32)
    return v / n if n > 0 else v


def query_command(args) -> int:
    from .engine import NO_ID, ChronoMindB200, CmError
    try:
        st = ChronoMindB200.from_snapshot(args.file)
        sidecar = args.file + ".graph"
        if os.path.exists(sidecar):
            st.load_graph(sidecar)
        elif not args.context:                       # search_in_context never touches the index
            if args.build == "gpu":                   # candidates from the device's all-pairs kNN, linking on the host
                st.build_graph_gpu(args.seed, args.device, args.threads)
            else:                                     # the reference's load path: LockFreeHnsw::insert, one by one
                st.build_graph(args.seed, max(1, args.threads))
        st.upload(args.device)
        q = parse_vector(args.vector)
        if args.normalize:
            q = normalize_vector(q)
        if args.context:
            ids, scores = st.search_in_context(args.context, q, args.limit)
        else:
            ids, scores = st.search(q, args.limit)
    except CmError as e:
        print("Error: %s" % e, file=sys.stderr)
        return 1
    hits = [(int(h), float(s)) for h, s in zip(ids[0], scores[0]) if h != NO_ID]
    if not hits:
        print("No results.")
        return 0
    print("Found %d result(s):" % len(hits))
    for rank, (h, s) in enumerate(hits):
        print("%d. id: %s  score: %.4f  importance: %.2f  context: %s"
              % (rank + 1, st.external_id(h), s, st.importance(h), st.context_label(h)))
    return 0


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="chronomind_b200", description=__doc__.split("\n")[0])
    sub = ap.add_subparsers(dest="cmd", required=True)
    q = sub.add_parser("query", help="Query a snapshot for the nearest memories")
    q.add_argument("-f", "--file", required=True)
    q.add_argument("-v", "--vector", required=True)
    q.add_argument("-l", "--limit", type=int, default=10)
    q.add_argument("-c", "--context", default=None)
    q.add_argument("-n", "--normalize", action="store_true")
    q.add_argument("--device", type=int, default=0)
    q.add_argument("--seed", type=int, default=0x5EED, help="layer seed of the rebuilt graph (the reference draws a random one)")
    q.add_argument("--threads", type=int, default=0, help="host threads for the graph build (0: all cores)")
    q.add_argument("--build", default="gpu", choices=["gpu", "host"], help="graph construction when no <file>.graph sidecar exists")
    args = ap.parse_args(argv)
    return query_command(args)


if __name__ == "__main__":
    sys.exit(main())
