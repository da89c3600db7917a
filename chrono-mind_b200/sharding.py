"""Host-side plumbing of the sharded read path (one process per GPU, `torch.distributed`).

The corpus shards by vector id exactly like `ShardedRwLockHnsw` (`src/index/sharded_rwlock.rs:54-74`), generalised
from 16 shards to G ranks: `shard = id mod G`, `local = id div G`, `global = local * G + shard`. Every rank searches its
own shard for the whole query batch; the per-shard `[nq, len]` lists are all-gathered (NCCL over NVLink on GPUs, gloo
in the CPU tests) into the `[G, nq, len]` layout the device merge kernel (`cm_merge_topk_dev`, K5) consumes. Nothing in
here computes distances.
"""
from __future__ import annotations

import numpy as np


def shard_rows(x: np.ndarray, rank: int, world: int) -> np.ndarray:
    """Rows of shard `rank`: global ids rank, rank + G, rank + 2G, ... (round-robin insert, sharded_rwlock.rs:70-74)."""
    return np.ascontiguousarray(x[rank::world]) if world > 1 else x


def shard_attrs(a, rank: int, world: int):
    return None if a is None else (np.ascontiguousarray(a[rank::world]) if world > 1 else a)


def global_handle(local, rank: int, world: int):
    """`join_handle` (sharded_rwlock.rs:54-58) for G shards."""
    return local * world + rank


def split_handle(handle, world: int):
    """`split` (sharded_rwlock.rs:60-66): (shard, local)."""
    return handle % world, handle // world


def allgather_lists(ids, dist, group=None):
    """All-gather per-shard result lists. ids/dist: `[nq, len]` tensors (CUDA with NCCL, CPU with gloo) of LOCAL
    handles. Returns `([G, nq, len] ids, [G, nq, len] dist)`, shard-major — rank r's list at index r."""
    import torch
    import torch.distributed as td
    world = td.get_world_size(group)
    nq = ids.shape[0]
    # concatenated along dim 0 (the one layout both NCCL and gloo accept), then viewed shard-major
    ids_g = torch.empty((world * nq,) + tuple(ids.shape[1:]), dtype=ids.dtype, device=ids.device)
    dist_g = torch.empty((world * nq,) + tuple(dist.shape[1:]), dtype=dist.dtype, device=dist.device)
    td.all_gather_into_tensor(ids_g, ids.contiguous(), group=group)
    td.all_gather_into_tensor(dist_g, dist.contiguous(), group=group)
    return ids_g.view((world,) + tuple(ids.shape)), dist_g.view((world,) + tuple(dist.shape))


def allgather_keys(keys, group=None):
    """All-gather packed (distance, local handle) keys (`engine.pack_topk_cuda`): `[nq, len]` int64 per rank ->
    `[G, nq, len]`, shard-major. One collective per step instead of the two of `allgather_lists`."""
    import torch
    import torch.distributed as td
    world = td.get_world_size(group)
    out = torch.empty((world * keys.shape[0],) + tuple(keys.shape[1:]), dtype=keys.dtype, device=keys.device)
    td.all_gather_into_tensor(out, keys.contiguous(), group=group)
    return out.view((world,) + tuple(keys.shape))
