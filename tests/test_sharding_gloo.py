"""N > 1 host logic on CPU: two `gloo` ranks shard a corpus by id, all-gather their per-shard top-k lists with the
same helper `bench.py` uses on NCCL, and the merged result (merge arithmetic = the oracle's restatement of
src/index/sharded_rwlock.rs:81-94, since the merge KERNEL needs a GPU) must equal the unsharded brute force."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from chronomind_b200 import datasets, sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, nq, dim, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data, q = datasets.uniform_dataset(n, nq, dim, 0xC0FFEE)
        shard = sharding.shard_rows(oracle.preprocess_rows(data), rank, world)
        li, ld = oracle.brute_force(shard, q, k, mode="prepared")            # this rank's shard, LOCAL handles
        ids_g, dist_g = sharding.allgather_lists(torch.from_numpy(li.view(np.int32)), torch.from_numpy(ld))
        assert ids_g.shape == (world, nq, k)
        assert np.array_equal(ids_g[rank].numpy().view(np.uint32), li)       # rank r's list sits at index r
        mi, md = oracle.sharded_merge(ids_g.numpy().view(np.uint32), dist_g.numpy(), k)
        # the one-collective form bench.py uses: packed (TotalF32 distance << 32 | local handle) keys
        from bench import total_keys
        keys = torch.from_numpy(total_keys(ld, li).view(np.int64))
        kg = sharding.allgather_keys(keys).numpy().view(np.uint64)
        assert kg.shape == (world, nq, k)
        assert np.array_equal((kg & np.uint64(0xFFFFFFFF)).astype(np.uint32), ids_g.numpy().view(np.uint32))
        assert np.array_equal(kg[rank], total_keys(ld, li))
        np.save(os.path.join(out_dir, "ids%d.npy" % rank), mi)
        np.save(os.path.join(out_dir, "dist%d.npy" % rank), md)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,nq,dim,k", [(1001, 17, 32, 10)])
def test_two_rank_sharded_search_equals_unsharded(tmp_path, n, nq, dim, k):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n, nq, dim, k, str(tmp_path)), nprocs=world, join=True)
    data, q = datasets.uniform_dataset(n, nq, dim, 0xC0FFEE)
    wi, wd = oracle.brute_force(oracle.preprocess_rows(data), q, k, mode="prepared")
    for r in range(world):
        mi = np.load(tmp_path / ("ids%d.npy" % r))
        md = np.load(tmp_path / ("dist%d.npy" % r))
        assert np.array_equal(mi, wi), "merged global handles differ from the unsharded scan"
        assert np.array_equal(md.view(np.uint32), wd.view(np.uint32))


def test_handle_mapping_round_trip():
    for world in (1, 2, 4, 8, 16):
        h = np.arange(1000, dtype=np.uint32)
        s, l = sharding.split_handle(h, world)
        assert np.array_equal(sharding.global_handle(l, s, world), h)
        x = np.arange(1000)[:, None]
        for r in range(world):
            assert np.array_equal(sharding.shard_rows(x, r, world)[:, 0], h[s == r])
