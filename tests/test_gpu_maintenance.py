"""SURVEY.md §8f N4 on the device: the all-pairs `similarity > threshold` test of `ChronoMind::consolidate`
(src/store.rs:516-568) as a tensor-core threshold self-join, and one `apply_decay` sweep (src/store.rs:428-458),
against the oracle's restatements."""
import numpy as np
import pytest

import oracle
from chronomind_b200 import datasets
from chronomind_b200.engine import ChronoMindB200, CmError, decay_sweep_cuda, default_config, last_counters

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _with_duplicates(n, dim, n_dup, seed, noise):
    rng = np.random.default_rng(seed)
    raw = (rng.normal(size=(n, dim)) * rng.uniform(0.5, 4.0, size=(n, 1))).astype(np.float32)
    src = rng.choice(n - n_dup, size=n_dup, replace=False)
    raw[n - n_dup:] = raw[src] * rng.uniform(0.5, 2.0, size=(n_dup, 1)).astype(np.float32) \
        + noise * rng.normal(size=(n_dup, dim)).astype(np.float32)
    return raw


@pytest.mark.parametrize("n,dim,n_dup,thr,noise", [(5000, 64, 300, 0.95, 0.05), (7000, 96, 500, 0.99, 0.3), (3000, 768, 100, 0.95, 0.02),
                                                   (2500, 33, 200, 0.5, 0.2)])
def test_similar_pairs_equals_the_oracle_pair_test(n, dim, n_dup, thr, noise):
    raw = _with_duplicates(n, dim, n_dup, n + dim, noise)
    dele = (np.random.default_rng(1).uniform(size=n) < 0.05).astype(np.uint8)
    st = ChronoMindB200.from_matrix(raw, default_config(dimensions=dim), deleted=dele).set_contexts(raw).upload(0)
    gi, gj, gs = st.similar_pairs(thr)
    c = last_counters()
    wi, wj, ws = oracle.similar_pairs(raw, thr, deleted=dele, cap=1 << 22)
    assert len(wi) >= n_dup // 2                                   # the planted duplicates are there
    assert np.array_equal(gi, wi) and np.array_equal(gj, wj)
    assert np.array_equal(bits(gs), bits(ws))                       # same arithmetic, bit for bit
    assert c["rescored"] >= len(wi)
    assert not dele[gi].any() and not dele[gj].any() and np.all(gi < gj)


def test_similar_pairs_small_capacity_and_state_errors():
    raw = _with_duplicates(3000, 32, 400, 3, 0.01)
    st = ChronoMindB200.from_matrix(raw, default_config(dimensions=32)).set_contexts(raw).upload(0)
    full = st.similar_pairs(0.9)
    grown = st.similar_pairs(0.9, cap=7)                            # wrapper retries with the reported count
    assert all(np.array_equal(a, b) for a, b in zip(full, grown)) and len(full[0]) > 7
    plain = ChronoMindB200.from_matrix(raw, default_config(dimensions=32)).upload(0)
    with pytest.raises(CmError) as e:
        plain.similar_pairs(0.9)
    assert e.value.kind == "State"


def test_decay_sweep_matches_oracle():
    import torch
    rng = np.random.default_rng(9)
    n = 100_003
    h = 3600 * 10**9
    now = 500 * h
    imp = rng.uniform(0, 1, n).astype(np.float32)
    la = rng.integers(0, 600 * h, n).astype(np.uint64)               # some accessed after `now`
    dt = np.minimum(la, rng.integers(0, 600 * h, n).astype(np.uint64))
    rate = np.where(rng.uniform(size=n) < 0.5, 0.0, rng.uniform(0.01, 0.3, n)).astype(np.float32)
    want_imp, want_dt = oracle.decay_sweep(imp, dt, la, rate, 0.1, now)
    t_imp = torch.from_numpy(imp.copy()).cuda()
    t_dt = torch.from_numpy(dt.view(np.int64).copy()).cuda()
    decay_sweep_cuda(t_imp, t_dt, torch.from_numpy(la.view(np.int64)).cuda(), torch.from_numpy(rate).cuda(), 0.1, now)
    torch.cuda.synchronize()
    got_imp, got_dt = t_imp.cpu().numpy(), t_dt.cpu().numpy().view(np.uint64)
    assert np.array_equal(got_dt, want_dt)
    assert np.allclose(got_imp, want_imp, rtol=2e-7, atol=0)          # libm expf vs correctly rounded exp: <= 1 ulp
    assert np.array_equal(got_imp[now <= np.maximum(dt, la)], imp[now <= np.maximum(dt, la)])
    # a second sweep at the same instant claims nothing
    decay_sweep_cuda(t_imp, t_dt, torch.from_numpy(la.view(np.int64)).cuda(), torch.from_numpy(rate).cuda(), 0.1, now)
    torch.cuda.synchronize()
    assert np.array_equal(t_imp.cpu().numpy(), got_imp)
