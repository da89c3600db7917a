"""Pins the oracle's metric surface to the reference's own known-answer tests
(`src/metric.rs:227-288`) and to closed-form values."""
import numpy as np
import pytest

import oracle

EPS = 1e-5


def test_identical_vectors_have_zero_distance():          # src/metric.rs:233-239
    v = [1.0, 0.5, -0.25, 2.0]
    assert abs(oracle.distance(v, v)) < EPS
    assert abs(oracle.similarity(v, v) - 1.0) < EPS


def test_orthogonal_vectors_have_unit_distance():          # src/metric.rs:241-248
    a, b = [1.0, 0.0, 0.0], [0.0, 1.0, 0.0]
    assert abs(oracle.distance(a, b) - 1.0) < EPS
    assert abs(oracle.similarity(a, b)) < EPS


def test_opposite_vectors_have_max_distance():             # src/metric.rs:250-256
    assert abs(oracle.distance([1.0, 2.0, 3.0], [-1.0, -2.0, -3.0]) - 2.0) < EPS


def test_degenerate_inputs_yield_max_distance():           # src/metric.rs:258-265 (exact equality)
    assert oracle.distance([], []) == 2.0
    assert oracle.distance([1.0], [1.0, 2.0]) == 2.0
    assert oracle.distance([0.0, 0.0], [0.0, 0.0]) == 2.0
    assert oracle.similarity([0.0, 0.0], [0.0, 0.0]) == 0.0
    assert oracle.distance_prepared([], []) == 2.0          # src/metric.rs:220-222
    assert oracle.distance_prepared([1.0], [1.0, 2.0]) == 2.0


def _closed_form(n):
    a = np.array([float((i * 37 % 19)) - 9.0 for i in range(n)], dtype=np.float32)
    b = np.array([float((i * 53 % 23)) - 11.0 for i in range(n)], dtype=np.float32)
    return a, b


@pytest.mark.parametrize("n", [1, 7, 8, 9, 15, 16, 17, 768])
def test_simd_and_scalar_paths_agree(n):                   # src/metric.rs:267-287
    a, b = _closed_form(n)
    # integer-valued inputs: every partial sum is exactly representable, so the exact answer is known.
    dot = int(np.dot(a.astype(np.int64), b.astype(np.int64)))
    na = int(np.dot(a.astype(np.int64), a.astype(np.int64)))
    nb = int(np.dot(b.astype(np.int64), b.astype(np.int64)))
    assert oracle.dot_scalar(a, b) == float(dot)
    assert oracle.dot_avx2(a, b) == float(dot)
    denom = np.sqrt(np.float32(na) * np.float32(nb))
    if denom <= np.finfo(np.float32).eps:
        return
    expected = np.float32(1.0) - np.clip(np.float32(dot) / denom, np.float32(-1), np.float32(1))
    got = oracle.distance(a, b)
    assert abs(got - oracle.distance_scalar(a, b)) < 1e-4   # the reference's own assertion
    assert got == float(expected)                           # stronger: exact, since sums are exact here


def test_dot_avx2_lane_and_hsum_order():
    """dot_avx2's summation tree (src/metric.rs:129-147) replayed in numpy float32/float64-FMA."""
    rng = np.random.default_rng(7)
    for n in (8, 24, 768, 771):
        a = rng.standard_normal(n).astype(np.float32)
        b = rng.standard_normal(n).astype(np.float32)
        acc = np.zeros(8, dtype=np.float32)
        chunks = n // 8 * 8
        for i in range(0, chunks, 8):
            # fma in one rounding: exact product in float64 (24x24 bits fit), add, round once.
            acc = (a[i:i + 8].astype(np.float64) * b[i:i + 8].astype(np.float64)
                   + acc.astype(np.float64)).astype(np.float32)
        s128 = acc[:4] + acc[4:]
        s64 = np.array([s128[0] + s128[2], s128[1] + s128[3]], dtype=np.float32)
        s = np.float32(s64[0] + s64[1])
        for i in range(chunks, n):
            s = np.float32(s + np.float32(a[i] * b[i]))
        assert oracle.dot_avx2(a, b) == float(s)


def test_preprocess_is_sequential_sum_sqrt_divide():       # src/metric.rs:208-214
    rng = np.random.default_rng(3)
    v = rng.uniform(-100, 100, 768).astype(np.float32)
    s = np.float32(0)
    for x in v:
        s = np.float32(s + np.float32(x * x))
    want = (v / np.sqrt(s)).astype(np.float32)
    got = oracle.preprocess(v)
    assert np.array_equal(got, want)
    z = np.zeros(5, dtype=np.float32)
    assert np.array_equal(oracle.preprocess(z), z)          # degenerate: unchanged
    tiny = np.full(4, 1e-9, dtype=np.float32)
    assert np.array_equal(oracle.preprocess(tiny), tiny)


def test_distance_prepared_clamps_and_matches_distance():
    rng = np.random.default_rng(5)
    a = oracle.preprocess(rng.standard_normal(768))
    b = oracle.preprocess(rng.standard_normal(768))
    dp = oracle.distance_prepared(a, b)
    assert 0.0 <= dp <= 2.0
    assert abs(dp - oracle.distance(a, b)) < 1e-5
    assert oracle.distance_prepared(a, a) >= 0.0            # clamp(1 - dot) never negative
    assert np.isnan(oracle.distance_prepared([np.nan, 1.0], [1.0, 1.0]))   # f32::clamp passes NaN


def test_total_key_is_total_cmp():                         # src/index/mod.rs:37-52
    vals = [-np.inf, -2.0, -0.0, 0.0, 1e-40, 0.5, 2.0, np.inf, np.nan]
    keys = [oracle.total_key(v) for v in vals]
    assert keys == sorted(keys) and len(set(keys)) == len(keys)


def test_splitmix64_known_answers():                       # src/index/lockfree_hnsw.rs:71-77
    # SplitMix64 reference stream from seed 0 (first outputs of the published generator).
    assert oracle.splitmix64(0) == 0xE220A8397B1DCDAF
    assert oracle.splitmix64(0x9E3779B97F4A7C15) == 0x6E789E6AA1B965F4


def test_layer_draw_distribution():                        # src/index/lockfree_hnsw.rs:126-132
    layers = np.array([oracle.random_layer(0xBEEF, t, 16) for t in range(200000)])
    assert layers.max() <= 31
    frac0 = (layers == 0).mean()
    assert abs(frac0 - (1 - 1 / 16)) < 0.005               # P(layer >= 1) = 1/M
    assert abs((layers >= 2).mean() - 1 / 256) < 0.001
