"""K1's per-thread running thresholds, restated in numpy and checked for the ONE property the exactness argument of the
tensor-core scan rests on (csrc/scan_tc.cu header): at every moment

    tau  <=  s_(k) - 2 eps          s_(k) = k-th largest score among the distinct live rows seen so far,

so every row of the final exact top-k is emitted as a candidate. Three structures carry tau in the kernel:

  * Ladder  (k > 32): a histogram over fixed bounds b(j) = j * 2 / n - 1; buckets at or below the current bound are not
    recorded any more; "k recorded scores reach b(j)" raises the bound.
  * LocalTopK with best-four parking (k <= 32): hits of a tile are offered to the list only after the tile, and only the
    tile's four best improving scores are offered at all.
  * the cold pre-pass for k <= 16: only the maximum of every 8 columns of the first tile is offered.

The models below follow the device code statement by statement (fp32 arithmetic, same clamps and slacks); they are
not used by the product. The GPU tests check results; this file checks the invariant on streams built to stress it."""
import numpy as np
import pytest

F = np.float32
SLACK = F(9.5367431640625e-07)   # 2^-20, the allowance for (s + 1) * scale rounding a score up into the next bucket


class Ladder:
    """struct Ladder (scan_tc.cu)."""

    def __init__(self, n, k):
        self.n, self.k = n, k
        self.scale, self.step = F(0.5) * F(n), F(2.0) / F(n)
        self.hist = np.zeros(n, np.int64)
        self.above, self.j, self.b_next = 0, -1, F(-np.inf)

    def record(self, s):
        i = int(min(max((F(s) + F(1.0)) * self.scale, F(0.0)), F(self.n - 1)))
        self.hist[i] += 1
        self.above += 1

    def tau(self, eps2):
        return F(self.j) * self.step - F(1.0) - F(eps2) - SLACK if self.j >= 0 else F(-np.inf)

    def advance(self):
        moved = False
        while self.above >= self.k and self.j < self.n - 1:
            self.j += 1
            self.above -= int(self.hist[self.j])
            moved = True
        if moved:
            self.b_next = F(self.j + 1) * self.step - F(1.0) if self.j < self.n - 1 else F(np.inf)
        return moved


def kth_largest(seen, k):
    return np.sort(np.asarray(seen, F))[-k] if len(seen) >= k else F(-np.inf)


def ladder_stream(scores, n, k, eps2, shared=None):
    """One thread's view of a unit: tile 0 cold (every live score recorded), then per 32-column chunk the hits >= tau are
    recorded when they reach b_next; the bound may rise after every chunk with a hit. `shared[t]`: a threshold another
    unit of the same query published before tile t (it only filters hits). Returns the emitted score set."""
    lad, seen, emitted = Ladder(n, k), [], []
    tiles = scores.reshape(-1, 256)
    for t, tile in enumerate(tiles):
        g = F(shared[t]) if shared is not None else F(-np.inf)
        tau = max(g, lad.tau(eps2))
        listed = False
        if tau == -np.inf:                                    # cold pre-pass: no emission
            for s in tile[~np.isnan(tile)]:
                lad.record(s)
            listed = True
            if lad.advance():
                tau = lad.tau(eps2)
        for c in range(0, 256, 32):
            hit = False
            for s in tile[c:c + 32]:
                if s >= tau:                                  # NaN never passes
                    emitted.append(s)
                    hit = True
                    if not listed and s >= lad.b_next:
                        lad.record(s)
            if hit and lad.above >= lad.k and lad.advance():
                tau = max(tau, lad.tau(eps2))
        seen.extend(tile[~np.isnan(tile)].tolist())
        # the invariant, after every tile (own bound only: `g` belongs to another unit's proof)
        assert lad.tau(eps2) <= kth_largest(seen, k) - F(eps2) + F(1e-7), (t, lad.j, lad.tau(eps2), kth_largest(seen, k))
    return np.asarray(emitted, F), np.asarray(seen, F), lad


def streams(rng):
    yield "uniform 32-d like", np.clip(rng.normal(0, 0.18, 40 * 256), -1, 1).astype(F)
    yield "all negative", (-0.9 + 0.05 * rng.random(24 * 256)).astype(F)
    yield "one bucket", (0.70312 + 1e-4 * rng.random(16 * 256)).astype(F)          # everything inside one ladder step
    yield "duplicates of 1.0", np.where(rng.random(20 * 256) < 0.3, 1.0, rng.uniform(-1, 1, 20 * 256)).astype(F)
    yield "exact bounds", (rng.integers(0, 257, 12 * 256) / 128.0 - 1.0).astype(F)  # scores ON the ladder's bounds
    yield "ascending", np.sort(rng.uniform(-1, 1, 16 * 256)).astype(F)
    yield "descending", np.sort(rng.uniform(-1, 1, 16 * 256))[::-1].astype(F)
    x = rng.uniform(-1, 1, 16 * 256).astype(F)
    x[rng.random(x.size) < 0.4] = np.nan                                            # tombstoned / padding rows
    yield "40% NaN", x
    y = np.full(8 * 256, np.nan, F)
    y[::50] = rng.uniform(-1, 1, y[::50].size)                                      # fewer live rows per tile than k
    yield "sparse", y
    yield "slightly outside [-1, 1]", np.clip(rng.normal(0, 0.6, 12 * 256), -1.0000002, 1.0000002).astype(F)


@pytest.mark.parametrize("n,k", [(256, 33), (256, 100), (128, 64), (128, 128)])
def test_ladder_threshold_is_a_lower_bound_and_loses_no_top_k_row(n, k):
    eps2 = F(2 * 0.0042)
    rng = np.random.default_rng(n * 1000 + k)
    for name, sc in streams(rng):
        emitted, seen, lad = ladder_stream(sc, n, k, eps2)
        if len(seen) >= k:
            need = np.sort(seen)[-k]
            # every seen score within 2 eps of the k-th best must have been emitted (what K2 relies on)
            must = np.sort(seen[seen >= need - eps2])
            got = np.sort(emitted[emitted >= need - eps2])
            assert np.array_equal(must, got), name
            # and the bound is not absurdly loose once k scores were seen: within two ladder steps + margin of s_(k),
            # unless the scores sit below the ladder's last raise (one-bucket and top-bucket streams)
            assert lad.j >= 0, name


def test_ladder_with_a_threshold_published_by_another_unit():
    """A higher shared threshold only filters the hits this thread records; its own bound stays valid for its own rows and
    climbs through the empty buckets once k hits above it accumulate."""
    eps2, k, n = F(2 * 0.0042), 40, 256
    rng = np.random.default_rng(5)
    sc = np.clip(rng.normal(0, 0.25, 30 * 256), -1, 1).astype(F)
    full = np.sort(sc)[-k]
    shared = np.full(30, -np.inf, F)
    shared[3:] = full - eps2 - F(0.01)         # valid for the QUERY (a unit that saw the same rows published it)
    emitted, seen, lad = ladder_stream(sc, n, k, eps2, shared)
    must = np.sort(seen[seen >= full - eps2])
    got = np.sort(emitted[emitted >= full - eps2])
    assert np.array_equal(must, got)


class LocalTopK:
    """struct LocalTopK: unsorted list of the k largest scores offered so far."""

    def __init__(self, k):
        self.k, self.list = k, []

    @property
    def filled(self):
        return len(self.list)

    @property
    def kth(self):
        return min(self.list) if len(self.list) == self.k else F(-np.inf)

    def insert(self, s):
        if len(self.list) < self.k:
            self.list.append(s)
            return len(self.list) == self.k
        if s > min(self.list):
            self.list[int(np.argmin(self.list))] = s
            return True
        return False


@pytest.mark.parametrize("k", [1, 10, 16, 32])
def test_list_with_group_maxima_pre_pass_and_best_four_parking(k):
    """The list path as the kernel runs it: cold tile = maxima of 8-column groups only (k <= 16) or every score; later
    tiles = hits offered after the tile, at most the four best improving ones. The list must always hold scores of k
    DISTINCT seen rows (so its minimum is a valid bound), and no row within 2 eps of the final k-th best may be missed."""
    eps2 = F(2 * 0.0042)
    rng = np.random.default_rng(k)
    for name, sc in streams(rng):
        top, seen, emitted = LocalTopK(k), [], []
        for t, tile in enumerate(sc.reshape(-1, 256)):
            tau = top.kth - eps2 if top.filled == k else F(-np.inf)
            listed = False
            if tau == -np.inf:                                   # cold pre-pass
                if k <= 16:
                    for g in range(0, 256, 8):
                        grp = tile[g:g + 8]
                        if not np.all(np.isnan(grp)):
                            top.insert(F(np.nanmax(grp)))
                else:
                    for s in tile[~np.isnan(tile)]:
                        top.insert(s)
                listed = True
                if top.filled == k:
                    tau = top.kth - eps2
            pend = []                                             # the tile's four best improving scores
            for s in tile:
                if s >= tau:
                    emitted.append(s)
                    if not listed:
                        if top.filled < k:
                            top.insert(s)
                            if top.filled == k:
                                tau = max(tau, top.kth - eps2)
                        elif s > top.kth and (len(pend) < 4 or s > min(pend)):
                            pend.append(s)
                            pend = sorted(pend, reverse=True)[:4]
            for s in pend:
                top.insert(s)
            seen.extend(tile[~np.isnan(tile)].tolist())
            # the list is a multiset of seen scores, each seen row used at most once
            pool = sorted(seen)
            for v in sorted(top.list):
                assert v in pool, name
                pool.remove(v)
            assert top.kth <= kth_largest(seen, k), name
        seen, emitted = np.asarray(seen, F), np.asarray(emitted, F)
        if len(seen) >= k:
            need = np.sort(seen)[-k]
            assert np.array_equal(np.sort(seen[seen >= need - eps2]), np.sort(emitted[emitted >= need - eps2])), name
