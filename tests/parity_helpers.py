"""Shared checks of the parity tests (north_star: ids bit-exact whenever neighbouring gaps exceed 1e-5 relative,
scores within 1e-5 relative)."""
import numpy as np


def assert_order_matches_up_to_ties(got_ids, want_ids, got_scores, want_scores, rel=1e-5, atol=1e-7):
    """Row by row: scores agree position-wise within `rel`, and a position whose id differs from the oracle's must be
    EXPLAINED by a near-tie in the oracle's own list — its score is within `rel` of a neighbouring position's (a swap
    the last ulp of `expf` can cause), or it is the last position (the rival sits just beyond the cut). Returns the
    number of such positions so a caller can bound it."""
    got_ids, want_ids = np.asarray(got_ids), np.asarray(want_ids)
    got_scores, want_scores = np.asarray(got_scores, np.float64), np.asarray(want_scores, np.float64)
    assert got_ids.shape == want_ids.shape
    fin = np.isfinite(want_scores)
    assert np.array_equal(fin, np.isfinite(got_scores)), "padding differs"
    tol = rel * np.abs(want_scores) + atol
    assert np.all(np.abs(got_scores - want_scores)[fin] <= tol[fin]), "scores differ by more than %g relative" % rel
    explained = 0
    k = want_ids.shape[1]
    for r, j in zip(*np.nonzero(got_ids != want_ids)):
        near = j == k - 1
        if j > 0:
            near |= abs(want_scores[r, j] - want_scores[r, j - 1]) <= tol[r, j]
        if j + 1 < k:
            near |= abs(want_scores[r, j] - want_scores[r, j + 1]) <= tol[r, j]
        assert near, "row %d position %d: id %d vs oracle %d with no near-tie around it (scores %r)" % (
            r, j, got_ids[r, j], want_ids[r, j], want_scores[r, max(0, j - 1):j + 2].tolist())
        explained += 1
    return explained
