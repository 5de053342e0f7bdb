"""
Parity rules shared by the GPU tests, smoke() and bench self-checks.  TEST INFRASTRUCTURE.

BASELINE.json north_star: indices bit-exact with tie order "lower index first", except candidates
whose score gap to the competing candidate is below tolerance; scores within 1e-5 relative for f32
working precision and 1e-12 for f64.

"Relative" is ill-posed where the reference's own arithmetic cancels (a dot product near zero, a
distance near zero), so the bound is stated against the natural magnitude of the rounding error:

    dot / cosine :  |a - b|   <= rtol * max(|b|,   FLOOR * scale_ij)     scale = |q_i||c_j| (dot), 1 (cosine)
    euclidean    :  |a^2-b^2| <= rtol * max(b^2,   FLOOR * scale_ij^2)   scale^2 = |q_i|^2 + |c_j|^2
                    (the quantity that is actually accumulated is the squared distance; sqrt is monotone)

with rtol = 1e-5 (f32) / 1e-12 (f64) and FLOOR = 0.05.
"""
from __future__ import annotations

import numpy as np

RTOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}
FLOOR = 0.05
COSINE, DOT, EUCLIDEAN = 0, 1, 2
METRIC_CODE = {"cosine": COSINE, "dot": DOT, "euclidean": EUCLIDEAN, "l2": EUCLIDEAN}


def _sqnorms(x):
    x = np.asarray(x)
    return np.einsum("ij,ij->i", x, x, dtype=np.float64)  # no full f64 copy of a large corpus


def _scale(q, c, metric):
    qn2 = _sqnorms(q)[:, None]
    cn2 = _sqnorms(c)[None, :]
    if metric == DOT:
        return np.sqrt(qn2 * cn2)
    if metric == COSINE:
        return np.ones((qn2.shape[0], cn2.shape[1]))
    return qn2 + cn2  # squared scale for euclidean


def score_close(a, b, scale, metric, rtol):
    """Elementwise closeness of scores a (ours) to b (oracle) under the rule above."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    if metric == EUCLIDEAN:
        ok = np.abs(a * a - b * b) <= rtol * np.maximum(b * b, FLOOR * scale)
    else:
        ok = np.abs(a - b) <= rtol * np.maximum(np.abs(b), FLOOR * scale)
    return ok | both_nan | (a == b)


def check_matmul(out, q, c, oracle_out, working_dtype):
    rtol = RTOL[np.dtype(working_dtype)]
    ok = score_close(out, oracle_out, _scale(q, c, DOT), DOT, rtol)
    assert ok.all(), f"matmul: {(~ok).sum()} of {ok.size} entries outside tolerance; worst abs diff {np.abs(out - oracle_out).max():.3e}"


def check_topk(index, score, q, c, k, metric, oracle, working_dtype=None, exact=False):
    """Compares a (index, score) result with the oracle on the same inputs.

    exact=True : indices identical and scores bit-identical (the SIMT path restates the oracle's order).
    exact=False: scores within tolerance position by position; an index may differ from the oracle's
                 only where the two candidates' ORACLE scores are within tolerance of each other
                 (a near-tie, the north_star exemption).  Returns the fraction of exactly equal indices.
    """
    m = METRIC_CODE[metric.lower()] if isinstance(metric, str) else metric
    wd = np.dtype(working_dtype or oracle.working_dtype(q.dtype, c.dtype))
    oi, osc = oracle.topk(q, c, k, m)
    assert index.shape == oi.shape and score.shape == osc.shape, (index.shape, oi.shape)
    assert index.dtype == np.uint32 and score.dtype == np.float64
    if index.size == 0:
        return 1.0
    if exact:
        assert np.array_equal(index, oi), f"indices differ at {np.argwhere(index != oi)[:5].tolist()}"
        same = (score == osc) | (np.isnan(score) & np.isnan(osc))
        assert same.all(), f"scores not bit-identical: max diff {np.nanmax(np.abs(score - osc)):.3e}"
        return 1.0
    rtol = RTOL[wd]
    full = oracle.scores(q, c, m).astype(np.float64)            # oracle score of every pair
    scale = _scale(q, c, m)
    rows = np.arange(index.shape[0])[:, None]
    ok = score_close(score, osc, scale[rows, oi], m, rtol)
    assert ok.all(), (f"{(~ok).sum()} scores outside tolerance; first at {np.argwhere(~ok)[0].tolist()}: "
                      f"ours {score[~ok][0]!r} oracle {osc[~ok][0]!r}")
    mism = index != oi
    if mism.any():
        # the candidate we returned must tie (within tolerance) with the oracle's pick at that rank
        ours_oracle_score = full[rows, index.astype(np.int64)]
        tie = score_close(ours_oracle_score, osc, scale[rows, oi], m, 2 * rtol)
        bad = mism & ~tie
        assert not bad.any(), (f"{bad.sum()} index mismatches are not near-ties; first at "
                               f"{np.argwhere(bad)[0].tolist()}: ours {index[bad][0]} oracle {oi[bad][0]}")
    # best-first order under the metric, and no duplicate index in a row
    s = score if m != EUCLIDEAN else -score
    finite = ~np.isnan(s)
    d = np.diff(np.where(finite, s, -np.inf), axis=1)
    assert (d <= 0).all(), "scores are not sorted best-first"
    srt = np.sort(index, axis=1)
    assert (np.diff(srt.astype(np.int64), axis=1) > 0).all(), "duplicate index within a row"
    return float((~mism).mean())
