"""
The losslessness proof of the tensor-core filter rests on ONE inequality: |filter value - exact value| <= E, with E from
`filter_error_bound` (polars_matmul_b200/csrc/pmm_kernels.h).  The constants in E are derived by hand (operand
rounding, one f32 ulp of truncation per tcgen05 accumulate step, rounding of the exact sum); this file MEASURES the
filter's real error against that bound on inputs built to stress each term:

  * all-positive rows      every product has the same sign: truncation bias accumulates instead of cancelling;
  * mixed-sign Gaussian    the common case;
  * large magnitudes       just inside the f16 range (row norms of a few thousand, elements up to ~700);
  * small magnitudes       elements around and below the f16 normal range (2^-14): the absolute error term;
  * D in {768, 2048, 4096, 8192} for every level: f16-rounded (default first level), exact f16 planes, TF32 x1, 3xTF32.

The filter values come out of the library through pmm_dev_filter_candidates (N <= list capacity, so EVERY pair is
exposed, not only the best ones) and the bound through pmm_filter_error_bound - the same function the proof calls.
The exact value is the oracle's working-precision score (sequential FMA), mapped to the filter's units in float64.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

COSINE, DOT, EUCLIDEAN = 0, 1, 2
DIMS = [768, 2048, 4096, 8192]


@pytest.fixture(scope="module")
def native():
    from polars_matmul_b200 import _native
    _native.lib()
    assert _native.device_count() > 0
    return _native


def _data(kind, nq, n, d, rng):
    if kind == "positive":
        q = rng.uniform(0.5, 1.5, size=(nq, d))
        c = rng.uniform(0.5, 1.5, size=(n, d))
    elif kind == "gauss":
        q = rng.standard_normal((nq, d))
        c = rng.standard_normal((n, d))
    elif kind == "large":      # |x| up to ~700/sqrt(d/768): row norms stay below 65504 (the level's max_norm)
        s = 40.0 * np.sqrt(768.0 / d)
        q = rng.standard_normal((nq, d)) * s
        c = np.abs(rng.standard_normal((n, d))) * s
    elif kind == "small":      # around the f16 normal limit 6.1e-5: part of every row is subnormal in f16
        q = rng.standard_normal((nq, d)) * 1e-4
        c = rng.standard_normal((n, d)) * 3e-5
    elif kind == "mixed":      # per-row scales spread over 4 decades inside the range
        q = rng.standard_normal((nq, d)) * 10.0 ** rng.uniform(-2, 1.5, size=(nq, 1))
        c = rng.standard_normal((n, d)) * 10.0 ** rng.uniform(-2, 1.5, size=(n, 1))
    else:
        raise ValueError(kind)
    return q.astype(np.float32), c.astype(np.float32)


def _filter_values(native, q, c, metric, level, kp=256):
    """[Q, N] filter values (NaN where a pair is missing from the kept list)."""
    import torch
    from polars_matmul_b200.sharded import unpack_candidates
    code = {np.dtype(np.float16): 0, np.dtype(np.float32): 1, np.dtype(np.float64): 2}
    tq, tc = torch.from_numpy(q).cuda(), torch.from_numpy(c).cuda()
    kept = torch.zeros((q.shape[0], kp), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    native.dev_filter_candidates(native.dev_matrix(tq.data_ptr(), q.shape[0], q.shape[1], code[q.dtype]),
                                 native.dev_matrix(tc.data_ptr(), c.shape[0], c.shape[1], code[c.dtype]),
                                 metric, level, kp, kept.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    packed = kept.cpu().numpy().view(np.uint64)
    idx, f = unpack_candidates(packed, True)
    out = np.full((q.shape[0], c.shape[0]), np.nan)
    valid = packed != 0
    rows = np.broadcast_to(np.arange(q.shape[0])[:, None], packed.shape)
    out[rows[valid], idx[valid].astype(np.int64)] = f[valid]
    return out


def _exact_in_filter_units(oracle, q32, c32, metric):
    """The oracle's f32 scores mapped to the filter's units (float64 arithmetic for the mapping only)."""
    dot = oracle.scores(q32, c32, DOT).astype(np.float64)
    if metric == DOT:
        return dot
    cn = oracle.norms(c32).astype(np.float64)
    if metric == COSINE:
        return dot / cn[None, :]
    qsq = oracle.norms(q32, squared=True).astype(np.float64)
    csq = oracle.norms(c32, squared=True).astype(np.float64)
    return -np.maximum(qsq[:, None] + csq[None, :] - 2.0 * dot, 0.0)


def _check(native, oracle, q, c, level, metric, label):
    wd = np.float64 if (q.dtype == np.float64 or c.dtype == np.float64) else np.float32   # working precision of the exact score
    q32, c32 = q.astype(wd), c.astype(wd)
    f = _filter_values(native, q, c, metric, level)
    assert not np.isnan(f).any(), f"{label}: pairs missing from the kept lists"
    exact = _exact_in_filter_units(oracle, q32, c32, metric)
    qn = np.sqrt(oracle.norms(q32, squared=True).astype(np.float64))
    cn = np.sqrt(oracle.norms(c32, squared=True).astype(np.float64))
    code = {np.dtype(np.float16): 0, np.dtype(np.float32): 1, np.dtype(np.float64): 2}
    worst = 0.0
    for i in range(q.shape[0]):
        e, max_norm = native.filter_error_bound(level, code[q.dtype], code[c.dtype], q.shape[1], metric, float(qn[i]),
                                                float(cn.max()), float(cn.min()))
        if max_norm > 0:
            assert qn[i] <= max_norm and cn.max() <= max_norm, f"{label}: test data left the level's range"
        err = np.abs(f[i] - exact[i]).max()
        worst = max(worst, err / e)
    print(f"{label}: max |filter - exact| / bound = {worst:.3f}")
    assert worst < 1.0, f"{label}: the filter's error exceeds the proof's bound ({worst:.3f} x)"
    return worst


@pytest.mark.parametrize("d", DIMS)
@pytest.mark.parametrize("kind", ["positive", "gauss", "large", "small", "mixed"])
def test_f16_rounded_level_within_bound(native, oracle, d, kind):
    rng = np.random.default_rng(d + len(kind))
    q, c = _data(kind, 96, 256, d, rng)
    for metric in (DOT, COSINE, EUCLIDEAN):
        _check(native, oracle, q, c, 0, metric, f"f16r d={d} {kind} metric={metric}")


@pytest.mark.parametrize("d", DIMS)
@pytest.mark.parametrize("kind", ["positive", "gauss", "mixed"])
def test_3xtf32_level_within_bound(native, oracle, d, kind):
    rng = np.random.default_rng(3 * d + len(kind))
    q, c = _data(kind, 96, 256, d, rng)
    for metric in (DOT, COSINE, EUCLIDEAN):
        _check(native, oracle, q, c, 3, metric, f"tf32x3 d={d} {kind} metric={metric}")


@pytest.mark.parametrize("d", DIMS)
@pytest.mark.parametrize("kind", ["positive", "gauss"])
def test_tf32x1_level_within_bound(native, oracle, d, kind):
    rng = np.random.default_rng(5 * d + len(kind))
    q, c = _data(kind, 96, 256, d, rng)
    for metric in (DOT, COSINE, EUCLIDEAN):
        _check(native, oracle, q, c, 1, metric, f"tf32x1 d={d} {kind} metric={metric}")


@pytest.mark.parametrize("d", DIMS)
@pytest.mark.parametrize("kind", ["positive", "gauss", "small"])
def test_exact_f16_planes_within_bound(native, oracle, d, kind):
    """f16-stored inputs: the planes are exact, the bound is accumulation + exact-sum rounding only - the tightest test of
    the 'one ulp of truncation per accumulate step' term."""
    rng = np.random.default_rng(7 * d + len(kind))
    q, c = _data(kind, 96, 256, d, rng)
    if kind == "small":
        q, c = q * 10, c * 10           # keep most elements representable in f16 storage
    q16, c16 = q.astype(np.float16), c.astype(np.float16)
    for metric in (DOT, COSINE, EUCLIDEAN):
        _check(native, oracle, q16, c16, 0, metric, f"f16 d={d} {kind} metric={metric}")


@pytest.mark.parametrize("d", [768, 4096])
@pytest.mark.parametrize("kind", ["positive", "gauss", "small"])
def test_f64_sources_within_bound(native, oracle, d, kind):
    """f64 working precision: the planes are rounded from the f64 value (f16 level) or split after an f64 -> f32 rounding
    (3xTF32 level); the exact counterpart is the oracle's f64 score."""
    rng = np.random.default_rng(11 * d + len(kind))
    q, c = _data(kind, 96, 256, d, rng)
    q64 = q.astype(np.float64) * (1.0 + 1e-9 * rng.standard_normal(q.shape))    # not representable in f32
    c64 = c.astype(np.float64) * (1.0 + 1e-9 * rng.standard_normal(c.shape))
    for level in (0, 3):
        for metric in (DOT, COSINE, EUCLIDEAN):
            _check(native, oracle, q64, c64, level, metric, f"f64 level={level} d={d} {kind} metric={metric}")


def test_bound_is_not_vacuous(native, oracle):
    """Sanity of the measurement itself: the f16-rounded level's real error is a visible fraction of its bound on
    adversarial data (if the ratio were ~0 the test above would prove nothing about the constants)."""
    rng = np.random.default_rng(99)
    q, c = _data("positive", 96, 256, 4096, rng)
    w = _check(native, oracle, q, c, 0, DOT, "f16r positive d=4096 (vacuity check)")
    assert w > 0.01
