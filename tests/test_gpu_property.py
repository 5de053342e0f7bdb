"""
Property tests on the GPU (hypothesis): random shapes, k, metrics, dtypes and containers through the C
ABI against the oracle — the items the reference's own tests never pin (tie order, zero norms, nulls,
ragged rows, k edge cases).  Top-k results must be bit-identical to the oracle.
"""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from tests import parity

pytestmark = pytest.mark.gpu
METRICS = ["cosine", "dot", "euclidean", "L2", "Cosine"]
EXAMPLES = int(os.environ.get("PMM_PROPERTY_EXAMPLES", "40"))   # raise for a longer soak on the GPU box


@pytest.fixture(scope="module")
def native():
    from polars_matmul_b200 import _native
    _native.lib()
    assert _native.device_count() > 0
    return _native


def _hm(a):
    from polars_matmul_b200.arrow import to_host_matrix
    return to_host_matrix(a)


@st.composite
def problems(draw):
    nq = draw(st.integers(1, 300))
    n = draw(st.integers(1, 2500))
    d = draw(st.one_of(st.integers(1, 200), st.integers(1, 200), st.integers(201, 2048)))   # a third of the draws: long vectors
    k = draw(st.one_of(st.integers(0, 140), st.integers(0, 140), st.integers(141, 300)))   # <= 248 fused path, above: slab path
    metric = draw(st.sampled_from(METRICS))
    seed = draw(st.integers(0, 2**31 - 1))
    kind = draw(st.sampled_from(["gauss", "ints", "dups", "zeros", "scaled", "tiny", "huge"]))
    return nq, n, d, k, metric, seed, kind


def _data(nq, n, d, seed, kind, dtype):
    rng = np.random.default_rng(seed)
    if kind == "ints":      # small integers: exact arithmetic everywhere, many exact ties
        q = rng.integers(-2, 3, size=(nq, d)).astype(dtype)
        c = rng.integers(-2, 3, size=(n, d)).astype(dtype)
    else:
        q = rng.standard_normal((nq, d)).astype(dtype)
        c = rng.standard_normal((n, d)).astype(dtype)
    if kind == "dups" and n > 3:     # duplicated corpus rows: exact ties at arbitrary ranks
        src = rng.integers(0, n, size=n // 2)
        dst = rng.integers(0, n, size=n // 2)
        c[dst] = c[src]
    # magnitudes outside the f16 range: the first filter level rounds f32 operands to f16 (subnormals below 6e-5,
    # overflow above 65504); its losslessness proof must hand such rows to the next level
    if kind == "scaled":             # every row at its own scale, 1e-8 .. 1e5
        q *= (10.0 ** rng.uniform(-8, 5, size=(nq, 1))).astype(dtype)
        c *= (10.0 ** rng.uniform(-8, 5, size=(n, 1))).astype(dtype)
    if kind == "tiny":
        q *= dtype(1e-6)
        c *= dtype(3e-7)
    if kind == "huge":
        c[rng.integers(0, n, size=max(1, n // 20))] *= dtype(1e6)
        q[rng.integers(0, nq, size=max(1, nq // 20))] *= dtype(2e5)
    if kind == "zeros":              # zero vectors: the cosine guards, euclidean cancellation
        c[rng.integers(0, n, size=max(1, n // 10))] = 0
        q[rng.integers(0, nq, size=max(1, nq // 10))] = 0
    return q, c


@settings(max_examples=EXAMPLES, deadline=None, suppress_health_check=list(HealthCheck))
@given(problems())
def test_topk_f32_bit_exact_vs_oracle(native, oracle, p):
    nq, n, d, k, metric, seed, kind = p
    q, c = _data(nq, n, d, seed, kind, np.float32)
    idx, sc = native.topk(_hm(q), _hm(c), k, metric)
    assert idx.shape == (nq, min(k, n))
    parity.check_topk(idx, sc, q, c, k, metric, oracle, exact=True)


@settings(max_examples=EXAMPLES, deadline=None, suppress_health_check=list(HealthCheck))
@given(problems(), st.sampled_from([100, 150, 300]))
def test_topk_chunked_host_path_bit_exact(native, oracle, p, ratio_pct):
    """The chunked upload path on tiny corpora (256-row chunks, forced): candidate lists carried from launch to
    launch, ties across chunk boundaries, re-query levels against the whole-corpus planes."""
    nq, n, d, k, metric, seed, kind = p
    n = 700 + n * 3                               # 700 .. 8200 rows -> 2 to 8 chunks
    q, c = _data(nq, n, d, seed, kind, np.float32)
    native.set_option("host_chunk_min_mb", 0)
    native.set_option("host_chunk_min_rows", 256)
    native.set_option("host_chunk_ratio_pct", ratio_pct)
    try:
        idx, sc = native.topk(_hm(q), _hm(c), k, metric)
    finally:
        native.set_option("host_chunk_min_mb", 64)
        native.set_option("host_chunk_min_rows", 16384)
        native.set_option("host_chunk_ratio_pct", 0)
    parity.check_topk(idx, sc, q, c, k, metric, oracle, exact=True)


@settings(max_examples=max(15, EXAMPLES // 2), deadline=None, suppress_health_check=list(HealthCheck))
@given(problems())
def test_topk_f16_input_bit_exact(native, oracle, p):
    """f16-stored inputs (README contract: upcast exactly to f32, then the f32 path): one kind::f16 MMA per K-step on
    the exact planes, re-scoring and proof as for f32."""
    nq, n, d, k, metric, seed, kind = p
    if kind in ("scaled", "huge"):       # would not fit f16 storage itself
        kind = "gauss"
    q, c = _data(nq, n, d, seed, kind, np.float32)
    if kind == "tiny":
        q, c = q * np.float32(1e3), c * np.float32(1e3)      # f16 subnormal range, still non-zero
    q16, c16 = q.astype(np.float16), c.astype(np.float16)
    idx, sc = native.topk(_hm(q16), _hm(c16), k, metric)
    parity.check_topk(idx, sc, q16.astype(np.float32), c16.astype(np.float32), k, metric, oracle, exact=True)


@settings(max_examples=max(15, EXAMPLES // 2), deadline=None, suppress_health_check=list(HealthCheck))
@given(problems())
def test_topk_f64_vs_oracle(native, oracle, p):
    nq, n, d, k, metric, seed, kind = p
    q, c = _data(min(nq, 64) if n * d > 200_000 else nq, n, d, seed, kind, np.float64)
    idx, sc = native.topk(_hm(q), _hm(c), k, metric)
    # f64 top-k = tensor-core filter + exact f64 re-scoring (sequential FMA): bit-identical to the oracle for every kind
    parity.check_topk(idx, sc, q, c, k, metric, oracle, exact=True)


@settings(max_examples=max(15, EXAMPLES // 4), deadline=None, suppress_health_check=list(HealthCheck))
@given(problems())
def test_matmul_vs_oracle(native, oracle, p):
    nq, n, d, k, metric, seed, kind = p
    for dtype in (np.float32, np.float64):
        q, c = _data(nq, n, d, seed, kind, dtype)
        out = native.matmul(_hm(q), _hm(c))
        assert out.dtype == dtype and out.shape == (nq, n)
        ref = oracle.matmul(q, c)
        if kind == "ints":
            assert np.array_equal(out, ref)
        else:
            parity.check_matmul(out, q, c, ref, dtype)


@settings(max_examples=max(15, EXAMPLES // 4), deadline=None, suppress_health_check=list(HealthCheck))
@given(st.integers(1, 60), st.integers(1, 400), st.integers(1, 40), st.integers(1, 30), st.sampled_from(["cosine", "dot", "euclidean"]),
       st.integers(0, 2**31 - 1))
def test_list_container_nulls_and_ragged(native, oracle, nq, n, d, k, metric, seed):
    """pl.List semantics (src/matmul.rs:231-286): row 0 defines dim, short rows zero padded, null element -> 0,
    null row -> zeros."""
    import pyarrow as pa
    import polars_matmul_b200 as pmm
    rng = np.random.default_rng(seed)
    dense = rng.standard_normal((n, d)).astype(np.float32)
    rows = [r.tolist() for r in dense]
    for i in rng.integers(1, n, size=n // 5) if n > 1 else []:
        cut = int(rng.integers(0, d + 1))
        rows[i] = rows[i][:cut]
        dense[i, cut:] = 0
    for i in rng.integers(1, n, size=n // 7) if n > 1 else []:
        if rows[i] is not None and len(rows[i]) > 0:
            j = int(rng.integers(0, len(rows[i])))
            rows[i][j] = None
            dense[i, j] = 0
    for i in rng.integers(1, n, size=n // 9) if n > 1 else []:
        rows[i] = None
        dense[i] = 0
    arr = pa.array(rows, type=pa.large_list(pa.float32()))
    q = rng.standard_normal((nq, d)).astype(np.float32)
    idx, sc = pmm.topk_arrays(q, arr, k, metric)
    parity.check_topk(idx, sc, q, dense, k, metric, oracle, exact=True)
    out = pmm.matmul_array(q, arr)
    parity.check_matmul(out, q, dense, oracle.matmul(q, dense), np.float32)
