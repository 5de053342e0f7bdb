"""
GPU parity tests: the CUDA path, called through the C ABI (include/pmm.h), against the CPU oracle on
the same seeded inputs, plus the committed golden fixtures.  Run on the B200 box: pytest -m gpu.
Tolerances are stated in tests/parity.py (1e-5 relative f32, 1e-12 f64; indices exact up to near-ties).
"""
import json
import os

import numpy as np
import pytest

from tests import parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def native():
    from polars_matmul_b200 import _native
    _native.lib()
    if _native.device_count() == 0:
        pytest.fail("no CUDA device visible: the gpu tests must run on the GPU box")
    _native.set_option("force_generic", 0)
    return _native


@pytest.fixture(scope="module")
def pmm():
    import polars_matmul_b200
    return polars_matmul_b200


def _randn(rng, *shape, dtype=np.float32):
    return rng.standard_normal(shape).astype(dtype)


def _hm(a):
    from polars_matmul_b200.arrow import to_host_matrix
    return to_host_matrix(a)


# ---------------------------------------------------------------------------------------------- golden vectors
def test_reference_known_answers_topk(native, oracle, golden_dir):
    known = json.load(open(os.path.join(golden_dir, "reference_known_answers.json")))
    for case in known["topk"]:
        for dt in (np.float64, np.float32):
            q, c = np.array(case["query"], dt), np.array(case["corpus"], dt)
            idx, sc = native.topk(_hm(q), _hm(c), case["k"], case["metric"])
            assert idx.shape[1] == case["n_results"], case["src"]
            assert idx[:, 0].tolist() == case["top1_index"], case["src"]
            np.testing.assert_allclose(sc[:, 0], case["top1_score"], atol=case["atol"])
            parity.check_topk(idx, sc, q, c, case["k"], case["metric"], oracle)


def test_reference_known_answers_matmul(native, oracle, golden_dir):
    known = json.load(open(os.path.join(golden_dir, "reference_known_answers.json")))
    for case in known["matmul"]:
        dt = np.float32 if case["dtype"] == "f32" else np.float64
        out = native.matmul(_hm(np.array(case["left"], dt)), _hm(np.array(case["right"], dt)))
        assert out.dtype == dt
        np.testing.assert_allclose(out, np.array(case["expect"]), rtol=case["rtol"])
        if "flat" in case:
            np.testing.assert_allclose(out.reshape(-1), case["flat"], rtol=case["rtol"])


def test_reference_errors(native, golden_dir):
    known = json.load(open(os.path.join(golden_dir, "reference_known_answers.json")))
    for case in known["errors"]:
        q = np.array(case["query"], np.float64).reshape(len(case["query"]), -1)
        c = np.array(case["corpus"], np.float64).reshape(len(case["corpus"]), -1 if case["corpus"] else q.shape[1])
        with pytest.raises(RuntimeError, match=case["match"]):
            if case["call"] == "topk":
                native.topk(_hm(q), _hm(c), case["k"], case["metric"])
            else:
                native.matmul(_hm(q), _hm(c))


def test_seeded_numpy_fixtures(native, oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "matmul_seed42_10x20x32_f64.npz"))
    out = native.matmul(_hm(g["left"]), _hm(g["right"]))
    np.testing.assert_allclose(out, g["expect"], rtol=float(g["rtol"]))
    g = np.load(os.path.join(golden_dir, "cosine_seed42_5x20x16_f64.npz"))
    idx, sc = native.topk(_hm(g["query"]), _hm(g["corpus"]), int(g["k"]), "cosine")
    np.testing.assert_allclose(sc, g["expect_sorted_desc"], rtol=float(g["rtol"]))
    g = np.load(os.path.join(golden_dir, "bench_selfcheck_seed42_100x500x64_f64.npz"))
    for dt in (np.float64, np.float32):
        idx, sc = native.topk(_hm(g["query"].astype(dt)), _hm(g["corpus"].astype(dt)), int(g["k"]), "cosine")
        np.testing.assert_allclose(sc, g["expect_topk_scores"], rtol=float(g["rtol"]))


# ---------------------------------------------------------------------------------------------- SIMT path: bit-exact
@pytest.mark.parametrize("metric", ["cosine", "dot", "euclidean"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_generic_path_bit_exact(native, oracle, metric, dtype):
    rng = np.random.default_rng(7)
    q, c = _randn(rng, 37, 50, dtype=dtype), _randn(rng, 1000, 50, dtype=dtype)
    native.set_option("force_generic", 1)
    native.set_option("f64_simt", 1)          # f64: sequential-FMA kernel instead of DMMA
    try:
        idx, sc = native.topk(_hm(q), _hm(c), 17, metric)
    finally:
        native.set_option("force_generic", 0)
        native.set_option("f64_simt", 0)
    parity.check_topk(idx, sc, q, c, 17, metric, oracle, exact=True)


def test_generic_path_large_k_and_clamp(native, oracle):
    rng = np.random.default_rng(8)
    q, c = _randn(rng, 9, 24), _randn(rng, 700, 24)
    for k in (129, 200, 248, 249, 300, 700, 5000):  # <= 248: fused path with 256-entry lists; above: SIMT slab path;
        idx, sc = native.topk(_hm(q), _hm(c), k, "dot")   # 5000 clamps to N (src/matmul.rs:443)
        assert idx.shape == (9, min(k, 700))
        parity.check_topk(idx, sc, q, c, k, "dot", oracle, exact=True)
    q2, c2 = _randn(rng, 140, 72), _randn(rng, 20_000, 72)
    c2[rng.integers(0, 20_000, size=400)] = c2[rng.integers(0, 20_000, size=400)]
    for metric, k in (("cosine", 130), ("euclidean", 248), ("dot", 199)):
        native.reset_stats()
        native.set_option("profile", 1)
        try:
            idx, sc = native.topk(_hm(q2), _hm(c2), k, metric)
            assert native.get_stat("select_f32_launches") == 0 and native.get_stat("tc_topk_f16r_kp256_launches") >= 1
        finally:
            native.set_option("profile", 0)
        parity.check_topk(idx, sc, q2, c2, k, metric, oracle, exact=True)


def test_large_k_runs_several_fused_passes(native, oracle):
    """248 < k <= 2000 (f32): P passes of the fused filter with 256-entry lists, pass p+1 restricted to candidates below
    the worst one pass p kept; all 256 P candidates re-scored exactly, sorted as one, proven against the last pass.
    No score slab (the select kernel never runs); beyond 2000 the slab path takes over. Always the oracle's answer."""
    rng = np.random.default_rng(81)
    q, c = _randn(rng, 150, 80), _randn(rng, 30_000, 80)
    c[rng.integers(0, 30_000, size=600)] = c[rng.integers(0, 30_000, size=600)]      # exact ties at arbitrary ranks
    for metric, k, passes in (("dot", 249, 2), ("cosine", 1000, 4), ("euclidean", 600, 3), ("dot", 2000, 8)):
        native.reset_stats()
        native.set_option("profile", 1)
        try:
            idx, sc = native.topk(_hm(q), _hm(c), k, metric)
            assert native.get_stat("tc_topk_f16r_kp256_launches") == passes, (k, native.get_stat("tc_topk_f16r_kp256_launches"))
            assert native.get_stat("sort_lists_launches") == 1
            # the score-slab kernels run only for queries the proof hands to the exact fallback (ties beyond the lists)
            assert native.get_stat("select_f32_launches") == (1 if native.get_stat("fallback_queries") > 0 else 0)
        finally:
            native.set_option("profile", 0)
        parity.check_topk(idx, sc, q, c, k, metric, oracle, exact=True)
    native.set_option("profile", 1)
    native.reset_stats()
    try:
        idx, sc = native.topk(_hm(q[:20]), _hm(c), 2001, "dot")                       # beyond the multi-pass range: slab path
        assert native.get_stat("select_f32_launches") >= 1
    finally:
        native.set_option("profile", 0)
    parity.check_topk(idx, sc, q[:20], c, 2001, "dot", oracle, exact=True)
    # massive duplication: more exact ties than all passes collect -> the proof fails -> exact fallback, lower index first
    base = _randn(rng, 3, 80)
    cd = np.concatenate([base[rng.integers(0, 3, size=4000)], _randn(rng, 500, 80)])
    native.reset_stats()
    idx, sc = native.topk(_hm(q[:16]), _hm(cd), 300, "cosine")
    assert native.get_stat("fallback_queries") > 0
    parity.check_topk(idx, sc, q[:16], cd, 300, "cosine", oracle, exact=True)
    # the resident handle serves large k from the same planes
    h = native.ResidentCorpus(_hm(c), native.DTYPE_F32)
    try:
        i2, s2 = h.topk(_hm(q), 500, "dot")
    finally:
        h.close()
    parity.check_topk(i2, s2, q, c, 500, "dot", oracle, exact=True)
    # switched off: the slab path, same answer
    native.set_option("multipass", 0)
    try:
        i3, s3 = native.topk(_hm(q), _hm(c), 500, "dot")
    finally:
        native.set_option("multipass", 1)
    assert np.array_equal(i3, i2) and np.array_equal(s3, s2)


def test_f64_path(native, oracle):
    rng = np.random.default_rng(9)
    q, c = _randn(rng, 64, 64, dtype=np.float64), _randn(rng, 2000, 64, dtype=np.float64)
    # default: tensor-core filter on operands rounded to f16 + exact f64 re-scoring (sequential FMA, the oracle's order)
    # + per-query losslessness proof: scores and indices bit-identical to the oracle, no Q x N slab
    native.reset_stats()
    native.set_option("profile", 1)
    try:
        for metric in ("cosine", "dot", "euclidean"):
            idx, sc = native.topk(_hm(q), _hm(c), 10, metric)
            parity.check_topk(idx, sc, q, c, 10, metric, oracle, exact=True)
        assert native.get_stat("rescore_f64_launches") == 3 and native.get_stat("select_f64_launches") == 0
        assert native.get_stat("scores_f64_dmma_launches") == 0
    finally:
        native.set_option("profile", 0)
    # the slab path (DMMA scores + select): within 1e-12 relative, indices equal up to near-ties
    native.set_option("f64_tc", 0)
    try:
        for metric in ("cosine", "dot", "euclidean"):
            idx, sc = native.topk(_hm(q), _hm(c), 10, metric)
            frac = parity.check_topk(idx, sc, q, c, 10, metric, oracle)
            assert frac == 1.0
    finally:
        native.set_option("f64_tc", 1)
    out = native.matmul(_hm(q), _hm(c))
    assert out.dtype == np.float64
    parity.check_matmul(out, q, c, oracle.matmul(q, c), np.float64)
    # odd shapes through the DMMA tiles (row/column/k tails)
    q2, c2 = _randn(rng, 131, 37, dtype=np.float64), _randn(rng, 77, 37, dtype=np.float64)
    parity.check_matmul(native.matmul(_hm(q2), _hm(c2)), q2, c2, oracle.matmul(q2, c2), np.float64)
    # sequential-FMA kernel: bit-identical
    native.set_option("f64_simt", 1)
    native.set_option("f64_tc", 0)
    try:
        for metric in ("cosine", "dot", "euclidean"):
            idx, sc = native.topk(_hm(q), _hm(c), 10, metric)
            parity.check_topk(idx, sc, q, c, 10, metric, oracle, exact=True)
        assert np.array_equal(native.matmul(_hm(q), _hm(c)), oracle.matmul(q, c))
    finally:
        native.set_option("f64_simt", 0)
        native.set_option("f64_tc", 1)


@pytest.mark.parametrize("metric", ["cosine", "dot", "euclidean"])
def test_f64_fused_path_shapes_and_levels(native, oracle, metric):
    """f64 working precision on the fused path: shapes with row/column/k tails, mixed storage dtypes (widened exactly,
    src/matmul.rs:308), k up to 248, magnitudes outside the f16 and f32 ranges (the proof must escalate: f16-rounded ->
    3xTF32 -> exact f64 SIMT), exact ties. Always bit-identical to the oracle."""
    rng = np.random.default_rng(90)
    for nq, n, d, k in ((300, 5000, 96, 10), (33, 700, 257, 100), (5, 3000, 64, 200), (1, 40, 3, 40)):
        q, c = _randn(rng, nq, d, dtype=np.float64), _randn(rng, n, d, dtype=np.float64)
        idx, sc = native.topk(_hm(q), _hm(c), k, metric)
        parity.check_topk(idx, sc, q, c, k, metric, oracle, exact=True)
    q, c = _randn(rng, 40, 48, dtype=np.float64), _randn(rng, 3000, 48, dtype=np.float64)
    # mixed f32 / f64 and f16 / f64 storage
    for qd, cd in ((np.float32, np.float64), (np.float64, np.float32), (np.float16, np.float64)):
        qq, cc = q.astype(qd), c.astype(cd)
        idx, sc = native.topk(_hm(qq), _hm(cc), 7, metric)
        parity.check_topk(idx, sc, qq.astype(np.float64), cc.astype(np.float64), 7, metric, oracle, working_dtype=np.float64, exact=True)
    # beyond the f16 range (1e6), beyond the f32 range (1e60), far below both (1e-50), mixed in one corpus
    big = c.copy()
    big[::5] *= 1e6
    huge = c.copy()
    huge[::11] *= 1e60
    tiny = np.concatenate([c[:1500], c[1500:] * 1e-50])
    native.reset_stats()
    for corpus in (big, huge, tiny):
        idx, sc = native.topk(_hm(q), _hm(corpus), 9, metric)
        parity.check_topk(idx, sc, q, corpus, 9, metric, oracle, exact=True)
    assert native.get_stat("requeried_tf32x3") > 0 and native.get_stat("fallback_queries") > 0
    # exact ties: integer data and duplicated rows
    qi = rng.integers(-3, 4, size=(30, 20)).astype(np.float64)
    ci = rng.integers(-3, 4, size=(900, 20)).astype(np.float64)
    ci[500:700] = ci[:200]
    idx, sc = native.topk(_hm(qi), _hm(ci), 50, metric)
    parity.check_topk(idx, sc, qi, ci, 50, metric, oracle, exact=True)


def test_mixed_dtype_uses_f64(native, oracle):
    rng = np.random.default_rng(10)
    q, c = _randn(rng, 5, 16, dtype=np.float32), _randn(rng, 40, 16, dtype=np.float64)
    out = native.matmul(_hm(q), _hm(c))
    assert out.dtype == np.float64
    parity.check_matmul(out, q, c, oracle.matmul(q, c), np.float64)
    idx, sc = native.topk(_hm(q), _hm(c), 3, "cosine")
    assert parity.check_topk(idx, sc, q, c, 3, "cosine", oracle) == 1.0


# ---------------------------------------------------------------------------------------------- tensor-core path
@pytest.mark.parametrize("metric", ["cosine", "dot", "euclidean"])
@pytest.mark.parametrize("shape", [(300, 5000, 96, 10), (257, 3000, 100, 100), (64, 9000, 768, 128),
                                   (1, 300, 8, 1), (130, 257, 33, 64)])
def test_tc_topk_f32(native, oracle, metric, shape):
    nq, n, d, k = shape
    rng = np.random.default_rng(nq + n + d)
    q, c = _randn(rng, nq, d), _randn(rng, n, d)
    idx, sc = native.topk(_hm(q), _hm(c), k, metric)
    # tensor-core filter + exact re-scoring: bit-identical to the oracle (scores AND indices)
    parity.check_topk(idx, sc, q, c, k, metric, oracle, exact=True)


def test_tc_topk_mid_size_long_scan(native, oracle):
    """300k corpus rows (1172 corpus tiles per query tile): the per-row thresholds go through their whole life -
    open lists, hard and soft merges, settled phase - on both epilogue warp sets; every query compared with the oracle."""
    rng = np.random.default_rng(2024)
    q, c = _randn(rng, 384, 64), _randn(rng, 300_000, 64)
    c[rng.integers(0, 300_000, size=2000)] = c[rng.integers(0, 300_000, size=2000)]   # exact ties at arbitrary ranks
    for metric, k in (("dot", 100), ("cosine", 10), ("euclidean", 33)):
        idx, sc = native.topk(_hm(q), _hm(c), k, metric)
        parity.check_topk(idx, sc, q, c, k, metric, oracle, exact=True)
    h, hc = q.astype(np.float16), c[:200_000].astype(np.float16)
    idx, sc = native.topk(_hm(h), _hm(hc), 100, "dot")
    parity.check_topk(idx, sc, h.astype(np.float32), hc.astype(np.float32), 100, "dot", oracle, exact=True)


def test_tc_readme_config_c1(native, oracle):
    # BASELINE.json configs[0]: 1000 x 10000, 256-d f32, cosine, k=10; same generator as
    # examples/benchmark_topk.py:69-71
    np.random.seed(42)
    q = np.random.randn(1000, 256).astype(np.float32)
    c = np.random.randn(10000, 256).astype(np.float32)
    idx, sc = native.topk(_hm(q), _hm(c), 10, "cosine")
    parity.check_topk(idx, sc, q, c, 10, "cosine", oracle, exact=True)
    # and against the reference's own NumPy comparator (sorted scores, rtol 1e-4, benchmark_topk.py:122-138)
    _, ns = oracle.numpy_topk_cosine(q, c, 10)
    np.testing.assert_allclose(sc, ns, rtol=1e-4)


def test_tc_matmul_f32_c2(native, oracle):
    # BASELINE.json configs[1]: raw matmul 1000 x 10000 x 256
    np.random.seed(42)
    q = np.random.randn(1000, 256).astype(np.float32)
    c = np.random.randn(10000, 256).astype(np.float32)
    out = native.matmul(_hm(q), _hm(c))
    assert out.dtype == np.float32 and out.shape == (1000, 10000)
    parity.check_matmul(out, q, c, oracle.matmul(q, c), np.float32)
    q64, c64 = q[:200].astype(np.float64), c[:3000].astype(np.float64)
    out = native.matmul(_hm(q64), _hm(c64))
    parity.check_matmul(out, q64, c64, oracle.matmul(q64, c64), np.float64)


@pytest.mark.parametrize("shape", [(7, 13, 5), (129, 1000, 257), (300, 257, 64)])
def test_tc_matmul_odd_shapes(native, oracle, shape):
    nq, n, d = shape
    rng = np.random.default_rng(nq * n)
    q, c = _randn(rng, nq, d), _randn(rng, n, d)
    out = native.matmul(_hm(q), _hm(c))
    parity.check_matmul(out, q, c, oracle.matmul(q, c), np.float32)


@pytest.mark.parametrize("d", [1, 8, 9, 32, 33, 64, 65, 130, 192, 255, 256])
def test_matmul_f16_split_planes(native, oracle, d):
    """Raw f32 matmul on the tensor cores runs on row-scaled hi/lo f16 planes (one, several and an odd number of 64-element
    K-blocks; the second sweep packs two K-blocks per stage) for 32 < D <= 256; up to 8 elements the exact kernel, up to 32
    the TF32 planes.  Per-row magnitudes over 13 decades exercise the power-of-two scaling; small integers must come out
    exactly; the 3xTF32 planes stay available as an option."""
    rng = np.random.default_rng(1000 + d)
    q, c = _randn(rng, 150, d), _randn(rng, 1100, d)
    native.set_option("profile", 1)
    try:
        native.reset_stats()
        out = native.matmul(_hm(q), _hm(c))
        assert native.get_stat("scores_f32_launches" if d <= 8 else "tc_matmul_tf32x3_launches" if d <= 32 else "tc_matmul_f16x3_launches") == 1
    finally:
        native.set_option("profile", 0)
    ref = oracle.matmul(q, c)
    if d <= 8:
        assert np.array_equal(out, ref)
    parity.check_matmul(out, q, c, ref, np.float32)
    qs = q * (10.0 ** rng.uniform(-8, 5, size=(150, 1))).astype(np.float32)
    cs = c * (10.0 ** rng.uniform(-8, 5, size=(1100, 1))).astype(np.float32)
    parity.check_matmul(native.matmul(_hm(qs), _hm(cs)), qs, cs, oracle.matmul(qs, cs), np.float32)
    qi = rng.integers(-3, 4, size=(150, d)).astype(np.float32)
    ci = rng.integers(-3, 4, size=(1100, d)).astype(np.float32)
    assert np.array_equal(native.matmul(_hm(qi), _hm(ci)), oracle.matmul(qi, ci))
    native.set_option("matmul_split16", 0)
    native.set_option("profile", 1)
    try:
        native.reset_stats()
        out3 = native.matmul(_hm(q), _hm(c))
        assert native.get_stat("tc_matmul_tf32x3_launches") == (1 if d > 8 else 0)
    finally:
        native.set_option("matmul_split16", 1)
        native.set_option("profile", 0)
    assert np.abs(out3 - ref).max() <= 1e-5 * np.abs(ref).max()


@pytest.mark.parametrize("shape", [(2600, 1500, 96), (300, 5000, 200), (5000, 300, 24), (10000, 3000, 96), (12000, 2000, 20)])
def test_matmul_tile_schedules_agree(native, oracle, shape):
    """The flat tile schedule (equal shares of the tile list per CTA pair: shares start and end inside a query tile, the
    resident query planes are reloaded at the boundary), the hybrid one (main sweeps + helper pairs on the sweeps' tails,
    for 37 < query tiles < 74: the last two shapes) and the classic one give the same matrix."""
    nq, n, d = shape
    rng = np.random.default_rng(nq + n)
    q, c = _randn(rng, nq, d), _randn(rng, n, d)
    outs = []
    for flat in (1, 2, 0):
        native.set_option("matmul_flat", flat)
        try:
            outs.append(native.matmul(_hm(q), _hm(c)))
        finally:
            native.set_option("matmul_flat", -1)
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    parity.check_matmul(outs[0], q, c, oracle.matmul(q, c), np.float32)


def test_matmul_f16_split_extreme_rows(native, oracle):
    """Rows the power-of-two scaling cannot serve - a largest element beyond 2^+-60, inf / NaN, all-zero rows, elements
    25 binades below the row's largest - keep the reference's result: the first two are recomputed with IEEE arithmetic
    (bit-identical to the oracle), zero rows give zeros, tiny elements are off by far less than the tolerance."""
    rng = np.random.default_rng(77)
    q, c = _randn(rng, 70, 96), _randn(rng, 900, 96)
    q[3] *= np.float32(1e25)
    q[4] *= np.float32(1e-25)
    q[5] = 0
    q[6, ::2] *= np.float32(3e-8)
    q[7, 10] = np.inf
    c[11] *= np.float32(1e-30)
    c[12] *= np.float32(1e22)
    c[13] = 0
    c[14, 5] = np.nan
    c[15, 1::3] *= np.float32(1e-9)
    out = native.matmul(_hm(q), _hm(c))
    ref = oracle.matmul(q, c)
    special_q, special_c = [3, 4, 7], [11, 12, 14]
    assert np.array_equal(out[special_q], ref[special_q], equal_nan=True)
    assert np.array_equal(out[:, special_c], ref[:, special_c], equal_nan=True)
    keep_q = [i for i in range(70) if i not in special_q]
    keep_c = [j for j in range(900) if j not in special_c]
    assert not out[5][keep_c].any() and not out[keep_q, 13].any()
    parity.check_matmul(out[np.ix_(keep_q, keep_c)], q[keep_q], c[keep_c], ref[np.ix_(keep_q, keep_c)], np.float32)


def test_matmul_f16_split_mixed_storage_and_lists(native, oracle):
    """f16-stored rows beside f32 rows (working precision f32, src/matmul.rs:13-19 as relaxed by the README contract) and a
    pl.List column with nulls and short rows go through the same split planes."""
    import pyarrow as pa
    rng = np.random.default_rng(5)
    q, c = _randn(rng, 90, 100), _randn(rng, 700, 100).astype(np.float16)
    out = native.matmul(_hm(q), _hm(c))
    parity.check_matmul(out, q, c.astype(np.float32), oracle.matmul(q, c.astype(np.float32)), np.float32)
    rows, dense = [], np.zeros((90, 100), np.float32)
    for i in range(90):
        if i % 11 == 5:
            rows.append(None)
            continue
        ln = 100 if i == 0 or i % 7 else int(rng.integers(0, 100))
        vals = rng.standard_normal(ln).astype(np.float32)
        dense[i, :ln] = vals
        r = vals.tolist()
        if ln > 3 and i % 5 == 0:
            r[2] = None
            dense[i, 2] = 0
        rows.append(r)
    col = pa.array(rows, type=pa.large_list(pa.float32()))
    cf = c.astype(np.float32)
    out = native.matmul(_hm(col), _hm(cf))
    parity.check_matmul(out, dense, cf, oracle.matmul(dense, cf), np.float32)


@pytest.mark.parametrize("d", [320, 385, 768, 1024, 2048])
def test_matmul_long_vectors_stay_within_tolerance(native, oracle, d):
    """tcgen05 accumulates f32 with truncation, and the raw matmul has no exact re-scoring behind it: beyond D = 256 (f32,
    3xTF32) / 1024 (f16 storage) it runs on the exact sequential-FMA kernel instead - bit-identical to the oracle - so that
    the stated 1e-5 holds for every vector length, on same-sign data (where the truncation bias accumulates) as well."""
    rng = np.random.default_rng(d)
    for kind in ("gauss", "positive"):
        q = (_randn(rng, 200, d) if kind == "gauss" else rng.uniform(0.5, 1.5, size=(200, d)).astype(np.float32))
        c = (_randn(rng, 1500, d) if kind == "gauss" else rng.uniform(0.5, 1.5, size=(1500, d)).astype(np.float32))
        out = native.matmul(_hm(q), _hm(c))
        ref = oracle.matmul(q, c)
        assert np.array_equal(out, ref), (kind, np.abs(out - ref).max())
        h, hc = q.astype(np.float16), c.astype(np.float16)
        out16 = native.matmul(_hm(h), _hm(hc))
        parity.check_matmul(out16, h.astype(np.float32), hc.astype(np.float32), oracle.matmul(h.astype(np.float32), hc.astype(np.float32)), np.float32)
    # the limit is an option: forcing the tensor cores at this length still gives a result close to the oracle (diagnostics)
    native.set_option("matmul_tc_max_dim", 1 << 20)
    try:
        out_tc = native.matmul(_hm(q), _hm(c))
    finally:
        native.set_option("matmul_tc_max_dim", 0)
    np.testing.assert_allclose(out_tc, ref, rtol=1e-4)


def test_matmul_infinite_inputs_propagate(native, oracle):
    """+-inf elements: the 3xTF32 split would turn them into NaN (0 * inf in the lo*hi term); the rows are marked by the prep
    pass and recomputed with IEEE arithmetic, so the result equals the reference's propagation (src/metrics.rs:160-202)."""
    rng = np.random.default_rng(91)
    q, c = _randn(rng, 130, 64), _randn(rng, 700, 64)
    q[3, 5] = np.inf
    q[77, 0] = -np.inf
    c[10, 63] = np.inf
    c[300, 1] = np.nan
    out = native.matmul(_hm(q), _hm(c))
    ref = oracle.matmul(q, c)
    same = (out == ref) | (np.isnan(out) & np.isnan(ref))
    marked_rows, marked_cols = [3, 77], [10, 300]
    assert same[marked_rows].all() and same[:, marked_cols].all()
    assert np.isinf(out[3]).sum() > 600 and np.isnan(out[:, 300]).all()
    rest = np.ones_like(same)
    rest[marked_rows] = False
    rest[:, marked_cols] = False
    parity.check_matmul(np.where(rest, out, 0), np.where(np.isfinite(q), q, 0), np.where(np.isfinite(c), c, 0),
                        np.where(rest, ref, 0), np.float32)


def test_tie_stress(native, oracle):
    # exact ties: duplicated corpus rows, small-integer vectors, zero vectors (north_star tie rule)
    rng = np.random.default_rng(3)
    base = rng.integers(-3, 4, size=(50, 32)).astype(np.float32)
    c = np.concatenate([base, base, np.zeros((5, 32), np.float32), base[:10]], 0)   # many duplicates
    q = np.concatenate([base[:20], np.zeros((2, 32), np.float32)], 0)
    for metric in ("cosine", "dot", "euclidean"):
        for k in (1, 7, 40, 115):
            idx, sc = native.topk(_hm(q), _hm(c), k, metric)
            # integer-valued data: every product and partial sum is exact in f32, so the tensor-core path
            # must reproduce the oracle's indices exactly, ties included
            oi, osc = oracle.topk(q, c, k, metric)
            assert np.array_equal(idx, oi), (metric, k, np.argwhere(idx != oi)[:4].tolist())
            parity.check_topk(idx, sc, q, c, k, metric, oracle)


def test_filter_levels_and_fallback_are_exercised(native, oracle):
    """First-level filter (f16-rounded operands, 11 significant bits) -> 3xTF32 re-query -> exact SIMT fallback: craft inputs that need each level and check
    via the library's statistics that the level actually ran; the result must still equal the oracle."""
    rng = np.random.default_rng(123)
    d, n, k = 64, 4000, 10
    # (a) well separated scores: level 0 suffices
    q = _randn(rng, 64, d)
    c = _randn(rng, n, d)
    native.reset_stats()
    idx, sc = native.topk(_hm(q), _hm(c), k, "dot")
    parity.check_topk(idx, sc, q, c, k, "dot", oracle, exact=True)
    assert native.get_stat("requeried_tf32x3") == 0 and native.get_stat("fallback_queries") == 0
    # (b) 60 corpus rows spaced 1e-5 relative at the top: rank 10 and rank 32 are 2.2e-4 apart, below the TF32 x1
    #     bound (~1e-3) but far above the 3xTF32 bound (~1e-5): level 0 must hand these queries to level 1
    base = _randn(rng, 1, d)
    c2 = c.copy()
    c2[:60] = base * (1.0 + 1e-5 * np.arange(60, dtype=np.float32)[:, None]) * 3.0
    q2 = np.repeat(base, 8, axis=0) + 1e-3 * _randn(rng, 8, d)
    native.reset_stats()
    idx, sc = native.topk(_hm(q2), _hm(c2), k, "dot")
    parity.check_topk(idx, sc, q2, c2, k, "dot", oracle, exact=True)
    assert native.get_stat("requeried_f16_wide") > 0        # first retry: same filter, 256-entry lists (rank 256 is far enough)
    assert native.get_stat("requeried_tf32x3") == 0
    native.set_option("f16r_wide", 0)                        # without that retry the queries go straight to 3xTF32
    try:
        native.reset_stats()
        i0, s0 = native.topk(_hm(q2), _hm(c2), k, "dot")
        assert native.get_stat("requeried_tf32x3") > 0 and native.get_stat("requeried_f16_wide") == 0
        assert np.array_equal(i0, idx) and np.array_equal(s0, sc)
    finally:
        native.set_option("f16r_wide", 1)
    # (b2) 400 rows spaced 2e-6: even rank 256 is within the first level's bound -> wide retry fails too -> 3xTF32
    c2b = c.copy()
    c2b[:400] = base * (1.0 + 2e-6 * np.arange(400, dtype=np.float32)[:, None]) * 3.0
    native.reset_stats()
    idx, sc = native.topk(_hm(q2), _hm(c2b), k, "dot")
    parity.check_topk(idx, sc, q2, c2b, k, "dot", oracle, exact=True)
    assert native.get_stat("requeried_f16_wide") > 0 and native.get_stat("requeried_tf32x3") > 0
    # (c) 300 exact duplicates at the top (more than the longest candidate list): no filter can prove anything about
    #     exact ties it did not keep -> exact SIMT path. (60 duplicates are settled by the 256-entry retry.)
    c3 = c.copy()
    c3[100:160] = base * 3.0
    native.reset_stats()
    idx, sc = native.topk(_hm(q2), _hm(c3), k, "cosine")
    parity.check_topk(idx, sc, q2, c3, k, "cosine", oracle, exact=True)
    assert native.get_stat("requeried_f16_wide") > 0 and native.get_stat("fallback_queries") == 0
    c3[100:400] = base * 3.0
    native.reset_stats()
    idx, sc = native.topk(_hm(q2), _hm(c3), k, "cosine")
    parity.check_topk(idx, sc, q2, c3, k, "cosine", oracle, exact=True)
    assert native.get_stat("fallback_queries") > 0
    assert (idx[:, 0] == 100).all()          # lowest index among the exact ties first
    # (d) every choice of first level (3: f16-rounded operands [default], 2: TF32 x1, 1: 3xTF32) gives the same answers
    results = []
    try:
        for levels in (1, 2, 3):
            native.set_option("tc_levels", levels)
            results.append(native.topk(_hm(q2), _hm(c2), k, "dot"))
    finally:
        native.set_option("tc_levels", 3)
    for i1, s1 in results[1:]:
        assert np.array_equal(i1, results[0][0]) and np.array_equal(s1, results[0][1])


def test_f16_rounded_level_handles_range(native, oracle):
    """The default first level rounds f32 operands to f16: values beyond 65504 overflow there and values below 6e-5
    become subnormal. The losslessness proof must notice both (norm limit, absolute error term) and hand the queries
    to the 3xTF32 level; results stay the oracle's for every metric."""
    rng = np.random.default_rng(321)
    d, n, k = 96, 6000, 10
    q, c = _randn(rng, 48, d), _randn(rng, n, d)
    # (a) huge corpus values (f16 overflow): every query is re-run one level up
    big = c.copy()
    big[::7] *= 3.0e5
    native.reset_stats()
    for metric in ("dot", "cosine", "euclidean"):
        idx, sc = native.topk(_hm(q), _hm(big), k, metric)
        parity.check_topk(idx, sc, q, big, k, metric, oracle, exact=True)
    assert native.get_stat("requeried_tf32x3") >= 3 * q.shape[0]
    # (b) huge query rows only: only those queries are re-run
    qb = q.copy()
    qb[:5] *= 1.0e6
    native.reset_stats()
    idx, sc = native.topk(_hm(qb), _hm(c), k, "dot")
    parity.check_topk(idx, sc, qb, c, k, "dot", oracle, exact=True)
    assert 5 <= native.get_stat("requeried_tf32x3") < q.shape[0]
    # (c) tiny magnitudes (f16 subnormals / flush): whole corpus scaled by 1e-6, and a mix of scales within one corpus
    for corpus in (c * np.float32(1e-6), np.concatenate([c[:3000], c[3000:] * np.float32(1e-7)])):
        for metric in ("cosine", "dot", "euclidean"):
            idx, sc = native.topk(_hm(q), _hm(corpus), k, metric)
            parity.check_topk(idx, sc, q, corpus, k, metric, oracle, exact=True)
    # (e) regression (found by the property soak): fewer corpus rows than the candidate list holds, one of them beyond
    #     the f16 range. Nothing is dropped by the filter, but the re-scoring must not skip candidates on the strength
    #     of filter values that overflowed.
    small = _randn(rng, 6, d)
    small[2] *= 1.0e5
    for metric in ("euclidean", "dot", "cosine"):
        idx, sc = native.topk(_hm(q[:2]), _hm(small), 1, metric)
        parity.check_topk(idx, sc, q[:2], small, 1, metric, oracle, exact=True)
    # (d) tiny queries against a normal corpus
    qs = q * np.float32(1e-7)
    for metric in ("cosine", "dot"):
        idx, sc = native.topk(_hm(qs), _hm(c), k, metric)
        parity.check_topk(idx, sc, qs, c, k, metric, oracle, exact=True)


def test_nan_and_inf_rank_last(native, oracle):
    c = np.ones((300, 16), np.float32)
    c[5, 0] = np.nan
    c[9, 3] = np.inf
    c[11] *= 2
    q = np.ones((3, 16), np.float32)
    idx, sc = native.topk(_hm(q), _hm(c), 300, "dot")          # k > 128: SIMT path, IEEE semantics throughout
    oi, osc = oracle.topk(q, c, 300, "dot")
    assert np.array_equal(idx, oi)
    assert np.isnan(sc[:, -1]).all() and idx[0, 0] == 9 and idx[0, 1] == 11
    # tensor-core path: NaN ranks last as well (an infinite INPUT turns into NaN under the 3xTF32 split,
    # 0 * inf in the lo*hi term — documented deviation)
    c[9, 3] = 1.0
    idx, sc = native.topk(_hm(q), _hm(c), 100, "dot")
    oi, osc = oracle.topk(q, c, 100, "dot")
    assert np.array_equal(idx, oi) and idx[0, 0] == 11
    idx, sc = native.topk(_hm(q), _hm(c[:40]), 40, "dot")
    assert idx[0, -1] == 5 and np.isnan(sc[0, -1])


def test_zero_norm_cosine(native, oracle):
    c = np.array([[0, 0, 0, 0], [1, 0, 0, 0], [0, 2, 0, 0]], np.float32)
    q = np.array([[1, 1, 0, 0], [0, 0, 0, 0]], np.float32)
    idx, sc = native.topk(_hm(q), _hm(c), 3, "cosine")
    oi, osc = oracle.topk(q, c, 3, "cosine")
    assert np.array_equal(idx, oi) and np.array_equal(sc, osc)


# ---------------------------------------------------------------------------------------------- containers / dtypes
def test_list_input_with_nulls_and_short_rows(pmm, native, oracle):
    import pyarrow as pa
    rng = np.random.default_rng(5)
    dense = _randn(rng, 40, 12)
    rows = [r.tolist() for r in dense]
    rows[3] = rows[3][:7]            # short row -> zero padded (src/matmul.rs:247-254)
    rows[5][2] = None                # null element -> 0.0
    rows[8] = None                   # null row -> zeros
    arr = pa.array(rows, type=pa.large_list(pa.float32()))
    dense[3, 7:] = 0
    dense[5, 2] = 0
    dense[8] = 0
    q = _randn(rng, 6, 12)
    idx, sc = pmm.topk_arrays(q, arr, 5, "dot")
    parity.check_topk(idx, sc, q, dense, 5, "dot", oracle)
    out = pmm.matmul_array(arr, q)
    parity.check_matmul(out, dense, q, oracle.matmul(dense, q), np.float32)
    # Array (fixed-size list) container and int32 offsets
    fsl = pa.FixedSizeListArray.from_arrays(pa.array(dense.reshape(-1)), 12)
    idx2, sc2 = pmm.topk_arrays(q, fsl, 5, "dot")
    assert np.array_equal(idx, idx2)
    with pytest.raises(RuntimeError, match="ragged"):
        bad = pa.array([[1.0, 2.0], [1.0, 2.0, 3.0]], type=pa.list_(pa.float32()))
        pmm.topk_arrays(np.ones((1, 2), np.float32), bad, 1, "dot")


def test_integer_columns_are_cast_to_f64(pmm, oracle):
    import pyarrow as pa
    arr = pa.array([[1, 2], [3, 4], [5, 6]], type=pa.list_(pa.int64()))
    out = pmm.matmul_array(arr, arr)
    assert out.dtype == np.float64
    assert out.tolist() == [[5, 11, 17], [11, 25, 39], [17, 39, 61]]


def test_f16_storage(native, oracle):
    rng = np.random.default_rng(11)
    q16, c16 = _randn(rng, 70, 128).astype(np.float16), _randn(rng, 4000, 128).astype(np.float16)
    q32, c32 = q16.astype(np.float32), c16.astype(np.float32)   # reference contract: exact upcast, then f32 path
    for metric in ("cosine", "dot", "euclidean"):
        idx, sc = native.topk(_hm(q16), _hm(c16), 10, metric)
        parity.check_topk(idx, sc, q32, c32, 10, metric, oracle, working_dtype=np.float32, exact=True)
    out = native.matmul(_hm(q16), _hm(c16))
    assert out.dtype == np.float32
    parity.check_matmul(out, q32, c32, oracle.matmul(q32, c32), np.float32)
    # mixed f16 / f32 stays on the f32 path (planes built from f16 AND f32 sources; D = 128: the vectorised plane pass)
    idx, sc = native.topk(_hm(q32), _hm(c16), 10, "cosine")
    parity.check_topk(idx, sc, q32, c32, 10, "cosine", oracle, working_dtype=np.float32, exact=True)
    out = native.matmul(_hm(q32), _hm(c16))
    parity.check_matmul(out, q32, c32, oracle.matmul(q32, c32), np.float32)
    out = native.matmul(_hm(q16), _hm(c32))
    parity.check_matmul(out, q32, c32, oracle.matmul(q32, c32), np.float32)


def test_result_containers(pmm):
    import pyarrow as pa
    q = np.eye(3, dtype=np.float32)
    c = np.eye(3, dtype=np.float32)[[2, 0, 1]]
    r = pmm._topk(q, c, 2, "cosine")
    assert r.type == pa.large_list(pa.struct([("index", pa.uint32()), ("score", pa.float64())]))
    assert [x[0]["index"] for x in r.to_pylist()] == [1, 2, 0]
    m = pmm._matmul(q, c)
    assert m.type == pa.list_(pa.float32(), 3)
    assert pmm._matmul(np.empty((0, 3), np.float32), c).type == pa.large_list(pa.float32())
    assert len(pmm._topk(np.empty((0, 3)), c, 2, "not-a-metric")) == 0   # empty query short-circuits first
    assert len(pmm._topk(q, c, 0, "dot").to_pylist()[0]) == 0             # k = 0 -> empty lists


def test_results_come_from_the_pinned_pool(native, oracle):
    """Large result buffers are page-locked blocks from pmm_host_alloc, recycled once the arrays are dropped."""
    rng = np.random.default_rng(5)
    q, c = _randn(rng, 3000, 32), _randn(rng, 500, 32)
    idx, sc = native.topk(_hm(q), _hm(c), 100, "dot")           # 1.2 MB + 2.4 MB of results
    parity.check_topk(idx[:50], sc[:50], q[:50], c, 100, "dot", oracle, exact=True)
    addr = sc.ctypes.data
    keep = sc.copy()
    del idx, sc
    idx2, sc2 = native.topk(_hm(q), _hm(c), 100, "dot")
    assert sc2.ctypes.data == addr                               # same block again
    assert np.array_equal(sc2, keep)
    small_i, small_s = native.topk(_hm(q[:4]), _hm(c), 5, "dot")  # tiny results stay ordinary arrays
    assert small_s.base is None


def test_resident_corpus_handle(native, oracle):
    rng = np.random.default_rng(12)
    q, c = _randn(rng, 50, 64), _randn(rng, 3000, 64)
    h = native.ResidentCorpus(_hm(c), native.DTYPE_F32)
    try:
        for metric in ("cosine", "euclidean", "dot"):
            idx, sc = h.topk(_hm(q), 10, metric)
            i2, s2 = native.topk(_hm(q), _hm(c), 10, metric)
            assert np.array_equal(idx, i2) and np.array_equal(sc, s2)
    finally:
        h.close()


@pytest.mark.parametrize("seed", range(int(os.environ.get("PMM_WARM_CASES", "12"))))   # raise for a soak on the GPU box
def test_warm_seeds_randomised(native, oracle, seed):
    """Seeded random shapes above the warm-seed thresholds (the hypothesis suite stays below 2048 queries): k, metric,
    vector length, corpus size and data kind vary - Gaussian, clustered (queries sit on cluster centres, the corpus is
    ordered cluster by cluster), duplicated rows, zero rows, per-row scales over six decades - alternating between the
    resident-style and the chunked host path.  Bit-identical to the oracle every time."""
    rng = np.random.default_rng(1000 + seed)
    nq = int(rng.integers(2048, 2600))
    n = int(rng.integers(70_000, 140_000))
    d = int(rng.choice([16, 24, 40, 64]))
    k = int(rng.choice([1, 5, 10, 24, 25, 56, 100]))
    metric = ["dot", "cosine", "euclidean"][seed % 3]
    kind = ["gauss", "clustered", "dups", "zeros", "scaled", "clustered"][seed % 6]
    q, c = _randn(rng, nq, d), _randn(rng, n, d)
    if kind == "clustered":
        centres = _randn(rng, 40, d) * np.float32(3.0)
        c = (np.repeat(centres, -(-n // 40), axis=0)[:n] + np.float32(0.3) * c).astype(np.float32)      # cluster by cluster
        q = (centres[rng.integers(0, 40, size=nq)] + np.float32(0.3) * q).astype(np.float32)
    elif kind == "dups":
        src = rng.integers(0, n, size=n // 3)
        c[rng.integers(0, n, size=n // 3)] = c[src]
    elif kind == "zeros":
        c[rng.integers(0, n, size=n // 10)] = 0
        q[rng.integers(0, nq, size=nq // 10)] = 0
    elif kind == "scaled":
        c *= (10.0 ** rng.uniform(-3, 3, size=(n, 1))).astype(np.float32)
        q *= (10.0 ** rng.uniform(-3, 3, size=(nq, 1))).astype(np.float32)
    if seed % 2:
        native.set_option("host_chunk_min_mb", 0)
    native.set_option("profile", 1)
    native.reset_stats()
    try:
        idx, sc = native.topk(_hm(q), _hm(c), k, metric)
        assert native.get_stat("tc_topk_warm_launches") >= 1, (n, k)
    finally:
        native.set_option("host_chunk_min_mb", 64)
        native.set_option("profile", 0)
    parity.check_topk(idx, sc, q, c, k, metric, oracle, exact=True)


@pytest.mark.parametrize("chunked", [False, True])
def test_warm_seeds_on_a_sorted_corpus(native, oracle, chunked):
    """A corpus sorted by norm (dot metric: every query's best rows come first) must not make the seeds too aggressive: the
    sample is strided over the whole corpus - on the device, or gathered from the host buffer before the first chunk is
    uploaded - so only a handful of rows need the re-query levels, and the result is the oracle's."""
    rng = np.random.default_rng(31)
    nq, n, d, k = 2304, 120_000, 48, 10
    q, c = _randn(rng, nq, d), _randn(rng, n, d)
    c *= np.linspace(4.0, 0.25, n, dtype=np.float32)[:, None]
    if chunked:
        native.set_option("host_chunk_min_mb", 0)
    native.set_option("profile", 1)
    native.reset_stats()
    try:
        idx, sc = native.topk(_hm(q), _hm(c), k, "dot")
        assert native.get_stat("tc_topk_warm_launches") >= 1
        flagged = native.get_stat("requeried_f16_wide") + native.get_stat("requeried_tf32x3") + native.get_stat("fallback_queries")
        assert flagged < 0.02 * nq, flagged
    finally:
        native.set_option("host_chunk_min_mb", 64)
        native.set_option("profile", 0)
    parity.check_topk(idx, sc, q, c, k, "dot", oracle, exact=True)


@pytest.mark.parametrize("case", ["k10", "k100", "sample_dominates", "host_chunks", "f16"])
def test_warm_seeds_of_the_first_level(native, oracle, case):
    """Large calls start the first filter level from thresholds of a sample pre-pass (the r-th best of the first 4096
    corpus rows): same bits as without, for every metric; a sample that is far better than the rest of the corpus makes
    the seeds too aggressive - lists do not fill, the losslessness check re-queries those rows - and the result is
    still the oracle's; zero queries and duplicated top rows ride along."""
    rng = np.random.default_rng(len(case))
    if case == "k100":
        nq, n, d, k = 2304, 210_000, 32, 100
    else:
        nq, n, d, k = 2304, 100_000, 48, 10
    q, c = _randn(rng, nq, d), _randn(rng, n, d)
    q[5] = 0
    c[70_000:70_040] = c[3]                       # 41 identical rows: ties inside the top ranks of some queries
    if case == "sample_dominates":
        # dot / euclidean: all the extreme scores lie inside the (strided) sample - every 97th corpus tile of 256 rows at this
        # size - whose 8th best then exceeds the true 10th best: the lists cannot fill
        c[:256] *= np.float32(6.0)
    if case == "f16":
        q, c = q.astype(np.float16), c.astype(np.float16)
    if case == "host_chunks":
        native.set_option("host_chunk_min_mb", 0)
    try:
        for metric in ("dot", "cosine", "euclidean"):
            outs = []
            for warm in (1, 0):
                native.set_option("warm_seed", warm)
                native.set_option("profile", 1)
                native.reset_stats()
                try:
                    outs.append(native.topk(_hm(q), _hm(c), k, metric))
                    assert (native.get_stat("tc_topk_warm_launches") >= 1) == bool(warm)
                    if case == "sample_dominates" and warm and metric == "dot":
                        assert native.get_stat("requeried_f16_wide") + native.get_stat("requeried_tf32x3") + native.get_stat("fallback_queries") > 100
                finally:
                    native.set_option("warm_seed", 1)
                    native.set_option("profile", 0)
            assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1]), metric
            parity.check_topk(outs[0][0], outs[0][1], q.astype(np.float32), c.astype(np.float32), k, metric, oracle, exact=True)
    finally:
        native.set_option("host_chunk_min_mb", 64)


def test_results_written_straight_into_page_locked_buffers(native, oracle):
    """Host path with page-locked result buffers: the re-scoring kernel (and the scatter of re-queried rows) store into them
    directly, no trailing device->host copy; pageable buffers take the staged copy.  Same bits either way, incl. queries
    that go through the re-query levels (duplicated corpus rows: ties beyond the list)."""
    import ctypes
    rng = np.random.default_rng(123)
    q, c = _randn(rng, 3000, 192), _randn(rng, 90_000, 192)          # 69 MB corpus -> chunked host path
    c[1000:1300] = c[7]                                               # 300 identical rows: some queries need the exact fallback
    k = 40
    hq, hc = _hm(q), _hm(c)
    outs = {}
    for label, direct, pinned in (("direct", 1, True), ("copied", 0, True), ("pageable", 1, False)):
        if pinned:
            blocks = []
            for nbytes in (3000 * k * 4, 3000 * k * 8):
                p = ctypes.c_void_p()
                native.check(native.lib().pmm_host_alloc(nbytes, ctypes.byref(p)))
                blocks.append(p)
            idx = np.ctypeslib.as_array(ctypes.cast(blocks[0], ctypes.POINTER(ctypes.c_uint32)), shape=(3000, k))
            sc = np.ctypeslib.as_array(ctypes.cast(blocks[1], ctypes.POINTER(ctypes.c_double)), shape=(3000, k))
        else:
            blocks, idx, sc = [], np.empty((3000, k), np.uint32), np.empty((3000, k), np.float64)
        idx[:] = 0xdeadbeef
        sc[:] = -7.0
        ka = ctypes.c_int64(0)
        qs, cs = hq.c_struct(), hc.c_struct()
        native.set_option("d2h_direct", direct)
        try:
            native.check(native.lib().pmm_topk(ctypes.byref(qs), ctypes.byref(cs), k, b"dot", idx.ctypes.data, sc.ctypes.data, ctypes.byref(ka)))
        finally:
            native.set_option("d2h_direct", 1)
        assert ka.value == k
        outs[label] = (idx.copy(), sc.copy())
        for p in blocks:
            native.check(native.lib().pmm_host_free(p))
    parity.check_topk(outs["direct"][0], outs["direct"][1], q, c, k, "dot", oracle, exact=True)
    for label in ("copied", "pageable"):
        assert np.array_equal(outs[label][0], outs["direct"][0]) and np.array_equal(outs[label][1], outs["direct"][1])


def test_host_chunked_upload_path(pmm, native, oracle):
    """Corpora >= 64 MB take the chunked host path (upload of chunk i+1 overlaps compute of chunk i):
    same result as the oracle, for fixed-size rows and for list offsets with nulls."""
    import pyarrow as pa
    rng = np.random.default_rng(77)
    q, c = _randn(rng, 200, 192), _randn(rng, 90_000, 192)          # 69 MB corpus -> 2 chunks
    c[20_000] = c[3]                                                  # an exact tie across the chunk boundary
    for metric in ("cosine", "dot", "euclidean"):
        idx, sc = native.topk(_hm(q), _hm(c), 20, metric)
        parity.check_topk(idx, sc, q, c, 20, metric, oracle, exact=True)
    native.set_option("host_chunked", 0)
    try:
        i0, s0 = native.topk(_hm(q), _hm(c), 20, "cosine")
    finally:
        native.set_option("host_chunked", 1)
    i1, s1 = native.topk(_hm(q), _hm(c), 20, "cosine")
    assert np.array_equal(i0, i1) and np.array_equal(s0, s1)
    # f64 working precision takes the same overlapped path (f16-rounded planes for the filter, exact f64 norms and
    # re-scoring): 74 MB of f64 rows, and an f32 corpus queried with f64 vectors
    c64 = c[:48_000].astype(np.float64)
    q64 = q.astype(np.float64)
    for metric in ("cosine", "euclidean"):
        idx, sc = native.topk(_hm(q64), _hm(c64), 20, metric)
        parity.check_topk(idx, sc, q64, c64, 20, metric, oracle, exact=True)
    idx, sc = native.topk(_hm(q64), _hm(c), 20, "dot")
    parity.check_topk(idx, sc, q64, c.astype(np.float64), 20, "dot", oracle, working_dtype=np.float64, exact=True)
    # list layout: offsets + a null row + a null element + a short row, spread over both chunks
    flat = pa.array(c.reshape(-1))
    offsets = np.arange(0, (c.shape[0] + 1) * 192, 192, dtype=np.int64)
    lst = pa.LargeListArray.from_arrays(pa.array(offsets), flat)
    rows = lst.to_pylist()[:3]
    dense = c.copy()
    mask = np.ones(c.shape[0], bool)
    mask[[5, 50_000]] = False                                         # null rows -> zeros
    lst = pa.LargeListArray.from_arrays(pa.array(offsets), flat, mask=pa.array(~mask))
    dense[[5, 50_000]] = 0
    del rows
    idx, sc = pmm.topk_arrays(q, lst, 20, "dot")
    parity.check_topk(idx, sc, q, dense, 20, "dot", oracle, exact=True)


def test_host_chunked_many_chunks_carry_lists(native, oracle):
    """Eight equal corpus chunks: the candidate lists are carried from launch to launch (no per-chunk list warm-up,
    no cross-chunk merge). Ties across chunk boundaries, f32 (one epilogue set) and f16 (two sets)."""
    rng = np.random.default_rng(78)
    native.set_option("host_chunk_first_div", 8)
    native.set_option("host_chunk_ratio_pct", 100)
    try:
        q, c = _randn(rng, 150, 128), _randn(rng, 140_000, 128)     # 72 MB -> 8 chunks of 17408 rows (+ remainder)
        c[[17_500, 60_000, 139_999]] = c[7]                          # exact ties in different chunks
        q[3] = c[7]
        for metric, k in (("cosine", 10), ("dot", 100), ("euclidean", 30)):
            idx, sc = native.topk(_hm(q), _hm(c), k, metric)
            parity.check_topk(idx, sc, q, c, k, metric, oracle, exact=True)
        h, hc = _randn(rng, 150, 256).astype(np.float16), _randn(rng, 140_000, 256).astype(np.float16)   # 72 MB of f16
        hc[[100, 70_000]] = hc[139_000]
        idx, sc = native.topk(_hm(h), _hm(hc), 10, "cosine")
        parity.check_topk(idx, sc, h.astype(np.float32), hc.astype(np.float32), 10, "cosine", oracle, exact=True)
    finally:
        native.set_option("host_chunk_first_div", 0)
        native.set_option("host_chunk_ratio_pct", 0)


def test_host_chunked_requery_runs_piecewise(native, oracle):
    """Chunked upload + queries the first filter level cannot prove: the 3xTF32 level rebuilds the corpus planes
    piece by piece (131072 rows at a time) and carries the lists; massive exact ties push some queries on to the
    exact fallback. Result must still be the oracle's, tie order included."""
    rng = np.random.default_rng(79)
    base = _randn(rng, 40, 128)
    c = base[rng.integers(0, 40, size=150_000)]                      # 150k rows, only 40 distinct vectors: 77 MB
    c[::1000] += 1e-3 * _randn(rng, 150, 128)                        # and some near ties
    q = _randn(rng, 64, 128)
    native.reset_stats()
    idx, sc = native.topk(_hm(q), _hm(c), 10, "dot")
    parity.check_topk(idx, sc, q, c, 10, "dot", oracle, exact=True)
    assert native.get_stat("requeried_tf32x3") > 0
    assert native.get_stat("fallback_queries") > 0


# ---------------------------------------------------------------------------------------------- device entry points
def test_device_level_shards_merge(native, oracle):
    """Two corpus shards on one GPU -> packed candidates with global indices -> merge == unsharded."""
    import torch
    rng = np.random.default_rng(13)
    q, c = _randn(rng, 200, 96), _randn(rng, 6000, 96)
    dq = torch.from_numpy(q).cuda()
    k = 20
    for metric_name, metric in (("cosine", 0), ("dot", 1), ("euclidean", 2)):
        lists = torch.zeros((2, 200, k), dtype=torch.int64, device="cuda")
        bounds = [(0, 2500), (2500, 6000)]
        for s, (lo, hi) in enumerate(bounds):
            dc = torch.from_numpy(c[lo:hi]).cuda()
            native.dev_topk(native.dev_matrix(dq.data_ptr(), 200, 96, native.DTYPE_F32),
                            native.dev_matrix(dc.data_ptr(), hi - lo, 96, native.DTYPE_F32), k, metric,
                            index_base=lo, cand_ptr=lists[s].data_ptr(),
                            stream=torch.cuda.current_stream().cuda_stream)
        idx = torch.empty((200, k), dtype=torch.int32, device="cuda")
        sc = torch.empty((200, k), dtype=torch.float64, device="cuda")
        native.dev_merge_candidates(lists.data_ptr(), 2, 200, k, k, metric, idx.data_ptr(), sc.data_ptr(),
                                    stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        i_np = idx.cpu().numpy().view(np.uint32)
        parity.check_topk(i_np, sc.cpu().numpy(), q, c, k, metric_name, oracle)
        i1, s1 = native.topk(_hm(q), _hm(c), k, metric_name)
        assert np.array_equal(i1, i_np) and np.array_equal(s1, sc.cpu().numpy())


def test_host_shard_entry_point(native, oracle):
    """pmm_topk_shard: host shard in, exact candidates on the device; two shards merged == unsharded oracle.
    The second shard is large enough (>= 64 MB) to take the chunked/overlapped upload path."""
    import torch
    from polars_matmul_b200.sharded import shard_bounds
    rng = np.random.default_rng(31)
    q, c = _randn(rng, 150, 256), _randn(rng, 80_000, 256)
    k = 12
    bounds = [(0, 10_000), (10_000, 80_000)]                     # 10 MB (simple path) + 72 MB (chunked path)
    for metric_name, metric in (("cosine", 0), ("euclidean", 2)):
        lists = torch.zeros((2, 150, k), dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        for s_, (lo, hi) in enumerate(bounds):
            native.topk_shard(_hm(q), _hm(c[lo:hi]), k, metric, lo, lists[s_].data_ptr())
        idx = torch.empty((150, k), dtype=torch.int32, device="cuda")
        sc = torch.empty((150, k), dtype=torch.float64, device="cuda")
        native.dev_merge_candidates(lists.data_ptr(), 2, 150, k, k, metric, idx.data_ptr(), sc.data_ptr(),
                                    stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        parity.check_topk(idx.cpu().numpy().view(np.uint32), sc.cpu().numpy(), q, c, k, metric_name, oracle, exact=True)
    # a one-rank group (world size 1) end to end, host and device variants
    from polars_matmul_b200.sharded import RankGroup, unique_id
    grp = RankGroup(unique_id(), 0, 1)
    try:
        i2, s2 = grp.topk_host(q, c, 0, c.shape[0], k, "cosine")
        parity.check_topk(i2, s2, q, c, k, "cosine", oracle, exact=True)
        dq, dc = torch.from_numpy(q).cuda(), torch.from_numpy(c).cuda()
        di = torch.empty((150, k), dtype=torch.int32, device="cuda")
        ds = torch.empty((150, k), dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        grp.topk_device(dq.data_ptr(), 150, 256, 1, dc.data_ptr(), c.shape[0], 1, 0, c.shape[0], k, "euclidean",
                        di.data_ptr(), ds.data_ptr())
        parity.check_topk(di.cpu().numpy().view(np.uint32), ds.cpu().numpy(), q, c, k, "euclidean", oracle, exact=True)
    finally:
        grp.close()
    assert shard_bounds(80_000, 2) == [(0, 40_000), (40_000, 80_000)]


def _gpu_count(native):
    return native.device_count()


def test_single_process_group_over_all_gpus(native, oracle):
    """The multi-GPU driver behind the C ABI (north_star item 6): one process, one host thread per GPU, corpus rows
    sharded, NCCL all-to-all of packed candidates, merge. Needs >= 2 GPUs (skipped on the single-GPU test box; run with
    gpurun --gpus 2)."""
    if _gpu_count(native) < 2:
        pytest.skip("needs at least 2 GPUs")
    from polars_matmul_b200.sharded import LocalGroup
    rng = np.random.default_rng(50)
    q, c = _randn(rng, 333, 128), _randn(rng, 150_000, 128)          # 77 MB: every shard takes the chunked host path
    c[[10, 80_000, 149_999]] = c[3]                                   # exact ties across shards
    g = LocalGroup()
    try:
        assert g.size == _gpu_count(native)
        for metric, k in (("cosine", 10), ("dot", 100), ("euclidean", 200)):
            idx, sc = g.topk(q, c, k, metric)
            parity.check_topk(idx, sc, q, c, k, metric, oracle, exact=True)
        # fewer corpus rows than GPUs x 256: the last shards are empty, k > rows of a shard
        small = _randn(rng, 300, 128)
        idx, sc = g.topk(q[:20], small, 200, "dot")
        parity.check_topk(idx, sc, q[:20], small, 200, "dot", oracle, exact=True)
    finally:
        g.close()
    # the plugin call itself spreads over the box once the call is large enough (threshold lowered for the test)
    native.set_option("multi_gpu_min_gflop", 1)
    native.set_option("profile", 1)
    native.reset_stats()
    try:
        idx, sc = native.topk(_hm(q), _hm(c), 50, "cosine")
        assert native.get_stat("group_merge_launches") == _gpu_count(native)
        parity.check_topk(idx, sc, q, c, 50, "cosine", oracle, exact=True)
        # raw matmul: output rows sharded, no collective
        big_q = _randn(rng, 4096, 128)
        out = native.matmul(_hm(big_q), _hm(c[:20_000]))
        parity.check_matmul(out, big_q, c[:20_000], oracle.matmul(big_q, c[:20_000]), np.float32)
        out64 = native.matmul(_hm(big_q[:2048].astype(np.float64)), _hm(c[:5000].astype(np.float64)))
        parity.check_matmul(out64, big_q[:2048].astype(np.float64), c[:5000].astype(np.float64),
                            oracle.matmul(big_q[:2048].astype(np.float64), c[:5000].astype(np.float64)), np.float64)
    finally:
        native.set_option("multi_gpu_min_gflop", 4000)
        native.set_option("profile", 0)
    native.set_option("multi_gpu", 0)
    try:
        i1, s1 = native.topk(_hm(q), _hm(c), 50, "cosine")           # single GPU: identical answer
    finally:
        native.set_option("multi_gpu", 1)
    assert np.array_equal(i1, idx) and np.array_equal(s1, sc)


def test_device_norms_bit_exact(native, oracle):
    import torch
    rng = np.random.default_rng(14)
    for dt, tdt, code in ((np.float32, torch.float32, 1), (np.float64, torch.float64, 2)):
        x = _randn(rng, 777, 203, dtype=dt)
        dx = torch.from_numpy(x).cuda()
        out = torch.empty(777, dtype=tdt, device="cuda")
        for squared in (False, True):
            native.dev_norms(native.dev_matrix(dx.data_ptr(), 777, 203, code), squared, out.data_ptr(),
                             stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert np.array_equal(out.cpu().numpy(), oracle.norms(x, squared=squared))
    # f16 storage (working type f32): even dims take the paired-load path (whole 64-element steps, then the scalar tail),
    # an odd dim the element-wise one; both must reproduce ndarray's summation order on the upcast values
    for dim in (64, 200, 1024, 203):
        h = _randn(rng, 515, dim).astype(np.float16)
        dh = torch.from_numpy(h).cuda()
        out = torch.empty(515, dtype=torch.float32, device="cuda")
        for squared in (False, True):
            native.dev_norms(native.dev_matrix(dh.data_ptr(), 515, dim, 0), squared, out.data_ptr(),
                             stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert np.array_equal(out.cpu().numpy(), oracle.norms(h.astype(np.float32), squared=squared)), dim


def _full_size_device_case(native, oracle, Q, N, D, k, metrics, seed):
    """Full-size configuration with device-resident Gaussian inputs: size-independent properties on the whole result plus
    bit-identity with the oracle for 16 sampled queries against the FULL corpus."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    dq = torch.randn((Q, D), generator=g, device="cuda", dtype=torch.float32)
    dc = torch.randn((N, D), generator=g, device="cuda", dtype=torch.float32)
    idx = torch.empty((Q, k), dtype=torch.int32, device="cuda")
    sc = torch.empty((Q, k), dtype=torch.float64, device="cuda")
    ch = dc.cpu().numpy()
    sample = np.arange(0, Q, Q // 16)[:16]
    ts = torch.from_numpy(sample).cuda()
    qs = dq[ts].cpu().numpy()
    for metric_name, metric in metrics:
        native.dev_topk(native.dev_matrix(dq.data_ptr(), Q, D, 1), native.dev_matrix(dc.data_ptr(), N, D, 1), k, metric,
                        index_ptr=idx.data_ptr(), score_ptr=sc.data_ptr(),
                        stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        i64 = idx.long() & 0xFFFFFFFF
        # sorted best-first, indices in range and unique per row
        d = sc[:, 1:] - sc[:, :-1]
        assert bool((d >= 0).all()) if metric == 2 else bool((d <= 0).all())
        assert int(i64.max()) < N
        srt, _ = torch.sort(i64, dim=1)
        assert bool((srt[:, 1:] > srt[:, :-1]).all())
        # scores recomputed in f64 from the returned indices (a checksum that does not need the oracle)
        rows = torch.arange(0, Q, 997, device="cuda")
        cq = dq[rows].double()
        cc = dc[i64[rows]].double()                              # [r, k, D]
        dot = torch.einsum("rd,rkd->rk", cq, cc)
        if metric == 1:
            ref = dot
        elif metric == 0:
            ref = dot / (cq.norm(dim=1)[:, None] * cc.norm(dim=2))
        else:
            ref = torch.sqrt(torch.clamp((cq * cq).sum(1)[:, None] + (cc * cc).sum(2) - 2 * dot, min=0))
        assert torch.allclose(sc[rows], ref, rtol=1e-5, atol=1e-6 if metric == 0 else 0)
        # the oracle on a sample of queries against the FULL corpus: bit-identical scores and indices
        parity.check_topk(idx[ts].cpu().numpy().view(np.uint32), sc[ts].cpu().numpy(), qs, ch, k, metric_name, oracle, exact=True)


def test_full_size_c3_properties(native, oracle):
    """BASELINE.json configs[2] at full size (100k x 1M x 768 f32, k=100), all three metrics."""
    _full_size_device_case(native, oracle, 100_000, 1_000_000, 768, 100, (("dot", 1), ("euclidean", 2), ("cosine", 0)), 42)


def test_full_size_c4_shard_shape(native, oracle):
    """BASELINE.json configs[3], the share of one of 8 GPUs: 100k queries x 1.25M corpus rows x 768 f32, cosine, k=100."""
    _full_size_device_case(native, oracle, 100_000, 1_250_000, 768, 100, (("cosine", 0),), 43)


def test_full_size_c5_shard_shape_f16_list(pmm, native, oracle):
    """BASELINE.json configs[4], the share of one of 8 GPUs, in its stated container: f16-stored 1024-d embeddings as
    pl.List (Arrow LargeList<halffloat> with i64 offsets, flattened on the device), 1M queries x 125k corpus rows,
    cosine, k=10 - through the host entry point (chunked, staged upload). Oracle on 16 sampled queries, exact."""
    import pyarrow as pa
    Q, N, D, k = 1_000_000, 125_000, 1024, 10
    rng = np.random.default_rng(44)

    def f16_list(rows):
        vals = np.empty((rows, D), np.float16)
        for lo in range(0, rows, 65536):
            hi = min(rows, lo + 65536)
            vals[lo:hi] = rng.standard_normal((hi - lo, D), dtype=np.float32).astype(np.float16)
        offsets = np.arange(rows + 1, dtype=np.int64) * D
        return vals, pa.LargeListArray.from_arrays(pa.array(offsets), pa.array(vals.reshape(-1)))

    qv, qa = f16_list(Q)
    cv, ca = f16_list(N)
    assert qa.type == pa.large_list(pa.float16())
    idx, sc = pmm.topk_arrays(qa, ca, k, "cosine")
    assert idx.shape == (Q, k)
    assert (np.diff(sc, axis=1) <= 0).all()
    sample = np.arange(0, Q, Q // 16)[:16]
    parity.check_topk(idx[sample], sc[sample], qv[sample].astype(np.float32), cv.astype(np.float32), k, "cosine", oracle,
                      working_dtype=np.float32, exact=True)


def test_f16_list_container_with_offsets_nulls_and_short_rows(pmm, native, oracle):
    """f16 storage COMBINED with list offsets (never covered in round 1): ragged rows, a null row, a null element."""
    import pyarrow as pa
    rng = np.random.default_rng(45)
    dense = _randn(rng, 3000, 40).astype(np.float16)
    rows = [r.tolist() for r in dense]
    rows[7] = rows[7][:11]
    rows[9] = None
    rows[12][3] = None
    dense[7, 11:] = 0
    dense[9] = 0
    dense[12, 3] = 0
    arr = pa.array(rows, type=pa.large_list(pa.float16()))
    q16 = _randn(rng, 50, 40).astype(np.float16)
    qarr = pa.array([r.tolist() for r in q16], type=pa.list_(pa.float16()))
    for metric in ("cosine", "dot", "euclidean"):
        idx, sc = pmm.topk_arrays(qarr, arr, 12, metric)
        parity.check_topk(idx, sc, q16.astype(np.float32), dense.astype(np.float32), 12, metric, oracle,
                          working_dtype=np.float32, exact=True)


def test_pageable_inputs_are_staged_by_the_library(native, oracle):
    """Pageable host buffers (what Polars / Arrow hand over) go through the library's page-locked staging ring; buffers
    that are already page-locked are copied directly. Same results either way, and with staging switched off."""
    rng = np.random.default_rng(46)
    q, c = _randn(rng, 400, 256), _randn(rng, 120_000, 256)          # 123 MB corpus: chunked host path
    c[60_000] = c[5]
    native.reset_stats()
    idx, sc = native.topk(_hm(q), _hm(c), 20, "cosine")
    staged = native.get_stat("staged_h2d_bytes")
    assert staged >= c.nbytes, staged                                  # the corpus went through the ring
    parity.check_topk(idx, sc, q, c, 20, "cosine", oracle, exact=True)
    cp = native.result_empty(c.shape, np.float32)                      # page-locked (pmm_host_alloc)
    cp[...] = c
    native.reset_stats()
    i2, s2 = native.topk(_hm(q), _hm(cp), 20, "cosine")
    assert native.get_stat("staged_h2d_bytes") < c.nbytes / 2          # only the (pageable) queries
    assert np.array_equal(i2, idx) and np.array_equal(s2, sc)
    native.set_option("stage", 0)
    try:
        native.reset_stats()
        i3, s3 = native.topk(_hm(q), _hm(c), 20, "cosine")
        assert native.get_stat("staged_h2d_bytes") == 0
    finally:
        native.set_option("stage", 1)
    assert np.array_equal(i3, idx) and np.array_equal(s3, sc)
    # small ring (4 slots of 1 MB), one copy thread: many pieces per chunk, slots recycled while the GPU works
    native.set_option("stage_slot_mb", 1)
    native.set_option("stage_threads", 1)
    try:
        i4, s4 = native.topk(_hm(q), _hm(c), 20, "cosine")
        out = native.matmul(_hm(q), _hm(c[:9000]))                    # 14 MB result into a pageable buffer
    finally:
        native.set_option("stage_slot_mb", 32)
        native.set_option("stage_threads", 0)
    assert np.array_equal(i4, idx) and np.array_equal(s4, sc)
    parity.check_matmul(out, q, c[:9000], oracle.matmul(q, c[:9000]), np.float32)
    # pageable result buffers through the C ABI directly (a binding that allocates a plain Vec)
    import ctypes
    oi = np.empty((400, 20), np.uint32)
    os_ = np.empty((400, 20), np.float64)
    ka = ctypes.c_int64(0)
    qs, cs = _hm(q).c_struct(), _hm(c).c_struct()
    native.check(native.lib().pmm_topk(ctypes.byref(qs), ctypes.byref(cs), 20, b"cosine", oi.ctypes.data, os_.ctypes.data, ctypes.byref(ka)))
    assert np.array_equal(oi, idx) and np.array_equal(os_, sc)


def test_two_threads_call_concurrently(native, oracle):
    """The reference releases the GIL and is re-entrant (src/lib.rs:25,45; tests/test_polars_matmul.py:551-572 runs
    queries from a thread pool). Two threads, different corpora and per-thread options, interleaved calls: every
    result exact."""
    import threading
    rng = np.random.default_rng(47)
    jobs = [(_randn(rng, 300, 96), _randn(rng, 40_000, 96), 10, "cosine", 3),
            (_randn(rng, 200, 128), _randn(rng, 30_000, 128), 50, "euclidean", 2)]
    errors = []

    def work(q, c, k, metric, levels):
        try:
            native.set_thread_option("tc_levels", levels)           # per-thread: the other thread keeps its own
            for _ in range(6):
                idx, sc = native.topk(_hm(q), _hm(c), k, metric)
                parity.check_topk(idx, sc, q, c, k, metric, oracle, exact=True)
            native.set_thread_option(None)
        except Exception as e:  # pragma: no cover
            errors.append(repr(e))

    ts = [threading.Thread(target=work, args=j) for j in jobs]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors


def test_pipelined_rounds_overlap_rescoring(native, oracle):
    """Large query batches run one filter launch per round of query tiles while the merge + exact re-scoring of the
    previous round run on a second stream. Forced onto small data here (two scheduling units, no size threshold): 600
    queries -> two 512-row parts... same answers as the single-launch path, flagged queries still escalate."""
    import torch
    rng = np.random.default_rng(49)
    q, c = _randn(rng, 1700, 96), _randn(rng, 30_000, 96)
    c[:40] = c[40:80] * np.float32(1.00001)                    # near ties: some queries need the retry levels
    q[5] = c[41]
    dq, dc = torch.from_numpy(q).cuda(), torch.from_numpy(c).cuda()
    res = {}
    for pipe in (1, 0):
        native.set_option("pipeline", pipe)
        native.set_option("pipeline_min_gflop", 0)
        native.set_option("tc_max_units", 2)                   # a round = 2 units x 256 query rows
        native.set_option("profile", 1)
        native.reset_stats()
        try:
            for metric_name, metric, k in (("dot", 1, 100), ("cosine", 0, 10), ("euclidean", 2, 30)):
                idx = torch.empty((1700, k), dtype=torch.int32, device="cuda")
                sc = torch.empty((1700, k), dtype=torch.float64, device="cuda")
                native.dev_topk(native.dev_matrix(dq.data_ptr(), 1700, 96, 1), native.dev_matrix(dc.data_ptr(), 30_000, 96, 1), k, metric,
                                index_ptr=idx.data_ptr(), score_ptr=sc.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                res[(pipe, metric_name)] = (idx.cpu().numpy().view(np.uint32), sc.cpu().numpy())
            n_filter = native.get_stat("tc_topk_f16r_launches")
            assert n_filter == (3 * 4 if pipe else 3), n_filter     # 1700 rows = 3 rounds of 512 + a remainder
            assert native.get_stat("rescore_launches") >= n_filter
        finally:
            native.set_option("pipeline", 0)
            native.set_option("pipeline_min_gflop", 2000)
            native.set_option("tc_max_units", 0)
            native.set_option("profile", 0)
    for metric_name, k in (("dot", 100), ("cosine", 10), ("euclidean", 30)):
        i1, s1 = res[(1, metric_name)]
        i0, s0 = res[(0, metric_name)]
        assert np.array_equal(i1, i0) and np.array_equal(s1, s0)
        parity.check_topk(i1, s1, q, c, k, metric_name, oracle, exact=True)
    # f64 working precision through the same pipeline
    native.set_option("pipeline", 1)
    native.set_option("pipeline_min_gflop", 0)
    native.set_option("tc_max_units", 2)
    try:
        q64, c64 = q[:1100].astype(np.float64), c[:8000].astype(np.float64)
        idx, sc = native.topk(_hm(q64), _hm(c64), 10, "cosine")
    finally:
        native.set_option("pipeline", 0)
        native.set_option("pipeline_min_gflop", 2000)
        native.set_option("tc_max_units", 0)
    parity.check_topk(idx, sc, q64, c64, 10, "cosine", oracle, exact=True)


def test_rescore_cp_async_variant(native, oracle):
    """The cp.async gather of the re-scoring kernel (option rescore_fixed, off by default): identical results, on row lengths
    with and without a partial last step, skipped candidates and every list capacity."""
    rng = np.random.default_rng(52)
    native.set_option("rescore_fixed", 1)
    try:
        for nq, n, d, k, metric in ((300, 9000, 768, 100, "dot"), (77, 5000, 100, 10, "cosine"), (130, 7000, 36, 50, "euclidean"),
                                    (64, 3000, 4, 200, "dot"), (9, 400, 1028, 7, "cosine")):
            q, c = _randn(rng, nq, d), _randn(rng, n, d)
            idx, sc = native.topk(_hm(q), _hm(c), k, metric)
            parity.check_topk(idx, sc, q, c, k, metric, oracle, exact=True)
    finally:
        native.set_option("rescore_fixed", 0)


def test_seeded_requery_levels(native, oracle):
    """Queries the first level cannot prove are re-run from SEEDED thresholds (the exact k-th score at hand minus the
    next level's error bound): same answers with and without seeding, and the seeded launch is the one that ran."""
    rng = np.random.default_rng(48)
    d, n, k = 64, 50_000, 10
    c = _randn(rng, n, d)
    base = _randn(rng, 1, d)
    c[:60] = base * (1.0 + 1e-5 * np.arange(60, dtype=np.float32)[:, None]) * 3.0     # near ties at the top
    c[1000:1400] = base * (1.0 + 2e-6 * np.arange(400, dtype=np.float32)[:, None]) * 2.9
    q = np.concatenate([np.repeat(base, 8, axis=0) + 1e-3 * _randn(rng, 8, d), _randn(rng, 120, d)])
    res = {}
    for seeded in (1, 0):
        native.set_option("seed_retry", seeded)
        native.set_option("profile", 1)
        native.reset_stats()
        try:
            for metric in ("dot", "cosine", "euclidean"):
                res[(seeded, metric)] = native.topk(_hm(q), _hm(c), k, metric)
            assert native.get_stat("requeried_f16_wide") > 0
            if seeded:
                assert native.get_stat("tc_topk_f16r_seeded_launches") >= 1 and native.get_stat("tc_topk_f16r_kp256_launches") == 0
            else:
                assert native.get_stat("tc_topk_f16r_seeded_launches") == 0 and native.get_stat("tc_topk_f16r_kp256_launches") >= 1
        finally:
            native.set_option("seed_retry", 1)
            native.set_option("profile", 0)
    for metric in ("dot", "cosine", "euclidean"):
        i1, s1 = res[(1, metric)]
        i0, s0 = res[(0, metric)]
        assert np.array_equal(i1, i0) and np.array_equal(s1, s0)
        parity.check_topk(i1, s1, q, c, k, metric, oracle, exact=True)


def test_corpus_cache_in_the_plugin_call(pmm, native, oracle):
    """`_topk` keeps a corpus resident from its second sighting on: five calls on the same Arrow
    corpus upload it twice (streamed, then made resident) and afterwards only the queries cross PCIe. Results identical
    to the oracle every time; writable NumPy corpora are never cached; explicit invalidation works."""
    import pyarrow as pa
    rng = np.random.default_rng(60)
    c = _randn(rng, 30_000, 64)
    carr = pa.FixedSizeListArray.from_arrays(pa.array(c.reshape(-1)), 64)
    pmm.corpus_cache_clear()
    growth = []
    for i in range(5):
        q = _randn(rng, 100, 64)
        native.reset_stats()
        out = pmm._topk(q, carr, 5, "cosine")
        growth.append(native.get_stat("h2d_bytes"))
        idx = np.asarray(out.values.field("index")).reshape(100, 5)
        sc = np.asarray(out.values.field("score")).reshape(100, 5)
        parity.check_topk(idx, sc, q, c, 5, "cosine", oracle, exact=True)
    assert growth[0] >= c.nbytes and growth[1] >= c.nbytes
    assert all(g < c.nbytes / 10 for g in growth[2:]), growth              # only the queries from the third call on
    # pl.List layout (i64 offsets): keyed on the Arrow offsets buffer, so it is recognised from call to call as well
    larr = pa.LargeListArray.from_arrays(pa.array(np.arange(c.shape[0] + 1, dtype=np.int64) * 64), pa.array(c.reshape(-1)))
    l32 = pa.ListArray.from_arrays(pa.array(np.arange(c.shape[0] + 1, dtype=np.int32) * 64), pa.array(c.reshape(-1)))
    for arr_ in (larr, l32):
        seen = []
        for i in range(4):
            native.reset_stats()
            out = pmm._topk(q, arr_, 5, "cosine")
            seen.append(native.get_stat("h2d_bytes"))
        assert seen[0] >= c.nbytes and all(g < c.nbytes / 10 for g in seen[2:]), seen
        idx = np.asarray(out.values.field("index")).reshape(100, 5)
        sc = np.asarray(out.values.field("score")).reshape(100, 5)
        parity.check_topk(idx, sc, q, c, 5, "cosine", oracle, exact=True)
    # a different query dtype is a different entry (working precision f64): streamed again, still correct
    q64 = _randn(rng, 20, 64, dtype=np.float64)
    out = pmm._topk(q64, carr, 5, "dot")
    idx = np.asarray(out.values.field("index")).reshape(20, 5)
    sc = np.asarray(out.values.field("score")).reshape(20, 5)
    parity.check_topk(idx, sc, q64, c.astype(np.float64), 5, "dot", oracle, working_dtype=np.float64, exact=True)
    # writable NumPy corpus: mutation in place must be seen
    cw = c.copy()
    for _ in range(3):
        native.reset_stats()
        pmm._topk(q, cw, 5, "dot")
        assert native.get_stat("h2d_bytes") >= cw.nbytes
    cw[17] = q[0] * 100
    out = pmm._topk(q[:1], cw, 1, "dot")
    assert out.to_pylist()[0][0]["index"] == 17
    pmm.corpus_cache_clear()
    native.reset_stats()
    pmm._topk(q, carr, 5, "cosine")
    assert native.get_stat("h2d_bytes") >= c.nbytes                         # invalidated: streamed again
    pmm.corpus_cache_configure(enabled=False)
    try:
        for _ in range(3):
            native.reset_stats()
            pmm._topk(q, carr, 5, "cosine")
            assert native.get_stat("h2d_bytes") >= c.nbytes
    finally:
        pmm.corpus_cache_configure(enabled=True)


def test_multi_chunk_columns_upload_chunk_by_chunk(pmm, native, oracle):
    """A Series with several chunks (the reference's zero-copy path gives up at src/matmul.rs:53):
    the chunks go to the device one after the other, for queries and corpus, on the plain and on the chunked/overlapped
    host path, for top-k, matmul and the resident handle."""
    import pyarrow as pa
    rng = np.random.default_rng(61)

    def chunked(a, cuts):
        parts = [a[lo:hi] for lo, hi in zip([0] + cuts, cuts + [a.shape[0]])]
        return pa.chunked_array([pa.FixedSizeListArray.from_arrays(pa.array(p.reshape(-1)), a.shape[1]) for p in parts])

    q, c = _randn(rng, 300, 96), _randn(rng, 7000, 96)
    qa, ca = chunked(q, [1, 130]), chunked(c, [999, 1000, 4097])
    from polars_matmul_b200.arrow import to_host_matrix
    assert to_host_matrix(ca).chunks is not None
    for metric in ("cosine", "dot", "euclidean"):
        idx, sc = pmm.topk_arrays(qa, ca, 9, metric)
        parity.check_topk(idx, sc, q, c, 9, metric, oracle, exact=True)
    out = pmm.matmul_array(qa, ca)
    parity.check_matmul(out, q, c, oracle.matmul(q, c), np.float32)
    # large corpus: the overlapped host path cuts ITS chunks across the column's chunk boundaries
    big = _randn(rng, 200_000, 96)                                         # 77 MB
    big[[3, 70_001, 199_999]] = big[5]
    ba = chunked(big, [50_000, 70_000, 70_001, 160_000])
    idx, sc = pmm.topk_arrays(qa, ba, 20, "dot")
    parity.check_topk(idx, sc, q, big, 20, "dot", oracle, exact=True)
    pmm.corpus_cache_clear()
    for _ in range(3):                                                     # third call: resident handle built from chunks
        out = pmm._topk(qa, ba, 20, "dot")
    idx2 = np.asarray(out.values.field("index")).reshape(300, 20)
    assert np.array_equal(idx2, idx)
    pmm.corpus_cache_clear()


def test_flatten_returns_the_flat_buffer(pmm, oracle):
    import pyarrow as pa
    rng = np.random.default_rng(62)
    q, c = _randn(rng, 17, 8), _randn(rng, 23, 8)
    flat = pmm._matmul(q, c, flatten=True)
    assert flat.type == pa.float32() and len(flat) == 17 * 23
    arr = pmm._matmul(q, c)
    assert np.array_equal(np.asarray(flat), np.asarray(arr.values))         # row-major flatten == explode() of Array[T, N]
    assert len(pmm._matmul(np.empty((0, 8), np.float32), c, flatten=True)) == 0
    parity.check_matmul(np.asarray(flat).reshape(17, 23), q, c, oracle.matmul(q, c), np.float32)
