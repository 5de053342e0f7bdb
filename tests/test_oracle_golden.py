"""
Pins the CPU oracle against every known-answer vector the reference's own tests hold for the
pmm.topk / pmm.matmul path. CPU-only.
"""
import json
import os

import numpy as np
import pytest

NP = {"f32": np.float32, "f64": np.float64}


@pytest.fixture(scope="module")
def known(golden_dir):
    with open(os.path.join(golden_dir, "reference_known_answers.json")) as f:
        return json.load(f)


def test_select_known_answers(oracle, known):
    for case in known["select"]:
        m = np.array(case["matrix"], dtype=NP[case["dtype"]])
        idx, sc = oracle.select(m, case["k"], case["higher"])
        assert idx.tolist() == case["index"], case["src"]
        # scores come back best-first
        s = sc if case["higher"] else -sc
        assert np.all(np.diff(s, axis=1) <= 0)


def test_scores_known_answers(oracle, known):
    for case in known["scores"]:
        dt = NP[case["dtype"]]
        q, c = np.array(case["query"], dt), np.array(case["corpus"], dt)
        out = oracle.scores(q, c, oracle.metric_from_str(case["metric"]))
        assert out.dtype == dt
        for key, val in case["expect"].items():
            i, j = map(int, key.split(","))
            assert abs(out[i, j] - val) < case["atol"], (case["src"], key)


def test_list_to_dense_known_answers(oracle, known):
    for case in known["list_to_dense"]:
        dt = NP[case["dtype"]]
        m = oracle.list_to_dense(np.array(case["values"], dt), np.array(case["offsets"], np.int64))
        assert list(m.shape) == case["shape"]
        for key, val in case["expect"].items():
            i, j = map(int, key.split(","))
            assert abs(m[i, j] - val) < case["atol"]


def test_topk_known_answers(oracle, known):
    for case in known["topk"]:
        dt = NP[case["dtype"]]
        idx, sc = oracle.topk(np.array(case["query"], dt), np.array(case["corpus"], dt), case["k"], case["metric"])
        assert idx.dtype == np.uint32 and sc.dtype == np.float64
        assert idx.shape[1] == case["n_results"], case["src"]
        assert idx[:, 0].tolist() == case["top1_index"], case["src"]
        np.testing.assert_allclose(sc[:, 0], case["top1_score"], atol=case["atol"])


def test_matmul_known_answers(oracle, known):
    for case in known["matmul"]:
        dt = NP[case["dtype"]]
        out = oracle.matmul(np.array(case["left"], dt), np.array(case["right"], dt))
        assert out.dtype == dt
        np.testing.assert_allclose(out, np.array(case["expect"]), rtol=case["rtol"])
        if "flat" in case:
            np.testing.assert_allclose(out.reshape(-1), case["flat"], rtol=case["rtol"])


def test_dtype_dispatch(oracle, known):
    for a, b, want in known["dtype_dispatch"][0]["cases"]:
        assert oracle.working_dtype(NP[a], NP[b]) == NP[want]
        out = oracle.matmul(np.ones((1, 2), NP[a]), np.ones((1, 2), NP[b]))
        assert out.dtype == NP[want]


def test_errors(oracle, known):
    for case in known["errors"]:
        q = np.array(case["query"], np.float64).reshape(len(case["query"]), -1)
        c = np.array(case["corpus"], np.float64).reshape(len(case["corpus"]), -1 if case["corpus"] else q.shape[1])
        with pytest.raises(Exception, match=case["match"]):
            if case["call"] == "topk":
                oracle.topk(q, c, case["k"], case["metric"])
            else:
                oracle.matmul(q, c)


def test_empty_query_short_circuits_before_metric_parse(oracle):
    # src/matmul.rs:480-490: an invalid metric with zero queries is NOT an error
    idx, sc = oracle.topk(np.empty((0, 2)), np.ones((1, 2)), 1, "invalid_metric")
    assert idx.shape[0] == 0 and sc.shape[0] == 0


def test_metric_strings(oracle, known):
    ms = known["metric_strings"]
    for s, want in ms["ok"].items():
        assert oracle.metric_from_str(s) == want
    for s in ms["bad"]:
        with pytest.raises(RuntimeError, match="Unknown metric"):
            oracle.metric_from_str(s)
    assert oracle.higher_is_better(0) and oracle.higher_is_better(1) and not oracle.higher_is_better(2)


def test_matmul_seed42_vs_numpy(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "matmul_seed42_10x20x32_f64.npz"))
    out = oracle.matmul(g["left"], g["right"])
    np.testing.assert_allclose(out, g["expect"], rtol=float(g["rtol"]))
    # f32 working precision of the same data, judged with the cancellation-aware bound
    out32 = oracle.matmul(g["left"].astype(np.float32), g["right"].astype(np.float32))
    scale = oracle.score_scale(g["left"], g["right"], oracle.DOT)
    assert np.all(np.abs(out32 - g["expect"]) <= 1e-5 * np.maximum(np.abs(g["expect"]), 0.05 * scale))


def test_cosine_seed42_sorted_scores(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "cosine_seed42_5x20x16_f64.npz"))
    idx, sc = oracle.topk(g["query"], g["corpus"], int(g["k"]), "cosine")
    np.testing.assert_allclose(sc, g["expect_sorted_desc"], rtol=float(g["rtol"]))
    # every corpus row appears exactly once per query when k == N
    assert np.all(np.sort(idx, axis=1) == np.arange(20))


def test_bench_selfcheck_seed42(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "bench_selfcheck_seed42_100x500x64_f64.npz"))
    idx, sc = oracle.topk(g["query"], g["corpus"], int(g["k"]), "cosine")
    np.testing.assert_allclose(sc, g["expect_topk_scores"], rtol=float(g["rtol"]))
    ni, ns = oracle.numpy_topk_cosine(g["query"], g["corpus"], int(g["k"]))
    assert np.array_equal(ni.astype(np.uint32), idx)


def test_tie_rule_lower_index_first(oracle):
    # duplicated corpus rows -> exact ties; north_star rule: lower index first
    c = np.array([[1, 0], [0, 1], [1, 0], [0, 1], [1, 0]], np.float32)
    q = np.array([[1, 0]], np.float32)
    idx, sc = oracle.topk(q, c, 4, "dot")
    assert idx.tolist() == [[0, 2, 4, 1]]
    idx, sc = oracle.topk(q, c, 5, "euclidean")
    assert idx.tolist() == [[0, 2, 4, 1, 3]]


def test_zero_norm_and_k_edge_cases(oracle):
    c = np.array([[0, 0], [1, 0], [0, 2]], np.float32)
    q = np.array([[1, 1], [0, 0]], np.float32)
    idx, sc = oracle.topk(q, c, 3, "cosine")
    assert sc[1].tolist() == [0.0, 0.0, 0.0] and idx[1].tolist() == [0, 1, 2]  # zero query -> row of zeros
    assert idx[0].tolist() == [1, 2, 0] and sc[0, 2] == 0.0                    # zero corpus row scores 0
    idx, sc = oracle.topk(q, c, 0, "cosine")
    assert idx.shape == (2, 0)                                                 # k = 0 -> empty lists


def test_nan_ranks_last(oracle):
    c = np.array([[np.nan, 0], [1, 0], [2, 0]], np.float32)
    q = np.array([[1, 0]], np.float32)
    idx, sc = oracle.topk(q, c, 3, "dot")
    assert idx.tolist() == [[2, 1, 0]] and np.isnan(sc[0, 2])


def test_list_marshalling_semantics(oracle):
    vals = np.array([1, 2, 3, 4, 5], np.float64)
    off = np.array([0, 3, 5], np.int64)                       # second row shorter -> zero padded
    m = oracle.list_to_dense(vals, off)
    assert m.tolist() == [[1, 2, 3], [4, 5, 0]]
    validity = np.array([0b11101], np.uint8)                  # element 1 is null -> 0.0
    m = oracle.list_to_dense(vals, off, validity=validity)
    assert m.tolist() == [[1, 0, 3], [4, 5, 0]]
    with pytest.raises(RuntimeError, match="ragged"):
        oracle.list_to_dense(vals, np.array([0, 2, 5], np.int64))


def test_f16_upcast_exact(oracle):
    h = np.arange(0, 65536, dtype=np.uint16).view(np.float16)
    mine = oracle.f16_to_f32(h)
    ref = h.astype(np.float32)
    assert np.array_equal(mine.view(np.uint32)[~np.isnan(ref)], ref.view(np.uint32)[~np.isnan(ref)])


def test_gap_reporting(oracle):
    q = np.array([[1.0, 0.0]], np.float32)
    c = np.array([[3, 0], [2, 0], [1.5, 0]], np.float32)
    idx, sc, gap = oracle.topk(q, c, 2, "dot", with_gap=True)
    assert idx.tolist() == [[0, 1]] and gap[0] == pytest.approx(0.5)
    _, _, gap = oracle.topk(q, c, 3, "dot", with_gap=True)
    assert np.isinf(gap[0])
