"""
Polars-level tests: the reference's own integration tests (tests/test_polars_matmul.py, classes TestTopk,
TestMatmul, TestNumpyEquivalence, TestErrorHandling, TestFloat32Support, TestLazyFrameEdgeCases) restated
against `import polars_matmul_b200`.  They need the `polars` package, which this image does not ship, so
the whole module is skipped here; the same behaviours are covered at the Arrow level (pyarrow
FixedSizeList == pl.Array, LargeList == pl.List) by tests/test_gpu_parity.py and tests/test_host_logic.py.
"""
import numpy as np
import pytest

pl = pytest.importorskip("polars")
pytestmark = pytest.mark.gpu

import polars_matmul_b200  # noqa: E402,F401  registers the .pmm namespace


def _queries():
    return pl.DataFrame({"query_id": [0, 1], "embedding": [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]]})


def _corpus():
    return pl.DataFrame({"corpus_id": [0, 1, 2],
                         "embedding": [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]],
                         "label": ["a", "b", "c"]})


class TestTopk:  # tests/test_polars_matmul.py:10-163
    def test_basic_cosine(self):
        result = _queries().with_columns(pl.col("embedding").pmm.topk(_corpus()["embedding"], k=2).alias("matches"))
        assert len(result) == 2
        assert result["matches"].dtype == pl.List(pl.Struct({"index": pl.UInt32, "score": pl.Float64}))
        m0 = result.filter(pl.col("query_id") == 0)["matches"][0][0]
        assert m0["index"] == 0 and abs(m0["score"] - 1.0) < 1e-6
        m1 = result.filter(pl.col("query_id") == 1)["matches"][0][0]
        assert m1["index"] == 1 and abs(m1["score"] - 1.0) < 1e-6

    def test_explode_unnest_pattern(self):
        q = pl.DataFrame({"query_id": [0, 1], "embedding": [[1.0, 0.0], [0.0, 1.0]]})
        c = pl.Series("e", [[1.0, 0.0], [0.0, 1.0], [0.5, 0.5]])
        r = q.with_columns(pl.col("embedding").pmm.topk(c, k=2).alias("matches")).explode("matches").unnest("matches")
        assert len(r) == 4 and "index" in r.columns and "score" in r.columns

    def test_dot_product(self):
        q = pl.DataFrame({"embedding": [[2.0, 0.0]]})
        c = pl.Series("e", [[1.0, 0.0], [3.0, 0.0]])
        r = q.with_columns(pl.col("embedding").pmm.topk(c, k=2, metric="dot").alias("m")).explode("m").unnest("m")
        top = r.sort("score", descending=True).row(0)
        assert top[1] == 1 and abs(top[2] - 6.0) < 1e-6

    def test_euclidean(self):
        q = pl.DataFrame({"embedding": [[0.0, 0.0]]})
        c = pl.Series("e", [[3.0, 4.0], [1.0, 0.0]])
        r = q.with_columns(pl.col("embedding").pmm.topk(c, k=2, metric="euclidean").alias("m")).explode("m").unnest("m")
        top = r.sort("score").row(0)
        assert top[1] == 1 and abs(top[2] - 1.0) < 1e-6

    def test_k_larger_than_corpus(self):
        q = pl.DataFrame({"embedding": [[1.0, 0.0]]})
        c = pl.Series("e", [[1.0, 0.0], [0.0, 1.0]])
        r = q.with_columns(pl.col("embedding").pmm.topk(c, k=10).alias("m")).explode("m").unnest("m")
        assert len(r) == 2

    def test_join_with_corpus_metadata(self):
        corpus = _corpus()
        r = (pl.DataFrame({"query_id": [0], "embedding": [[1.0, 0.0, 0.0]]})
             .with_columns(pl.col("embedding").pmm.topk(corpus["embedding"], k=2).alias("m"))
             .explode("m").unnest("m").join(corpus.with_row_index("index"), on="index"))
        assert {"label", "corpus_id", "score"} <= set(r.columns)


class TestMatmul:  # tests/test_polars_matmul.py:166-258
    def test_basic(self):
        df = pl.DataFrame({"embedding": [[1.0, 2.0], [3.0, 4.0]]})
        c = pl.Series("e", [[1.0, 0.0], [0.0, 1.0]])
        r = df.select(pl.col("embedding").pmm.matmul(c).alias("scores"))
        assert r["scores"][0].to_list() == pytest.approx([1.0, 2.0])
        assert r["scores"][1].to_list() == pytest.approx([3.0, 4.0])

    def test_against_numpy(self):
        np.random.seed(42)
        a, b = np.random.randn(10, 32), np.random.randn(20, 32)
        r = pl.DataFrame({"embedding": a.tolist()}).select(pl.col("embedding").pmm.matmul(pl.Series("e", b.tolist())).alias("s"))
        for i in range(10):
            np.testing.assert_allclose(r["s"][i].to_list(), (a @ b.T)[i], rtol=1e-5)

    def test_flatten_mode(self):
        df = pl.DataFrame({"embedding": [[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]]})
        c = pl.Series("e", [[1.0, 0.0], [0.0, 1.0]])
        r = df.select(pl.col("embedding").pmm.matmul(c, flatten=True).alias("flat"))
        assert len(r) == 6 and r["flat"].dtype == pl.Float64
        np.testing.assert_allclose(r["flat"].to_list(), [1.0, 0.0, 0.0, 1.0, 1.0, 1.0], rtol=1e-5)

    def test_list_and_array_input_types(self):
        df = pl.DataFrame({"embedding": [[1.0, 2.0, 3.0, 4.0], [5.0, 6.0, 7.0, 8.0]]})
        c = pl.Series("e", [[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0]])
        r = df.select(pl.col("embedding").pmm.matmul(c).alias("s"))
        assert r["s"].dtype == pl.Array(pl.Float64, 2)
        df2 = df.with_columns(pl.col("embedding").cast(pl.Array(pl.Float64, 4)))
        r2 = df2.select(pl.col("embedding").pmm.matmul(c.cast(pl.Array(pl.Float64, 4))).alias("s"))
        assert r2["s"].dtype == pl.Array(pl.Float64, 2)
        np.testing.assert_allclose(r2["s"][1].to_list(), [5.0, 6.0], rtol=1e-5)


class TestNumpyEquivalence:  # :261-296
    def test_cosine_similarity_matches_numpy(self):
        np.random.seed(42)
        q, c = np.random.randn(5, 16), np.random.randn(20, 16)
        expected = (q / np.linalg.norm(q, axis=1, keepdims=True)) @ (c / np.linalg.norm(c, axis=1, keepdims=True)).T
        r = (pl.DataFrame({"embedding": q.tolist()}).with_row_index("qid")
             .with_columns(pl.col("embedding").pmm.topk(pl.Series("e", c.tolist()), k=20).alias("m")).explode("m").unnest("m"))
        for i in range(5):
            got = sorted(r.filter(pl.col("qid") == i)["score"].to_list(), reverse=True)
            np.testing.assert_allclose(got, sorted(expected[i].tolist(), reverse=True), rtol=1e-5)


class TestErrorHandling:  # :299-363
    def test_invalid_metric(self):
        with pytest.raises(Exception, match="Unknown metric"):
            pl.DataFrame({"embedding": [[1.0, 0.0]]}).select(pl.col("embedding").pmm.topk(pl.Series("e", [[1.0, 0.0]]), k=1, metric="invalid_metric"))

    def test_corpus_expression_raises_error(self):
        with pytest.raises(TypeError, match="corpus must be a Polars Series"):
            pl.DataFrame({"embedding": [[1.0, 0.0]]}).select(pl.col("embedding").pmm.topk(pl.col("embedding"), k=1))

    def test_empty_query(self):
        df = pl.DataFrame({"embedding": []}).cast({"embedding": pl.List(pl.Float64)})
        assert len(df.select(pl.col("embedding").pmm.topk(pl.Series("e", [[1.0, 0.0]]), k=1))) == 0

    def test_empty_corpus(self):
        with pytest.raises(Exception, match="Empty"):
            pl.DataFrame({"embedding": [[1.0, 0.0]]}).select(pl.col("embedding").pmm.topk(pl.Series("e", [], dtype=pl.List(pl.Float64)), k=1))

    def test_dimension_mismatch(self):
        df = pl.DataFrame({"embedding": [[1.0, 2.0]]})
        c = pl.Series("e", [[1.0, 2.0, 3.0]])
        with pytest.raises(Exception, match="Dimension mismatch"):
            df.select(pl.col("embedding").pmm.matmul(c))
        with pytest.raises(Exception, match="Dimension mismatch"):
            df.select(pl.col("embedding").pmm.topk(c, k=1))


class TestFloat32Support:  # :366-464
    def test_matmul_dtypes(self):
        df32 = pl.DataFrame({"embedding": [[1.0, 2.0], [3.0, 4.0]]}).with_columns(pl.col("embedding").cast(pl.List(pl.Float32)))
        c32 = pl.Series("e", [[1.0, 0.0], [0.0, 1.0]]).cast(pl.List(pl.Float32))
        assert df32.select(pl.col("embedding").pmm.matmul(c32).alias("s"))["s"].dtype == pl.Array(pl.Float32, 2)
        df64 = pl.DataFrame({"embedding": [[1.0, 2.0], [3.0, 4.0]]})
        c64 = pl.Series("e", [[1.0, 0.0], [0.0, 1.0]])
        assert df64.select(pl.col("embedding").pmm.matmul(c64).alias("s"))["s"].dtype == pl.Array(pl.Float64, 2)
        # mixed f32 query / f64 corpus -> f64 (src/matmul.rs:308)
        assert df32.select(pl.col("embedding").pmm.matmul(pl.Series("e", [[1.0, 0.0]])).alias("s"))["s"].dtype == pl.Array(pl.Float64, 1)
        # f16-stored corpus (not in the reference, whose README says "cast to f32 first"): storage is upcast exactly and
        # the result is f32 - the dtype the namespace declares must be the one that comes back (round-1 advisor finding)
        if hasattr(pl, "Float16"):
            c16 = c32.cast(pl.List(pl.Float16))
            assert df32.select(pl.col("embedding").pmm.matmul(c16).alias("s"))["s"].dtype == pl.Array(pl.Float32, 2)
            assert df32.select(pl.col("embedding").pmm.matmul(c16, flatten=True).alias("s"))["s"].dtype == pl.Float32

    def test_topk_f32(self):
        np.random.seed(42)
        df = pl.DataFrame({"query_id": [0, 1], "embedding": [np.random.randn(32).tolist(), np.random.randn(32).tolist()]}
                          ).with_columns(pl.col("embedding").cast(pl.List(pl.Float32)))
        c = pl.Series("e", [np.random.randn(32).tolist() for _ in range(10)]).cast(pl.List(pl.Float32))
        r = df.with_columns(pl.col("embedding").pmm.topk(c, k=2).alias("m")).explode("m").unnest("m")
        assert len(r) == 4 and all(-1.01 <= s <= 1.01 for s in r["score"].to_list())


class TestLazyFrameEdgeCases:  # :467-771 (structure of the plan around the expression)
    def test_lazy_topk_with_filter_select_head(self):
        q = pl.DataFrame({"qid": list(range(6)), "embedding": np.eye(6, 4).tolist()}).lazy()
        c = pl.Series("e", np.eye(5, 4).tolist())
        r = (q.filter(pl.col("qid") < 4).with_columns(pl.col("embedding").pmm.topk(c, k=2).alias("m"))
             .select("qid", "m").head(3).collect())
        assert len(r) == 3 and "m" in r.columns

    def test_lazy_two_expressions_and_group_by(self):
        q = pl.DataFrame({"g": [0, 0, 1], "embedding": [[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]]}).lazy()
        c = pl.Series("e", [[1.0, 0.0], [0.0, 1.0]])
        r = q.with_columns(pl.col("embedding").pmm.topk(c, k=1).alias("a"),
                           pl.col("embedding").pmm.matmul(c).alias("b")).collect()
        assert len(r) == 3 and {"a", "b"} <= set(r.columns)
        g = q.with_columns(pl.col("embedding").pmm.topk(c, k=1).alias("a")).group_by("g").agg(pl.len()).collect()
        assert len(g) == 2

    def test_lazy_zero_vector_corpus(self):
        q = pl.DataFrame({"embedding": [[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]]}).lazy()
        c = pl.Series("e", [[0.0, 0.0], [1.0, 0.0]]).cast(pl.Array(pl.Float64, 2))
        assert len(q.with_columns(pl.col("embedding").pmm.topk(c, k=2).alias("m")).collect()) == 3
