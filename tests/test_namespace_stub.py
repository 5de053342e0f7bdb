"""
`PmmNamespace` executed against a minimal Polars stand-in (tests/fake_polars): the image has no Polars wheel, so the real
Polars tests (tests/test_polars_api.py, restating tests/test_polars_matmul.py of the reference) are skipped here; this
file still runs the namespace's three `map_batches` calls (python/polars_matmul/__init__.py:109-196), the Series branch
of `_topk` / `_matmul` (src/lib.rs:15-55) and the declared-vs-returned dtype contract.  Each case runs in a fresh
interpreter with the stub on PYTHONPATH, so the rest of the suite keeps seeing "no Polars".
"""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "tests", "fake_polars")


def _run(body: str):
    try:
        import polars  # noqa: F401  (real Polars present: test_polars_api.py covers this, nothing to stub)
        pytest.skip("real Polars is installed; see tests/test_polars_api.py")
    except ImportError:
        pass
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([STUB, ROOT, os.environ.get("PYTHONPATH", "")]))
    prog = "import numpy as np, polars as pl, pyarrow as pa\nimport polars_matmul_b200 as pmm\n" + textwrap.dedent(body)
    r = subprocess.run([sys.executable, "-c", prog], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return r.stdout


def test_namespace_registers_and_declares_the_reference_dtypes():
    _run("""
        assert pmm.pl is pl and hasattr(pl.Expr, "pmm")                      # registration == the reference's decorator
        corpus32 = pl.Series("c", np.eye(4, dtype=np.float32))
        corpus64 = pl.Series("c", np.eye(4, dtype=np.float64))
        corpus16 = pl.Series("c", np.eye(4, dtype=np.float16))
        e = pl.col("emb").pmm.topk(corpus32, 2, "dot")
        assert e.is_elementwise and e.return_dtype == pl.List(pl.Struct({"index": pl.UInt32, "score": pl.Float64}))
        e = pl.col("emb").pmm.matmul(corpus32)
        assert e.is_elementwise and e.return_dtype == pl.Array(pl.Float32, 4)
        assert pl.col("emb").pmm.matmul(corpus64).return_dtype == pl.Array(pl.Float64, 4)
        assert pl.col("emb").pmm.matmul(corpus16).return_dtype == pl.Array(pl.Float32, 4)   # f16 storage computes in f32
        e = pl.col("emb").pmm.matmul(corpus32, flatten=True)
        assert not e.is_elementwise and e.return_dtype == pl.Float32()
        for call in (lambda: pl.col("emb").pmm.topk(pl.col("other"), 2), lambda: pl.col("emb").pmm.matmul(pl.col("other"))):
            try:
                call()
                raise SystemExit("an Expr corpus must be rejected")
            except TypeError as ex:
                assert "corpus must be a Polars Series" in str(ex)
    """)


@pytest.mark.gpu
def test_namespace_end_to_end_on_the_gpu(oracle):
    _run("""
        from oracle import pmm_oracle as oracle
        rng = np.random.default_rng(3)
        q, c = rng.standard_normal((50, 32)).astype(np.float32), rng.standard_normal((400, 32)).astype(np.float32)
        df = pl.DataFrame({"emb": pl.Series("emb", q)})
        corpus = pl.Series("c", c)
        for metric in ("cosine", "dot", "euclidean"):
            out = df.select(pl.col("emb").pmm.topk(corpus, 5, metric))["topk"]
            assert out.name == "topk" and out.dtype == pl.List(pl.Struct({"index": pl.UInt32, "score": pl.Float64}))
            rows = out.to_list()
            oi, osc = oracle.topk(q, c, 5, metric)
            assert [[m["index"] for m in r] for r in rows] == oi.tolist()
            assert [[m["score"] for m in r] for r in rows] == osc.tolist()
        mm = df.select(pl.col("emb").pmm.matmul(corpus))["matmul"]
        assert mm.dtype == pl.Array(pl.Float32, 400) and len(mm) == 50
        flat = df.select(pl.col("emb").pmm.matmul(corpus, flatten=True))["matmul"]
        assert flat.dtype == pl.Float32() and len(flat) == 50 * 400
        assert np.array_equal(np.asarray(flat.to_arrow()), np.asarray(mm.to_arrow().values))
        # dtype contract across storage types: what map_batches declared is what comes back (the stub raises otherwise)
        for qd, cd, want in ((np.float64, np.float64, pl.Float64), (np.float16, np.float16, pl.Float32), (np.float32, np.float16, pl.Float32)):
            d2 = pl.DataFrame({"emb": pl.Series("emb", q.astype(qd))})
            r = d2.select(pl.col("emb").pmm.matmul(pl.Series("c", c.astype(cd))))["matmul"]
            assert r.dtype == pl.Array(want, 400), (qd, cd, r.dtype)
        # list-of-floats columns (Float64, the reference's default test input, tests/test_polars_matmul.py:13-53)
        dl = pl.DataFrame({"emb": pl.Series("emb", pa.array([[1.0, 0.0], [0.0, 1.0]], type=pa.large_list(pa.float64())))})
        cl = pl.Series("c", pa.array([[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]], type=pa.large_list(pa.float64())))
        res = dl.select(pl.col("emb").pmm.topk(cl, 1, "cosine"))["topk"].to_list()
        assert [r[0]["index"] for r in res] == [0, 1] and abs(res[0][0]["score"] - 1.0) < 1e-12
        empty = pl.DataFrame({"emb": pl.Series("emb", pa.array([], type=pa.large_list(pa.float64())))})
        assert len(empty.select(pl.col("emb").pmm.topk(cl, 1, "no-such-metric"))["topk"]) == 0   # empty query short-circuits first
    """)
