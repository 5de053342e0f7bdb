"""
polars_matmul_b200 — B200-native drop-in for polars-matmul's similarity-search hot path.

Same surface as the reference's python/polars_matmul/__init__.py: importing the package registers the
`pmm` expression namespace (when Polars is installed) with

    pl.col("emb").pmm.topk(corpus, k, metric="cosine")   -> List[Struct{index: u32, score: f64}]
    pl.col("emb").pmm.matmul(corpus, flatten=False)      -> Array[f32|f64, N]  (or flat f32|f64)

The native functions `_topk` / `_matmul` keep the reference's names and argument meaning
(src/lib.rs:15-55) but call the CUDA library through its C ABI (include/pmm.h).  There is no CPU
fallback: without libpmm_b200.so or without a GPU every call raises.
"""
from __future__ import annotations

from typing import Any, Literal

import numpy as np

from . import _native
from . import arrow as _arrow

try:
    import polars as pl
except Exception:  # Polars is optional in this image; the Arrow-level API below works without it
    pl = None

__version__ = "0.1.4"
__all__ = ["PmmNamespace"]

Metric = Literal["cosine", "dot", "euclidean"]


def _is_polars_series(x: Any) -> bool:
    return pl is not None and isinstance(x, pl.Series)


def topk_arrays(left: Any, right: Any, k: int, metric: str = "cosine"):
    """Arrow/NumPy-level top-k: returns (index uint32 [Q,k_eff], score float64 [Q,k_eff]).
    Check order follows topk_impl (src/matmul.rs:473-519)."""
    q = _arrow.to_host_matrix(left)
    c = _arrow.to_host_matrix(right)
    return _native.topk(q, c, int(k), metric)


def matmul_array(left: Any, right: Any) -> np.ndarray:
    """Arrow/NumPy-level raw matmul: returns the [Q,N] matrix in the working dtype."""
    l = _arrow.to_host_matrix(left)
    r = _arrow.to_host_matrix(right)
    return _native.matmul(l, r)


def _topk(left: Any, right: Any, k: int, metric: str):
    """Mirror of `_topk` (src/lib.rs:33-55): Series in, Series `topk` of List[Struct{index,score}] out.
    With pyarrow / NumPy inputs a pyarrow LargeListArray is returned instead of a Polars Series."""
    q = _arrow.to_host_matrix(left)
    if q.n_rows == 0:  # before the metric is parsed (src/matmul.rs:480-490)
        out = _arrow.empty_topk_arrow()
    else:
        c = _arrow.to_host_matrix(right)
        idx, sc = _native.topk(q, c, int(k), metric)
        out = _arrow.topk_to_arrow(idx, sc)
    if _is_polars_series(left):
        return pl.Series("topk", out)
    return out


def _matmul(left: Any, right: Any):
    """Mirror of `_matmul` (src/lib.rs:15-30): Series in, Series `matmul` of Array[T, N] out."""
    l = _arrow.to_host_matrix(left)
    r = _arrow.to_host_matrix(right)
    if l.n_rows == 0:
        out = _arrow.empty_matmul_arrow(_native.working_dtype(l, r))
    else:
        out = _arrow.matmul_to_arrow(_native.matmul(l, r))
    if _is_polars_series(left):
        return pl.Series("matmul", out)
    return out


class PmmNamespace:
    """Polars expression namespace `pmm` (python/polars_matmul/__init__.py:39-196)."""

    def __init__(self, expr):
        self._expr = expr

    def topk(self, corpus, k: int, metric: Metric = "cosine"):
        """Top-k similar corpus rows per embedding -> List[Struct{index: u32, score: f64}]."""
        if pl is not None and isinstance(corpus, pl.Expr):
            raise TypeError(
                "corpus must be a Polars Series, not an Expression. "
                "Use corpus['column_name'] or corpus.get_column('column_name').")
        return self._expr.map_batches(
            lambda s: _topk(s, corpus, k, metric),
            is_elementwise=True,
            return_dtype=pl.List(pl.Struct({"index": pl.UInt32, "score": pl.Float64})),
        )

    def matmul(self, corpus, flatten: bool = False):
        """All pairwise dot products -> Array[f32|f64, len(corpus)] or, flattened, a flat column."""
        if pl is not None and isinstance(corpus, pl.Expr):
            raise TypeError(
                "corpus must be a Polars Series, not an Expression. "
                "Use corpus['column_name'] or corpus.get_column('column_name').")
        n_corpus = len(corpus)
        try:  # the declared dtype follows the CORPUS only (python/polars_matmul/__init__.py:166-171)
            is_f32 = corpus.dtype.inner == pl.Float32
        except Exception:
            is_f32 = False
        inner = pl.Float32 if is_f32 else pl.Float64
        if flatten:
            return self._expr.map_batches(
                lambda s: _matmul(s, corpus).explode(),
                is_elementwise=False,
                return_dtype=inner,
            )
        return self._expr.map_batches(
            lambda s: _matmul(s, corpus),
            is_elementwise=True,
            return_dtype=pl.Array(inner, n_corpus),
        )


if pl is not None:  # registration == the reference's decorator at __init__.py:39
    PmmNamespace = pl.api.register_expr_namespace("pmm")(PmmNamespace)
