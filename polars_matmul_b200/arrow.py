"""
Arrow-level marshalling: Polars Series / pyarrow arrays / NumPy matrices  <->  HostMatrix views.

Host-side mirror of `series_to_matrix[_f32]`, `try_extract_contiguous_*` and the result builders of
the reference (src/matmul.rs:22-286, :98-125, :497-518), except that nothing is copied here when the
input is already a float buffer: the C ABI reads the Arrow child buffer in place and the flattening of
List offsets / null handling happens on the device (pmm_prep.cu).

Dtype rule (src/matmul.rs:13-19, :143, :161, :179, :211): Float32 columns stay f32, Float16 columns
stay f16 (storage; upcast on the device), every other numeric dtype is cast to Float64.
"""
from __future__ import annotations

from typing import Any, Optional

import numpy as np

from ._native import HostMatrix

try:  # pyarrow is the interchange format; polars is optional in this image
    import pyarrow as pa
except Exception:  # pragma: no cover
    pa = None

try:
    import polars as pl
except Exception:  # pragma: no cover
    pl = None

_FLOATS = (np.dtype(np.float16), np.dtype(np.float32), np.dtype(np.float64))


def _child_to_numpy(child: "pa.Array") -> tuple[np.ndarray, Optional[np.ndarray], int]:
    """Child values array -> (values covering absolute positions [0, offset+len), validity bitmap
    aligned to absolute position 0, absolute offset of the child's first element)."""
    t = child.type
    if pa.types.is_float16(t) or pa.types.is_float32(t) or pa.types.is_float64(t):
        np_dt = {16: np.float16, 32: np.float32, 64: np.float64}[t.bit_width]
        bufs = child.buffers()
        n_abs = child.offset + len(child)
        if bufs[1] is None:
            vals = np.empty(0, np_dt)
        else:
            vals = np.frombuffer(bufs[1], dtype=np_dt, count=n_abs)
        validity = None
        if child.null_count and bufs[0] is not None:
            validity = np.frombuffer(bufs[0], dtype=np.uint8, count=(n_abs + 7) // 8)
        return vals, validity, child.offset
    # any other numeric dtype: cast to Float64 like the reference (nulls survive the cast)
    casted = child.cast(pa.float64())
    return _child_to_numpy(casted)


def _bitmap(arr: "pa.Array") -> Optional[np.ndarray]:
    """Row validity bitmap of `arr` re-based to bit 0, or None when there are no nulls."""
    if not arr.null_count:
        return None
    mask = np.asarray(arr.is_valid())
    return np.packbits(mask, bitorder="little")


def _chunked_fixed(arr: "pa.ChunkedArray") -> Optional[HostMatrix]:
    """A multi-chunk Array (fixed-size list) column of floats without nulls -> one HostMatrix over the chunks' buffers
    (PMM_MATRIX_CHUNKED: uploaded chunk by chunk, no host-side concatenation; the reference's zero-copy path gives up
    here, src/matmul.rs:53).  None when the column does not qualify (the caller then concatenates)."""
    t = arr.type
    if not pa.types.is_fixed_size_list(t) or arr.null_count:
        return None
    vt = t.value_type
    if not (pa.types.is_float16(vt) or pa.types.is_float32(vt) or pa.types.is_float64(vt)):
        return None
    dim = t.list_size
    views = []
    for ch in arr.chunks:
        if len(ch) == 0:
            continue
        child = ch.values.slice(ch.offset * dim, len(ch) * dim)
        if child.null_count:
            return None
        vals, _, off = _child_to_numpy(child)
        views.append(vals[off: off + len(ch) * dim])
    if len(views) < 2:
        return None
    return HostMatrix(np.empty(0, views[0].dtype), len(arr), dim, chunks=views, owner=arr)


def from_arrow(arr: Any) -> HostMatrix:
    if isinstance(arr, pa.ChunkedArray):
        if arr.num_chunks > 1:
            hm = _chunked_fixed(arr)
            if hm is not None:
                return hm
        arr = arr.combine_chunks() if arr.num_chunks != 1 else arr.chunk(0)
        if isinstance(arr, pa.ChunkedArray):   # combine_chunks of some pyarrow versions still returns a ChunkedArray
            arr = arr.chunk(0) if arr.num_chunks == 1 else pa.concat_arrays(arr.chunks)
    t = arr.type
    n_rows = len(arr)
    if pa.types.is_fixed_size_list(t):
        dim = t.list_size
        # `.values` ignores the parent's offset: slice the child explicitly
        child = arr.values.slice(arr.offset * dim, n_rows * dim)
        vals, validity, child_off = _child_to_numpy(child)
        vals = vals[child_off: child_off + n_rows * dim]
        if validity is not None:  # re-base the child bitmap to the sliced values
            bits = np.unpackbits(validity, bitorder="little")[child_off: child_off + n_rows * dim]
            validity = np.packbits(bits, bitorder="little")
        return HostMatrix(np.ascontiguousarray(vals), n_rows, dim, None, validity, _bitmap(arr), owner=arr)
    if pa.types.is_list(t) or pa.types.is_large_list(t):
        odt = np.int64 if pa.types.is_large_list(t) else np.int32
        obuf = arr.buffers()[1]
        vals, validity, child_off = _child_to_numpy(arr.values)
        offsets_id = None
        if obuf is None or n_rows == 0:
            offsets = np.zeros(1, np.int64)
        else:
            raw = np.frombuffer(obuf, dtype=odt, count=arr.offset + n_rows + 1)[arr.offset:]
            if odt is np.int64 and child_off == 0:
                offsets = raw                                   # Polars' List layout: the Arrow buffer itself, no copy
            else:                                               # 32-bit offsets or a sliced child: widened / re-based copy
                offsets = raw.astype(np.int64) + child_off      # absolute positions in the child buffer
                offsets_id = (obuf.address, arr.offset, child_off)   # identity of the SOURCE buffer (resident-corpus cache key)
        row_valid = _bitmap(arr)
        dim = 0
        if n_rows > 0:
            if row_valid is not None and not (row_valid[0] & 1):
                raise RuntimeError("First element is null")  # src/matmul.rs:238
            dim = int(offsets[1] - offsets[0])               # row 0 defines the dimension
        return HostMatrix(vals, n_rows, dim, np.ascontiguousarray(offsets), validity, row_valid, owner=arr, offsets_id=offsets_id)
    raise RuntimeError(f"expected a List or Array (fixed-size list) column of numbers, got {t}")


def from_numpy(a: np.ndarray) -> HostMatrix:
    a = np.asarray(a)
    if a.ndim != 2:
        raise RuntimeError(f"expected a 2-D matrix of embeddings, got shape {a.shape}")
    if a.dtype not in _FLOATS:
        a = a.astype(np.float64)
    a = np.ascontiguousarray(a)
    return HostMatrix(a.reshape(-1), a.shape[0], a.shape[1])


def to_host_matrix(x: Any) -> HostMatrix:
    """Polars Series | pyarrow (Chunked)Array | NumPy 2-D | list of lists -> HostMatrix."""
    if isinstance(x, HostMatrix):
        return x
    if pl is not None and isinstance(x, pl.Series):
        return from_arrow(x.to_arrow())
    if pa is not None and isinstance(x, (pa.Array, pa.ChunkedArray)):
        return from_arrow(x)
    if isinstance(x, np.ndarray):
        return from_numpy(x)
    if isinstance(x, (list, tuple)):
        if len(x) == 0:
            return HostMatrix(np.empty(0, np.float64), 0, 0)
        return from_numpy(np.asarray(x, dtype=np.float64))
    raise TypeError(f"unsupported embedding container: {type(x)!r}")


# ---------------------------------------------------------------------------------------------- results
def topk_to_arrow(index: np.ndarray, score: np.ndarray) -> "pa.Array":
    """[Q,k] index/score -> LargeList<Struct{index: u32, score: f64}> over the two flat child buffers
    (the layout the reference assembles row by row, src/matmul.rs:497-518). Zero-copy."""
    q, k = index.shape
    st = pa.StructArray.from_arrays(
        [pa.array(index.reshape(-1), type=pa.uint32()), pa.array(score.reshape(-1), type=pa.float64())],
        names=["index", "score"])
    offsets = pa.array(np.arange(q + 1, dtype=np.int64) * k, type=pa.int64())
    return pa.LargeListArray.from_arrays(offsets, st)


def matmul_flat_to_arrow(out: np.ndarray) -> "pa.Array":
    """[Q,N] -> the flat row-major column of flatten=True (python/polars_matmul/__init__.py:173-187), over the result
    buffer itself: no Array[T, N] wrapper and no explode() round trip."""
    return pa.array(out.reshape(-1))


def matmul_to_arrow(out: np.ndarray) -> "pa.Array":
    """[Q,N] -> FixedSizeList[T, N] (Array[T, N], src/matmul.rs:100-125). Zero-copy."""
    q, n = out.shape
    flat = pa.array(out.reshape(-1))
    return pa.FixedSizeListArray.from_arrays(flat, n)


def empty_topk_arrow() -> "pa.Array":
    st = pa.struct([("index", pa.uint32()), ("score", pa.float64())])
    return pa.array([], type=pa.large_list(st))


def empty_matmul_arrow(np_dtype) -> "pa.Array":
    """src/matmul.rs:297-305: an empty left side yields an empty *List*[inner] (sic), not Array."""
    return pa.array([], type=pa.large_list(pa.from_numpy_dtype(np_dtype)))
