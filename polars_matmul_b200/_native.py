"""
ctypes binding to libpmm_b200.so (C ABI: include/pmm.h).

This is the Python side of the boundary the reference crosses with PyO3 at src/lib.rs:15-55.
There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
ctypes releases the GIL for the duration of every foreign call, like `py.detach` in the reference
(src/lib.rs:25,45).
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpmm_b200.so")

PMM_OK, PMM_ERR_INVALID, PMM_ERR_CUDA, PMM_ERR_UNSUPPORTED = 0, 1, 2, 3
PMM_MATRIX_ON_DEVICE, PMM_MATRIX_CHUNKED = 1, 2
DTYPE_F16, DTYPE_F32, DTYPE_F64 = 0, 1, 2
METRIC_COSINE, METRIC_DOT, METRIC_EUCLIDEAN = 0, 1, 2

_NP_TO_CODE = {np.dtype(np.float16): DTYPE_F16, np.dtype(np.float32): DTYPE_F32, np.dtype(np.float64): DTYPE_F64}
_CODE_TO_NP = {v: k for k, v in _NP_TO_CODE.items()}


class PmmMatrix(ctypes.Structure):
    """struct pmm_matrix (include/pmm.h)."""
    _fields_ = [
        ("values", ctypes.c_void_p),
        ("offsets", ctypes.c_void_p),
        ("validity", ctypes.c_void_p),
        ("row_validity", ctypes.c_void_p),
        ("n_rows", ctypes.c_int64),
        ("dim", ctypes.c_int64),
        ("dtype", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


class PmmChunk(ctypes.Structure):
    """struct pmm_chunk (include/pmm.h)."""
    _fields_ = [("values", ctypes.c_void_p), ("n_rows", ctypes.c_int64)]


class PmmChunks(ctypes.Structure):
    """struct pmm_chunks (include/pmm.h)."""
    _fields_ = [("n_chunks", ctypes.c_int64), ("chunks", ctypes.POINTER(PmmChunk))]


class PmmError(RuntimeError):
    """Raised for every non-zero status; str(e) is pmm_last_error() — the same text the reference
    wraps into PyRuntimeError (src/lib.rs:28,53)."""

    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code


_lib: Optional[ctypes.CDLL] = None


def lib() -> ctypes.CDLL:
    """Loads libpmm_b200.so; raises ImportError (loudly) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m polars_matmul_b200.build` "
            "(nvcc, sm_100a). polars_matmul_b200 has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    P = ctypes.POINTER(PmmMatrix)
    i32, i64, vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p
    sigs = {
        "pmm_metric_from_str": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(i32)]),
        "pmm_higher_is_better": (ctypes.c_int, [i32]),
        "pmm_working_dtype": (ctypes.c_int, [i32, i32]),
        "pmm_topk": (ctypes.c_int, [P, P, i64, ctypes.c_char_p, vp, vp, ctypes.POINTER(i64)]),
        "pmm_matmul": (ctypes.c_int, [P, P, vp]),
        "pmm_corpus_create": (ctypes.c_int, [P, i32, ctypes.POINTER(vp)]),
        "pmm_corpus_destroy": (ctypes.c_int, [vp]),
        "pmm_corpus_rows": (i64, [vp]),
        "pmm_topk_corpus": (ctypes.c_int, [P, vp, i64, ctypes.c_char_p, vp, vp, ctypes.POINTER(i64)]),
        "pmm_dev_topk": (ctypes.c_int, [P, P, i64, i32, i64, vp, vp, vp, vp]),
        "pmm_topk_shard": (ctypes.c_int, [P, P, i64, i32, i64, vp]),
        "pmm_dev_merge_candidates": (ctypes.c_int, [vp, i64, i64, i64, i64, i32, vp, vp, vp]),
        "pmm_dev_matmul": (ctypes.c_int, [P, P, vp, vp]),
        "pmm_dev_norms": (ctypes.c_int, [P, i32, vp, vp]),
        "pmm_last_error": (ctypes.c_char_p, []),
        "pmm_version": (ctypes.c_char_p, []),
        "pmm_device_count": (ctypes.c_int, []),
        "pmm_set_device": (ctypes.c_int, [i32]),
        "pmm_kernel_launch_count": (i64, []),
        "pmm_reset_kernel_launch_count": (None, []),
        "pmm_set_option": (ctypes.c_int, [ctypes.c_char_p, i64]),
        "pmm_set_thread_option": (ctypes.c_int, [ctypes.c_char_p, i64]),
        "pmm_dev_filter_candidates": (ctypes.c_int, [P, P, i32, i32, i32, i64, vp, vp]),
        "pmm_group_unique_id": (ctypes.c_int, [ctypes.c_char_p]),
        "pmm_group_init_rank": (ctypes.c_int, [ctypes.c_char_p, i32, i32, ctypes.POINTER(vp)]),
        "pmm_group_init_local": (ctypes.c_int, [i32, ctypes.POINTER(vp)]),
        "pmm_group_destroy": (ctypes.c_int, [vp]),
        "pmm_group_size": (ctypes.c_int, [vp]),
        "pmm_group_topk": (ctypes.c_int, [vp, P, P, i64, ctypes.c_char_p, vp, vp, ctypes.POINTER(i64)]),
        "pmm_group_topk_shard": (ctypes.c_int, [vp, P, P, i64, i64, i64, i32, i32, vp, vp]),
        "pmm_filter_error_bound": (ctypes.c_int, [i32, i32, i32, i64, i32, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                                  ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]),
        "pmm_get_stat": (ctypes.c_double, [ctypes.c_char_p]),
        "pmm_reset_stats": (None, []),
        "pmm_thread_stream": (vp, []),
        "pmm_host_alloc": (ctypes.c_int, [i64, ctypes.POINTER(vp)]),
        "pmm_host_free": (ctypes.c_int, [vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)  # AttributeError if the .so does not export what include/pmm.h declares
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "pmm_metric_from_str", "pmm_higher_is_better", "pmm_working_dtype", "pmm_topk", "pmm_matmul",
    "pmm_corpus_create", "pmm_corpus_destroy", "pmm_corpus_rows", "pmm_topk_corpus", "pmm_dev_topk", "pmm_topk_shard",
    "pmm_dev_merge_candidates", "pmm_dev_matmul", "pmm_dev_norms", "pmm_last_error", "pmm_version",
    "pmm_device_count", "pmm_set_device", "pmm_kernel_launch_count", "pmm_reset_kernel_launch_count",
    "pmm_set_option", "pmm_set_thread_option", "pmm_get_stat", "pmm_reset_stats", "pmm_host_alloc", "pmm_host_free",
    "pmm_thread_stream", "pmm_dev_filter_candidates", "pmm_filter_error_bound", "pmm_group_unique_id", "pmm_group_init_rank",
    "pmm_group_init_local", "pmm_group_destroy", "pmm_group_size", "pmm_group_topk", "pmm_group_topk_shard",
]


def check(rc: int) -> None:
    if rc != PMM_OK:
        raise PmmError(rc, lib().pmm_last_error().decode("utf-8", "replace"))


# ------------------------------------------------------------------------------------------ host matrices
@dataclass
class HostMatrix:
    """An embedding column in Arrow layout held by NumPy views (see polars_matmul_b200.arrow)."""
    values: np.ndarray                       # 1-D child values, f16/f32/f64, C-contiguous
    n_rows: int
    dim: int
    offsets: Optional[np.ndarray] = None     # int64 [n_rows+1] or None (fixed-size rows)
    validity: Optional[np.ndarray] = None    # uint8 Arrow bitmap over child values
    row_validity: Optional[np.ndarray] = None
    chunks: Optional[list] = None            # multi-chunk column: list of 1-D value views (fixed-size rows, no nulls);
                                             # `values` is then an empty array that only carries the dtype
    owner: object = None                     # whatever keeps the viewed buffers alive (an Arrow array, a Series)
    offsets_id: object = None                # identity of the source offsets buffer when `offsets` is a derived copy

    @property
    def dtype_code(self) -> int:
        return _NP_TO_CODE[self.values.dtype]

    def c_struct(self) -> PmmMatrix:
        def ptr(a):
            return None if a is None else a.ctypes.data

        if self.chunks is not None:   # PMM_MATRIX_CHUNKED: values -> pmm_chunks_t; the ctypes objects live on the struct
            arr = (PmmChunk * len(self.chunks))(*[PmmChunk(c.ctypes.data if c.size else None, c.size // max(1, self.dim))
                                                  for c in self.chunks])
            desc = PmmChunks(len(self.chunks), arr)
            m = PmmMatrix(ctypes.addressof(desc), None, None, None, self.n_rows, self.dim, self.dtype_code, PMM_MATRIX_CHUNKED)
            m._keep = (arr, desc)
            return m
        return PmmMatrix(ptr(self.values) if self.values.size else None, ptr(self.offsets), ptr(self.validity),
                         ptr(self.row_validity), self.n_rows, self.dim, self.dtype_code, 0)

    def cache_key(self):
        """Identity of the underlying buffers (addresses, shape, dtype) - what the resident-corpus cache keys on."""
        def addr(a):
            return 0 if a is None else a.ctypes.data
        vals = tuple((c.ctypes.data, c.size) for c in self.chunks) if self.chunks is not None else ((addr(self.values), self.values.size),)
        offs = self.offsets_id if self.offsets_id is not None else addr(self.offsets)
        return (vals, self.n_rows, self.dim, self.dtype_code, offs, addr(self.validity), addr(self.row_validity))


def metric_from_str(name: str) -> int:
    m = ctypes.c_int32(-1)
    check(lib().pmm_metric_from_str(str(name).encode(), ctypes.byref(m)))
    return m.value


def working_dtype(left: HostMatrix, right: HostMatrix):
    return _CODE_TO_NP[lib().pmm_working_dtype(left.dtype_code, right.dtype_code)]


class _PinnedBlock:
    """One page-locked block; exposes itself to NumPy through the array interface, so arrays built on it keep it
    alive and it goes back to the pool when the last of them is collected."""
    __slots__ = ("ptr", "cap", "nbytes", "__weakref__")

    def __init__(self, ptr: int, cap: int, nbytes: int):
        self.ptr, self.cap, self.nbytes = ptr, cap, nbytes

    @property
    def __array_interface__(self):
        return {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 3}

    def __del__(self):
        try:
            _pinned_pool.give(self.ptr, self.cap)
        except Exception:  # interpreter shutdown
            pass


class _PinnedPool:
    """Result buffers in page-locked memory (pmm_host_alloc), recycled: cudaHostAlloc costs tens of ms per 100 MB,
    a pooled block nothing. Blocks of at least 1 MB; at most MAX_FREE bytes are kept while unused."""
    MIN_BYTES = 1 << 20
    MAX_BYTES = 1 << 30   # larger results (raw matmul slabs) stay pageable: pinning GBs costs seconds
    MAX_FREE = 4 << 30

    def __init__(self):
        import threading
        self._lock = threading.RLock()   # give() runs from __del__, possibly during a GC triggered under the lock
        self._free = {}      # capacity -> [ptr]
        self._free_bytes = 0

    def take(self, nbytes: int):
        cap = (nbytes + self.MIN_BYTES - 1) // self.MIN_BYTES * self.MIN_BYTES
        with self._lock:
            lst = self._free.get(cap)
            if lst:
                self._free_bytes -= cap
                return lst.pop(), cap
        p = ctypes.c_void_p(None)
        if lib().pmm_host_alloc(cap, ctypes.byref(p)) != PMM_OK:
            # page-locked memory exhausted: give back what the pool holds and let the caller use pageable memory
            self.clear()
            return None, cap
        return p.value, cap

    def give(self, ptr: int, cap: int) -> None:
        with self._lock:
            if self._free_bytes + cap <= self.MAX_FREE:
                self._free.setdefault(cap, []).append(ptr)
                self._free_bytes += cap
                return
        lib().pmm_host_free(ptr)

    def clear(self) -> None:
        with self._lock:
            blocks, self._free, self._free_bytes = self._free, {}, 0
        for lst in blocks.values():
            for ptr in lst:
                lib().pmm_host_free(ptr)


_pinned_pool = _PinnedPool()


def result_empty(shape, dtype) -> np.ndarray:
    """np.empty for result buffers: page-locked (pooled) when large, so the device->host copy runs at PCIe rate."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    if nbytes < _PinnedPool.MIN_BYTES or nbytes > _PinnedPool.MAX_BYTES:
        return np.empty(shape, dtype)
    ptr, cap = _pinned_pool.take(nbytes)
    if ptr is None:
        return np.empty(shape, dtype)
    return np.asarray(_PinnedBlock(ptr, cap, nbytes)).view(dtype).reshape(shape)


def topk(queries: HostMatrix, corpus: HostMatrix, k: int, metric: str):
    """pmm_topk. Returns (index uint32 [Q, k_eff], score float64 [Q, k_eff])."""
    if k < 0:
        raise OverflowError("can't convert negative int to unsigned")  # what PyO3 raises for usize
    keff = min(int(k), corpus.n_rows)
    idx = result_empty((queries.n_rows, keff), np.uint32)
    sc = result_empty((queries.n_rows, keff), np.float64)
    ka = ctypes.c_int64(0)
    q, c = queries.c_struct(), corpus.c_struct()
    check(lib().pmm_topk(ctypes.byref(q), ctypes.byref(c), int(k), str(metric).encode(),
                         idx.ctypes.data, sc.ctypes.data, ctypes.byref(ka)))
    assert ka.value == keff
    return idx, sc


def matmul(left: HostMatrix, right: HostMatrix) -> np.ndarray:
    """pmm_matmul. Returns [Q, N] in the working dtype."""
    out = result_empty((left.n_rows, right.n_rows), working_dtype(left, right))
    l, r = left.c_struct(), right.c_struct()
    check(lib().pmm_matmul(ctypes.byref(l), ctypes.byref(r), out.ctypes.data))
    return out


class ResidentCorpus:
    """pmm_corpus_t: the prepared corpus kept in HBM across calls."""

    def __init__(self, corpus: HostMatrix, query_dtype_code: int = DTYPE_F32):
        self._h = ctypes.c_void_p(None)
        c = corpus.c_struct()
        check(lib().pmm_corpus_create(ctypes.byref(c), query_dtype_code, ctypes.byref(self._h)))
        self.n_rows = int(lib().pmm_corpus_rows(self._h))

    def topk(self, queries: HostMatrix, k: int, metric: str):
        if k < 0:
            raise OverflowError("can't convert negative int to unsigned")
        keff = min(int(k), self.n_rows)
        idx = result_empty((queries.n_rows, keff), np.uint32)
        sc = result_empty((queries.n_rows, keff), np.float64)
        ka = ctypes.c_int64(0)
        q = queries.c_struct()
        check(lib().pmm_topk_corpus(ctypes.byref(q), self._h, int(k), str(metric).encode(),
                                    idx.ctypes.data, sc.ctypes.data, ctypes.byref(ka)))
        return idx, sc

    def close(self):
        if self._h:
            lib().pmm_corpus_destroy(self._h)
            self._h = ctypes.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------ device level

def dev_matrix(data_ptr: int, n_rows: int, dim: int, dtype_code: int, offsets_ptr: int = 0, flags: int = 0) -> PmmMatrix:
    """Describes a device-resident matrix (e.g. a torch CUDA tensor's data_ptr()). flags: PMM_MATRIX_ON_DEVICE for
    the entry points that take host descriptors by default (pmm_topk_shard's queries)."""
    return PmmMatrix(data_ptr or None, offsets_ptr or None, None, None, n_rows, dim, dtype_code, flags)


def dev_topk(dq: PmmMatrix, dc: PmmMatrix, k: int, metric: int, index_base: int = 0, index_ptr: int = 0,
             score_ptr: int = 0, cand_ptr: int = 0, stream: int = 0) -> None:
    check(lib().pmm_dev_topk(ctypes.byref(dq), ctypes.byref(dc), k, metric, index_base, index_ptr or None,
                             score_ptr or None, cand_ptr or None, stream or None))


def topk_shard(queries, corpus_shard: HostMatrix, k: int, metric: int, index_base: int, cand_ptr: int) -> None:
    """pmm_topk_shard: host buffers in (queries: a HostMatrix, or a PmmMatrix flagged PMM_MATRIX_ON_DEVICE), exact
    packed candidates left on the device at cand_ptr."""
    q = queries if isinstance(queries, PmmMatrix) else queries.c_struct()
    c = corpus_shard.c_struct()
    check(lib().pmm_topk_shard(ctypes.byref(q), ctypes.byref(c), int(k), int(metric), int(index_base), cand_ptr))


def dev_merge_candidates(lists_ptr: int, n_lists: int, n_queries: int, k_in: int, k_out: int, metric: int,
                         index_ptr: int, score_ptr: int, stream: int = 0) -> None:
    check(lib().pmm_dev_merge_candidates(lists_ptr, n_lists, n_queries, k_in, k_out, metric, index_ptr or None,
                                         score_ptr or None, stream or None))


def dev_matmul(dl: PmmMatrix, dr: PmmMatrix, out_ptr: int, stream: int = 0) -> None:
    check(lib().pmm_dev_matmul(ctypes.byref(dl), ctypes.byref(dr), out_ptr, stream or None))


def dev_norms(dx: PmmMatrix, squared: bool, out_ptr: int, stream: int = 0) -> None:
    check(lib().pmm_dev_norms(ctypes.byref(dx), int(bool(squared)), out_ptr, stream or None))


def set_option(key: str, value: int) -> None:
    check(lib().pmm_set_option(key.encode(), int(value)))


def set_thread_option(key, value: int = 0) -> None:
    """Per-thread override of an option (key=None drops the calling thread's overrides)."""
    check(lib().pmm_set_thread_option(None if key is None else key.encode(), int(value)))


def dev_filter_candidates(dq: PmmMatrix, dc: PmmMatrix, metric: int, level: int, kp: int, kept_ptr: int,
                          index_base: int = 0, stream: int = 0) -> None:
    """Diagnostics: the raw output of one tensor-core filter level (packed candidates keyed by FILTER value)."""
    check(lib().pmm_dev_filter_candidates(ctypes.byref(dq), ctypes.byref(dc), metric, level, kp, index_base, kept_ptr,
                                          stream or None))


def filter_error_bound(level: int, q_dtype: int, c_dtype: int, dim: int, metric: int, q_norm: float, c_norm_max: float,
                       c_norm_min: float):
    """(E, max_norm): the bound the losslessness proof uses for |filter value - exact value| (filter units)."""
    e, mx = ctypes.c_float(0), ctypes.c_float(0)
    check(lib().pmm_filter_error_bound(level, q_dtype, c_dtype, dim, metric, q_norm, c_norm_max, c_norm_min,
                                       ctypes.byref(e), ctypes.byref(mx)))
    return e.value, mx.value


def get_stat(name: str) -> float:
    return float(lib().pmm_get_stat(name.encode()))


def reset_stats() -> None:
    lib().pmm_reset_stats()


def kernel_launch_count() -> int:
    return int(lib().pmm_kernel_launch_count())


def reset_kernel_launch_count() -> None:
    lib().pmm_reset_kernel_launch_count()


def thread_stream() -> int:
    """cudaStream_t of the calling thread's host entry points / group calls (0 without a device)."""
    return int(lib().pmm_thread_stream() or 0)


def device_count() -> int:
    return int(lib().pmm_device_count())


def set_device(device: int) -> None:
    check(lib().pmm_set_device(device))
