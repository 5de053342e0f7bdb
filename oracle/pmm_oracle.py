"""
CPU oracle for the polars-matmul hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` leg may import
this module.  The product package (polars_matmul_b200) never imports it and has no CPU fallback.

Two layers:
  * ctypes binding to oracle/_build/libpmm_oracle.so (C restatement, oracle/pmm_oracle.c — see its
    header for the reference file:line each function follows and for the parity pin status);
  * small NumPy helpers: an f64 "truth" (every op in float64) used to adjudicate tolerance, and the
    reference's own NumPy comparator (examples/benchmark_topk.py:14-33) restated for the CPU baseline.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpmm_oracle.so")

COSINE, DOT, EUCLIDEAN = 0, 1, 2


def build(force: bool = False) -> str:
    """Compile the C oracle (gcc). Returns the .so path."""
    src = os.path.join(_HERE, "pmm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        i64, p = ctypes.c_int64, ctypes.c_void_p
        L.pmm_oracle_metric_from_str.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.c_char_p, ctypes.c_int]
        L.pmm_oracle_metric_from_str.restype = ctypes.c_int
        L.pmm_oracle_higher_is_better.argtypes = [ctypes.c_int]
        L.pmm_oracle_higher_is_better.restype = ctypes.c_int
        L.pmm_oracle_num_threads.restype = ctypes.c_int
        L.pmm_oracle_set_num_threads.argtypes = [ctypes.c_int]
        L.pmm_oracle_set_num_threads.restype = None
        for t in ("f32", "f64"):
            getattr(L, f"pmm_oracle_norms_{t}").argtypes = [p, i64, i64, p]
            getattr(L, f"pmm_oracle_sqnorms_{t}").argtypes = [p, i64, i64, p]
            getattr(L, f"pmm_oracle_scores_{t}").argtypes = [p, p, i64, i64, i64, ctypes.c_int, p]
            getattr(L, f"pmm_oracle_matmul_{t}").argtypes = [p, p, i64, i64, i64, p]
            f = getattr(L, f"pmm_oracle_topk_{t}")
            f.argtypes = [p, p, i64, i64, i64, i64, ctypes.c_int, p, p, p]
            f.restype = i64
            getattr(L, f"pmm_oracle_select_{t}").argtypes = [p, i64, i64, i64, ctypes.c_int, p, p]
            f = getattr(L, f"pmm_oracle_list_to_dense_{t}")
            f.argtypes = [p, p, p, p, i64, i64, p]
            f.restype = ctypes.c_int
        L.pmm_oracle_f16_to_f32.argtypes = [p, i64, p]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _sfx(dtype) -> str:
    return "f32" if np.dtype(dtype) == np.float32 else "f64"


def metric_from_str(s: str) -> int:
    """src/metrics.rs:19-27. Raises RuntimeError with the reference's message."""
    m = ctypes.c_int(-1)
    buf = ctypes.create_string_buffer(256)
    rc = lib().pmm_oracle_metric_from_str(s.encode(), ctypes.byref(m), buf, 256)
    if rc != 0:
        raise RuntimeError(buf.value.decode())
    return m.value


def higher_is_better(metric: int) -> bool:
    return bool(lib().pmm_oracle_higher_is_better(metric))


def num_threads() -> int:
    return int(lib().pmm_oracle_num_threads())


def set_num_threads(n: int) -> None:
    """OpenMP thread count of the oracle (torchrun exports OMP_NUM_THREADS=1; the CPU arms set their own)."""
    lib().pmm_oracle_set_num_threads(int(n))


def working_dtype(q_dtype, c_dtype):
    """src/matmul.rs:308,427: f32 iff BOTH sides f32 (f16 storage counts as f32 after upcast)."""
    qd, cd = np.dtype(q_dtype), np.dtype(c_dtype)
    f32ish = (np.dtype(np.float32), np.dtype(np.float16))
    return np.float32 if (qd in f32ish and cd in f32ish) else np.float64


def _prep(q, c):
    wd = working_dtype(q.dtype, c.dtype)
    q = np.ascontiguousarray(q, dtype=wd)
    c = np.ascontiguousarray(c, dtype=wd)
    if q.ndim != 2 or c.ndim != 2:
        raise ValueError("2-D matrices expected")
    if c.shape[0] == 0:
        raise RuntimeError("Empty series")  # src/matmul.rs:134,152
    if q.shape[1] != c.shape[1]:
        raise RuntimeError(
            f"Dimension mismatch: left has {q.shape[1]} dimensional vectors, "
            f"right has {c.shape[1]} dimensional vectors")  # src/matmul.rs:435-439
    return q, c, wd


def norms(x, squared=False):
    x = np.ascontiguousarray(x)
    out = np.empty(x.shape[0], dtype=x.dtype)
    fn = getattr(lib(), f"pmm_oracle_{'sq' if squared else ''}norms_{_sfx(x.dtype)}")
    fn(_ptr(x), x.shape[0], x.shape[1], _ptr(out))
    return out


def scores(q, c, metric: int):
    """Full similarity matrix, src/metrics.rs:258-365."""
    q, c, wd = _prep(q, c)
    out = np.empty((q.shape[0], c.shape[0]), dtype=wd)
    getattr(lib(), f"pmm_oracle_scores_{_sfx(wd)}")(_ptr(q), _ptr(c), q.shape[0], c.shape[0], q.shape[1], metric, _ptr(out))
    return out


def matmul(q, c):
    """Raw q @ c.T, src/matmul.rs:295-417. Empty left -> empty (src/matmul.rs:297-305)."""
    if q.shape[0] == 0:
        return np.empty((0, c.shape[0]), dtype=working_dtype(q.dtype, c.dtype))
    q, c, wd = _prep(q, c)
    out = np.empty((q.shape[0], c.shape[0]), dtype=wd)
    getattr(lib(), f"pmm_oracle_matmul_{_sfx(wd)}")(_ptr(q), _ptr(c), q.shape[0], c.shape[0], q.shape[1], _ptr(out))
    return out


def topk(q, c, k: int, metric="cosine", with_gap=False):
    """src/matmul.rs:473-519 + :420-469. Returns (index u32 [Q,k_eff], score f64 [Q,k_eff][, gap [Q]])."""
    if q.shape[0] == 0:  # empty queries short-circuit BEFORE metric parse (src/matmul.rs:480-490)
        e = (np.empty((0, 0), np.uint32), np.empty((0, 0), np.float64))
        return e + (np.empty(0),) if with_gap else e
    m = metric_from_str(metric) if isinstance(metric, str) else int(metric)
    q, c, wd = _prep(q, c)
    if k < 0:
        raise OverflowError("can't convert negative int to unsigned")  # PyO3 usize extraction
    keff = min(int(k), c.shape[0])
    idx = np.empty((q.shape[0], keff), np.uint32)
    sc = np.empty((q.shape[0], keff), np.float64)
    gap = np.empty(q.shape[0], np.float64)
    getattr(lib(), f"pmm_oracle_topk_{_sfx(wd)}")(
        _ptr(q), _ptr(c), q.shape[0], c.shape[0], q.shape[1], int(k), m, _ptr(idx), _ptr(sc), _ptr(gap))
    return (idx, sc, gap) if with_gap else (idx, sc)


def select(matrix, k: int, higher: bool):
    """src/topk.rs:6-75 on a given score matrix."""
    m = np.ascontiguousarray(matrix)
    keff = min(k, m.shape[1])
    idx = np.empty((m.shape[0], keff), np.int64)
    sc = np.empty((m.shape[0], keff), m.dtype)
    getattr(lib(), f"pmm_oracle_select_{_sfx(m.dtype)}")(_ptr(m), m.shape[0], m.shape[1], k, int(higher), _ptr(idx), _ptr(sc))
    return idx, sc


def list_to_dense(values, offsets, dim=None, validity=None, row_validity=None):
    """src/matmul.rs:231-286. dim defaults to the length of row 0."""
    values = np.ascontiguousarray(values)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    if n <= 0:
        raise RuntimeError("Empty series")
    if dim is None:
        dim = int(offsets[1] - offsets[0])
    if dim == 0:
        raise RuntimeError("Zero-dimensional vectors")
    out = np.empty((n, dim), dtype=values.dtype)
    rc = getattr(lib(), f"pmm_oracle_list_to_dense_{_sfx(values.dtype)}")(
        _ptr(values), _ptr(offsets), _ptr(validity), _ptr(row_validity), n, dim, _ptr(out))
    if rc == 2:
        raise RuntimeError("ragged list: a row is longer than row 0 (reference panics: ndarray index out of bounds)")
    return out


def f16_to_f32(h):
    h = np.ascontiguousarray(h).view(np.uint16)
    out = np.empty(h.shape, np.float32)
    lib().pmm_oracle_f16_to_f32(_ptr(h), h.size, _ptr(out))
    return out


# ----------------------------------------------------------------------------- f64 truth
def truth_scores(q, c, metric: int, eps: float = 0.0):
    """Every operation in float64, regardless of input dtype: adjudicates f32 tolerance.
    eps = the zero-norm guard of the working precision (1e-6 f32 / 1e-10 f64)."""
    q = np.asarray(q, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64)
    dot = q @ c.T
    if metric == DOT:
        return dot
    if metric == COSINE:
        qn = np.sqrt((q * q).sum(1))[:, None]
        cn = np.sqrt((c * c).sum(1))[None, :]
        with np.errstate(divide="ignore", invalid="ignore"):
            out = dot / (qn * cn)
        out[np.broadcast_to(qn <= eps, out.shape)] = 0.0
        out[np.broadcast_to(cn <= eps, out.shape)] = 0.0
        return out
    sq = (q * q).sum(1)[:, None] + (c * c).sum(1)[None, :] - 2.0 * dot
    return np.sqrt(np.maximum(sq, 0.0))


def score_scale(q, c, metric: int):
    """Natural magnitude of the rounding error source for each (i,j): |q_i||c_j| for dot, 1 for
    cosine, for euclidean the distance itself floored by sqrt(eps_cancel) (cancellation)."""
    q = np.asarray(q, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64)
    qn = np.sqrt((q * q).sum(1))[:, None]
    cn = np.sqrt((c * c).sum(1))[None, :]
    if metric == DOT:
        return qn * cn
    if metric == COSINE:
        return np.ones((q.shape[0], c.shape[0]))
    return np.sqrt(qn * qn + cn * cn)


# ----------------------------------------------------------------------------- NumPy comparator
def numpy_topk_cosine(query, corpus, k):
    """The reference's own comparator, restated (examples/benchmark_topk.py:14-33): normalise,
    BLAS matmul, argpartition, argsort. Used only as a CPU timing data point."""
    qn = query / np.sqrt(np.sum(query ** 2, axis=1, keepdims=True))
    cn = corpus / np.sqrt(np.sum(corpus ** 2, axis=1, keepdims=True))
    sim = qn @ cn.T
    part = np.argpartition(sim, -k, axis=1)[:, -k:]
    rows = np.arange(len(query))[:, None]
    top = sim[rows, part]
    order = np.argsort(-top, axis=1)
    return part[rows, order], top[rows, order]
