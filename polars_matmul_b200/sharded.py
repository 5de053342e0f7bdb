"""
Multi-GPU top-k: corpus rows sharded over the GPUs of one box, queries replicated (BASELINE.json
north_star item 6).  The reference has no distributed code at all; the corpus partitions naturally.

All of the multi-GPU work lives behind the C ABI (include/pmm.h, "groups of GPUs"): every GPU runs the fused kernel on
its shard with global row numbers, the GPUs exchange Q x k packed candidates (8 bytes each) with NCCL over NVLink — an
all-to-all, so each GPU receives and merges only its slice of the queries — and merge by u64 max.  Because the packed
order (score key, then lower index) is a total order, the result does not depend on the number of shards.

Two process models (both thin ctypes wrappers, no torch, no NCCL binding of their own):

  * `LocalGroup`  one process drives all GPUs (one host thread per GPU inside the library).  This is what the plugin
                  call uses on its own: `pmm_topk` / `pmm_matmul` spread large calls over the box.
  * `RankGroup`   one process per GPU (torchrun, MPI ...).  The host application only has to hand the 128-byte id
                  that `unique_id()` returns on one rank to every other rank.
"""
from __future__ import annotations

import ctypes
from typing import List, Tuple

import numpy as np

from . import _native

OUT_HOST_SLICE, OUT_DEVICE_FULL, OUT_DEVICE_SLICE, OUT_HOST_FULL = 1, 2, 3, 4
QUERIES_FROM_ROOT = 256
MAX_K = 248


def shard_bounds(n_rows: int, world_size: int, align: int = 1) -> List[Tuple[int, int]]:
    """Contiguous row ranges: rank g owns [g*per, min(N, (g+1)*per)), per = ceil(N/G) rounded up to `align`.
    align=1 is the library's split of the QUERY rows (who merges what); the single-process driver shards CORPUS rows
    with align=256 so that validity bitmaps slice by whole bytes."""
    per = -(-n_rows // world_size) if world_size > 0 else n_rows
    per = max(align, -(-per // align) * align)
    return [(min(n_rows, g * per), min(n_rows, (g + 1) * per)) for g in range(world_size)]


# ---- packed candidate format (host mirror of pmm_common.cuh pack_candidate / score_key) ------------
def pack_candidates(index: np.ndarray, score_f32: np.ndarray, higher_is_better: bool) -> np.ndarray:
    """(u32 index, f32 score) -> u64 candidates: (ordered key << 32) | ~index. Data-format helper."""
    s = np.asarray(score_f32, np.float32) + np.float32(0.0)            # -0.0 -> +0.0
    u = s.view(np.uint32)
    key = np.where(u & np.uint32(0x80000000), ~u, u | np.uint32(0x80000000)).astype(np.uint32)
    if not higher_is_better:
        key = ~key
    key = np.where(np.isnan(s), np.uint32(0), key).astype(np.uint64)
    return (key << np.uint64(32)) | (~np.asarray(index, np.uint32)).astype(np.uint64)


def unpack_candidates(cand: np.ndarray, higher_is_better: bool):
    """u64 candidates -> (u32 index, f64 score). Empty slots (0) decode to index 2^32-1, score NaN."""
    cand = np.asarray(cand, np.uint64)
    index = ~(cand & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    key = (cand >> np.uint64(32)).astype(np.uint32)
    u = key if higher_is_better else ~key
    bits = np.where(u & np.uint32(0x80000000), u ^ np.uint32(0x80000000), ~u).astype(np.uint32)
    score = bits.view(np.float32).astype(np.float64)
    score = np.where(key == 0, np.nan, score)
    return index, score


def unique_id() -> bytes:
    """pmm_group_unique_id: call on ONE rank and distribute the bytes to the others."""
    buf = ctypes.create_string_buffer(128)
    _native.check(_native.lib().pmm_group_unique_id(buf))
    return buf.raw


def bind_near_gpu(device_index: int):
    """One process per GPU: restrict this process to the CPUs of the GPU's NUMA node, so that the staging buffers it
    allocates afterwards (and the threads that fill them) are local to the GPU's PCIe root. Best effort: returns a
    short description, or None when the topology cannot be read (nothing is changed then)."""
    import os
    try:
        import subprocess
        bdf = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(device_index)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if not bdf:
            return None
        if len(bdf.split(":")[0]) == 8:      # 00000000:1B:00.0 -> 0000:1b:00.0
            bdf = bdf[4:]
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return f"gpu {device_index} ({bdf}) -> NUMA node {node}, {len(allowed)} cpus"
    except Exception:
        return None


class _Group:
    def __init__(self):
        self._h = ctypes.c_void_p(None)

    @property
    def size(self) -> int:
        return int(_native.lib().pmm_group_size(self._h))

    def close(self):
        if self._h:
            _native.lib().pmm_group_destroy(self._h)
            self._h = ctypes.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LocalGroup(_Group):
    """All (or the first n) GPUs of the box driven by this process."""

    def __init__(self, n_devices: int = 0):
        super().__init__()
        _native.check(_native.lib().pmm_group_init_local(int(n_devices), ctypes.byref(self._h)))

    def topk(self, queries, corpus, k: int, metric: str):
        """Host buffers in (NumPy / Arrow / HostMatrix), (index u32 [Q,k_eff], score f64 [Q,k_eff]) out."""
        from .arrow import to_host_matrix
        q, c = to_host_matrix(queries), to_host_matrix(corpus)
        if k < 0:
            raise OverflowError("can't convert negative int to unsigned")
        keff = min(int(k), c.n_rows)
        idx = _native.result_empty((q.n_rows, keff), np.uint32)
        sc = _native.result_empty((q.n_rows, keff), np.float64)
        ka = ctypes.c_int64(0)
        qs, cs = q.c_struct(), c.c_struct()
        _native.check(_native.lib().pmm_group_topk(self._h, ctypes.byref(qs), ctypes.byref(cs), int(k), str(metric).encode(),
                                                   idx.ctypes.data, sc.ctypes.data, ctypes.byref(ka)))
        return idx, sc


class RankGroup(_Group):
    """This process is rank `rank` of `world`; its GPU is the calling thread's current device (pmm_set_device)."""

    def __init__(self, uid: bytes, rank: int, world: int):
        super().__init__()
        self.rank, self.world = int(rank), int(world)
        _native.check(_native.lib().pmm_group_init_rank(uid, self.rank, self.world, ctypes.byref(self._h)))

    def query_slice(self, n_queries: int) -> Tuple[int, int]:
        """Rows of the result this rank merges (and, in the *_SLICE output modes, the only rows it fills)."""
        return shard_bounds(n_queries, self.world)[self.rank]

    def topk_shard_raw(self, q_desc, c_desc, index_base: int, n_total: int, k: int, metric: int, flags: int,
                       index_ptr: int, score_ptr: int) -> None:
        _native.check(_native.lib().pmm_group_topk_shard(self._h, ctypes.byref(q_desc), ctypes.byref(c_desc), int(index_base),
                                                         int(n_total), int(k), int(metric), int(flags), index_ptr, score_ptr))

    def topk_device(self, q_ptr: int, n_queries: int, dim: int, q_dtype: int, c_ptr: int, n_local: int, c_dtype: int,
                    index_base: int, n_total: int, k: int, metric: str, index_ptr: int, score_ptr: int,
                    full: bool = True) -> None:
        """Device-resident queries and shard; index/score device buffers [Q, k_eff] (u32 / f64).  full=True: every rank
        receives the whole result, else only this rank's query slice is written."""
        m = _native.metric_from_str(metric)
        on = _native.PMM_MATRIX_ON_DEVICE
        self.topk_shard_raw(_native.dev_matrix(q_ptr, n_queries, dim, q_dtype, flags=on),
                            _native.dev_matrix(c_ptr, n_local, dim, c_dtype, flags=on), index_base, n_total, k, m,
                            OUT_DEVICE_FULL if full else OUT_DEVICE_SLICE, index_ptr, score_ptr)

    def topk_host(self, queries, corpus_shard, index_base: int, n_total: int, k: int, metric: str, full: bool = False,
                  queries_from_root: bool = True):
        """Host buffers in, host arrays [Q, k_eff] out.  full=False (default): only this rank's query slice
        (`query_slice`) is filled — the whole-job result is the union over the ranks and crosses PCIe once;
        full=True: every rank reads the whole result back."""
        from .arrow import to_host_matrix
        m = _native.metric_from_str(metric)
        q, c = to_host_matrix(queries), to_host_matrix(corpus_shard)
        keff = min(int(k), int(n_total))
        idx = _native.result_empty((q.n_rows, keff), np.uint32)
        sc = _native.result_empty((q.n_rows, keff), np.float64)
        flags = (OUT_HOST_FULL if full else OUT_HOST_SLICE) | (QUERIES_FROM_ROOT if queries_from_root else 0)
        self.topk_shard_raw(q.c_struct(), c.c_struct(), index_base, n_total, k, m, flags, idx.ctypes.data, sc.ctypes.data)
        return idx, sc
