"""
Multi-GPU driver for the top-k path: one process per GPU, corpus rows sharded contiguously, queries
replicated (BASELINE.json north_star item 6).

The reference has no distributed code at all; the corpus partitions naturally, so every rank runs the
fused kernel on its shard with global row numbers (index_base), the ranks exchange only Q x k packed
candidates (8 bytes each) with one NCCL all-gather over NVLink, and every rank merges the gathered lists
with the same merge kernel the single-GPU path uses between corpus pieces.  Because the packed order
(score key, then lower index) is a total order, the result does not depend on the number of shards.

torch / torch.distributed are plumbing here (device buffers, the NCCL communicator, streams); the
compute is libpmm_b200.so.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from . import _native


def shard_bounds(n_rows: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous row ranges: rank g owns [g*ceil(N/G), min(N, (g+1)*ceil(N/G)))."""
    per = -(-n_rows // world_size) if world_size > 0 else n_rows
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


def all_gather_candidates(cand, group=None):
    """[Q, k] int64 tensor per rank -> [G, Q, k] on every rank (NCCL on GPU tensors, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(cand.shape), dtype=cand.dtype, device=cand.device)
    if cand.is_cuda:
        dist.all_gather_into_tensor(out, cand.contiguous(), group=group)
    else:  # gloo has no all_gather_into_tensor for every build: use the list form
        parts = [torch.empty_like(cand) for _ in range(world)]
        dist.all_gather(parts, cand.contiguous(), group=group)
        out = torch.stack(parts, 0)
    return out


def bind_near_gpu(device_index: int):
    """One process per GPU: restrict this process to the CPUs of the GPU's NUMA node, so that the page-locked host
    buffers it allocates afterwards (and the threads that fill them) are local to the GPU's PCIe root. With 8 ranks
    uploading their corpus shards at once the host side is the bottleneck and remote-socket pages cost bandwidth.
    Best effort: returns a short description, or None when the topology cannot be read (nothing is changed then)."""
    import os
    try:
        import torch
        props = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
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


class ShardedTopk:
    """Top-k of replicated queries against a corpus sharded over the ranks of `group`."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def topk_device(self, d_queries, d_corpus_shard, index_base: int, n_total: int, k: int, metric: str):
        """d_queries [Q, D], d_corpus_shard [n_local, D]: CUDA tensors (f32 or f16) on this rank's GPU.
        Returns (index int32-as-u32 [Q, k_eff], score f64 [Q, k_eff]) CUDA tensors, identical on every rank."""
        import torch
        m = _native.metric_from_str(metric)
        Q, D = d_queries.shape
        n_local = d_corpus_shard.shape[0]
        k_eff = min(int(k), int(n_total))
        if k_eff > 128:
            raise _native.PmmError(_native.PMM_ERR_UNSUPPORTED, "sharded top-k supports k <= 128")
        code = {torch.float16: _native.DTYPE_F16, torch.float32: _native.DTYPE_F32}[d_queries.dtype]
        ccode = {torch.float16: _native.DTYPE_F16, torch.float32: _native.DTYPE_F32}[d_corpus_shard.dtype]
        stream = torch.cuda.current_stream().cuda_stream
        k_local = min(k_eff, n_local)
        cand = torch.zeros((Q, k_eff), dtype=torch.int64, device=d_queries.device)  # 0 = empty slot
        if k_local > 0:
            local = cand if k_local == k_eff else torch.empty((Q, k_local), dtype=torch.int64, device=d_queries.device)
            _native.dev_topk(_native.dev_matrix(d_queries.data_ptr(), Q, D, code),
                             _native.dev_matrix(d_corpus_shard.data_ptr(), n_local, D, ccode),
                             k_local, m, index_base=index_base, cand_ptr=local.data_ptr(), stream=stream)
            if local is not cand:
                cand[:, :k_local] = local
        gathered = all_gather_candidates(cand, self.group) if self.world > 1 else cand.unsqueeze(0)
        idx = torch.empty((Q, k_eff), dtype=torch.int32, device=d_queries.device)
        sc = torch.empty((Q, k_eff), dtype=torch.float64, device=d_queries.device)
        _native.dev_merge_candidates(gathered.data_ptr(), gathered.shape[0], Q, k_eff, k_eff, m,
                                     idx.data_ptr(), sc.data_ptr(), stream=stream)
        return idx, sc

    def topk_host(self, queries, corpus_shard, index_base: int, n_total: int, k: int, metric: str):
        """End-to-end variant: host buffers in (NumPy / Arrow / HostMatrix; pinned memory is copied without
        staging), host arrays out. The shard upload overlaps the fused kernel inside libpmm_b200
        (pmm_topk_shard); only the Q x k candidates cross NVLink."""
        import torch
        import torch.distributed as dist
        from .arrow import to_host_matrix
        m = _native.metric_from_str(metric)
        q, c = to_host_matrix(queries), to_host_matrix(corpus_shard)
        Q = q.n_rows
        k_eff = min(int(k), int(n_total))
        if k_eff > 128:
            raise _native.PmmError(_native.PMM_ERR_UNSUPPORTED, "sharded top-k supports k <= 128")
        k_local = min(k_eff, c.n_rows)
        dev = torch.device("cuda", torch.cuda.current_device())
        cand = torch.zeros((Q, k_eff), dtype=torch.int64, device=dev)
        if k_local > 0:
            local = cand if k_local == k_eff else torch.empty((Q, k_local), dtype=torch.int64, device=dev)
            q_arg = q
            if self.world > 1 and q.offsets is None and q.validity is None and q.row_validity is None and q.values.size:
                # the query batch is the same on every rank: one rank sends it through its host link, the others get
                # it over NVLink (8 ranks: 2.5 GB less through the host per C3 step)
                tq = torch.empty((Q, q.dim), dtype=torch.from_numpy(q.values[:1]).dtype, device=dev)
                src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
                if self.rank == 0:
                    tq.copy_(torch.from_numpy(q.values).view(Q, q.dim), non_blocking=True)
                dist.broadcast(tq, src=src, group=self.group)
                q_arg = _native.dev_matrix(tq.data_ptr(), Q, q.dim, q.dtype_code, flags=_native.PMM_MATRIX_ON_DEVICE)
            torch.cuda.current_stream().synchronize()          # cand is zeroed (and the queries have arrived) before the library runs
            _native.topk_shard(q_arg, c, k_local, m, index_base, local.data_ptr())
            if local is not cand:
                cand[:, :k_local] = local
        gathered = all_gather_candidates(cand, self.group) if self.world > 1 else cand.unsqueeze(0)
        idx = torch.empty((Q, k_eff), dtype=torch.int32, device=dev)
        sc = torch.empty((Q, k_eff), dtype=torch.float64, device=dev)
        _native.dev_merge_candidates(gathered.data_ptr(), gathered.shape[0], Q, k_eff, k_eff, m,
                                     idx.data_ptr(), sc.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
        # results land in pooled page-locked buffers (device->host at PCIe rate, no pageable staging)
        out_i = _native.result_empty((Q, k_eff), np.int32)
        out_s = _native.result_empty((Q, k_eff), np.float64)
        torch.from_numpy(out_i).copy_(idx)
        torch.from_numpy(out_s).copy_(sc)
        return out_i.view(np.uint32), out_s
