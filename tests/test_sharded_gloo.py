"""
World-size-2 test of the multi-GPU host logic on CPU (gloo): shard bounds, global index offsets, the
packed candidate exchange and the shard-count independence of the merged result.
The per-shard candidates come from the CPU oracle and the merge is a NumPy sort — both are test
stand-ins for the CUDA kernels (which the -m gpu tests cover); what is under test here is the
plumbing in polars_matmul_b200/sharded.py.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, c, k, metric, ret):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from oracle import pmm_oracle as oracle
    from polars_matmul_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        higher = metric != "euclidean"
        lo, hi = sharded.shard_bounds(c.shape[0], world)[rank]
        k_eff = min(k, c.shape[0])
        cand = np.zeros((q.shape[0], k_eff), np.uint64)
        if hi > lo:
            li, ls = oracle.topk(q, c[lo:hi], k_eff, metric)
            cand[:, : li.shape[1]] = sharded.pack_candidates(li + np.uint32(lo), ls.astype(np.float32), higher)
        gathered = sharded.all_gather_candidates(torch.from_numpy(cand.view(np.int64)))
        g = gathered.numpy().view(np.uint64)                       # [G, Q, k]
        assert g.shape == (world, q.shape[0], k_eff)
        merged = np.sort(np.transpose(g, (1, 0, 2)).reshape(q.shape[0], -1), axis=1)[:, ::-1][:, :k_eff]
        idx, sc = sharded.unpack_candidates(merged, higher)
        if rank == 0:
            ret["idx"], ret["sc"] = idx, sc
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_two_rank_candidate_exchange_matches_unsharded(oracle, metric):
    import torch.multiprocessing as mp
    rng = np.random.default_rng(21)
    q = rng.standard_normal((13, 24)).astype(np.float32)
    c = rng.standard_normal((301, 24)).astype(np.float32)
    c[150] = c[10]                                                  # a cross-shard exact tie
    k = 9
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, c, k, metric, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    oi, osc = oracle.topk(q, c, k, metric)
    assert np.array_equal(ret["idx"], oi)
    assert np.array_equal(ret["sc"], osc)
