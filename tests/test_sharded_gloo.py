"""
World-size-2 test of the multi-GPU host logic on CPU (gloo): shard bounds, global index offsets, the
packed candidate exchange and the shard-count independence of the merged result.
The per-shard candidates come from the CPU oracle, the exchange is a gloo all-to-all and the merge a NumPy
sort — test stand-ins for the CUDA kernels and the NCCL exchange inside libpmm_b200 (which the -m gpu
tests cover on 2+ GPU boxes); what is under test here is the host-side logic the library and
polars_matmul_b200/sharded.py share: shard bounds (including EMPTY shards when there are more ranks than
rows), global index offsets, the packed candidate format, the query-slice all-to-all layout
([source rank][query of my slice][k]) and the shard-count independence of the merged result.
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
        nq = q.shape[0]
        cand = np.zeros((nq, k_eff), np.uint64)                     # 0 = empty slot (shard smaller than k, or empty)
        if hi > lo:
            li, ls = oracle.topk(q, c[lo:hi], k_eff, metric)
            cand[:, : li.shape[1]] = sharded.pack_candidates(li + np.uint32(lo), ls.astype(np.float32), higher)
        # all-to-all: rank r merges the queries of ITS slice and receives, from every rank, only those rows
        qb = sharded.shard_bounds(nq, world)
        send = [torch.from_numpy(cand[a:b].view(np.int64).copy()) for a, b in qb]
        q0, q1 = qb[rank]
        recv = [torch.empty((q1 - q0, k_eff), dtype=torch.int64) for _ in range(world)]
        dist.all_to_all(recv, send) if dist.get_backend() != "gloo" else _gloo_all_to_all(dist, recv, send, rank, world)
        lists = np.stack([r.numpy().view(np.uint64) for r in recv], 0)             # [G, qn, k]
        merged = np.sort(np.transpose(lists, (1, 0, 2)).reshape(q1 - q0, world * k_eff), axis=1)[:, ::-1][:, :k_eff]
        idx, sc = sharded.unpack_candidates(merged, higher)
        ret[rank] = (q0, q1, idx, sc)
    finally:
        dist.destroy_process_group()


def _gloo_all_to_all(dist, recv, send, rank, world):
    """gloo has no all_to_all: pairwise isend/irecv."""
    reqs = []
    for r in range(world):
        if r == rank:
            recv[r].copy_(send[r])
        else:
            reqs.append(dist.isend(send[r], r))
            reqs.append(dist.irecv(recv[r], r))
    for rq in reqs:
        rq.wait()


def _run(world, q, c, k, metric):
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, c, k, metric, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    k_eff = min(k, c.shape[0])
    idx = np.zeros((q.shape[0], k_eff), np.uint32)
    sc = np.zeros((q.shape[0], k_eff), np.float64)
    for r in range(world):
        q0, q1, i, s = ret[r]
        idx[q0:q1], sc[q0:q1] = i, s
    return idx, sc


@pytest.mark.parametrize("metric", ["cosine", "euclidean"])
def test_two_rank_candidate_exchange_matches_unsharded(oracle, metric):
    rng = np.random.default_rng(21)
    q = rng.standard_normal((13, 24)).astype(np.float32)
    c = rng.standard_normal((301, 24)).astype(np.float32)
    c[150] = c[10]                                                  # a cross-shard exact tie
    k = 9
    idx, sc = _run(2, q, c, k, metric)
    oi, osc = oracle.topk(q, c, k, metric)
    assert np.array_equal(idx, oi)
    assert np.array_equal(sc, osc)


def test_more_ranks_than_rows_leaves_empty_shards(oracle):
    """3 corpus rows over 4 ranks: shard_bounds gives rank 3 an EMPTY shard (and the queries, 2 over 4 ranks, leave two
    ranks without a query slice). Every rank must still take part in the exchange (round-1 advisor finding: a collective
    skipped on the empty rank hangs the others)."""
    from polars_matmul_b200 import sharded
    assert sharded.shard_bounds(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    assert sharded.shard_bounds(41, 8)[7] == (41, 41)
    assert sharded.shard_bounds(1000, 3, align=256) == [(0, 512), (512, 1000), (1000, 1000)]
    rng = np.random.default_rng(22)
    q = rng.standard_normal((2, 8)).astype(np.float32)
    c = rng.standard_normal((3, 8)).astype(np.float32)
    idx, sc = _run(4, q, c, 5, "dot")                               # k clamps to 3
    oi, osc = oracle.topk(q, c, 5, "dot")
    assert idx.shape == (2, 3)
    assert np.array_equal(idx, oi) and np.array_equal(sc, osc)
