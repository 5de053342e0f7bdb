"""torchrun check (one process per GPU): RankGroup.topk_host / topk_device (libpmm_b200 groups: NCCL all-to-all of packed
candidates) against the oracle on the whole corpus, including an EMPTY shard and shards smaller than k.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/check_sharded_gpu.py
torch.distributed only carries the 128-byte group id and the final comparison."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pmm_oracle as oracle
from polars_matmul_b200 import _native, sharded
from tests import parity

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
_native.set_device(lr)
_native.set_option("multi_gpu", 0)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
uid = [sharded.unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
grp = sharded.RankGroup(uid[0], rank, world)
rng = np.random.default_rng(5)
ok = True
for (nq, n, d, k, metric) in ((700, 400_000, 64, 100, "cosine"), (333, 90_000, 128, 20, "euclidean"), (50, 3 * world - 1, 16, 7, "dot"), (50, max(1, world - 1), 16, 7, "cosine"),
                              (64, 40_000, 96, 200, "dot")):
    q = rng.standard_normal((nq, d)).astype(np.float32)
    c = rng.standard_normal((n, d)).astype(np.float32)
    if n > 1000:
        c[n // 2 + 3] = c[7]                                  # a cross-shard exact tie
    lo, hi = sharded.shard_bounds(n, world)[rank]
    shard = np.ascontiguousarray(c[lo:hi])
    # host variant: every rank fills its query slice
    i_, s_ = grp.topk_host(q, shard, lo, n, k, metric, full=False)
    q0, q1 = grp.query_slice(nq)
    keff = min(k, n)
    full_i = torch.zeros((nq, keff), dtype=torch.int64, device="cuda")
    full_s = torch.zeros((nq, keff), dtype=torch.float64, device="cuda")
    full_i[q0:q1] = torch.from_numpy(i_[q0:q1].astype(np.int64)).cuda()
    full_s[q0:q1] = torch.from_numpy(s_[q0:q1]).cuda()
    dist.all_reduce(full_i)
    dist.all_reduce(full_s)
    # device variant: every rank receives the whole result
    dq, dc = torch.from_numpy(q).cuda(), torch.from_numpy(shard).cuda() if hi > lo else torch.empty((0, d), device="cuda")
    di = torch.empty((nq, keff), dtype=torch.int32, device="cuda")
    ds = torch.empty((nq, keff), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    grp.topk_device(dq.data_ptr(), nq, d, 1, dc.data_ptr(), hi - lo, 1, lo, n, k, metric, di.data_ptr(), ds.data_ptr(), full=True)
    if rank == 0:
        try:
            parity.check_topk(full_i.cpu().numpy().astype(np.uint32), full_s.cpu().numpy(), q, c, k, metric, oracle, exact=True)
            parity.check_topk(di.cpu().numpy().view(np.uint32), ds.cpu().numpy(), q, c, k, metric, oracle, exact=True)
            print(f"OK  {nq}x{n}x{d} k={k} {metric} over {world} ranks (shard of rank {world - 1}: {sharded.shard_bounds(n, world)[-1]})", flush=True)
        except AssertionError as e:
            ok = False
            print(f"FAIL {nq}x{n}x{d} k={k} {metric}: {e}", flush=True)
grp.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
