"""torchrun check (one process per GPU): ShardedTopk.topk_host / topk_device against the oracle on the whole corpus.
Run: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/check_sharded_gpu.py"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from polars_matmul_b200 import _native, sharded
from oracle import pmm_oracle as oracle

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); _native.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
drv = sharded.ShardedTopk()
rng = np.random.default_rng(7)
Q, N, D = 257, 6000 * world, 96
q = rng.standard_normal((Q, D)).astype(np.float32)
c = rng.standard_normal((N, D)).astype(np.float32)
c[N // 2 + 5] = c[3]; c[N - 1] = c[3]                      # exact ties across shards
lo, hi = sharded.shard_bounds(N, world)[rank]
ok = True
for chunked in (0, 1):                                      # single upload and the chunked host path (256-row chunks)
    _native.set_option("host_chunk_min_mb", 0 if chunked else 64); _native.set_option("host_chunk_min_rows", 256 if chunked else 16384)
    for metric, k in (("cosine", 10), ("dot", 100), ("euclidean", 7)):
        idx, sc = drv.topk_host(q, c[lo:hi], lo, N, k, metric)
        oi, osc = oracle.topk(q, c, k, metric)
        good = np.array_equal(idx, oi) and np.array_equal(sc, osc)
        ok &= good
        if rank == 0:
            print(f"chunked={chunked} {metric} k={k}: {'ok' if good else 'MISMATCH'}", flush=True)
flag = torch.tensor([1 if ok else 0], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SHARDED CHECK", "PASSED" if int(flag.item()) else "FAILED", "world", world, flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
