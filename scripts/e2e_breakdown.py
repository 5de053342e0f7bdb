"""Where the end-to-end (host buffers) C3 step spends its time: per-kernel CUDA-event sums vs wall clock."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from polars_matmul_b200 import _native
from polars_matmul_b200.arrow import to_host_matrix
Q, N, D, k = 100000, 1000000, 768, 100
g = torch.Generator().manual_seed(0)
q = torch.randn((Q, D), generator=g).pin_memory(); c = torch.randn((N, D), generator=g).pin_memory()
hq, hc = to_host_matrix(q.numpy()), to_host_matrix(c.numpy())
import itertools
x = torch.empty(1 << 28, dtype=torch.float32).pin_memory(); dx = torch.empty(1 << 28, dtype=torch.float32, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter(); dx.copy_(x, non_blocking=True); torch.cuda.synchronize()
print("H2D GB/s", x.numel() * 4 / (time.perf_counter() - t0) / 1e9)
dc = torch.empty((N, D), device="cuda")
def h2d_times(n):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); dc.copy_(c, non_blocking=True); torch.cuda.synchronize()
        ts.append(round((time.perf_counter() - t0) * 1e3, 1))
    return ts
print("pure H2D 3.07 GB ms:", h2d_times(12), flush=True)
for div, ratio, verify in [(16, 0, 1), (32, 0, 1), (16, 350, 1), (8, 100, 1), (16, 0, 1)]:
    _native.set_option("host_chunk_first_div", div); _native.set_option("host_chunk_ratio_pct", ratio); _native.set_option("verify", verify)
    ts = []
    for it in range(12):
        t0 = time.perf_counter()
        idx, sc = _native.topk(hq, hc, k, "dot")
        ts.append(round((time.perf_counter() - t0) * 1e3, 1))
    print("first_div", div, "ratio_pct", ratio, "verify", verify, "wall ms", ts, flush=True)
