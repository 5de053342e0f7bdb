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
import ctypes
oi = torch.empty((Q, k), dtype=torch.int32).pin_memory(); osc = torch.empty((Q, k), dtype=torch.float64).pin_memory()
ka = ctypes.c_int64(0)
qs_, cs_ = hq.c_struct(), hc.c_struct()
for label, ip, sp in (("numpy (pageable) outputs", None, None), ("pinned outputs", oi.data_ptr(), osc.data_ptr())):
    ts = []
    for it in range(8):
        t0 = time.perf_counter()
        if ip is None:
            idx, sc = _native.topk(hq, hc, k, "dot")
        else:
            _native.check(_native.lib().pmm_topk(ctypes.byref(qs_), ctypes.byref(cs_), k, b"dot", ip, sp, ctypes.byref(ka)))
        ts.append(round((time.perf_counter() - t0) * 1e3, 1))
    print(label, ts, flush=True)
