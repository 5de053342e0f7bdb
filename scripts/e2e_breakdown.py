"""Where the end-to-end (host buffers) C3 call spends its time: pure H2D rate, wall clock of pmm_topk with pageable
and with page-locked result buffers, per-kernel CUDA-event sums, and the resident call for comparison.
Options to try: host_chunk_first_div, host_chunk_ratio_pct, verify, f16r_wide, tc_levels (see include/pmm.h)."""
import ctypes
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
from polars_matmul_b200 import _native
from polars_matmul_b200.arrow import to_host_matrix

Q, N, D, k = 100000, 1000000, 768, 100
g = torch.Generator().manual_seed(0)
q = torch.randn((Q, D), generator=g).pin_memory(); c = torch.randn((N, D), generator=g).pin_memory()
hq, hc = to_host_matrix(q.numpy()), to_host_matrix(c.numpy())

dc = torch.empty((N, D), device="cuda")
ts = []
for _ in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter(); dc.copy_(c, non_blocking=True); torch.cuda.synchronize()
    ts.append(round((time.perf_counter() - t0) * 1e3, 1))
print("pure H2D of the corpus (3.07 GB) ms:", ts, flush=True)
del dc

def wall(fn, n=8):
    out = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); out.append(round((time.perf_counter() - t0) * 1e3, 1))
    return out

print("pmm_topk, results in pooled page-locked buffers (the shim's default):", wall(lambda: _native.topk(hq, hc, k, "dot")), flush=True)
oi = np.empty((Q, k), np.uint32); osc = np.empty((Q, k), np.float64)   # pageable
ka = ctypes.c_int64(0); qs_, cs_ = hq.c_struct(), hc.c_struct()
def pageable():
    _native.check(_native.lib().pmm_topk(ctypes.byref(qs_), ctypes.byref(cs_), k, b"dot", oi.ctypes.data, osc.ctypes.data, ctypes.byref(ka)))
print("pmm_topk, results in pageable NumPy arrays:", wall(pageable), flush=True)

names = ["prep_ms", "tc_topk_f16r_ms", "tc_topk_f16r_kp256_ms", "tc_topk_tf32x3_ms", "merge_ms", "rescore_ms", "gather_ms", "scatter_ms",
         "requeried_f16_wide", "requeried_tf32x3", "fallback_queries"]
_native.set_option("profile", 1); _native.reset_stats()
t0 = time.perf_counter(); _native.topk(hq, hc, k, "dot"); w = (time.perf_counter() - t0) * 1e3
st = {n: round(_native.get_stat(n), 2) for n in names}
_native.set_option("profile", 0)
print("one profiled call: wall ms", round(w, 1), "kernel sum", round(sum(v for n, v in st.items() if n.endswith("_ms")), 1), st, flush=True)

dq, dcc = q.cuda(), c.cuda()
idx_d = torch.empty((Q, k), dtype=torch.int32, device="cuda"); sc_d = torch.empty((Q, k), dtype=torch.float64, device="cuda")
def resident():
    _native.dev_topk(_native.dev_matrix(dq.data_ptr(), Q, D, 1), _native.dev_matrix(dcc.data_ptr(), N, D, 1), k, 1, index_ptr=idx_d.data_ptr(),
                     score_ptr=sc_d.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
print("resident (pmm_dev_topk) wall ms:", wall(resident, 4), flush=True)
