"""Does a host->device copy slow down while the fused top-k kernel runs? (CUDA events on the copy stream)"""
import sys, time
import torch
sys.path.insert(0, ".")
from polars_matmul_b200 import _native
Q, N, D, k = 100000, 1000000, 768, 100
g = torch.Generator(device="cuda").manual_seed(0)
dq = torch.randn((Q, D), generator=g, device="cuda"); dc = torch.randn((N, D), generator=g, device="cuda")
idx = torch.empty((Q, k), dtype=torch.int32, device="cuda"); sc = torch.empty((Q, k), dtype=torch.float64, device="cuda")
host = torch.empty((N // 2, D)).pin_memory(); dst = torch.empty((N // 2, D), device="cuda")
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
def run_kernel():
    _native.dev_topk(_native.dev_matrix(dq.data_ptr(), Q, D, 1), _native.dev_matrix(dc.data_ptr(), N, D, 1), k, 1,
                     index_ptr=idx.data_ptr(), score_ptr=sc.data_ptr(), stream=sa.cuda_stream)
run_kernel(); torch.cuda.synchronize()
for load in (0, 1, 1, 1, 1, 1, 1, 0):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if load:
        k0.record(sa); run_kernel(); k1.record(sa)
        time.sleep(0.02)
    with torch.cuda.stream(sb):
        e0.record(sb); dst.copy_(host, non_blocking=True); e1.record(sb)
    torch.cuda.synchronize()
    print("kernel running" if load else "idle GPU      ", "copy of 1.54 GB: %.1f ms" % e0.elapsed_time(e1), ("kernel step %.1f ms" % k0.elapsed_time(k1)) if load else "", flush=True)
