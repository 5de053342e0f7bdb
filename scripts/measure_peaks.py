"""Measures cuBLAS TF32 / bf16 GEMM throughput and a device copy on this GPU (denominators for the
3xTF32 roofline: MEASURED_PEAKS.json holds bf16 only)."""
import json
import torch

def gemm_tflops(dtype, allow_tf32, n=8192, iters=20):
    torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(60):
        a @ b
    e1.record(); torch.cuda.synchronize()
    sustained = e0.elapsed_time(e1) / 60
    return 2 * n ** 3 / best / 1e9, 2 * n ** 3 / sustained / 1e9

out = {}
out["tf32_tflops_burst"], out["tf32_tflops_sustained"] = gemm_tflops(torch.float32, True)
out["fp32_simt_tflops_burst"], out["fp32_simt_tflops_sustained"] = gemm_tflops(torch.float32, False, n=4096, iters=5)
out["bf16_tflops_burst"], out["bf16_tflops_sustained"] = gemm_tflops(torch.bfloat16, True)
x = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda"); y = torch.empty_like(x)
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); y.copy_(x); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
out["hbm_copy_gbs"] = 2 * x.numel() * 2 / best / 1e6
print(json.dumps(out))
