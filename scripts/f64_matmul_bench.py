"""f64 raw matmul (DMMA kernel), kernel-only, several shapes + cuBLAS (torch.matmul) on the same shapes. One JSON object."""
import json
import sys

import torch

sys.path.insert(0, ".")
from polars_matmul_b200 import _native

st = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device="cuda").manual_seed(1)
out = {}
for (Q, N, D) in ((1000, 10000, 256), (4096, 16384, 256), (8192, 8192, 1024), (2000, 100000, 64)):
    a = torch.randn((Q, D), generator=g, device="cuda", dtype=torch.float64)
    b = torch.randn((N, D), generator=g, device="cuda", dtype=torch.float64)
    o = torch.empty((Q, N), device="cuda", dtype=torch.float64)
    fn = lambda: _native.dev_matmul(_native.dev_matrix(a.data_ptr(), Q, D, 2), _native.dev_matrix(b.data_ptr(), N, D, 2), o.data_ptr(), st)
    res = {}
    for mode in (1, 3, 4, 0, 1, 3, 4, 0):
        _native.set_option("f64_dmma_async", mode)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        _native.set_option("profile", 1)
        _native.reset_stats()
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        res.setdefault(mode, []).append(round(_native.get_stat("scores_f64_dmma_ms") / 10, 4))
        _native.set_option("profile", 0)
    _native.set_option("f64_dmma_async", 1)
    ms = min(res[1] + res[3] + res[4])
    bt = b.t().contiguous()
    torch.matmul(a, bt, out=o)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        torch.matmul(a, bt, out=o)
    e1.record()
    torch.cuda.synchronize()
    cms = e0.elapsed_time(e1) / 10
    ref = a[:64] @ b.t()
    fn()
    torch.cuda.synchronize()
    err = float((o[:64] - ref).abs().max() / ref.abs().max())
    fl = 2.0 * Q * N * D
    out[f"{Q}x{N}x{D}"] = {"async_ms": res[1], "regstaged_ms": res[0], "async16x32_4cta_ms": res[3], "async16x32_3cta_ms": res[4], "dmma_ms": round(ms, 4), "dmma_TF": round(fl / ms / 1e9, 2), "cublas_ms": round(cms, 4), "cublas_TF": round(fl / cms / 1e9, 2), "rel_err": err}
    del a, b, o, bt
print(json.dumps(out, indent=1))
