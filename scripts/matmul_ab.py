"""Raw f32 matmul, kernel-only (CUDA events recorded by the library, inputs resident, L2 flushed between iterations):
hi/lo f16 planes (default) against the 3xTF32 planes, per shape.  Prints one JSON object."""
import json
import sys

import torch

sys.path.insert(0, ".")
from polars_matmul_b200 import _native

PEAK_HBM = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
st = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device="cuda").manual_seed(1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {"hbm_peak_gbs": PEAK_HBM}
for (Q, N, D) in ((1000, 10000, 256), (1000, 10000, 64), (16384, 65536, 32), (16384, 65536, 64), (16384, 65536, 128), (16384, 65536, 256), (10000, 65536, 128)):
    a = torch.randn((Q, D), generator=g, device="cuda")
    b = torch.randn((N, D), generator=g, device="cuda")
    o = torch.empty((Q, N), device="cuda")
    fn = lambda: _native.dev_matmul(_native.dev_matrix(a.data_ptr(), Q, D, 1), _native.dev_matrix(b.data_ptr(), N, D, 1), o.data_ptr(), st)
    gb = (Q * N * 4 + (Q + N) * D * 4) / 1e9
    row = {}
    for split in (1, 0):
        _native.set_option("matmul_split16", split)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        _native.set_option("profile", 1)
        _native.reset_stats()
        for _ in range(10):
            flush.zero_()
            fn()
        torch.cuda.synchronize()
        name = next(n for n in ("tc_matmul_f16x3", "tc_matmul_tf32x3", "scores_f32") if _native.get_stat(n + "_ms") > 0)
        ms = _native.get_stat(name + "_ms") / 10
        prep = _native.get_stat("prep_ms") / 10
        _native.set_option("profile", 0)
        row[f"split16={split}"] = {"kernel": name, "kernel_ms": ms, "prep_ms": prep, "GBps": gb / ms * 1e3, "frac_hbm": gb / ms * 1e3 / PEAK_HBM, "TFLOPs": 2.0 * Q * N * D / ms / 1e9}
    _native.set_option("matmul_split16", 1)
    out[f"f32_{Q}x{N}x{D}"] = row
    del a, b, o
print(json.dumps(out, indent=1))
