"""Raw f32 matmul: classic / flat / hybrid tile schedules in the same run (kernel-only, L2 flushed). One JSON object."""
import json
import sys

import torch

sys.path.insert(0, ".")
from polars_matmul_b200 import _native

st = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device="cuda").manual_seed(1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {}
for (Q, N, D) in ((16384, 65536, 32), (16384, 65536, 64), (16384, 65536, 128), (16384, 65536, 256), (12000, 65536, 128)):
    a = torch.randn((Q, D), generator=g, device="cuda")
    b = torch.randn((N, D), generator=g, device="cuda")
    o = torch.empty((Q, N), device="cuda")
    fn = lambda: _native.dev_matmul(_native.dev_matrix(a.data_ptr(), Q, D, 1), _native.dev_matrix(b.data_ptr(), N, D, 1), o.data_ptr(), st)
    row = {}
    for rep in range(2):
        for flat in (0, 2, 1):
            _native.set_option("matmul_flat", flat)
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            _native.set_option("profile", 1)
            _native.reset_stats()
            for _ in range(8):
                flush.zero_()
                fn()
            torch.cuda.synchronize()
            name = next(n for n in ("tc_matmul_f16x3", "tc_matmul_tf32x3") if _native.get_stat(n + "_ms") > 0)
            row.setdefault(f"flat={flat}", []).append(round(_native.get_stat(name + "_ms") / 8, 4))
            _native.set_option("profile", 0)
    _native.set_option("matmul_flat", -1)
    out[f"{Q}x{N}x{D}"] = row
    del a, b, o
print(json.dumps(out))
