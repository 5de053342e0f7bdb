"""Where the MMA warps of the raw-matmul kernel wait (tc_debug_skip = 8): python scripts/matmul_waits.py Q N D"""
import sys

import torch

sys.path.insert(0, ".")
from polars_matmul_b200 import _native

Q, N, D = (int(x) for x in sys.argv[1:4])
g = torch.Generator(device="cuda").manual_seed(1)
a = torch.randn((Q, D), generator=g, device="cuda")
b = torch.randn((N, D), generator=g, device="cuda")
o = torch.empty((Q, N), device="cuda")
st = torch.cuda.current_stream().cuda_stream
fn = lambda: _native.dev_matmul(_native.dev_matrix(a.data_ptr(), Q, D, 1), _native.dev_matrix(b.data_ptr(), N, D, 1), o.data_ptr(), st)
for split in (1, 0):
    _native.set_option("matmul_split16", split)
    fn(); fn()
    torch.cuda.synchronize()
    _native.set_option("tc_debug_skip", 8)
    _native.get_stat("tc_dbg_wait0")
    fn()
    torch.cuda.synchronize()
    w = [_native.get_stat(f"tc_dbg_wait{i}") for i in range(3)]
    _native.set_option("tc_debug_skip", 0)
    print(f"split16={split} D={D}: MMA warps wait for a free accumulator {w[0] / w[2]:.3f}, for operands {w[1] / w[2]:.3f} of their time")
