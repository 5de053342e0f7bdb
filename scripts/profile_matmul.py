"""One raw f32 matmul of the given shape (device-resident), for ncu captures: python scripts/profile_matmul.py Q N D [iters]"""
import sys

import torch

sys.path.insert(0, ".")
from polars_matmul_b200 import _native

Q, N, D = (int(x) for x in sys.argv[1:4])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
g = torch.Generator(device="cuda").manual_seed(1)
a = torch.randn((Q, D), generator=g, device="cuda")
b = torch.randn((N, D), generator=g, device="cuda")
o = torch.empty((Q, N), device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(iters):
    _native.dev_matmul(_native.dev_matrix(a.data_ptr(), Q, D, 1), _native.dev_matrix(b.data_ptr(), N, D, 1), o.data_ptr(), st)
torch.cuda.synchronize()
print("ok", float(o[0, 0]))
