"""C3 filter launch with tc_debug_skip = 8: where the MMA warps wait, by tile position inside an item (bucket b = tiles 2^b .. 2^(b+1)-1)."""
import sys

import torch

sys.path.insert(0, ".")
from polars_matmul_b200 import _native

Q, N, D, k = 100_000, 1_000_000, 768, 100
g = torch.Generator(device="cuda").manual_seed(0)
dq = torch.randn((Q, D), generator=g, device="cuda")
dc = torch.randn((N, D), generator=g, device="cuda")
idx = torch.empty((Q, k), dtype=torch.int32, device="cuda")
sc = torch.empty((Q, k), dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
fn = lambda: _native.dev_topk(_native.dev_matrix(dq.data_ptr(), Q, D, 1), _native.dev_matrix(dc.data_ptr(), N, D, 1), k, 1,
                              index_ptr=idx.data_ptr(), score_ptr=sc.data_ptr(), stream=st)
fn(); fn()
torch.cuda.synchronize()
_native.set_option("tc_debug_skip", 8)
_native.get_stat("tc_dbg_wait0")
fn()
torch.cuda.synchronize()
w = [_native.get_stat(f"tc_dbg_wait{i}") for i in range(52)]
_native.set_option("tc_debug_skip", 0)
tot = w[2]
print(f"MMA warps: total {tot:.3e} cycles; wait for a free accumulator {w[0] / tot:.3f}, for operands {w[1] / tot:.3f}; epilogue warp 2 flush cycles {w[3]:.3e}")
print("bucket: tiles        acc-wait share of total | epilogue filter cycles | flush cycles (warp 2 of leader CTAs)")
for b in range(13):
    print(f"{b:2d}: {2**b:5d}..{2**(b+1)-1:5d}   {w[4 + b] / tot:8.4f}   {w[20 + b]:.3e}   {w[36 + b]:.3e}")
