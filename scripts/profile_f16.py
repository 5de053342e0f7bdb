"""One C5-shaped f16 top-k call (for ncu): Q x 125k x 1024 f16 cosine k=10."""
import sys
import torch
sys.path.insert(0, ".")
from polars_matmul_b200 import _native
Q, N, D, k = int(sys.argv[1]) if len(sys.argv) > 1 else 131072, 125_000, 1024, 10
g = torch.Generator(device="cuda").manual_seed(1)
a = torch.randn((Q, D), generator=g, device="cuda").half()
b = torch.randn((N, D), generator=g, device="cuda").half()
idx = torch.empty((Q, k), dtype=torch.int32, device="cuda"); sc = torch.empty((Q, k), dtype=torch.float64, device="cuda")
for _ in range(2):
    _native.dev_topk(_native.dev_matrix(a.data_ptr(), Q, D, 0), _native.dev_matrix(b.data_ptr(), N, D, 0), k, 0,
                     index_ptr=idx.data_ptr(), score_ptr=sc.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("done")
