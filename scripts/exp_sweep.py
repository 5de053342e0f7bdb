"""Experiment: merge-policy knobs of the fused top-k kernel (kernel-only, CUDA events): tc_soft_at, tc_max_flush.
Earlier sweeps (schedule group, pacing, debug counters) are recorded in profiles/sweep_r1.md."""
import json
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from polars_matmul_b200 import _native
from polars_matmul_b200.arrow import to_host_matrix
from oracle import pmm_oracle as oracle

rng = np.random.default_rng(1)
q = rng.standard_normal((300, 200)).astype(np.float32); c = rng.standard_normal((5000, 200)).astype(np.float32)
for metric in ("cosine", "dot"):
    idx, sc = _native.topk(to_host_matrix(q), to_host_matrix(c), 10, metric)
    oi, osc = oracle.topk(q, c, 10, metric)
    assert np.array_equal(idx, oi) and np.array_equal(sc, osc)
print("parity ok", flush=True)

Q, N, D, k = 100000, 1000000, 768, 100
g = torch.Generator(device="cuda").manual_seed(0)
dq = torch.randn((Q, D), generator=g, device="cuda"); dc = torch.randn((N, D), generator=g, device="cuda")
idx = torch.empty((Q, k), dtype=torch.int32, device="cuda"); sc = torch.empty((Q, k), dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def run():
    _native.dev_topk(_native.dev_matrix(dq.data_ptr(), Q, D, 1), _native.dev_matrix(dc.data_ptr(), N, D, 1), k, 1,
                     index_ptr=idx.data_ptr(), score_ptr=sc.data_ptr(), stream=st)
def timed(fn, stat):
    fn(); torch.cuda.synchronize()
    _native.set_option("profile", 1); _native.reset_stats()
    fn(); fn(); torch.cuda.synchronize()
    ms = _native.get_stat(stat) / 2
    _native.set_option("profile", 0)
    return ms
res = []
for soft, mf in [(0, 0), (24, 0), (32, 0), (64, 0), (80, 0), (0, 2), (0, 4), (0, 32), (32, 4), (64, 4)]:
    _native.set_option("tc_soft_at", soft); _native.set_option("tc_max_flush", mf)
    ms = timed(run, "tc_topk_f16r_ms")
    res.append({"case": "C3 f16r", "soft_at": soft, "max_flush": mf, "kernel_ms": ms, "tflops": 2.0 * Q * N * D / ms / 1e9})
    print(res[-1], flush=True)
del dq, dc
# C5-shaped f16 shard: 1M x 125k x 1024 f16 cosine k=10
Qh, Nh, Dh, kh = 1_000_000, 125_000, 1024, 10
ha = torch.randn((Qh, Dh), generator=g, device="cuda").half(); hb = torch.randn((Nh, Dh), generator=g, device="cuda").half()
hidx = torch.empty((Qh, kh), dtype=torch.int32, device="cuda"); hsc = torch.empty((Qh, kh), dtype=torch.float64, device="cuda")
def runh():
    _native.dev_topk(_native.dev_matrix(ha.data_ptr(), Qh, Dh, 0), _native.dev_matrix(hb.data_ptr(), Nh, Dh, 0), kh, 0,
                     index_ptr=hidx.data_ptr(), score_ptr=hsc.data_ptr(), stream=st)
for soft, mf in [(0, 0), (24, 0), (32, 0), (64, 0), (0, 2), (0, 4)]:
    _native.set_option("tc_soft_at", soft); _native.set_option("tc_max_flush", mf)
    ms = timed(runh, "tc_topk_f16_ms")
    res.append({"case": "C5 f16", "soft_at": soft, "max_flush": mf, "kernel_ms": ms, "tflops": 2.0 * Qh * Nh * Dh / ms / 1e9})
    print(res[-1], flush=True)
json.dump(res, open("gpurun_out/sweep.json", "w"), indent=1)
