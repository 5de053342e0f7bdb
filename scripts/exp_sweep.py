"""Experiment: fused kernel time vs schedule group size and pipeline shape (kernel-only, CUDA events)."""
import json
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from polars_matmul_b200 import _native
from polars_matmul_b200.arrow import to_host_matrix
from oracle import pmm_oracle as oracle

def check(rowb, clm=1, c4=0):
    _native.set_option("tc_cg", rowb); _native.set_option("tc_clm", clm); _native.set_option("tc_cluster4", c4)
    rng = np.random.default_rng(1)
    q = rng.standard_normal((300, 200)).astype(np.float32)
    c = rng.standard_normal((5000, 200)).astype(np.float32)
    for metric in ("cosine", "dot"):
        idx, sc = _native.topk(to_host_matrix(q), to_host_matrix(c), 10, metric)
        oi, osc = oracle.topk(q, c, 10, metric)
        assert np.array_equal(idx, oi) and np.array_equal(sc, osc), (rowb, metric)
    h = rng.standard_normal((300, 200)).astype(np.float16)
    ch = rng.standard_normal((5000, 200)).astype(np.float16)
    idx, sc = _native.topk(to_host_matrix(h), to_host_matrix(ch), 10, "cosine")
    oi, osc = oracle.topk(h.astype(np.float32), ch.astype(np.float32), 10, "cosine")
    assert np.array_equal(idx, oi) and np.array_equal(sc, osc), (rowb, "f16")
    print("parity ok cg", rowb, "clm", clm, flush=True)

check(2)
Q, N, D, k = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (100000, 1000000, 768, 100)))
g = torch.Generator(device="cuda").manual_seed(0)
dq = torch.randn((Q, D), generator=g, device="cuda")
dc = torch.randn((N, D), generator=g, device="cuda")
idx = torch.empty((Q, k), dtype=torch.int32, device="cuda")
sc = torch.empty((Q, k), dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def run():
    _native.dev_topk(_native.dev_matrix(dq.data_ptr(), Q, D, 1), _native.dev_matrix(dc.data_ptr(), N, D, 1), k, 1,
                     index_ptr=idx.data_ptr(), score_ptr=sc.data_ptr(), stream=st)
res = []
_native.set_option("verify", 0)
configs = [(2, 0, 32, 3, 0, 8), (2, 0, 32, 3, 0, 1), (2, 0, 32, 2, 0, 0)]
for rowb, grp, rs, lv, clm, c4 in configs:
    _native.set_option("tc_cg", rowb); _native.set_option("tc_group", grp); _native.set_option("tc_sync_tiles", rs); _native.set_option("tc_levels", lv); _native.set_option("tc_max_flush", clm); _native.set_option("tc_debug_skip", c4)
    run(); torch.cuda.synchronize()
    _native.set_option("profile", 1); _native.reset_stats()
    run(); run(); torch.cuda.synchronize()
    ms = (_native.get_stat("tc_topk_tf32x3_ms") + _native.get_stat("tc_topk_tf32x1_ms") + _native.get_stat("tc_topk_f16r_ms")) / 2
    rq = _native.get_stat("requeried_tf32x3") / 2
    _native.set_option("profile", 0)
    w = [_native.get_stat("tc_dbg_wait%d" % i) / 1e6 for i in range(52)]
    print("Mcycles tempty/full/total/flush:", [round(x) for x in w[:4]])
    print("  MMA stall by octave :", [round(x) for x in w[4:20]])
    print("  filter by octave    :", [round(x) for x in w[20:36]])
    print("  flush by octave     :", [round(x) for x in w[36:52]])
    tf = 2.0 * Q * N * D / ms / 1e9
    res.append({"cg": rowb, "sync_tiles": rs, "levels": lv, "max_flush": clm, "debug_skip": c4, "requeried": rq, "group": grp, "kernel_ms": ms, "tflops": tf})
    print(res[-1], flush=True)
# C5-shaped f16 shard: 1M x 125k x 1024 f16 cosine k=10
Qh, Nh, Dh, kh = 1_000_000, 125_000, 1024, 10
ha = torch.randn((Qh, Dh), generator=g, device="cuda").half(); hb = torch.randn((Nh, Dh), generator=g, device="cuda").half()
hidx = torch.empty((Qh, kh), dtype=torch.int32, device="cuda"); hsc = torch.empty((Qh, kh), dtype=torch.float64, device="cuda")
def runh():
    _native.dev_topk(_native.dev_matrix(ha.data_ptr(), Qh, Dh, 0), _native.dev_matrix(hb.data_ptr(), Nh, Dh, 0), kh, 0,
                     index_ptr=hidx.data_ptr(), score_ptr=hsc.data_ptr(), stream=st)
for dbg in (0, 1, 8):
    _native.set_option("tc_debug_skip", dbg)
    runh(); torch.cuda.synchronize()
    _native.set_option("profile", 1); _native.reset_stats()
    runh(); runh(); torch.cuda.synchronize()
    msh = _native.get_stat("tc_topk_f16_ms") / 2
    _native.set_option("profile", 0)
    print({"f16_C5_kernel_ms": msh, "debug_skip": dbg, "tflops": 2.0 * Qh * Nh * Dh / msh / 1e9}, flush=True)
    if dbg == 8:
        w = [_native.get_stat("tc_dbg_wait%d" % i) / 1e6 for i in range(52)]
        print("Mcycles tempty/full/total/flush:", [round(x) for x in w[:4]])
        print("  MMA stall by octave :", [round(x) for x in w[4:20]])
        print("  filter by octave    :", [round(x) for x in w[20:36]])
        print("  flush by octave     :", [round(x) for x in w[36:52]])
_native.set_option("tc_debug_skip", 0)
json.dump(res, open("gpurun_out/sweep.json", "w"), indent=1)
