"""Secondary measurements (kernel-only, CUDA events, inputs resident in HBM): the HBM-bound kernels
(norms, prep, raw matmul) and the other top-k configurations of BASELINE.json. Prints one JSON object."""
import json
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from polars_matmul_b200 import _native

PEAK_HBM = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if __import__("os").path.exists("MEASURED_PEAKS.json") else 6650.0
st = lambda: torch.cuda.current_stream().cuda_stream
CODE = {torch.float16: 0, torch.float32: 1, torch.float64: 2}
out = {"hbm_peak_gbs": PEAK_HBM}

def timed(fn, name, iters=5, warm=2, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    _native.set_option("profile", 1); _native.reset_stats()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        fn()
    ev[1].record(); torch.cuda.synchronize()
    stats = {}
    for k in ("prep", "norms", "tc_topk_f16r", "tc_topk_f16r_kp256", "tc_topk_tf32x1", "tc_topk_tf32x3", "tc_topk_f16", "tc_matmul_tf32x3", "tc_matmul_f16", "scores_f32", "scores_f64", "scores_f64_dmma",
              "select_f32", "select_f64", "merge", "rescore"):
        v = _native.get_stat(k + "_ms")
        if v:
            stats[k + "_ms"] = v / iters
    _native.set_option("profile", 0)
    return stats

g = torch.Generator(device="cuda").manual_seed(1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

# ---- norms: N x D streaming read
for (N, D, dt) in ((1_000_000, 768, torch.float32), (1_000_000, 1024, torch.float16), (4_000_000, 256, torch.float32)):
    x = torch.randn((N, D), generator=g, device="cuda", dtype=torch.float32).to(dt)
    o = torch.empty(N, dtype=torch.float32, device="cuda")
    s = timed(lambda: _native.dev_norms(_native.dev_matrix(x.data_ptr(), N, D, CODE[dt]), False, o.data_ptr(), st()), "norms", flush=flush)
    gb = N * D * x.element_size() / 1e9
    out[f"norms_{N}x{D}_{str(dt)[6:]}"] = {"ms": s["norms_ms"], "GBps": gb / s["norms_ms"] * 1e3, "frac_hbm": gb / s["norms_ms"] * 1e3 / PEAK_HBM,
                                            "algorithmic_GB": gb}
    del x

# ---- raw matmul
def matmul_case(Q, N, D, dt, label):
    a = torch.randn((Q, D), generator=g, device="cuda", dtype=torch.float32).to(dt)
    b = torch.randn((N, D), generator=g, device="cuda", dtype=torch.float32).to(dt)
    odt = torch.float64 if dt == torch.float64 else torch.float32
    o = torch.empty((Q, N), dtype=odt, device="cuda")
    s = timed(lambda: _native.dev_matmul(_native.dev_matrix(a.data_ptr(), Q, D, CODE[dt]), _native.dev_matrix(b.data_ptr(), N, D, CODE[dt]),
                                         o.data_ptr(), st()), label, flush=flush)
    kern = [k for k in s if k.startswith(("tc_matmul", "scores"))][0]
    gb = (Q * N * o.element_size() + (Q + N) * D * a.element_size()) / 1e9
    out[label] = {"kernel": kern[:-3], "kernel_ms": s[kern], "prep_ms": s.get("prep_ms"), "algorithmic_GB": gb,
                  "GBps_kernel": gb / s[kern] * 1e3, "frac_hbm": gb / s[kern] * 1e3 / PEAK_HBM,
                  "TFLOPs": 2.0 * Q * N * D / s[kern] / 1e9}
matmul_case(1000, 10000, 256, torch.float32, "matmul_C2_f32_1000x10000x256")
matmul_case(1000, 10000, 256, torch.float64, "matmul_C2_f64_1000x10000x256")
matmul_case(16384, 65536, 64, torch.float32, "matmul_f32_16384x65536x64")
matmul_case(16384, 65536, 32, torch.float32, "matmul_f32_16384x65536x32")
matmul_case(16384, 65536, 256, torch.float16, "matmul_f16_16384x65536x256")
matmul_case(16384, 65536, 256, torch.float32, "matmul_f32_16384x65536x256")

# ---- other top-k configurations
def topk_case(Q, N, D, k, metric, dt, label, iters=2):
    a = torch.randn((Q, D), generator=g, device="cuda", dtype=torch.float32).to(dt)
    b = torch.randn((N, D), generator=g, device="cuda", dtype=torch.float32).to(dt)
    idx = torch.empty((Q, k), dtype=torch.int32, device="cuda"); sc = torch.empty((Q, k), dtype=torch.float64, device="cuda")
    m = _native.metric_from_str(metric)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn = lambda: _native.dev_topk(_native.dev_matrix(a.data_ptr(), Q, D, CODE[dt]), _native.dev_matrix(b.data_ptr(), N, D, CODE[dt]), k, m,
                                  index_ptr=idx.data_ptr(), score_ptr=sc.data_ptr(), stream=st())
    fn(); torch.cuda.synchronize()
    e0.record()
    s = timed(fn, label, iters=iters, warm=0)
    e1.record(); torch.cuda.synchronize()
    total_ms = e0.elapsed_time(e1) / iters
    kern = max([k_ for k_ in s if k_.startswith(("tc_topk", "scores"))], key=lambda k_: s[k_])
    out[label] = {"step_ms": total_ms, "queries_per_s": Q / total_ms * 1e3, "kernel": kern[:-3], "kernel_ms": s[kern],
                  "kernel_TFLOPs": 2.0 * Q * N * D / s[kern] / 1e9, "per_kernel_ms": s,
                  "requeried_f16_wide": _native.get_stat("requeried_f16_wide") / iters, "requeried_tf32x3": _native.get_stat("requeried_tf32x3") / iters, "fallback_queries": _native.get_stat("fallback_queries") / iters}
topk_case(1000, 10000, 256, 10, "cosine", torch.float32, "topk_C1_1000x10000x256_cosine_k10", iters=20)
topk_case(100_000, 1_000_000, 768, 100, "euclidean", torch.float32, "topk_C3_euclidean")
topk_case(100_000, 1_000_000, 768, 100, "cosine", torch.float32, "topk_C3shape_cosine")
topk_case(100_000, 1_250_000, 768, 100, "cosine", torch.float32, "topk_C4_shard_1of8_cosine")
topk_case(1_000_000, 125_000, 1024, 10, "cosine", torch.float16, "topk_C5_shard_1of8_f16_cosine_k10")
topk_case(2000, 100_000, 256, 10, "cosine", torch.float64, "topk_f64_2000x100000x256", iters=1)
print(json.dumps(out, indent=1))
