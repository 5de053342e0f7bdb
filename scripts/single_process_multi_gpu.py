"""One process, the whole box: pmm_topk / pmm_matmul through the single-process GPU group inside libpmm_b200 (one host
thread per GPU, corpus rows resp. output rows sharded, NCCL all-to-all of packed candidates) against the same calls pinned
to one GPU. C3 (100k x 1M x 768 f32, k=100) end to end with pageable host buffers; prints one JSON object."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from polars_matmul_b200 import _native
from polars_matmul_b200.arrow import from_numpy

Q, N, D, k = 100_000, 1_000_000, 768, 100
rng = np.random.default_rng(42)
q = rng.standard_normal((Q, D), dtype=np.float32)
c = np.empty((N, D), np.float32)
for lo in range(0, N, 65536):
    c[lo:lo + 65536] = rng.standard_normal((min(N, lo + 65536) - lo, D), dtype=np.float32)
hq, hc = from_numpy(q), from_numpy(c)
out = {"gpus": _native.device_count(), "workload": f"C3 {Q}x{N}x{D} f32 k={k}, pageable host buffers, pmm_topk (C ABI)"}


def timed(fn, n=3):
    fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        r = fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts), r


for metric in ("dot", "cosine"):
    _native.set_option("multi_gpu", 0)
    ms1, (i1, s1) = timed(lambda: _native.topk(hq, hc, k, metric))
    _native.set_option("multi_gpu", 1)
    _native.topk(hq, hc, k, metric)                       # first call: NCCL communicator + connection set-up
    _native.set_option("profile", 1)
    _native.reset_stats()
    msg, (ig, sg) = timed(lambda: _native.topk(hq, hc, k, metric))
    stats = {n: _native.get_stat(n + "_ms") / 4 / max(1, _native.device_count()) for n in ("group_broadcast", "group_exchange", "group_merge")
             if _native.get_stat(n + "_ms") > 0}   # per call and GPU; the brackets include waiting for the slowest peer
    _native.set_option("profile", 0)
    out[metric] = {"one_gpu_ms": ms1, "all_gpus_ms": msg, "speedup": ms1 / msg, "identical": bool(np.array_equal(i1, ig) and np.array_equal(s1, sg)),
                   "collectives_ms_per_gpu": stats}
# raw matmul: output rows sharded, the device->host copy of the result runs on all host links at once
ql, cl = from_numpy(q[:16384, :256].copy()), from_numpy(c[:65536, :256].copy())
_native.set_option("multi_gpu", 0)
m1, r1 = timed(lambda: _native.matmul(ql, cl), 2)
_native.set_option("multi_gpu", 1)
mg, rg = timed(lambda: _native.matmul(ql, cl), 2)
out["matmul_16384x65536x256_f32"] = {"one_gpu_ms": m1, "all_gpus_ms": mg, "speedup": m1 / mg, "identical": bool(np.array_equal(r1, rg)),
                                     "result_bytes": int(r1.nbytes)}
print("RESULT " + json.dumps(out))   # (NCCL may print its version banner to stdout as well)
