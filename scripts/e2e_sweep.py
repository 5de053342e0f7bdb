"""End-to-end leg only (C3, pageable NumPy inputs through _topk -> Arrow), under a list of option sets.
    python scripts/e2e_sweep.py "stage_threads=4" "stage_threads=12 stage_slot_mb=16" ...      ("-" = defaults)"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import polars_matmul_b200 as pmm
from polars_matmul_b200 import _native

Q, N, D, k = 100_000, 1_000_000, 768, 100
rng = np.random.default_rng(42)
q = rng.standard_normal((Q, D), dtype=np.float32)
c = np.empty((N, D), np.float32)
for lo in range(0, N, 65536):
    c[lo:lo + 65536] = rng.standard_normal((min(N, lo + 65536) - lo, D), dtype=np.float32)
_native.set_option("multi_gpu", 0)
DEFAULTS = {"stage_threads": 0, "stage_slot_mb": 32, "stage_slots": 4, "host_chunk_first_div": 0, "host_chunk_ratio_pct": 0}
for spec in sys.argv[1:] or ["-"]:
    for key, val in DEFAULTS.items():
        _native.set_option(key, val)
    if spec != "-":
        for kv in spec.split():
            key, _, val = kv.partition("=")
            _native.set_option(key, int(val))
    for _ in range(2):
        pmm._topk(q, c, k, "dot")
    ts = []
    for _ in range(4):
        t0 = time.perf_counter()
        pmm._topk(q, c, k, "dot")
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"[{spec}] e2e pageable ms: min {min(ts):.1f} median {sorted(ts)[len(ts) // 2]:.1f} max {max(ts):.1f}", flush=True)
