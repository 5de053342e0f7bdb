"""
Raw f32 matmul on the tensor cores against the CPU oracle: how close does |ours - oracle| come to the parity tolerance
(tests/parity.py: 1e-5 * max(|x|, 0.05 |q||c|)) per vector length and data kind?  TEST INFRASTRUCTURE (runs the oracle).

    python scripts/matmul_error_soak.py [--cases N] [--max-dim D] [--option name=value ...] > gpurun_out/matmul_soak.json

Prints one JSON object: per (kind, dim bucket) the worst ratio |diff| / tolerance, the number of entries compared, and
every case whose ratio exceeded 0.5 with its parameters (so that it can be replayed).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tests import parity                                     # noqa: E402
from tests.test_gpu_property import _data                    # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=20000)
    ap.add_argument("--max-dim", type=int, default=256)
    ap.add_argument("--seconds", type=float, default=150.0)
    ap.add_argument("--option", action="append", default=[])
    args = ap.parse_args()
    from polars_matmul_b200 import _native
    from polars_matmul_b200.arrow import to_host_matrix
    from oracle import pmm_oracle as oracle
    oracle.build()
    for o in args.option:
        name, val = o.split("=")
        _native.set_option(name, int(val))
    rng = np.random.default_rng(20261018)
    kinds = ["gauss", "dups", "zeros", "scaled", "tiny", "huge", "positive"]
    buckets = {}
    hot = []
    t0 = time.time()
    done = 0
    for case in range(args.cases):
        if time.time() - t0 > args.seconds:
            break
        kind = kinds[case % len(kinds)]
        d = int(rng.integers(1, args.max_dim + 1)) if case % 3 else int(rng.choice([args.max_dim, args.max_dim - 1, args.max_dim // 2, 9, 10, 12, 16, 20, 31, 32, 33, 48]))
        nq, n = int(rng.integers(1, 301)), int(rng.integers(1, 2501))
        seed = int(rng.integers(0, 2**31 - 1))
        if kind == "positive":
            r2 = np.random.default_rng(seed)
            q = np.abs(r2.standard_normal((nq, d))).astype(np.float32)
            c = np.abs(r2.standard_normal((n, d))).astype(np.float32)
        else:
            q, c = _data(nq, n, d, seed, kind, np.float32)
        out = _native.matmul(to_host_matrix(q), to_host_matrix(c))
        ref = oracle.matmul(q, c)
        scale = parity._scale(q, c, parity.DOT)
        tol = 1e-5 * np.maximum(np.abs(ref.astype(np.float64)), parity.FLOOR * scale)
        diff = np.abs(out.astype(np.float64) - ref.astype(np.float64))
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = np.where(diff == 0, 0.0, diff / tol)
        ratio = np.nan_to_num(ratio, nan=0.0, posinf=1e30)
        worst = float(ratio.max())
        key = f"{kind}:" + next((f"<={b:03d}" for b in (4, 8, 16, 32, 64, 128, 192, 256, 512, 1024) if d <= b), ">1024")
        b = buckets.setdefault(key, {"worst_ratio": 0.0, "entries": 0, "cases": 0})
        b["worst_ratio"] = max(b["worst_ratio"], worst)
        b["entries"] += int(ratio.size)
        b["cases"] += 1
        if worst > 0.5:
            i, j = np.unravel_index(int(ratio.argmax()), ratio.shape)
            hot.append({"kind": kind, "nq": nq, "n": n, "d": d, "seed": seed, "ratio": worst, "at": [int(i), int(j)],
                        "ours": float(out[i, j]), "oracle": float(ref[i, j]), "scale": float(scale[i, j]),
                        "diff_over_scale": float(diff[i, j] / scale[i, j]) if scale[i, j] else None})
        done += 1
    print(json.dumps({"cases": done, "seconds": time.time() - t0, "max_dim": args.max_dim, "options": args.option,
                      "buckets": dict(sorted(buckets.items())), "above_half_tolerance": hot}, indent=1))


if __name__ == "__main__":
    main()
