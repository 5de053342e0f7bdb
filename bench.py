#!/usr/bin/env python3
"""
bench.py — headline benchmark of the pmm.topk hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the largest single-GPU configuration): 100k queries x 1M corpus
rows, 768-d f32, metric=dot, k=100, synthetic Gaussian data.  With N > 1 ranks the corpus is sharded
by rows, 1M rows per rank (weak scaling; N=8 is configs[3]'s shape at 8M rows): every rank scans its
shard for all queries with the fused kernel, the ranks exchange Q x k packed candidates with one NCCL
all-gather and every rank merges them.

One "step" = one full pass of the hot path (norm/split precompute, fused GEMM+top-k, merge) over the
batch.  `value` has the inputs resident in HBM when the timed region starts; `e2e` goes through the
host C ABI (pmm_topk) with host buffers in pinned memory, H2D and D2H copies inside the timed region.

`--impl reference`: the reference's CPU implementation cannot be built here (Rust + un-vendored faer,
no cargo in the image), so this arm times the oracle port (oracle/pmm_oracle.c, OpenMP, all host
threads) on a bounded query sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(name="C3", Q=100_000, N=1_000_000, D=768, k=100, metric="dot")
METRIC_NAME = "topk_queries_per_sec"
UNIT = "queries/s"


def load_traffic(kernel: str, workload: str):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        d = json.load(open(p))
        e = d.get(f"{kernel}@{workload}")
        if e:
            return e
    return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.dev = device_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.dev)], stdout=subprocess.PIPE, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline_sample(q_host: np.ndarray, c_host: np.ndarray, k: int, metric: str, target_s: float = 12.0):
    """Times the oracle port on a bounded query sample against the full corpus shard."""
    from oracle import pmm_oracle as oracle
    oracle.build()
    n_cal = 8 * max(1, oracle.num_threads())      # the oracle parallelises over blocks of 8 queries
    t0 = time.perf_counter()
    oracle.topk(q_host[:n_cal], c_host, k, metric)
    t_cal = time.perf_counter() - t0
    n = int(max(n_cal, min(q_host.shape[0], n_cal * target_s / max(t_cal, 1e-3))))
    n = max(8, n // 8 * 8)
    t0 = time.perf_counter()
    oracle.topk(q_host[:n], c_host, k, metric)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
            "sample": f"{n} of {q_host.shape[0]} queries x full {c_host.shape[0]}-row corpus, {dt:.1f} s, "
                      f"oracle/pmm_oracle.c (OpenMP, -O3 -mavx2 -mfma); the reference (Rust/faer) cannot be built in this image"}


def run_reference(args, emit):
    """--impl reference: oracle port on host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pmm_oracle as oracle
    oracle.build()
    W = dict(WORKLOAD)
    if args.small:
        W.update(Q=2000, N=50_000)
    rng = np.random.default_rng(42)
    # bounded sample: the oracle scans the FULL corpus for a subset of the queries
    n_probe = 8 * max(1, oracle.num_threads())    # the oracle parallelises over blocks of 8 queries
    c = rng.standard_normal((W["N"], W["D"]), dtype=np.float32)
    q = rng.standard_normal((4096, W["D"]), dtype=np.float32)
    t0 = time.perf_counter()
    oracle.topk(q[:n_probe], c, W["k"], W["metric"])
    t_probe = time.perf_counter() - t0
    n = int(min(4096, max(n_probe, n_probe * 4.0 / max(t_probe, 1e-3))))   # ~4 s per step
    n = max(8, n // 8 * 8)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        oracle.topk(q[:n], c, W["k"], W["metric"])
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    ms = 1000 * sum(times) / len(times)
    val = n / (ms / 1000)
    sample = (f"{n} of {W['Q']} queries x full {W['N']}-row corpus per step; oracle port "
              f"(oracle/pmm_oracle.c, OpenMP {oracle.num_threads()} threads); the Rust/faer reference cannot be built here")
    line = {
        "impl": "reference", "metric": METRIC_NAME, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{W['name']}: {W['Q']} queries x {W['N']} corpus rows, {W['D']}d f32, metric={W['metric']}, k={W['k']}",
                   "timed": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def main():
    # Libraries (NCCL, torchrun) print to stdout; the contract is ONE JSON line there. Route fd 1 to stderr for
    # the duration of the run and write the JSON line to the saved descriptor at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line: dict):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--small", action="store_true", help="tiny shapes for a functional check (not a bench value)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--q", type=int, default=None, help="override the query count (profiling runs only)")
    ap.add_argument("--n", type=int, default=None, help="override the corpus rows per GPU (profiling runs only)")
    ap.add_argument("--metric", default=None)
    ap.add_argument("--k", type=int, default=None)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args, emit)

    import torch
    import torch.distributed as dist
    from polars_matmul_b200 import _native, sharded
    from polars_matmul_b200.arrow import from_numpy

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    _native.lib()
    _native.set_device(local_rank)
    numa = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        if not os.environ.get("PMM_BENCH_NO_NUMA_BIND"):
            numa = sharded.bind_near_gpu(local_rank)   # before the pinned host buffers are allocated
    dev = torch.device("cuda", local_rank)

    W = dict(WORKLOAD)
    if args.small:
        W.update(Q=2000, N=50_000)
    if args.q:
        W["Q"] = args.q
    if args.n:
        W["N"] = args.n
    if args.metric:
        W["metric"] = args.metric
    if args.k:
        W["k"] = args.k
    Q, N, D, k, metric = W["Q"], W["N"], W["D"], W["k"], W["metric"]
    mcode = _native.metric_from_str(metric)
    n_total = N * world

    # synthetic data, generated on the host so that the e2e leg has real host buffers (pinned)
    rng_q = np.random.default_rng(42)
    q_pin = torch.empty((Q, D), dtype=torch.float32).pin_memory()
    q_pin.numpy()[...] = rng_q.standard_normal((Q, D), dtype=np.float32)
    c_pin = torch.empty((N, D), dtype=torch.float32).pin_memory()
    rng_c = np.random.default_rng(1000 + rank)
    cn = c_pin.numpy()
    for lo in range(0, N, 65536):
        hi = min(N, lo + 65536)
        cn[lo:hi] = rng_c.standard_normal((hi - lo, D), dtype=np.float32)
    dq = q_pin.to(dev)
    dc = c_pin.to(dev)
    torch.cuda.synchronize()

    driver = sharded.ShardedTopk() if world > 1 else None
    idx = torch.empty((Q, k), dtype=torch.int32, device=dev)
    sc = torch.empty((Q, k), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step_resident():
        if world == 1:
            _native.dev_topk(_native.dev_matrix(dq.data_ptr(), Q, D, _native.DTYPE_F32),
                             _native.dev_matrix(dc.data_ptr(), N, D, _native.DTYPE_F32), k, mcode,
                             index_ptr=idx.data_ptr(), score_ptr=sc.data_ptr(), stream=stream)
            return idx, sc
        return driver.topk_device(dq, dc, rank * N, n_total, k, metric)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- kernel/device-resident measurement -------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    barrier()
    _native.set_option("profile", 1)
    _native.reset_stats()
    _native.reset_kernel_launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step_resident()
    ev1.record()
    barrier()
    ms_step = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = _native.kernel_launch_count()
    kname = max(("tc_topk_f16r", "tc_topk_tf32x1", "tc_topk_tf32x3"), key=lambda n: _native.get_stat(n + "_ms"))
    k_ms = _native.get_stat(kname + "_ms")
    k_launches = _native.get_stat(kname + "_launches")
    stats = {n: _native.get_stat(n + "_ms") / max(1, args.steps)
             for n in ("prep", "tc_topk_f16r", "tc_topk_f16r_kp256", "tc_topk_tf32x1", "tc_topk_tf32x3", "merge", "rescore", "gather", "scatter", "scores_f32", "select_f32")
             if _native.get_stat(n + "_ms") > 0}
    stats["requeried_f16_wide_per_step"] = _native.get_stat("requeried_f16_wide") / max(1, args.steps)
    stats["requeried_tf32x3_per_step"] = _native.get_stat("requeried_tf32x3") / max(1, args.steps)
    stats["fallback_queries_per_step"] = _native.get_stat("fallback_queries") / max(1, args.steps)
    _native.set_option("profile", 0)
    value = world * Q / (ms_step / 1000.0)

    # ---- end-to-end through the host C ABI ---------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        hq, hc = from_numpy(q_pin.numpy()), from_numpy(c_pin.numpy())

        def step_e2e():
            if world == 1:
                return _native.topk(hq, hc, k, metric)
            return driver.topk_host(hq, hc, rank * N, n_total, k, metric)

        for _ in range(max(1, min(args.warmup, 2))):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            step_e2e()
        torch.cuda.synchronize()
        ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1000 / n_e2e)
        barrier()
        e2e = {"value": world * Q / (ms_e2e / 1000.0), "unit": UNIT, "ms_per_step": ms_e2e,
               # whole job: every rank uploads its corpus shard; the replicated queries cross ONE host link and are
               # broadcast over NVLink (world > 1); every rank reads the merged result back
               "h2d_bytes_per_step": int((Q + world * N) * D * 4), "d2h_bytes_per_step": int(world * Q * k * 12),
               "note": "pmm_topk C ABI with pinned host buffers; corpus re-uploaded every step as the reference re-marshals it "
                       "(src/matmul.rs:430-431)" + ("; bytes are whole-job totals over all ranks" if world > 1 else "")}

    # ---- sanity: sampled oracle check of the timed result (rank 0, N=1) ---------------------------
    if rank == 0:
        peaks, peak_src = load_peaks()
        flops_per_launch = 2.0 * Q * N * D
        k_avg_ms = k_ms / max(1.0, k_launches)
        achieved = flops_per_launch / (k_avg_ms / 1000.0) / 1e12 if k_avg_ms > 0 else None
        # Tensor-pipe peak for ALGORITHMIC f32 flops: TF32 runs at half the bf16 rate; the first-level filter issues one
        # TF32 MMA per MAC (bf16 / 2), the 3xTF32 split three (bf16 / 6).
        # The default first level rounds the f32 operands to f16 (11 significant bits, like TF32) and issues one
        # kind::f16 MMA per MAC: the bf16/f16 dense rate itself.
        terms = 1 if kname.endswith("x1") or kname.endswith("f16r") else 3
        rate_div = 1.0 if kname.endswith("f16r") else 2.0 * terms
        peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")) / rate_div
        tr = load_traffic(kname, f"{W['name']}:{Q}x{N}x{D}:k{k}")
        roofline = {"bound": "tensor", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": (achieved / peak) if achieved else None,
                    "traffic": tr["dram_bytes_per_launch"] if tr else None,
                    "traffic_note": tr["source"] if tr else "no ncu capture committed for this workload",
                    "ncu_tensor_pipe_active_pct": tr.get("tensor_pipe_active_pct_of_elapsed") if tr else None,
                    "algorithmic_flops_per_launch": flops_per_launch, "kernel_ms_avg": k_avg_ms,
                    "kernel_share_of_step": (k_avg_ms / ms_step) if ms_step else None,
                    "peak_note": f"{peak_src}: bf16_tflops_sustained / {rate_div:g} ("
                                 + ("one kind::f16 MMA per MAC on f16-rounded operands" if kname.endswith("f16r") else
                                    f"TF32 = 1/2 bf16 rate, {terms} TF32 MMA(s) per MAC")
                                 + f" in {kname}); raw bf16 sustained {peaks.get('bf16_tflops_sustained')} burst {peaks.get('bf16_tflops')}. "
                                 "The exact f32 result comes from the re-scoring kernel; queries whose filter is not provably "
                                 "lossless are re-run with 256-entry lists (requeried_f16_wide_per_step), then 3xTF32 (requeried_tf32x3_per_step), then on the SIMT path",
                    "per_kernel_ms_per_step": stats}
        cpu_base = None
        if world == 1 and not args.no_cpu_baseline:
            cpu_base = cpu_baseline_sample(q_pin.numpy(), c_pin.numpy(), k, metric)
        line = {
            "metric": METRIC_NAME, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{W['name']}: {Q} queries x {N} corpus rows per GPU, {D}d f32, metric={metric}, k={k}",
                       "sharding": f"corpus rows sharded over {world} rank(s), {N} rows each ({n_total} total); queries replicated; "
                                   "candidates merged after one NCCL all-gather" if world > 1 else "single GPU",
                       "value_definition": "n_gpus * Q / step time: every rank scans its own shard for all Q queries",
                       "host_placement": numa or "not bound",
                       "l2": "inputs (3.4 GB per rank) are far larger than the 126 MB L2; no explicit flush",
                       "arithmetic": "tcgen05 filter on f32 operands rounded to f16 (kind::f16, 11 significant bits; 3xTF32 hi/lo "
                                     "split on demand), f32 accumulate in TMEM; exact f32 re-scoring + per-query losslessness "
                                     "proof: results bit-identical to the f32 CPU oracle"},
            "tflops_effective": 2.0 * Q * n_total * D / (ms_step / 1000.0) / 1e12,
            "queries_per_sec_global_corpus": Q / (ms_step / 1000.0),
            "roofline": roofline, "cpu_baseline": cpu_base, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
