#!/bin/bash
# First-contact GPU run: each group in its own process so that one faulting kernel cannot hide the rest.
# Usage (under gpurun): bash scripts/gpu_check.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,driver_version --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/host.txt
run() { # name, timeout, cmd...
  local name=$1; local to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1
  echo "exit $? : $name" | tee -a gpurun_out/summary.txt
  tail -n 6 "gpurun_out/$name.log" | tee -a gpurun_out/summary.txt
}
run t_golden   300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "known_answers or errors or fixtures" --timeout 120
run t_generic  300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "generic or f64 or mixed or norms" --timeout 120
run t_tc_small 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tc_topk_f32" --timeout 120
run t_tc_rest  400 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "c1 or matmul or tie or nan or zero_norm" --timeout 120
run t_misc     400 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "list_input or integer or f16 or containers or resident or shards" --timeout 120
run smoke      200 python __graft_entry__.py smoke
run bench_small 300 python bench.py --small --steps 3 --warmup 3
