#!/bin/bash
# First GPU pass of round 2: the whole -m gpu suite (one process per file, so a CUDA fault in one file cannot poison the
# others), smoke, a short bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt 2>&1
nproc >> gpurun_out/r2a_gpu.txt; free -g >> gpurun_out/r2a_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2a_smoke.log
for f in test_gpu_parity test_gpu_bound test_gpu_property; do
  timeout 1500 python -m pytest tests/$f.py -m gpu -q -rf -s -p no:cacheprovider > gpurun_out/r2a_$f.log 2>&1
  echo "rc=$?" >> gpurun_out/r2a_$f.log
done
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?" >> gpurun_out/r2a_bench.err
tail -5 gpurun_out/r2a_smoke.log; for f in test_gpu_parity test_gpu_bound test_gpu_property; do tail -15 gpurun_out/r2a_$f.log; done; tail -3 gpurun_out/r2a_bench.err; head -c 1500 gpurun_out/r2a_bench.json
