#!/bin/bash
# Third GPU pass: whole parity file after the pipelining / matmul changes, bench with and without the pipelined first level.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -rf -p no:cacheprovider > gpurun_out/r2c_parity.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_parity.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --no-e2e > gpurun_out/r2c_bench_pipe.json 2> gpurun_out/r2c_bench_pipe.err; echo "rc=$?" >> gpurun_out/r2c_bench_pipe.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --no-e2e --opt pipeline=0 > gpurun_out/r2c_bench_nopipe.json 2> gpurun_out/r2c_bench_nopipe.err; echo "rc=$?" >> gpurun_out/r2c_bench_nopipe.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --no-e2e --metric cosine > gpurun_out/r2c_bench_cos.json 2> gpurun_out/r2c_bench_cos.err
tail -8 gpurun_out/r2c_parity.log
for f in pipe nopipe cos; do python - <<PY
import json
d=json.load(open('gpurun_out/r2c_bench_$f.json'))
print('$f', round(d['ms_per_step'],2), round(d['value']), d['selfcheck']['exact'], {k:round(v,2) for k,v in d['roofline']['per_kernel_ms_per_step'].items() if isinstance(v,float)})
PY
done
