#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -rf -p no:cacheprovider -x -k "warm_seeds or full_size or filter_levels or host_chunked" > gpurun_out/r2r_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2r_pytest.log
tail -n 15 gpurun_out/r2r_pytest.log
for w in 1 0 1 0; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --no-e2e --opt warm_seed=$w > gpurun_out/r2r_bench_w$w.json 2> gpurun_out/r2r_bench_w$w.err; echo "rc=$?" >> gpurun_out/r2r_bench_w$w.err
tail -n 1 gpurun_out/r2r_bench_w$w.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2r_bench_w$w.json').read().strip().splitlines()[-1])
print('warm_seed=$w', round(d['ms_per_step'],2), d['selfcheck']['exact'], {k:round(v,2) for k,v in d['roofline']['per_kernel_ms_per_step'].items()})
PY
done
