#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -rf -p no:cacheprovider -k "page_locked or host_chunked or c5_shard or namespace" > gpurun_out/r2n_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2n_pytest.log
tail -n 5 gpurun_out/r2n_pytest.log
for d in 1 0; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --opt d2h_direct=$d > gpurun_out/r2n_bench_d$d.json 2> gpurun_out/r2n_bench_d$d.err; echo "rc=$?" >> gpurun_out/r2n_bench_d$d.err
tail -n 2 gpurun_out/r2n_bench_d$d.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2n_bench_d$d.json').read().strip().splitlines()[-1])
print('d2h_direct=$d', round(d['ms_per_step'],2), d['selfcheck']['exact'], 'e2e pageable', round(d['e2e']['ms_per_step'],2), 'pinned', round(d['e2e']['pinned_inputs']['ms_per_step'],2))
PY
done
