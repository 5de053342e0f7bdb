#!/bin/bash
# Final check of the f16-split matmul: whole GPU suite, error soak, bench extras.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf -p no:cacheprovider > gpurun_out/r2m_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_pytest.log
timeout 300 python scripts/matmul_error_soak.py --seconds 90 > gpurun_out/r2m_matmul_soak.json 2> gpurun_out/r2m_matmul_soak.err; echo "rc=$?" >> gpurun_out/r2m_matmul_soak.err
timeout 900 python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "rc=$?" >> gpurun_out/r2m_bench.err
tail -n 5 gpurun_out/r2m_pytest.log; tail -n 2 gpurun_out/r2m_matmul_soak.err; tail -n 2 gpurun_out/r2m_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2m_matmul_soak.json'))
print(d['cases'], round(d['seconds']), 'worst', max(v['worst_ratio'] for v in d['buckets'].values()), 'entries', sum(v['entries'] for v in d['buckets'].values()))
d=json.loads(open('gpurun_out/r2m_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['selfcheck']['exact'], d['e2e']['value'])
for k,v in d['extra']['matmul'].items(): print(k, v['kernel'], round(v['kernel_ms'],4), round(v['frac_hbm'],3), round(v['TFLOPs'],1))
PY
