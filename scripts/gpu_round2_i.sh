#!/bin/bash
# Re-entry pass: where does the tensor-core raw matmul come close to the parity tolerance (soak failure of the property
# suite), the whole GPU suite, and the default bench line with the C5 extra.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python scripts/matmul_error_soak.py --seconds 150 > gpurun_out/r2i_matmul_soak.json 2> gpurun_out/r2i_matmul_soak.err; echo "rc=$?" >> gpurun_out/r2i_matmul_soak.err
timeout 1500 python -m pytest tests -m gpu -q -rf -p no:cacheprovider --durations=8 > gpurun_out/r2i_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_pytest.log
timeout 900 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "rc=$?" >> gpurun_out/r2i_bench.err
tail -5 gpurun_out/r2i_matmul_soak.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_matmul_soak.json'))
print(d['cases'], round(d['seconds']))
for k,v in d['buckets'].items(): print(k, round(v['worst_ratio'],3), v['entries'])
for h in d['above_half_tolerance'][:20]: print(h)
PY
tail -16 gpurun_out/r2i_pytest.log; tail -3 gpurun_out/r2i_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2i_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['selfcheck']['exact'], d['e2e']['value'], d['extra'].get('c5'))
PY
