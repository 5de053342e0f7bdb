#!/bin/bash
# Last pass of round 2: smoke, whole GPU suite, default bench line + reference arm, then a soak of the randomised suites.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2z_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2z_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -rf -p no:cacheprovider > gpurun_out/r2z_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2z_pytest.log
timeout 900 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "rc=$?" >> gpurun_out/r2z_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/r2z_ref.json 2> gpurun_out/r2z_ref.err; echo "rc=$?" >> gpurun_out/r2z_ref.err
PMM_PROPERTY_EXAMPLES=${1:-12000} PMM_WARM_CASES=${2:-500} timeout 900 python -m pytest tests/test_gpu_property.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "property or vs_oracle or bit_exact or container or warm_seeds_randomised" > gpurun_out/r2z_soak.log 2>&1; echo "rc=$?" >> gpurun_out/r2z_soak.log
tail -n 2 gpurun_out/r2z_smoke.log; tail -n 3 gpurun_out/r2z_pytest.log; tail -n 1 gpurun_out/r2z_bench.err; tail -n 1 gpurun_out/r2z_ref.err; tail -n 4 gpurun_out/r2z_soak.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z_bench.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],2), round(d['value']), d['selfcheck']['exact'], 'e2e', round(d['e2e']['ms_per_step'],2), round(d['e2e']['value']), 'pinned', round(d['e2e']['pinned_inputs']['ms_per_step'],2), 'frac', round(d['roofline']['frac'],3), d['clocks']['power_w_max'])
r=json.loads(open('gpurun_out/r2z_ref.json').read().strip().splitlines()[-1]); print('ref', r.get('value'))
PY
