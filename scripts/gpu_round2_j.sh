#!/bin/bash
# f16-split raw matmul: parity tests, error soak (default limit and forced up to D = 512), A/B timing against 3xTF32.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_property.py tests/test_gpu_bound.py -m gpu -q -rf -p no:cacheprovider -k "matmul or known_answers or smoke" > gpurun_out/r2j_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2j_pytest.log
timeout 300 python scripts/matmul_error_soak.py --seconds 120 > gpurun_out/r2j_matmul_soak.json 2> gpurun_out/r2j_matmul_soak.err; echo "rc=$?" >> gpurun_out/r2j_matmul_soak.err
timeout 300 python scripts/matmul_error_soak.py --seconds 50 --max-dim 512 --option matmul_tc_max_dim=512 > gpurun_out/r2j_matmul_soak512.json 2> gpurun_out/r2j_matmul_soak512.err
timeout 300 python scripts/matmul_ab.py > gpurun_out/r2j_matmul_ab.json 2> gpurun_out/r2j_matmul_ab.err; echo "rc=$?" >> gpurun_out/r2j_matmul_ab.err
tail -15 gpurun_out/r2j_pytest.log
for f in r2j_matmul_soak r2j_matmul_soak512; do python - <<PY
import json
d=json.load(open('gpurun_out/$f.json'))
print('$f', d['cases'], round(d['seconds']))
for k,v in d['buckets'].items(): print(' ', k, round(v['worst_ratio'],3), v['entries'])
for h in d['above_half_tolerance'][:6]: print(h)
PY
done
tail -n 3 gpurun_out/r2j_matmul_soak.err; tail -n 3 gpurun_out/r2j_matmul_ab.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_matmul_ab.json'))
for k,v in d.items():
    if isinstance(v,dict): print(k, {n:(round(x['kernel_ms'],4), round(x['prep_ms'],4), round(x['frac_hbm'],3), round(x['TFLOPs'],1)) for n,x in v.items()})
PY
