#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -rf -p no:cacheprovider -k "matmul" > gpurun_out/r2k_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_pytest.log
timeout 300 python scripts/matmul_ab.py > gpurun_out/r2k_matmul_ab.json 2> gpurun_out/r2k_matmul_ab.err; echo "rc=$?" >> gpurun_out/r2k_matmul_ab.err
tail -n 4 gpurun_out/r2k_pytest.log; tail -n 3 gpurun_out/r2k_matmul_ab.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2k_matmul_ab.json'))
for k,v in d.items():
    if isinstance(v,dict): print(k, {n:(x['kernel'], round(x['kernel_ms'],4), round(x['prep_ms'],4), round(x['frac_hbm'],3), round(x['TFLOPs'],1)) for n,x in v.items()})
PY
