#!/bin/bash
# Final pass of the round: smoke, whole GPU suite, default bench line, reference arm, launch list under ncu.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2f_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -rf -p no:cacheprovider > gpurun_out/r2f_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_pytest.log
( time timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err ) 2> gpurun_out/r2f_bench.time; echo "rc=$?" >> gpurun_out/r2f_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/r2f_ref.json 2> gpurun_out/r2f_ref.err; echo "rc=$?" >> gpurun_out/r2f_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2f_ncu.log 2>&1; echo "ncu rc=$?"
tail -n 3 gpurun_out/r2f_smoke.log; tail -n 4 gpurun_out/r2f_pytest.log; tail -n 2 gpurun_out/r2f_bench.err; cat gpurun_out/r2f_bench.time; tail -n 2 gpurun_out/r2f_ref.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],2), round(d['value']), d['selfcheck']['exact'], 'e2e', round(d['e2e']['ms_per_step'],2), round(d['e2e']['value']), 'pinned', round(d['e2e']['pinned_inputs']['ms_per_step'],2), 'frac', round(d['roofline']['frac'],3), d['clocks'])
print({k:round(v,3) for k,v in d['roofline']['per_kernel_ms_per_step'].items() if isinstance(v,float)})
for k,v in d['extra']['matmul'].items(): print(k, v['kernel'], round(v['kernel_ms'],4), round(v['prep_ms'],4), round(v['frac_hbm'],3), round(v['TFLOPs'],1), v.get('frac_fp64_cublas'))
for k,v in d['extra']['norms'].items(): print(k, round(v['kernel_ms'],3), round(v['frac_hbm'],3))
print(d['extra'].get('c1_e2e')); print({k:v for k,v in d['extra']['c5'].items() if k in ('ms_per_step','queries_per_sec','frac_of_bf16_sustained_per_gpu')}, d['extra']['c5']['selfcheck']['exact'])
print(d['extra'].get('f64_topk_2000x100000x256_cosine_k10'))
r=json.loads(open('gpurun_out/r2f_ref.json').read().strip().splitlines()[-1]); print('ref', r.get('value'), r.get('cpu_baseline'))
PY
