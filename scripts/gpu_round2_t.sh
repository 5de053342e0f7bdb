#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bound.py -m gpu -q -rf -p no:cacheprovider -x > gpurun_out/r2t_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2t_pytest.log
tail -n 6 gpurun_out/r2t_pytest.log
for w in 1 0; do
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --opt warm_seed=$w > gpurun_out/r2t_bench_w$w.json 2> gpurun_out/r2t_bench_w$w.err; echo "rc=$?" >> gpurun_out/r2t_bench_w$w.err
tail -n 1 gpurun_out/r2t_bench_w$w.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2t_bench_w$w.json').read().strip().splitlines()[-1])
p=d['roofline']['per_kernel_ms_per_step']
print('warm_seed=$w', round(d['ms_per_step'],2), d['selfcheck']['exact'], 'warm', round(p.get('tc_topk_warm',0),2), 'filter', round(p['tc_topk_f16r'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'pinned', round(d['e2e']['pinned_inputs']['ms_per_step'],2), 'frac', round(d['roofline']['frac'],3))
c5=d['extra']['c5']; print(' c5', round(c5['ms_per_step'],2), c5['selfcheck']['exact'], round(c5['filter_kernel_ms_rank0'],2))
print(' f64', d['extra']['f64_topk_2000x100000x256_cosine_k10']['step_ms'], 'c1', d['extra']['c1_e2e']['array_ms_median'])
PY
done
