#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "warm_rows=4096 warm_rank=32" "warm_rows=2048 warm_rank=16" "warm_rows=2048 warm_rank=8" "warm_rows=1024 warm_rank=8" "warm_rows=1024 warm_rank=4" "warm_rows=512 warm_rank=4" "warm_rows=4096 warm_rank=8" "warm_seed=0"; do
opts=""; for kv in $cfg; do opts="$opts --opt $kv"; done
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --no-e2e $opts > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2s_bench.json').read().strip().splitlines()[-1])
p=d['roofline']['per_kernel_ms_per_step']
print('$cfg', round(d['ms_per_step'],2), d['selfcheck']['exact'], 'warm', round(p.get('tc_topk_warm',0),2), 'filter', round(p['tc_topk_f16r'],2), 'requery', p['requeried_f16_wide_per_step'])
PY
done
