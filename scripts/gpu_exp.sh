#!/bin/bash
# Experiment runner: bench.py (resident only) under a list of option sets; prints one line per set.
# usage: gpu_exp.sh tag "opts1" "opts2" ...   (each opts = space separated key=value, or "-" for defaults)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=$1; shift
i=0
for o in "$@"; do
  args=""
  if [ "$o" != "-" ]; then for kv in $o; do args="$args --opt $kv"; done; fi
  timeout 600 python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline --no-extras --no-e2e ${EXTRA_ARGS:-} $args > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${tag}_$i.json'))
    print('[$o]', round(d['ms_per_step'],2), round(d['value']), d['selfcheck']['exact'], {k:round(v,2) for k,v in d['roofline']['per_kernel_ms_per_step'].items() if isinstance(v,float) and v>0.05}, d['clocks']['sm_mhz'], d['clocks']['power_w_max'])
except Exception as e:
    print('[$o] FAILED', e); print(open('gpurun_out/${tag}_$i.err').read()[-800:])
PY
  i=$((i+1))
done
