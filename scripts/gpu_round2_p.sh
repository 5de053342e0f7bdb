#!/bin/bash
# N-GPU sanity after the matmul / host-path changes: sharded parity script, single-process group test, bench at N.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
G=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29521 scripts/check_sharded_gpu.py > gpurun_out/r2p_check.log 2>&1; echo "rc=$?" >> gpurun_out/r2p_check.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -rf -p no:cacheprovider -k "single_process_group or matmul" > gpurun_out/r2p_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2p_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $G --steps 3 --warmup 3 > gpurun_out/r2p_bench$G.json 2> gpurun_out/r2p_bench$G.err; echo "rc=$?" >> gpurun_out/r2p_bench$G.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29523 bench.py --impl reference --gpus $G --steps 2 --warmup 1 > gpurun_out/r2p_ref$G.json 2> gpurun_out/r2p_ref$G.err; echo "rc=$?" >> gpurun_out/r2p_ref$G.err
tail -n 6 gpurun_out/r2p_check.log; tail -n 4 gpurun_out/r2p_pytest.log; tail -n 3 gpurun_out/r2p_bench$G.err; tail -n 2 gpurun_out/r2p_ref$G.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2p_bench$G.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['ms_per_step'],2), round(d['value']), d['selfcheck']['exact'], 'e2e', round(d['e2e']['value']))
x=d.get('extra',{})
for k in ('c4_strong','c5'):
    v=x.get(k)
    if v: print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms_per_step','queries_per_sec','frac_of_bf16_sustained_per_gpu','error')}, v.get('selfcheck',{}).get('exact'))
r=json.loads(open('gpurun_out/r2p_ref$G.json').read().strip().splitlines()[-1])
print('ref', r.get('value'), r.get('cpu_baseline',{}).get('cores'))
PY
