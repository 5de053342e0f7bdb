#!/bin/bash
# 8-GPU pass: one-rank-per-process group under torchrun (parity incl. empty shards), the single-process group, the 8-GPU
# bench line (weak scaling + C4 as stated), one call over the whole box.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
G=${1:-8}
nvidia-smi topo -m > gpurun_out/r2g_topo.txt 2>&1; nproc >> gpurun_out/r2g_topo.txt; free -g >> gpurun_out/r2g_topo.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29521 scripts/check_sharded_gpu.py > gpurun_out/r2g_check.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_check.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -rf -p no:cacheprovider -k "single_process_group" > gpurun_out/r2g_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $G --steps 3 --warmup 2 > gpurun_out/r2g_bench$G.json 2> gpurun_out/r2g_bench$G.err; echo "rc=$?" >> gpurun_out/r2g_bench$G.err
timeout 600 python scripts/single_process_multi_gpu.py > gpurun_out/r2g_single_process.json 2> gpurun_out/r2g_single_process.err; echo "rc=$?" >> gpurun_out/r2g_single_process.err
tail -9 gpurun_out/r2g_check.log; tail -4 gpurun_out/r2g_pytest.log; tail -3 gpurun_out/r2g_bench$G.err; head -c 300 gpurun_out/r2g_bench$G.json; echo; cat gpurun_out/r2g_single_process.json; tail -3 gpurun_out/r2g_single_process.err
