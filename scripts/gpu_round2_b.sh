#!/bin/bash
# Multi-GPU pass (gpurun --gpus 2): the single-process group, the one-rank-per-process group under torchrun, the 2-GPU
# bench line and its reference arm.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2b_gpu.txt 2>&1; nvidia-smi topo -m >> gpurun_out/r2b_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -rf -p no:cacheprovider -k "single_process_group or seeded or long_vectors or infinite or host_shard" > gpurun_out/r2b_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_pytest.log
timeout 600 python -m pytest tests/test_gpu_property.py -m gpu -q -rf -p no:cacheprovider -k "matmul" > gpurun_out/r2b_prop.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_prop.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_sharded_gpu.py > gpurun_out/r2b_check.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2b_bench2.json 2> gpurun_out/r2b_bench2.err; echo "rc=$?" >> gpurun_out/r2b_bench2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2b_ref2.json 2> gpurun_out/r2b_ref2.err; echo "rc=$?" >> gpurun_out/r2b_ref2.err
tail -12 gpurun_out/r2b_pytest.log; tail -6 gpurun_out/r2b_prop.log; tail -8 gpurun_out/r2b_check.log; tail -4 gpurun_out/r2b_bench2.err; head -c 600 gpurun_out/r2b_bench2.json; echo; head -c 400 gpurun_out/r2b_ref2.json
