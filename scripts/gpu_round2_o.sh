#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nproc; lscpu | grep -i "model name\|numa node(s)\|L3"
timeout 900 python scripts/e2e_sweep.py "-" "stage_nt=0" "stage_nt=1 stage_threads=12" "stage_nt=1 stage_threads=16" "stage_nt=0 stage_threads=16" "stage_nt=1 stage_threads=12 stage_slot_mb=16 stage_slots=6" "stage_nt=1 stage_threads=6" 2>&1 | tee gpurun_out/r2o_e2e_sweep.log | tail -12
