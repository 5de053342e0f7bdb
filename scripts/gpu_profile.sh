#!/bin/bash
# Full GPU pass: parity tests, headline bench, ncu launch list + one full capture of the fused kernel.
mkdir -p gpurun_out
R=${1:-r1}
echo "== pytest -m gpu" | tee gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt
tail -n 8 gpurun_out/pytest_gpu.log | tee -a gpurun_out/summary.txt
echo "== tf32/bf16 cuBLAS peaks" | tee -a gpurun_out/summary.txt
timeout 120 python scripts/measure_peaks.py > gpurun_out/peaks_$R.json 2> gpurun_out/peaks.err; cat gpurun_out/peaks_$R.json | tee -a gpurun_out/summary.txt
echo "== bench (headline)" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench.err; echo "exit $?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench_$R.json | tee -a gpurun_out/summary.txt; tail -n 3 gpurun_out/bench.err
echo "== ncu launch list" | tee -a gpurun_out/summary.txt
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain_launches.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "exit $?" | tee -a gpurun_out/summary.txt
echo "== ncu full capture (mid-size)" | tee -a gpurun_out/summary.txt
CMD2="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --q 20000 --n 200000"
timeout 300 $CMD2 > gpurun_out/plain_full.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:tc_kernel<.bool.1" -s 1 -c 1 -f -o gpurun_out/tc_topk_$R $CMD2 > gpurun_out/ncu_full.log 2>&1
echo "exit $?" | tee -a gpurun_out/summary.txt
tail -n 3 gpurun_out/ncu_full.log | tee -a gpurun_out/summary.txt
echo "== ncu DRAM traffic of the fused kernel at the bench size" | tee -a gpurun_out/summary.txt
CMD3="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed --clock-control none --kernel-name-base demangled -k "regex:tc_kernel<.bool.1" -s 1 -c 1 --csv --log-file gpurun_out/traffic_$R.csv $CMD3 > gpurun_out/ncu_traffic.log 2>&1
echo "exit $?" | tee -a gpurun_out/summary.txt
cat gpurun_out/traffic_$R.csv | tail -n 8 | tee -a gpurun_out/summary.txt
