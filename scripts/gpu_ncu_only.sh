#!/bin/bash
# ncu passes only (full capture of the first-level kernel at mid size + DRAM traffic at the bench size).
mkdir -p gpurun_out
R=${1:-r1}
CMD2="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --q 20000 --n 200000"
timeout 300 $CMD2 > gpurun_out/plain_full.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:tc_kernel<.bool.1" -s 1 -c 1 -f -o gpurun_out/tc_topk_$R $CMD2 > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
CMD3="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none --kernel-name-base demangled -k "regex:tc_kernel<.bool.1" -s 1 -c 1 --csv --log-file gpurun_out/traffic_$R.csv $CMD3 > gpurun_out/ncu_traffic.log 2>&1
echo "traffic exit $?"
tail -n 6 gpurun_out/traffic_$R.csv
