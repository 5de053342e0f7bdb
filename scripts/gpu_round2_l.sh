#!/bin/bash
# ncu: DRAM traffic + full capture of the f16-split matmul kernel at 16384 x 65536 x 256 and x 64.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for D in 256; do
CMD="python scripts/profile_matmul.py 16384 65536 $D 3"
timeout 200 $CMD > gpurun_out/r2l_plain_$D.log 2>&1 || { echo plain failed; tail gpurun_out/r2l_plain_$D.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:tc_kernel" -s 2 -c 1 -f -o gpurun_out/matmul_f16x3_d$D $CMD > gpurun_out/r2l_ncu_$D.log 2>&1
echo "ncu d=$D exit $?"
ncu -i gpurun_out/matmul_f16x3_d$D.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; r=rows[2]
for k in ('gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed','l1tex__m_xbar2l1tex_read_bytes.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio'):
    for i,x in enumerate(h):
        if x==k or x.endswith('.'+k): print(k, r[i]); break
"
done
