#!/bin/bash
# End-of-round-2 captures: the warm-seeded first-level filter at the full C3 size, its sample pre-pass, and the f16-split
# matmul kernel (resident query planes, eight epilogue warps) at 16384 x 65536 x 256.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=r2e
CMD1="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras"
cap() {  # name, kernel regex, skip, title, command
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/$1_$R $5 > gpurun_out/ncu_$1_$R.log 2>&1
  echo "$1 exit $? $(tail -n 1 gpurun_out/ncu_$1_$R.log)"
  python scripts/ncu_summary.py gpurun_out/$1_$R.ncu-rep gpurun_out/ncu_$1_$R.md "$4" > /dev/null 2>&1
  if [ $(stat -c %s gpurun_out/$1_$R.ncu-rep) -gt 12000000 ]; then rm -f gpurun_out/$1_$R.ncu-rep; fi
}
timeout 300 $CMD1 > gpurun_out/plain_$R.log 2>&1 || { echo plain failed; exit 1; }
cap filter "tc_kernel<.bool.1, .int.0, .int.4" 1 "first-level filter tc_kernel (f16-rounded, KP=128, started from warm seeds) at the full C3 size, end of round 2" "$CMD1"
cap warm "tc_kernel<.bool.1, .int.0, .int.1" 1 "warm-seed sample pre-pass (same kernel, 32-entry lists, 16 strided corpus tiles) at C3, end of round 2" "$CMD1"
cap matmul "tc_kernel<.bool.1, .int.1, .int.1, .int.128, .int.2, .int.2" 2 "raw f32 matmul 16384 x 65536 x 256, hi/lo f16 split, resident query planes, eight epilogue warps, end of round 2" "python scripts/profile_matmul.py 16384 65536 256 3"
grep -h "gpu__time_duration.sum\|dram__bytes_read.sum \|dram__bytes_write.sum \|pipe_tensor_cycles_active\|lts__t_sector_hit" gpurun_out/ncu_filter_$R.md gpurun_out/ncu_warm_$R.md gpurun_out/ncu_matmul_$R.md
du -sh gpurun_out
