#!/bin/bash
# Round-2 evidence pass on one GPU: the default bench line, the ncu launch list of the same command, full captures of the
# kernels VERDICT r1 singled out (plane building, re-scoring, the retry launch) and of the dominant filter kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=r2
timeout 900 python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench exit $?"
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extras"
timeout 600 $CMD > gpurun_out/plain_$R.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv $CMD > gpurun_out/ncu_launches_$R.log 2>&1
echo "launch list exit $?"
CMD1="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras"
cap() {  # name, kernel regex, skip, title.  Summaries are made here: gpurun_out/ may carry 64 MiB back at most.
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/$1_$R $CMD1 > gpurun_out/ncu_$1_$R.log 2>&1
  echo "$1 exit $? $(tail -n 1 gpurun_out/ncu_$1_$R.log)"
  python scripts/ncu_summary.py gpurun_out/$1_$R.ncu-rep gpurun_out/ncu_$1_$R.md "$4" > /dev/null 2>&1
  ncu -i gpurun_out/$1_$R.ncu-rep --page raw --csv > gpurun_out/ncu_$1_${R}_raw.csv 2>/dev/null
  if [ $(stat -c %s gpurun_out/$1_$R.ncu-rep) -gt 12000000 ]; then rm -f gpurun_out/$1_$R.ncu-rep; fi
}
cap prep "prep_fast_kernel" 1 "prep_fast_kernel, C3 corpus (1M x 768 f32 -> f16-rounded plane + norms), round 2"
cap rescore "rescore_kernel" 0 "rescore_kernel, C3 (100k queries x 128 candidates x 768 f32), round 2"
cap retry "tc_kernel<.bool.1, .int.0, .int.8" 0 "seeded retry launch (f16-rounded filter, 256-entry lists, 69 queries x 1M corpus), round 2"
cap filter "tc_kernel<.bool.1, .int.0, .int.4" 1 "first-level filter tc_kernel (f16-rounded, KP=128) at the full C3 size, round 2"
ls -la gpurun_out/ | tail -20; du -sh gpurun_out
