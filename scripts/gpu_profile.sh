#!/bin/bash
# Profiling pass: phase timestamps of the tcgen05 kernel, ncu full capture of two launches, launch list.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AB_TC_DEBUG_TIMING=1 timeout 300 python scripts/profile_forward.py 64 1024 tc_f16 1 2> gpurun_out/tc_timing.txt | tail -1
grep tc_timing gpurun_out/tc_timing.txt
for IDX in 8 15; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s $IDX -c 1 \
     -f -o gpurun_out/prof_tc_$IDX python scripts/profile_forward.py 16 1024 tc_f16 1 > gpurun_out/ncu_$IDX.log 2>&1
  tail -2 gpurun_out/ncu_$IDX.log
done
AB_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/launches.csv
