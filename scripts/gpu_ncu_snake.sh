#!/bin/bash
# ncu --set full on one activation1d launch (index $1, default 40) of a BigVGAN-base forward
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
IDX=${1:-40}
AB_BENCH_PROFILE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:activation1d -s $IDX -c 1 \
   -f -o gpurun_out/prof_snake_$IDX python bench.py --workload bigvgan_base --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_snake_$IDX.log 2>&1
tail -2 gpurun_out/ncu_snake_$IDX.log
