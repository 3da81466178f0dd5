#!/bin/bash
# ncu --set full on one tc_conv launch (index $1, default 15) of a B=$2 (default 16) forward
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
IDX=${1:-15}; B=${2:-16}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s $IDX -c 1 \
   -f -o gpurun_out/prof_tc_v3_$IDX python scripts/profile_forward.py $B 1024 tc_f16 1 > gpurun_out/ncu_v3_$IDX.log 2>&1
tail -2 gpurun_out/ncu_v3_$IDX.log
