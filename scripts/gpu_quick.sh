#!/bin/bash
# how long does the MMA issuer wait for weight stages?  (dual vs one CTA per SM)
mkdir -p gpurun_out
for V in "AB_TC_DUAL=1" "AB_TC_DUAL=0"; do
  echo "== $V"
  env $V AB_TC_DEBUG_TIMING=1 timeout 300 python scripts/profile_forward.py 16 1024 tc_f16 1 2>&1 | grep tc_timing | grep -E "C=(256|128|64|32) k=(3|11) d=1 " | cut -c14-330
done 2>&1 | tee gpurun_out/tc_phase_timing_v7.txt
