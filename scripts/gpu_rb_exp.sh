#!/bin/bash
# Timing experiments on the persistent ResBlock kernel (debug knobs; results of the SKIP runs are wrong by design).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # label, env...
  local label="$1"; shift
  for MODE in 3 1; do
    env AB_RB=$MODE AB_RB_DEBUG_TIMING=1 "$@" timeout 300 python scripts/profile_forward.py 16 1024 tc_f16 1 2>&1 \
      | grep -E "rb_timing" | grep -E "C=(128|32) k=(3|11) d=1," | sed "s/^/[$label mode$MODE] /" | cut -c1-400
  done
}
{
run base
run poll AB_RB_POLL=1
run cps1 AB_RB_CPS=1
run cps8 AB_RB_CPS=8
run skip_ld AB_RB_DEBUG_SKIP=1
run skip_st AB_RB_DEBUG_SKIP=2
run skip_tmem AB_RB_DEBUG_SKIP=4
run skip_smem AB_RB_DEBUG_SKIP=8
run skip_all AB_RB_DEBUG_SKIP=15
} | tee gpurun_out/rb_exp.txt
