#!/bin/bash
# Timing experiments on the persistent ResBlock kernel (debug knobs; results of the SKIP runs are wrong by design).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # label, env...
  local label="$1"; shift
  for MODE in ${MODES:-3 1}; do
    env AB_RB=$MODE AB_RB_DEBUG_TIMING=1 "$@" timeout 300 python scripts/profile_forward.py 16 1024 tc_f16 1 2>&1 \
      | grep -E "rb_timing" | grep -E "C=(128|64|32) k=(3|7|11) d=1," | sed "s/^/[$label mode$MODE] /" | cut -c1-400
  done
}
{
MODES="4 3" run base
MODES=4 run skip_all AB_RB_DEBUG_SKIP=15
} | tee gpurun_out/rb_exp.txt
echo "=== tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "fusion or fixture or mel" 2>&1 | tail -3
echo "=== bench per plan"
for MODE in 0 2 3 4; do echo "AB_RB=$MODE"; AB_RB=$MODE timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-also 2>>gpurun_out/bench.err | tee gpurun_out/bench_rb$MODE.json | cut -c1-200; done
