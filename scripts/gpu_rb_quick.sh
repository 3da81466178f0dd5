#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for MODE in 3 1; do
  AB_RB=$MODE AB_RB_DEBUG_TIMING=1 timeout 300 python scripts/profile_forward.py 16 1024 tc_f16 1 2>&1 | grep -E "rb_timing" | grep -E "C=(128|64|32) k=(3|7|11) d=1," | sed "s/^/[mode$MODE] /" | cut -c1-400
done | tee gpurun_out/rb_quick.txt
for MODE in 2; do echo "AB_RB=$MODE"; AB_RB=$MODE timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-also 2>>gpurun_out/bench.err | tee gpurun_out/bench_rbq$MODE.json | cut -c1-200; done
