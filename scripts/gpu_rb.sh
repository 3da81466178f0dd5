#!/bin/bash
# First check of the persistent fused ResBlock kernel: targeted tests, per-launch timing, bench per plan.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== rb tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "fusion or fixture or full_width" 2>&1 | tail -15 | tee gpurun_out/pytest_rb.log
echo "=== rb timing (B=16)"
for MODE in 1 3; do AB_RB=$MODE AB_RB_DEBUG_TIMING=1 timeout 300 python scripts/profile_forward.py 16 1024 tc_f16 1 2>&1 | grep -E "rb_timing|ok|rror" | cut -c1-420; done | tee gpurun_out/rb_timing.txt
echo "=== bench per plan"
for MODE in 0 1 2 3; do echo "AB_RB=$MODE"; AB_RB=$MODE timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_rb$MODE.json | cut -c1-330; done
