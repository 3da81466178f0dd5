#!/bin/bash
# final check of the unrolled MMA issue sequences: all GPU tests, smoke, bench, phase timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warn | tail -7 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench_native.json | cut -c1-200
AB_TC_DEBUG_TIMING=1 timeout 300 python scripts/profile_forward.py 16 1024 tc_f16 1 2>&1 | grep tc_timing | grep -E "C=(256|128|64|32) k=(3|11) d=1 " | cut -c14-230 | tee gpurun_out/tc_phase_timing_v11.txt
