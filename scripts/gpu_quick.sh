#!/bin/bash
# is the per-stage tcgen05.fence::after_thread_sync (after the TMA-completion barrier) the per-stage bubble?
mkdir -p gpurun_out
for V in 1 0; do
  echo "== AB_TC_STAGE_FENCE=$V"
  AB_TC_STAGE_FENCE=$V AB_TC_DEBUG_TIMING=1 timeout 300 python scripts/profile_forward.py 16 1024 tc_f16 1 2>&1 | grep tc_timing | grep -E "C=(256|128|32) k=(11) d=1 " | cut -c14-250
  AB_TC_STAGE_FENCE=$V timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('stage_fence $V', d['ms_per_step'], d['roofline']['classes']['tc_conv'])"
done 2>&1 | tee gpurun_out/tc_phase_timing_v10.txt
AB_TC_STAGE_FENCE=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tc_conv1d or generator_tensor_core or hifigan_v1 or config2" 2>&1 | tail -2
