#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for V in "AB_TC_WS=1 AB_TC_DUAL=1" "AB_TC_WS=1 AB_TC_DUAL=0"; do
  echo "=== $V"
  env $V timeout 300 python -m pytest tests -m gpu -q -k "tc_conv1d or tensor_core or full_width" -p no:cacheprovider 2>&1 | tail -5
  env $V AB_TC_DEBUG_TIMING=1 timeout 300 python scripts/profile_forward.py 64 1024 tc_f16 1 2>&1 | grep tc_timing | awk 'NR%9==0'
  env $V timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.3e samples/s  ms/step %.1f' % (d['value'], d['ms_per_step']), r['classes']['tc_conv'])"
done
