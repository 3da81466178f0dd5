#!/bin/bash
cd "$(dirname "$0")/.."
for N in 3 2; do
echo "=== AB_TC_NCTA=$N"
AB_TC_NCTA=$N timeout 600 python -m pytest tests -m gpu -q -x -k "tc_conv1d or tensor_core or full_width or properties" -p no:cacheprovider 2>&1 | tail -2
AB_TC_NCTA=$N AB_TC_DEBUG_TIMING=1 timeout 300 python scripts/profile_forward.py 64 1024 tc_f16 1 2>&1 | grep tc_timing | awk 'NR%3==0' | sed 's/nconv=2 //; s/img=1 staged=1 //' | cut -c1-200
AB_TC_NCTA=$N timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.3e samples/s  ms/step %.1f' % (d['value'], d['ms_per_step']), r['classes']['tc_conv'], r['classes']['tc_gemmconv'])"
done
