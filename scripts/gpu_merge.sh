#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -q -x -k "fixture or full_width or properties or plumbing or synthesis" -p no:cacheprovider 2>&1 | tail -2
for V in "AB_TC_MERGE=1" "AB_TC_MERGE=0"; do
echo "--- $V"
env $V timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.3e samples/s  ms/step %.1f launches %d' % (d['value'], d['ms_per_step'], d['gpu_launches']), {k:(v['launches'], round(v['ms'],1)) for k,v in r['classes'].items() if v['launches']})"
done
