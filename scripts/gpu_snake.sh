#!/bin/bash
cd "$(dirname "$0")/.."
for V in "AB_SNAKE_OCC=3" "AB_SNAKE_OCC=4" "AB_SNAKE_WARP=1"; do
echo "--- $V"
env $V timeout 300 python -m pytest tests -m gpu -q -x -k "activation1d or (fixture and bigvgan)" -p no:cacheprovider 2>&1 | tail -1
env $V timeout 600 python bench.py --workload bigvgan_base --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('ms/step %.1f' % d['ms_per_step'], {k:(v['launches'], round(v['ms'],1)) for k,v in r['classes'].items() if v['launches']})"
done
