#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "bigvgan or activation1d or tensor_core or full_width" -p no:cacheprovider 2>&1 | tail -5
for P in tc_f16 fp32; do
echo "--- bigvgan_base $P"
timeout 900 python bench.py --workload bigvgan_base --steps 3 --warmup 3 --precision $P --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_bigvgan_$P.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.3e samples/s  ms/step %.1f  e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value']))
print('dominant', r['kernel'], r['bound'], 'achieved %.1f %s frac %.3f share %.2f' % (r['achieved'], r['unit'], r['frac'], r['share_of_step']), r['classes'])"
done
