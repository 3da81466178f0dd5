#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for V in "AB_TC_WW=4" "AB_TC_WW=8"; do
echo "--- hifigan_v1 $V"
env $V timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.3e samples/s  ms/step %.1f' % (d['value'], d['ms_per_step']), {k:(v['launches'], round(v['ms'],1)) for k,v in r['classes'].items() if v['launches']})"
done
AB_TC_WW=8 AB_TC_DEBUG_TIMING=1 timeout 300 python scripts/profile_forward.py 64 1024 tc_f16 1 2>&1 | grep tc_timing | awk 'NR%9==0' | sed 's/nconv=2 //; s/img=1 staged=1 //' | cut -c1-200
for W in bigvgan_base bigvgan_large; do
echo "--- $W"
timeout 900 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_$W.err | tee gpurun_out/bench_$W.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.3e samples/s  ms/step %.1f  e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value']))
print('dominant', r['kernel'], r['bound'], 'achieved %.1f %s frac %.3f share %.2f' % (r['achieved'], r['unit'], r['frac'], r['share_of_step']), {k:(v['launches'], round(v['ms'],1)) for k,v in r['classes'].items() if v['launches']})"
tail -2 gpurun_out/bench_$W.err
done
python scripts/bench_mel.py 2>&1 | tail -2 | tee gpurun_out/bench_mel.json
