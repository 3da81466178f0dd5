#!/bin/bash
# Quick iteration: tensor-core parity + phase timing + bench line.
# env: PYTEST_K (test filter), BENCH_ARGS, VARIANTS (space separated "ENV=VAL" settings to A/B, default one run)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "${PYTEST_K:-tc_ or tensor_core or full_width or properties or transpose}" -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/pytest_quick.log
AB_TC_DEBUG_TIMING=1 timeout 300 python scripts/profile_forward.py 64 1024 tc_f16 1 2> gpurun_out/tc_timing.txt | tail -1
grep -E "tc_timing" gpurun_out/tc_timing.txt | awk 'NR%3==0'
i=0
for V in ${VARIANTS:-DEFAULT=1}; do
echo "--- $V"
env $V timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} 2>gpurun_out/bench_tc.err | tee gpurun_out/bench_tc_$i.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.3e samples/s  ms/step %.1f  e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value']))
print('dominant', r['kernel'], 'achieved %.1f TF/s frac %.3f share %.2f' % (r['achieved'], r['frac'], r['share_of_step']), r['classes'])
print('clocks', d['clocks'])
"
i=$((i+1))
done
tail -3 gpurun_out/bench_tc.err
