#!/bin/bash
# Quick iteration: tensor-core parity + phase timing + bench line.   env: PYTEST_K, BENCH_ARGS
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "${PYTEST_K:-tc_ or tensor_core or full_width or properties or transpose}" -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/pytest_quick.log
for ST in 0 2 4; do
echo "--- AB_TC_STAGGER=$ST"
AB_TC_STAGGER=$ST timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} 2>gpurun_out/bench_tc.err | tee gpurun_out/bench_tc_st$ST.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.3e samples/s  ms/step %.1f  e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value']))
print('dominant', r['kernel'], 'achieved %.1f TF/s frac %.3f share %.2f' % (r['achieved'], r['frac'], r['share_of_step']), r['classes'])
"
done
tail -3 gpurun_out/bench_tc.err
