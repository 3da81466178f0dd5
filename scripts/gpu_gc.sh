#!/bin/bash
# gemmconv: one CTA per SM (AB_GC_CTAS=1) vs two half-size CTAs (default), same box; plus the GPU tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 800 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -2
for c in 2 1 2 1; do
  AB_GC_CTAS=$c timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-also 2>>gpurun_out/bench.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ctas=$c', round(d['ms_per_step'],2), d['roofline']['classes'])"
done
for c in 2 1; do
  AB_GC_CTAS=$c timeout 600 python bench.py --workload bigvgan_base --steps 5 --warmup 3 --no-cpu-baseline --no-also 2>>gpurun_out/bench.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bigvgan_base ctas=$c', round(d['ms_per_step'],2), d['roofline']['classes'])"
done
