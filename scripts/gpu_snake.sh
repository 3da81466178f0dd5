#!/bin/bash
# Activation1d kernel iteration: parity subset + BigVGAN-base bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bigvgan or activation1d or generator" 2>&1 | tail -5 | tee gpurun_out/snake_pytest.log
timeout 600 python bench.py --workload bigvgan_base --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/snake_bigvgan_base.json
python - <<PY
import json
d = json.loads(open("gpurun_out/snake_bigvgan_base.json").read())
print("bigvgan_base", d["ms_per_step"], d["value"], d["roofline"]["classes"])
PY
