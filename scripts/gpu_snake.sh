#!/bin/bash
# Activation1d kernel iteration: parity subset + BigVGAN-base bench; DRAM traffic of every tc_conv launch of one config-2 forward
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bigvgan or activation1d or generator" 2>&1 | tail -3 | tee gpurun_out/snake_pytest.log
timeout 600 python bench.py --workload bigvgan_base --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/snake_bigvgan_base.json
python - <<PY
import json
d = json.loads(open("gpurun_out/snake_bigvgan_base.json").read())
print("bigvgan_base", d["ms_per_step"], d["value"], d["roofline"]["classes"])
PY
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:tc_conv_kernel -c 36 --csv \
  --log-file gpurun_out/tc_traffic.csv python scripts/profile_forward.py 64 1024 tc_f16 1 > gpurun_out/tc_traffic.log 2>&1
tail -3 gpurun_out/tc_traffic.csv | cut -c1-250
