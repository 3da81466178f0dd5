#!/bin/bash
# One gpurun call: parity tests (split so a trapped kernel cannot poison the rest), smoke, bench.
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [quick]'
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu.txt
SAFE='fp32 or transpose or activation1d or plumbing or mel or tacotron'
echo "=== A: CUDA-core kernels ===";
timeout 900 python -m pytest tests -m gpu -q -k "$SAFE" -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/pytest_A.log
echo "=== B: tcgen05 conv unit tests ===";
timeout 600 python -m pytest tests -m gpu -q -k "tc_conv1d" -p no:cacheprovider 2>&1 | tail -60 | tee gpurun_out/pytest_B.log
if grep -q failed gpurun_out/pytest_B.log; then
  echo "=== B': same with LBO/SBO swapped ===";
  AB_TC_SWAP_LBO_SBO=1 timeout 600 python -m pytest tests -m gpu -q -k "tc_conv1d" -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/pytest_Bswap.log
fi
echo "=== C: generators on tensor cores, properties ===";
timeout 900 python -m pytest tests -m gpu -q -k "not ($SAFE) and not tc_conv1d" -p no:cacheprovider 2>&1 | tail -60 | tee gpurun_out/pytest_C.log
echo "=== smoke ===";
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12 | tee gpurun_out/smoke.log
echo "=== bench fp32 ===";
timeout 600 python bench.py --steps 2 --warmup 3 --precision fp32 --no-cpu-baseline 2>gpurun_out/bench_fp32.err | tee gpurun_out/bench_fp32.json
echo "=== bench tc_f16 ===";
timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_tc.err | tee gpurun_out/bench_tc.json
tail -5 gpurun_out/bench_tc.err
