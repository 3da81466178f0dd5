#!/bin/bash
# Full validation: every GPU test (as the driver runs them), smoke, bench (both arms), launch list, ncu capture.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warn | tail -8 | tee gpurun_out/smoke.log
echo "=== bench native"; timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench_native.json | cut -c1-400
echo "=== bench reference"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/bench.err | tee gpurun_out/bench_reference.json | cut -c1-300
echo "=== bench BigVGAN-base (config 3)"; timeout 600 python bench.py --workload bigvgan_base --steps 5 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_bigvgan_base.json | cut -c1-300
echo "=== mel (config 4)"; timeout 300 python scripts/bench_mel.py 2>&1 | tail -3 | tee gpurun_out/bench_mel.log
echo "=== launch list"; AB_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
   --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -2 gpurun_out/launches.csv | cut -c1-200
echo "=== ncu full (stage-1 k=11 pair, stage-0 k=11 pair, stage-1 convT)"
for IDX in 8 15; do timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s $IDX -c 1 -f -o gpurun_out/r1_tc_conv_$IDX python scripts/profile_forward.py 16 1024 tc_f16 1 > /dev/null 2>&1; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemmconv_kernel -s 1 -c 1 -f -o gpurun_out/r1_gemmconv_1 python scripts/profile_forward.py 16 1024 tc_f16 1 > /dev/null 2>&1
AB_BENCH_PROFILE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:activation1d -s 40 -c 1 -f -o gpurun_out/r1_activation1d_40 python bench.py --workload bigvgan_base --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
