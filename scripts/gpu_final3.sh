#!/bin/bash
# closing numbers of the final build: GPU tests, both bench arms, conv_post capture
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -s -m gpu -p no:cacheprovider 2>&1 | grep -E "max\||full-size|apnet|passed|failed|error" | tee $O/r2_pytest_gpu.log | tail -3
echo "=== bench native"; timeout 900 python bench.py --steps 10 --warmup 3 2>$O/bench.err | tee $O/r2_bench_native.json | cut -c1-200
echo "=== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>>$O/bench.err | tee $O/r2_bench_reference.json | cut -c1-200
echo "=== launch list"; AB_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
   --log-file $O/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-also > $O/bench_under_ncu.log 2>&1; tail -1 $O/r2_launches.csv | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv_post" -s 0 -c 1 -f -o $O/r2_conv_post python scripts/profile_forward.py 64 1024 tc_f16 1 hifigan_v1 > /dev/null 2>&1
ls -la $O/r2_conv_post.ncu-rep
