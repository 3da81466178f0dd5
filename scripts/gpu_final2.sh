#!/bin/bash
# Round-2 closing evidence (after the activation / gemmconv / conv_post changes): GPU tests as the driver runs them,
# smoke, bench (both arms), launch list, ncu --set full of the kernels that changed since scripts/gpu_final.sh ran.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > $O/r2_gpu.txt 2>&1
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -s -m gpu -p no:cacheprovider 2>&1 | grep -E "max\||full-size|apnet|passed|failed|error" | tee $O/r2_pytest_gpu.log | tail -12
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warn | tail -9 | tee $O/r2_smoke.log
echo "=== bench native"; timeout 900 python bench.py --steps 10 --warmup 3 2>$O/bench.err | tee $O/r2_bench_native.json | cut -c1-300
echo "=== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>>$O/bench.err | tee $O/r2_bench_reference.json | cut -c1-300
echo "=== apnet"; timeout 300 python scripts/bench_apnet.py 16 861 2>&1 | grep "^apnet" | tee $O/r2_bench_apnet.txt
echo "=== launch list"; AB_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
   --log-file $O/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-also > $O/bench_under_ncu.log 2>&1; tail -1 $O/r2_launches.csv | cut -c1-200
echo "=== ncu --set full"
cap() {  # name, kernel regex, skip, then the command
  local name="$1" re="$2" skip="$3"; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$re" -s "$skip" -c 1 -f -o $O/r2_$name "$@" > /dev/null 2>&1
}
cap gemmconv_convT1 gemmconv_kernel 2 python scripts/profile_forward.py 64 1024 tc_f16 1 hifigan_v1
cap conv_post conv_post 0 python scripts/profile_forward.py 64 1024 tc_f16 1 hifigan_v1
cap activation1d activation1d 40 python scripts/profile_forward.py 32 1024 tc_f16 1 bigvgan_base
ls -la $O/*.ncu-rep
