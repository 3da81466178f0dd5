#!/bin/bash
# Round-2 evidence: every GPU test (as the driver runs them), smoke, bench (both arms), launch list, DRAM traffic of
# every kernel class, ncu --set full captures of each kernel on a BASELINE config at the bench batch, phase counters.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > $O/r2_gpu.txt 2>&1
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -s -m gpu -p no:cacheprovider 2>&1 | grep -E "max\||full-size|passed|failed|error" | tee $O/r2_pytest_gpu.log | tail -25
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warn | tail -9 | tee $O/r2_smoke.log
echo "=== bench native"; timeout 900 python bench.py --steps 10 --warmup 3 2>$O/bench.err | tee $O/r2_bench_native.json | cut -c1-300
echo "=== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>>$O/bench.err | tee $O/r2_bench_reference.json | cut -c1-300
echo "=== rb phase counters (B=64)"; AB_RB_DEBUG_TIMING=1 timeout 300 python scripts/profile_forward.py 64 1024 tc_f16 1 2>&1 | grep rb_timing | cut -c1-420 > $O/r2_rb_timing.txt; wc -l $O/r2_rb_timing.txt
echo "=== launch list"; AB_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
   --log-file $O/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-also > $O/bench_under_ncu.log 2>&1; tail -1 $O/r2_launches.csv | cut -c1-200
echo "=== DRAM traffic per kernel class"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 900 ncu --metrics $M --clock-control none --csv --log-file $O/traffic_v1.csv python scripts/profile_forward.py 64 1024 tc_f16 1 hifigan_v1 > /dev/null 2>&1
timeout 900 ncu --metrics $M --clock-control none --csv --log-file $O/traffic_bvb.csv python scripts/profile_forward.py 32 1024 tc_f16 1 bigvgan_base > /dev/null 2>&1
timeout 900 ncu --metrics $M --clock-control none --csv --log-file $O/traffic_mel.csv python scripts/profile_forward.py 64 220500 fp32 1 mel > /dev/null 2>&1
ls -la $O/traffic_*.csv
echo "=== ncu --set full"
cap() {  # name, kernel regex, skip, then the command
  local name="$1" re="$2" skip="$3"; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$re" -s "$skip" -c 1 -f -o $O/r2_$name "$@" > /dev/null 2>&1
}
cap rb_fused_c32 rb_kernel 12 python scripts/profile_forward.py 64 1024 tc_f16 1 hifigan_v1        # stage 3, k=3 block (fused)
cap rb_pair_c128_k11 rb_kernel 3 python scripts/profile_forward.py 64 1024 tc_f16 1 hifigan_v1     # stage 1, k=11 pair
cap tc_conv_c256_k11 tc_conv_kernel 8 python scripts/profile_forward.py 64 1024 tc_f16 1 hifigan_v1
cap gemmconv_convT1 gemmconv_kernel 2 python scripts/profile_forward.py 64 1024 tc_f16 1 hifigan_v1
cap conv_post conv_post 0 python scripts/profile_forward.py 64 1024 tc_f16 1 hifigan_v1
cap activation1d activation1d 40 python scripts/profile_forward.py 32 1024 tc_f16 1 bigvgan_base
cap gemmconv_stream gemmconv_stream 4 python scripts/profile_forward.py 8 2048 tc_f16 1 bigvgan_large
cap mel_fused mel_fused 0 python scripts/profile_forward.py 64 220500 fp32 1 mel
ls -la $O/*.ncu-rep
