#!/bin/bash
# Activation1d variants on one box (libvar_*.so built with -DAB_SNAKE_BLOCKS / -DAB_SNAKE_AHEAD) and an ncu capture
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
cp amphion_b200/libamphion_b200.so amphion_b200/libvar_cur.so
run() {
  cp amphion_b200/libvar_$1.so amphion_b200/libamphion_b200.so
  timeout 600 python bench.py --workload bigvgan_base --steps 5 --warmup 3 --no-cpu-baseline --no-also 2>>$O/bench.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],2), d['roofline']['classes']['activation1d'])"
}
for v in ${VARIANTS:-cur}; do run $v; done
cp amphion_b200/libvar_cur.so amphion_b200/libamphion_b200.so
if [ -n "${NCU:-}" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:activation1d" -s 40 -c 1 -f -o $O/r2_activation1d_v2 \
  python scripts/profile_forward.py 32 1024 tc_f16 1 bigvgan_base > /dev/null 2>&1
ls -la $O/*.ncu-rep
fi
