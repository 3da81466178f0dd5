#!/bin/bash
# A/B of two builds of the library on the same box: libamphion_b200.so (variant, current sources) vs libbase.so
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { for i in 1 2; do timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-also 2>>gpurun_out/bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],2), d['roofline']['classes'])"; done; }
run variant
cp amphion_b200/libamphion_b200.so amphion_b200/libvariant.so
cp amphion_b200/libbase.so amphion_b200/libamphion_b200.so
run base
cp amphion_b200/libvariant.so amphion_b200/libamphion_b200.so
run variant
