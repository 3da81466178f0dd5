#!/bin/bash
cd "$(dirname "$0")/.."
for SK in 0 1 2 3 4 7; do
  echo "=== AB_TC_DEBUG_SKIP=$SK"
  AB_TC_DEBUG_SKIP=$SK AB_TC_DEBUG_TIMING=1 timeout 300 python scripts/profile_forward.py 64 1024 tc_f16 1 2>&1 | grep tc_timing | awk 'NR==9 || NR==17 || NR==18 || NR==27 || NR==36' | sed 's/.*| cycles//'
done
