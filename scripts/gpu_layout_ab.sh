#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "1 0" "1 1" "0 0"; do
  set -- $cfg
  echo "=== AB_TC_LAYOUT=$1 AB_TC_BASE_OFFSET=$2 ==="
  AB_TC_LAYOUT=$1 AB_TC_BASE_OFFSET=$2 timeout 300 python -m pytest tests -m gpu -q -k "tc_conv1d" -p no:cacheprovider 2>&1 | tail -4
  AB_TC_LAYOUT=$1 AB_TC_BASE_OFFSET=$2 AB_TC_DEBUG_TIMING=1 timeout 300 python scripts/profile_forward.py 64 1024 tc_f16 1 2>&1 | grep tc_timing | awk 'NR%9==0'
done
