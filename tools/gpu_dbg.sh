#!/bin/bash
cd "$(dirname "$0")/.."
for cfg in "X=0" "WUN_EPI_TEAMS=1" "WUN_LIB=wave-u-net_b200/libwun_prev.so"; do
  echo "=== $cfg"
  env $cfg timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "batch16 or forced" --tb=short 2>&1 | tail -25
done
