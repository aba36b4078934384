#!/bin/bash
cd "$(dirname "$0")/.."
bash tools/gpu_ab.sh masks "X=0" "WUN_SIGN_MASK=0" "WUN_OUT_FUSE=0" "WUN_OUT_FUSE=0 WUN_SPLIT_COLSUM=1" "WUN_OUT_FUSE=0 WUN_SPLIT_COLSUM=2"
echo "=== ncu launch list (default build: fused output)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/masks_launches.csv \
   python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/masks_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/masks_launches.csv 1 > gpurun_out/masks_launches_summary.txt 2>&1
head -28 gpurun_out/masks_launches_summary.txt
grep -n "persistent" gpurun_out/masks_launches_summary.txt | head -40
