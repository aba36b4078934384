#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# one step eager under ncu: the first persistent launches are down1/2/3 forward; -s 3 skips the warm-up step's three
timeout 900 ncu --set full --clock-control none --import-source on -k regex:plane_conv_umma_persistent -s 13 -c 2 -o gpurun_out/c15_pers python bench.py --steps 1 --warmup 1 --no-graph --no-extras --no-cpu-baseline > gpurun_out/c15_ncu.log 2>&1
tail -2 gpurun_out/c15_ncu.log | cut -c1-300
ls -la gpurun_out/c15_pers.ncu-rep
