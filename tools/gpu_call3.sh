#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== cluster probe"; tools/cluster_probe
echo "=== adam test"
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k adam_step_matches 2>&1 | grep -v "^$" | head -60
echo "=== all gpu tests (no -x)"
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8
echo "=== bench"
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
tail -c 600 gpurun_out/c3_bench.json | head -c 300; echo
cp gpurun_out/layer_table_n1.json gpurun_out/c3_layers.json
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/c3_launches.csv python bench.py --steps 1 --warmup 1 --no-graph --no-extras --no-cpu-baseline > gpurun_out/c3_ncu.log 2>&1
tail -2 gpurun_out/c3_ncu.log
echo "=== ncu full: one fold launch (down11 fwd-like)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:plane_conv_umma_fold -s 4 -c 2 -o gpurun_out/c3_fold python bench.py --steps 1 --warmup 1 --no-graph --no-extras --no-cpu-baseline > gpurun_out/c3_ncu2.log 2>&1
tail -2 gpurun_out/c3_ncu2.log
