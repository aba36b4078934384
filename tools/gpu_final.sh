#!/bin/bash
# Last evidence call of the round (3.7 GPU-minutes were left): the full bench line and the ncu launch list of one step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 150 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
cp gpurun_out/layer_table_n1.json gpurun_out/r2_layer_table_n1.json 2>/dev/null
tail -c 600 gpurun_out/r2_bench_n1.json
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 760 --csv --log-file gpurun_out/r2_launches_bench_step.csv \
   python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-extras --no-prewarm > gpurun_out/r2_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/r2_launches_bench_step.csv > gpurun_out/r2_launches_bench_step_summary.txt 2>&1
head -24 gpurun_out/r2_launches_bench_step_summary.txt
