#!/bin/bash
# Round-2 evidence call: parity tests, the full bench line, the ncu launch list of one step, ncu --set full of the
# heaviest launch groups (raw + per-kernel csv exported on the box).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== gpu tests"
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
echo "=== warm-up bench (discarded)"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > /dev/null 2>&1
echo "=== bench (full line)"
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -2 gpurun_out/r2_bench_n1.err
cp gpurun_out/layer_table_n1.json gpurun_out/r2_layer_table_n1.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_n1.json").read().strip().splitlines()[-1])
print("ms/step %.3f value %.3e e2e %.3e families %s top %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], {k:round(v["us"]) for k,v in d.get("families",{}).items()}, d["top_launch"]))
print("extras", json.dumps(d.get("extra_configs")))
print("cpu", d.get("cpu_baseline"))
PY
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r2_launches_bench_step.csv \
   python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extras --no-prewarm > gpurun_out/r2_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/r2_launches_bench_step.csv > gpurun_out/r2_launches_bench_step_summary.txt 2>&1
head -30 gpurun_out/r2_launches_bench_step_summary.txt
echo "=== ncu --set full"
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -c 16 -o gpurun_out/r2_full -f \
   python tools/profile_passes.py ${PASSES:-1:2 3:2 1:1 3:0 8:0 24:1} > gpurun_out/r2_ncu_full.log 2>&1
tail -8 gpurun_out/r2_ncu_full.log
ncu -i gpurun_out/r2_full.ncu-rep --page raw --csv > gpurun_out/r2_full_raw.csv 2>/dev/null
ls -la gpurun_out | head -30
