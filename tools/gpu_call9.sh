#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== gpu tests"
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
echo "=== gpu tests, fused colsum + compact + prefetch paths"
WUN_SPLIT_COLSUM=1 WUN_FOLD_COMPACT=1 WUN_FOLD_PREFETCH=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3
echo "=== bench"
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err
tail -2 gpurun_out/c9_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/c9_bench.json").read().strip().splitlines()[-1])
print("ms/step %.3f  e2e %.3e  families %s" % (d["ms_per_step"], d["e2e"]["value"], {k:round(v["us"]) for k,v in d.get("families",{}).items()}))
PY
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/c9_launches.csv python bench.py --steps 1 --warmup 1 --no-graph --no-extras --no-cpu-baseline > gpurun_out/c9_ncu.log 2>&1
tail -1 gpurun_out/c9_ncu.log | cut -c1-200
