#!/bin/bash
# A/B helper for one GPU call: parity tests, then bench.py per environment setting, each leaving its per-layer table.
#   gpurun --timeout 900 -- 'bash tools/gpu_ab.sh TAG "X=0" "WUN_FOLD=0" > gpurun_out/TAG.log 2>&1; tail -40 gpurun_out/TAG.log'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=$1; shift
echo "=== parity tests"
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "=== warm-up bench (discarded: the first run on a fresh box is slower)"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > /dev/null 2>&1
i=0
for cfg in "$@"; do
  echo "=== bench $cfg"
  env $cfg timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_bench_$i.json 2> gpurun_out/${TAG}_bench_$i.err
  tail -3 gpurun_out/${TAG}_bench_$i.err
  cp gpurun_out/layer_table_n1.json gpurun_out/${TAG}_layers_$i.json 2>/dev/null
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_$i.json").read().strip().splitlines()[-1])
    print("ms/step %.3f  e2e %.3e  families %s" % (d["ms_per_step"], d["e2e"]["value"], {k:round(v["us"]) for k,v in d.get("families",{}).items()}))
except Exception as e:
    print("bench failed:", e)
PY
  i=$((i+1))
done
