#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== gpu tests"
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
for cfg in "X=0" "WUN_FUSE_N=0"; do
  echo "=== bench $cfg"
  env $cfg timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err
  tail -2 gpurun_out/c6_bench.err
  cp gpurun_out/layer_table_n1.json "gpurun_out/c6_layers_${cfg%%=*}.json"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c6_bench.json").read().strip().splitlines()[-1])
    print("ms/step %.3f  e2e %.3e  families %s" % (d["ms_per_step"], d["e2e"]["value"], {k:round(v["us"]) for k,v in d.get("families",{}).items()}))
except Exception as e:
    print("bench failed:", e)
PY
done
echo "=== fold trace"
WUN_LIB=$PWD/wave-u-net_b200/libwun_timing.so timeout 300 python tools/fold_trace.py 11 2>&1 | tail -36
