#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== gpu tests"
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
i=0
for cfg in "X=0" "WUN_WG_FUSE=0"; do
  echo "=== bench $cfg"
  env $cfg timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/c18_bench_$i.json 2> gpurun_out/c18_bench_$i.err
  tail -2 gpurun_out/c18_bench_$i.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c18_bench_$i.json").read().strip().splitlines()[-1])
    print("ms/step %.3f  e2e %.3e  families %s top %s" % (d["ms_per_step"], d["e2e"]["value"], {k:round(v["us"]) for k,v in d.get("families",{}).items()}, d["top_launch"]))
except Exception as e:
    print("bench failed:", e)
PY
  i=$((i+1))
done
echo "=== batch sweep M6"
timeout 600 python tools/batch_sweep.py full_multi_instrument 32 16 8 4 2>&1 | tail -5
echo "=== batch sweep M4"
timeout 600 python tools/batch_sweep.py baseline_stereo 16 4 2>&1 | tail -3
