#!/bin/bash
# One GPU call that answers the open questions round 1 left (run under gpurun from the repo root, ~4 min):
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/round2_first_call.sh > gpurun_out/r2_first.log 2>&1; tail -60 gpurun_out/r2_first.log'
# 1. are the bulk-copy-fed tcgen05 kernels (tools/presplit_probe.cu) correct, and how much faster than the converter-fed ones?
# 2. are the dedicated first-layer kernels (kernels_first.cu) correct and faster?  (probe, then the parity tests with the switch)
# 3. does WUN_PACK_EVENTS=1 help, and does the prefetch leg of bench.py's e2e work?
# 4. WUN_BULK_WGRAD=1: split pass + bulk-copy-fed wgrad inside the engine (parity, step time)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== presplit probe (correctness on small cases, then the M4 down3 layer)"
timeout 120 stdbuf -oL tools/presplit_probe small 2>&1 | grep -v "^$"
timeout 120 stdbuf -oL tools/presplit_probe wsmall 2>&1 | grep -v "^$"
timeout 120 stdbuf -oL tools/presplit_probe down3 2>&1 | grep -v "^$"
timeout 120 stdbuf -oL tools/presplit_probe down1 2>&1 | grep -v "^$"
timeout 120 stdbuf -oL tools/presplit_probe wdown3 2>&1 | grep -v "^$"
timeout 120 stdbuf -oL tools/presplit_probe wdown1 2>&1 | grep -v "^$"
echo "=== first-layer probe"
timeout 120 stdbuf -oL tools/first_layer_probe small 2>&1 | grep -v "^$"
timeout 120 stdbuf -oL tools/first_layer_probe m4 2>&1 | grep -v "^$"
echo "=== parity tests with the experimental switches"
WUN_FIRST_LAYER=1 WUN_PACK_EVENTS=1 timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "=== parity tests with the bulk-copy-fed wgrad"
WUN_BULK_WGRAD=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
echo "=== bench: defaults / first layer / + pack events / + bulk wgrad"
for cfg in "X=0" "WUN_FIRST_LAYER=1" "WUN_FIRST_LAYER=1 WUN_PACK_EVENTS=1" "WUN_BULK_WGRAD=1" "WUN_FIRST_LAYER=1 WUN_PACK_EVENTS=1 WUN_BULK_WGRAD=1"; do
  echo -n "$cfg  "
  env $cfg timeout 150 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step %.3f  e2e seq %.3e  prefetch %s  mode %s  err %s' % (d['ms_per_step'], d['e2e']['sequential_value'], d['e2e']['prefetch_value'], d['e2e']['mode'][:10], d['e2e']['prefetch_error']))"
done
