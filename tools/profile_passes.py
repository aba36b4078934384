"""Run one M4 (batch 16) training step, then the named (layer, pass) launch groups once each between
cudaProfilerStart/Stop - the target of `ncu --set full --profile-from-start off` captures (profiles/).
   python tools/profile_passes.py 3:0 1:2 3:2 1:1 12:0       (layer:pass, pass 0 fwd / 1 dgrad / 2 wgrad)"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "wave-u-net_b200"))
import torch
import bench

jobs = [tuple(int(x) for x in a.split(":")) for a in sys.argv[1:]] or [(3, 0)]
dev = torch.device("cuda:0")
run = bench.TrainingRun(bench.PRESET, bench.BATCH_PER_GPU, bench.BATCH_PER_GPU, 0, 1, dev, None)
with torch.cuda.stream(run.stream):
    run.step(); run.step()
    run.stream.synchronize()
    scratch = torch.zeros_like(run.sep.grads)
    for layer, p in jobs:                     # warm-up outside the profiled range
        run.eng.run_layer_pass(layer, p, 1, run.sep.params, run.mix_d, scratch)
    run.stream.synchronize()
    torch.cuda.profiler.start()
    for layer, p in jobs:
        fl = run.eng.run_layer_pass(layer, p, 1, run.sep.params, run.mix_d, scratch)
        run.stream.synchronize()
        print("layer %d pass %d kernel %s GFLOP %.3f" % (layer, p, run.eng.layer_kernel(layer, p), fl * 1e-9))
    torch.cuda.profiler.stop()
