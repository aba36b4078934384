"""Run one training step of M4 (batch 16), then the forward kernel of one conv layer N times - the target for
`ncu --set full -k regex:plane_conv_umma` captures of the dominant kernel (profiles/).
   python tools/profile_layer.py [layer=3] [iters=3]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "wave-u-net_b200"))
import torch
import Config
from Models.UnetAudioSeparator import UnetAudioSeparator
import bench

layer = int(sys.argv[1]) if len(sys.argv) > 1 else 3
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
run = bench.TrainingRun(bench.PRESET, bench.BATCH_PER_GPU, bench.BATCH_PER_GPU, 0, 1, dev, None)
with torch.cuda.stream(run.stream):
    run.step()
run.stream.synchronize()
eng, sep, mix_d = run.eng, run.sep, run.mix_d
torch.cuda.profiler.start()          # ncu --profile-from-start off: only the launches below are profiled
fl = eng.run_conv_layer(layer, iters, sep.params, mix_d)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("layer", layer, "GFLOP/launch", fl * 1e-9)
