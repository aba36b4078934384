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
cfg = Config.build_config([bench.PRESET], experiment_id=0)["model_config"]
t_in, t_out, mix, tg = bench.build_problem(cfg, bench.BATCH_PER_GPU, 1337)
sep = UnetAudioSeparator(cfg)
eng = sep.engine(input_frames=t_in)
sep._ensure_params(eng, torch.device("cuda"), create=True)
mix_d, tg_d = torch.from_numpy(mix).cuda(), torch.from_numpy(tg).cuda()
sep.loss_and_gradients(mix_d, tg_d)
torch.cuda.synchronize()
torch.cuda.profiler.start()          # ncu --profile-from-start off: only the launches below are profiled
fl = eng.run_conv_layer(layer, iters, sep.params, mix_d)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("layer", layer, "GFLOP/launch", fl * 1e-9)
