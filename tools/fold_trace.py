"""Timing build only (bash wave-u-net_b200/build.sh -DWUN_UMMA_TIMING -o libwun_timing.so; WUN_LIB=.../libwun_timing.so):
run one forward launch of the given layers alone and print the clock64 marks CTA (0,0,0) of plane_conv_umma_fold left.
   WUN_LIB=wave-u-net_b200/libwun_timing.so python tools/fold_trace.py 11 12 15"""
import ctypes, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "wave-u-net_b200"))
import torch
import Config
from Models.UnetAudioSeparator import UnetAudioSeparator
import bench, wun

dev = torch.device("cuda:0")
run = bench.TrainingRun(bench.PRESET, bench.BATCH_PER_GPU, bench.BATCH_PER_GPU, 0, 1, dev, None)
with torch.cuda.stream(run.stream):
    run.step()
run.stream.synchronize()
eng, sep, mix_d = run.eng, run.sep, run.mix_d
names = ["start", "setup done", "conv: 1st slab filled", "conv: all fills done", "team0: ACC_FULL seen", "team0: dumped", "after __syncthreads",
         "after cluster sync 1", "mma: 1st slab ready", "mma: weights stage 0", "mma: weights stage 1", "mma: weights stage 4",
         "mma: all issued", "reduce done", "after cluster sync 2"]
order = [0, 1, 2, 8, 9, 10, 11, 3, 12, 4, 5, 6, 7, 13, 14]
for layer in [int(a) for a in sys.argv[1:]] or [12]:
    for rep in range(2):
        with torch.cuda.stream(run.stream):
            eng.run_conv_layer(layer, 1, sep.params, mix_d)
        torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 32)()
    wun.lib.wun_debug_fold_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
    rc = wun.lib.wun_debug_fold_trace(buf, 32)
    t0 = buf[0]
    print("layer %d (%s) rc=%d" % (layer, eng.layer_kernel(layer, 0), rc))
    for i in order:
        print("   %-26s %8d cycles" % (names[i], buf[i] - t0 if buf[i] >= t0 else -1))
