"""Step time of one preset at several per-GPU batch sizes on ONE GPU (CUDA-graph replay, CUDA events) - the compute side of the
strong-scaling question (M6 batch 32 over 8 GPUs = 4 windows per GPU): python tools/batch_sweep.py full_multi_instrument 32 16 8 4"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "wave-u-net_b200"))
import torch
import bench

preset = sys.argv[1] if len(sys.argv) > 1 else "full_multi_instrument"
batches = [int(a) for a in sys.argv[2:]] or [32, 16, 8, 4]
dev = torch.device("cuda:0")
base = None
for B in batches:
    run = bench.TrainingRun(preset, B, B, 0, 1, dev, None)
    graphed = run.prepare(True)
    with torch.cuda.stream(run.stream):
        for _ in range(3):
            run.graph.replay() if graphed else run.step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(run.stream)
        for _ in range(20):
            run.graph.replay() if graphed else run.step()
        e1.record(run.stream)
        run.stream.synchronize()
    ms = e0.elapsed_time(e1) / 20
    base = base or (ms, B)
    print("%s batch %2d: %.3f ms/step (graph=%s)  windows/s %.0f  vs batch %d: %.2fx faster step, per-window efficiency %.2f"
          % (preset, B, ms, graphed, B / ms * 1e3, base[1], base[0] / ms, (B / ms) / (base[1] / base[0])))
    del run
    torch.cuda.empty_cache()
