#!/usr/bin/env python
"""bench.py - headline benchmark of the B200 Wave-U-Net engine.

Metric (BASELINE.json): audio samples/sec, fwd+bwd, M4 context model.
Pinned definition: OUTPUT FRAMES per second = B * T_out / step_time, a stereo frame counts
once; one step = forward + MSE loss + backward (+ NCCL gradient all-reduce when N > 1) + Adam, i.e. one
`sess.run([separator_solver, ...])` of /root/reference/Training.py:103-109, on synthetic windows.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Under torchrun (N > 1) every rank runs 16 windows (weak scaling) and all-reduces the flat gradient buffer.
`--impl reference` times the CPU restatement of the same step (oracle/, torch-CPU, all host cores): the
reference itself is TensorFlow 1.8 and cannot be installed here.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, "wave-u-net_b200")
for p in (REPO, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "audio samples/sec fwd+bwd, M4 context model (output frames/s)"
UNIT = "frames/s"
PRESET = "baseline_stereo"       # M4 (Config.py:71-78)
BATCH_PER_GPU = 16               # BASELINE.json configs[1]


def load_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(source="measured", hbm_gbs=p["hbm_gbs"], tf_burst=p["bf16_tflops"],
                    tf_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]))
    return dict(source="fallback", hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0)


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.lines = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin=None, t_end=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        inside = [ln for (ts, ln) in self.lines if (t_begin is None or ts >= t_begin) and (t_end is None or ts <= t_end + 0.06)]
        if not inside:                       # timed region shorter than one sample: take the samples closest to it
            inside = [ln for (ts, ln) in self.lines[-2:]]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def build_problem(cfg, batch, seed):
    import numpy as np
    from oracle import wave_unet_oracle as O     # data synthesis only; not on the timed path
    t_in, t_out = O.get_padding(cfg, cfg["num_frames"])
    mix, targets = O.synthetic_batch(cfg, batch, t_in, t_out, seed=seed)
    tg = np.stack([targets[s] for s in cfg["source_names"]])
    return t_in, t_out, mix, tg


def cpu_step_rate(cfg, seconds_budget, steps=None, warmup=1):
    """Times the oracle's restatement of one Training.py:103-109 step (fwd + MSE + bwd + TF-Adam) on the
    host cores, on a bounded sample (batch 1 window per step) of the same workload."""
    import numpy as np
    import torch
    from oracle import wave_unet_oracle as O
    t_in, t_out = O.get_padding(cfg, cfg["num_frames"])
    params = O.init_params(cfg, seed=1337)
    mix, targets = O.synthetic_batch(cfg, 1, t_in, t_out, seed=1)
    m = {k: np.zeros_like(v) for k, v in params.items()}
    v = {k: np.zeros_like(p) for k, p in params.items()}

    def one(step):
        _, _, grads = O.forward_backward(cfg, params, mix, targets)
        for k in params:
            params[k], m[k], v[k] = O.adam_update(params[k], grads[k], m[k], v[k], step, 1e-4)

    # "all the host threads it can use": one window's convs do not scale past a few dozen threads (oneDNN
    # oversubscribes badly on 100+ core hosts), so pick the fastest of a few thread counts, then time with that.
    ncpu = os.cpu_count() or 1
    best_t, best_n = None, 1
    for nthr in sorted(set([min(ncpu, c) for c in (8, 16, 32)])):     # >32 threads only loses (measured: 128 threads = 100x slower)
        torch.set_num_threads(nthr)
        one(1)
        t0 = time.perf_counter(); one(1); dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best_t, best_n = dt, nthr
    torch.set_num_threads(best_n)
    for i in range(warmup):
        one(i + 1)
    t0 = time.perf_counter()
    n = 0
    while True:
        one(warmup + n + 1)
        n += 1
        el = time.perf_counter() - t0
        if steps is not None:
            if n >= steps:
                break
        elif el >= seconds_budget and n >= 3:
            break
    el = time.perf_counter() - t0
    return dict(value=t_out * n / el, ms_per_step=1e3 * el / n, steps=n, cores=torch.get_num_threads(),
                sample="M4 window batch 1 (147443 in / 16389 out stereo), %d steps, fwd+loss+bwd+Adam" % n)


def run_reference(args, rank, world):
    if rank != 0:
        return
    import Config
    cfg = Config.build_config([PRESET], experiment_id=0)["model_config"]
    r = cpu_step_rate(cfg, 0, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "M4 baseline_stereo, L=12, 147443-in/16389-out stereo, one window per step",
                       "note": "CPU restatement (torch/oneDNN fp32) of the Training.py step, not TensorFlow 1.8"},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                             "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import Config
    from Models.UnetAudioSeparator import UnetAudioSeparator

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()                       # samples are time-stamped; only those inside the timed region are reported
    cfg = Config.build_config([PRESET], experiment_id=0)["model_config"]
    B = BATCH_PER_GPU
    t_in, t_out, mix_np, tg_np = build_problem(cfg, B, seed=1337 + rank)
    sep = UnetAudioSeparator(cfg)
    eng = sep.engine(input_frames=t_in)
    # same seed on every rank -> identical replicas (DP invariant)
    dummy = torch.zeros((1, t_in, mix_np.shape[2]), device=dev)
    sep._ensure_params(eng, dev, create=True)
    del dummy
    sep._ensure_training_state()
    lr = cfg["init_sup_sep_lr"]
    mix_h = torch.from_numpy(mix_np).pin_memory()
    tg_h = torch.from_numpy(tg_np).pin_memory()
    mix_d = mix_h.to(dev)
    tg_d = tg_h.to(dev)
    grad_scale = 1.0 / world

    def step_device():
        sep.loss_and_gradients(mix_d, tg_d, grad_scale=grad_scale)
        if world > 1:
            dist.all_reduce(sep.grads)
        sep.adam_step(lr)

    stream = torch.cuda.Stream(device=dev)
    graph = None
    with torch.cuda.stream(stream):
        for _ in range(2):
            step_device()
        stream.synchronize()
        if world == 1 and not args.no_graph:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                step_device()

        def run_step():
            if graph is not None:
                graph.replay()
            else:
                step_device()

        for _ in range(args.warmup):
            run_step()
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_begin = time.perf_counter()
        e0.record(stream)
        for _ in range(args.steps):
            run_step()
        e1.record(stream)
        stream.synchronize()
        torch.cuda.synchronize()
        t_end = time.perf_counter()
        ms_total = e0.elapsed_time(e1)
        clk = clocks.stop(t_begin, t_end) if rank == 0 else None
        if world > 1:
            t = torch.tensor([ms_total], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_total = float(t.item())
            dist.barrier()

        # ---- end to end through the public API with HOST buffers ---------------------------------
        def step_e2e():
            mix_d.copy_(mix_h, non_blocking=True)
            tg_d.copy_(tg_h, non_blocking=True)
            step_device()
            return float(sep._loss.item())      # device -> host read of the step's result

        for _ in range(max(1, min(3, args.warmup))):
            step_e2e()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e_steps = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(e_steps):
            last_loss = step_e2e()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e_s], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e_seq_s = e2e_s

        # ---- the same, with the package's input prefetcher: H2D of step i+1 overlaps the compute of step i, the step itself
        # is the captured CUDA graph when there is one.  Every step still copies its inputs from pinned host memory and
        # reads its loss back inside the timed region.  Any failure here leaves the sequential number standing.
        e2e_pipe_s, pipe_err = None, None
        try:
            from wun.prefetch import DevicePrefetcher
            pf = DevicePrefetcher([mix_d, tg_d])

            def run_pipelined(n):
                loss = None
                pf.issue([mix_h, tg_h])
                for i in range(n):
                    if i + 1 < n:
                        pf.issue([mix_h, tg_h])
                    pf.consume()
                    run_step()
                    loss = float(sep._loss.item())
                return loss

            run_pipelined(max(2, min(3, args.warmup)))
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            pipe_loss = run_pipelined(e_steps)
            torch.cuda.synchronize()
            e2e_pipe_s = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([e2e_pipe_s], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                e2e_pipe_s = float(t.item())
            if not (pipe_loss == pipe_loss and 0.0 < pipe_loss < 10.0):      # finite, plausible MSE
                raise RuntimeError("implausible loss %r from the pipelined loop" % (pipe_loss,))
            if e2e_pipe_s < e2e_s:
                e2e_s, last_loss = e2e_pipe_s, pipe_loss
        except Exception as ex:                                               # noqa: BLE001 - keep the bench line alive
            pipe_err = "%s: %s" % (type(ex).__name__, ex)

    ms_step = ms_total / args.steps
    frames = B * t_out * world
    value = frames / (ms_step * 1e-3)
    e2e_value = frames / (e2e_s / e_steps)
    peaks = load_peaks()
    flops = eng.forward_backward_flops(B)
    step_tf = flops / (ms_step * 1e-3) * 1e-12

    # ---- dominant kernel: the tcgen05 forward conv of the heaviest layer, timed alone with CUDA events --------------
    L = cfg["num_layers"]
    dom_layer = 3 if L > 3 else L - 1            # down3: 72->96 channels, 33.9 GFLOP at B=16 - the largest layer (ties down2)
    with torch.cuda.stream(stream):
        dom_flops = eng.run_conv_layer(dom_layer, 3, sep.params, mix_d)             # warm-up
        stream.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k_iters = 20
        k0.record(stream)
        eng.run_conv_layer(dom_layer, k_iters, sep.params, mix_d)
        k1.record(stream)
        stream.synchronize()
    dom_us = k0.elapsed_time(k1) * 1e3 / k_iters
    dom_tf = dom_flops / (dom_us * 1e-6) * 1e-12
    traffic = None
    tpath = os.path.join(REPO, "profiles", "dominant_kernel_r1.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "M4 baseline_stereo L=12, 147443-in/16389-out stereo, batch %d per GPU, "
                               "fwd+loss+bwd+Adam%s" % (B, "+NCCL all-reduce" if world > 1 else ""),
                   "global_batch": B * world, "parallelism": "dp%d" % world,
                   "l2_policy": "per-step working set (~1.2 GB activations + gradients) exceeds the 126 MB L2",
                   "cuda_graph": graph is not None,
                   "arithmetic": "fp32 in/out; tensor-core layers split every fp32 operand into bf16 hi+lo and issue "
                                 "3 bf16 MMAs per product (fp32 accumulate): 5e-6 rel. error vs the 1e-4 parity bar"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(mix_h.numel() * 4 + tg_h.numel() * 4),
                "d2h_bytes_per_step": 4, "steps": e_steps, "last_loss": last_loss,
                "sequential_value": frames / (e2e_seq_s / e_steps),
                "prefetch_value": (frames / (e2e_pipe_s / e_steps)) if e2e_pipe_s else None,
                "mode": "prefetch (wun.prefetch.DevicePrefetcher: H2D of step i+1 overlaps step i)"
                        if (e2e_pipe_s and e2e_pipe_s <= e2e_seq_s) else "sequential", "prefetch_error": pipe_err},
        "gpu_launches": int((eng.launches(True) + 1) * args.steps),
        "clocks": clk,
        "roofline": {"bound": "tensor", "achieved": dom_tf, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                     "frac": dom_tf / peaks["tf_burst"], "traffic": traffic,
                     "kernel": "plane_conv_umma_persistent (tcgen05), forward of down%d (%s rows x %d->%d ch, k=15), %.1f us/launch, "
                               "%.2f algorithmic GFLOP/launch (live positions, 2 FLOP/MAC, the 3 bf16 MMAs per product "
                               "count once)" % (dom_layer, "16x%d" % ((t_in >> (dom_layer + 1))), 24 * dom_layer,
                                                24 * (dom_layer + 1), dom_us, dom_flops * 1e-9),
                     "peak_source": "bf16 dense burst, %s; the fp32-accurate 3-MMA scheme caps frac at 1/3" % peaks["source"]},
        "step_roofline": {"bound": "tensor", "achieved": step_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                          "frac": step_tf / peaks["tf_sustained"],
                          "scope": "whole step: %.1f algorithmic GFLOP (fwd+bwd, live positions) / step time; peak = bf16 "
                                   "sustained, %s" % (flops * 1e-9, peaks["source"])},
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_step_rate(cfg, 12.0)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                    "sample": r["sample"]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
