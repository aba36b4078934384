"""Join the planner's audit (wun_debug_plan, no GPU needed) with a measured ncu launch list and print, per tensor-core
launch of one M4 training step, the measured time next to the tensor-pipe floor of the MMAs the launch issues:

    python tools/plan_model.py profiles/r1_launches_bench_step.csv [batch=16]

floor = mmas * cycles(N) / (148 SMs * f_clk), cycles(N) = the measured issue cost of one M=128, K=16 SS-mode MMA
(tools/umma_layout_bench: 46.7 for N <= 48, 48 @64, 56 @96, 64 @128, 96 @192), times the wave quantisation of the
launch (ceil(tiles / slots) / (tiles / slots)).  "eff" = floor / measured: what is left is operand staging, barrier
round trips, prologue / epilogue exposure.  The ncu times are cold-cache and serialised (no side-stream overlap)."""
import csv
import math
import os
import re
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "wave-u-net_b200"))

import Config                                                    # noqa: E402
from Models.UnetAudioSeparator import UnetAudioSeparator         # noqa: E402

F_CLK = 1.965e9
SMS = 148


def mma_cycles(n):
    return max(46.7, 32.0 + n / 4.0, n / 2.0)


def read_launches(path):
    rows = [l for l in open(path) if not l.startswith("==")]
    seq = []
    for r in csv.DictReader(rows):
        try:
            t = float(r["Metric Value"].replace(",", ""))
        except Exception:
            continue
        u = r["Metric Unit"]
        t = t / 1e3 if u == "ns" else (t * 1e3 if u == "ms" else t)
        seq.append((re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("wun::", ""), t))
    adam = [i for i, s in enumerate(seq) if s[0] == "adam_kernel"]
    if len(adam) >= 2:
        return seq[adam[0] + 1: adam[1] + 1]
    return seq[: adam[0] + 1] if adam else seq


def main():
    path = sys.argv[1]
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    mc = Config.build_config(["baseline_stereo"], experiment_id=0)["model_config"]
    eng = UnetAudioSeparator(mc).engine(num_frames=mc["num_frames"])
    audit = eng.plan_audit(batch)
    step = read_launches(path)
    conv_t = [t for n, t in step if n.startswith("plane_conv_umma")]
    conv_n = [n for n, t in step if n.startswith("plane_conv_umma")]
    wg_t = [t for n, t in step if n.startswith("wgrad_umma")]
    convs = [d for d in audit if d["op"] == "conv"]
    wg_layers = []
    for d in audit:
        if d["op"] == "wgrad" and (not wg_layers or wg_layers[-1][0]["layer"] != d["layer"]):
            wg_layers.append([d])
        elif d["op"] == "wgrad":
            wg_layers[-1].append(d)
    if len(conv_t) != len(convs) or len(wg_t) != len(wg_layers):
        print("launch list (%d conv, %d wgrad) does not match the plan (%d, %d) - different build / batch?"
              % (len(conv_t), len(wg_t), len(convs), len(wg_layers)))
        return 1
    L = mc["num_layers"]
    name = lambda i: ("down%d" % i) if i < L else ("bottleneck" if i == L else "up%d" % (i - L - 1))
    print("%-11s %-5s %-10s %4s %4s %2s %6s %6s %9s %9s %5s" % ("layer", "pass", "kernel", "N", "NPAD", "MT", "tiles", "waves",
                                                               "meas us", "floor us", "eff"))
    tot_m = tot_f = 0.0
    for d, t, kn in zip(convs, conv_t, conv_n):
        slots = SMS * (2 if d["kernel"] == "dense2" else 1)
        waves = d["tiles"] / slots
        quant = math.ceil(waves) / waves
        floor = d["mmas"] * mma_cycles(d["NPAD"]) / (SMS * F_CLK) * 1e6 * quant
        tot_m += t; tot_f += floor
        print("%-11s %-5s %-10s %4d %4d %2d %6d %6.2f %9.1f %9.1f %5.2f" % (name(d["layer"]), ("fwd", "dgrad")[d["pass"]], d["kernel"],
                                                                          d["N"], d["NPAD"], d["MT"], d["tiles"], waves, t, floor, floor / t))
    print("conv total: measured %.0f us, MMA floor %.0f us (%.2f)" % (tot_m, tot_f, tot_f / tot_m))
    tot_m = tot_f = 0.0
    print("\n%-11s %6s %9s %9s %5s   groups: (Cp x Cg, NT, taps/cta x tapsets)" % ("layer", "CTAs", "meas us", "floor us", "eff"))
    for groups, t in zip(wg_layers, wg_t):
        mmas = ctas = 0
        per_cta_cycles = 0.0
        for g in groups:
            n_cta = g["n_ctas_x"] * g["mtiles"] * g["ntiles"] * g["tapsets"]
            ctas += n_cta
            # every CTA: chunks_per_cta chunks x 4 K steps x taps x 3 MMAs with N = NT
            per_cta_cycles = max(per_cta_cycles, g["chunks_per_cta"] * 4 * g["taps_per_cta"] * 3 * mma_cycles(g["NT"]))
            mmas += g["chunks"] * g["mtiles"] * g["ntiles"] * 4 * g["ntaps"] * 3
        floor = per_cta_cycles / F_CLK * 1e6 * math.ceil(ctas / SMS)
        tot_m += t; tot_f += floor
        desc = " ".join("(%dx%d,%d,%dx%d)" % (g["Cp"], g["Cg"], g["NT"], g["taps_per_cta"], g["tapsets"]) for g in groups[:2])
        print("%-11s %6d %9.1f %9.1f %5.2f   %s%s" % (name(groups[0]["layer"]), ctas, t, floor, floor / t, desc, " ..." if len(groups) > 2 else ""))
    print("wgrad total: measured %.0f us, MMA floor %.0f us (%.2f)" % (tot_m, tot_f, tot_f / tot_m))
    return 0


if __name__ == "__main__":
    sys.exit(main())
