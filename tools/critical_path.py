#!/usr/bin/env python
"""Critical path of one training step through its launch DAG (no GPU needed to run this; it combines two committed artefacts).

The DAG - which launch runs on which stream, which event orders what - comes from the engine's real host code on the recording
CUDA runtime (tests/hostsim); the duration of every launch from the ncu launch list of the same step
(profiles/r2_launches_bench_step.csv: serialised, cold-cache times).  With unlimited SMs a launch starts when its predecessor on
its stream and every event its stream waited for have finished; the longest chain is the least time the step's DEPENDENCIES
allow.  Measured step time above it = launches of different streams competing for the same SMs / memory system; serialised sum
above the measured time = what the overlap of the streams buys.

usage: python tools/critical_path.py [launch list csv]  ->  text on stdout (profiles/r2_critical_path.txt)"""
import csv
import os
import re
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tests", "hostsim"))
sys.path.insert(0, os.path.join(REPO, "wave-u-net_b200"))
sys.path.insert(0, REPO)
import schedule  # noqa: E402


def short(name):
    n = re.sub(r"\(.*", "", name).replace("void_", "").replace("void ", "").replace("wun::", "").replace(", ", ",").replace(",_", ",")
    return n.replace("false", "0").replace("true", "1")          # (ncu prints bool template arguments as 0 / 1)


def ncu_step(path):
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    seq = []
    for row in csv.DictReader(lines):
        try:
            t = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        u = row["Metric Unit"]
        seq.append((short(row["Kernel Name"]), t / 1e3 if u == "ns" else (t * 1e3 if u == "ms" else t)))
    idx = [i for i, s in enumerate(seq) if s[0] == "adam_kernel"]
    return seq[idx[0] + 2: idx[1] + 2]                 # from the first launch after Adam's advance to this step's advance


def main():
    import Config
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "profiles", "r2_launches_bench_step.csv")
    measured_ms = float(sys.argv[2]) if len(sys.argv) > 2 else 4.83
    cfg = Config.build_config(["baseline_stereo"], experiment_id=0)["model_config"]
    meta, ops = schedule.trace("train", ["baseline_stereo"], {}, 16, cfg["num_frames"])
    adv = [i for i, o in enumerate(ops) if o[0] == "L" and "adam_advance_kernel" in o[2]]
    step_ops = ops[adv[0] + 1: adv[1] + 1]
    times = ncu_step(path)
    kernels = [o for o in step_ops if o[0] == "L" and o[2] != "memset"]
    assert len(kernels) == len(times), (len(kernels), len(times))
    for o, (n, _) in zip(kernels, times):
        assert short(o[2]) == n, (short(o[2]), n)
    dur = iter(t for _, t in times)
    finish_stream, finish_event, crit_stream, crit_event = {}, {}, {}, {}
    nodes = []                                         # (name, stream, start, finish, critical predecessor index)
    pending = {}                                       # stream -> (time, node index) from waited events
    for o in step_ops:
        if o[0] == "E":
            finish_event[o[1]] = finish_stream.get(o[2], 0.0)
            crit_event[o[1]] = crit_stream.get(o[2])
        elif o[0] == "S":
            t = finish_event.get(o[2], 0.0)            # events recorded before the step started count as time 0
            if t > pending.get(o[1], (0.0, None))[0]:
                pending[o[1]] = (t, crit_event.get(o[2]))
        elif o[0] == "L":
            s = o[1]
            d = 2.0 if o[2] == "memset" else next(dur)
            start, pred = finish_stream.get(s, 0.0), crit_stream.get(s)
            if s in pending and pending[s][0] > start:
                start, pred = pending[s]
            pending.pop(s, None)
            nodes.append((short(o[2]), s, start, start + d, pred))
            finish_stream[s] = start + d
            crit_stream[s] = len(nodes) - 1
    end = max(range(len(nodes)), key=lambda i: nodes[i][3])
    chain, i = [], end
    while i is not None:
        chain.append(i)
        i = nodes[i][4]
    chain.reverse()
    serial = sum(t for _, t in times)
    cp = nodes[end][3]
    names = {meta["main"]: "caller"}
    print("# python tools/critical_path.py   (DAG: tests/hostsim trace of M4 batch 16; durations: %s)" % os.path.relpath(path, REPO))
    print("serialised kernel time of the step   %8.1f us   (%d launches, cold-cache ncu times)" % (serial, len(times)))
    print("critical path through the launch DAG %8.1f us   (unlimited SMs: only stream order and events constrain)" % cp)
    print("measured step (CUDA graph, warm)     %8.1f us" % (measured_ms * 1e3))
    per_stream = {}
    for n in nodes:
        per_stream[n[1]] = per_stream.get(n[1], 0.0) + (n[3] - n[2])
    print("busy time per stream: " + ", ".join("%s %.0f us" % (names.get(s, "internal %d" % s), t) for s, t in sorted(per_stream.items(), key=lambda kv: -kv[1])))
    on = {}
    for i in chain:
        key = (names.get(nodes[i][1], "internal %d" % nodes[i][1]), re.sub(r"<.*", "", nodes[i][0]))
        on[key] = on.get(key, [0, 0.0])
        on[key][0] += 1
        on[key][1] += nodes[i][3] - nodes[i][2]
    print("critical path = %d launches:" % len(chain))
    for (s, k), (n, t) in sorted(on.items(), key=lambda kv: -kv[1][1]):
        print("   %-10s %-34s n=%3d %8.1f us  %5.1f %%" % (s, k, n, t, 100.0 * t / cp))
    waits = sum(max(0.0, nodes[b][2] - nodes[a][3]) for a, b in zip(chain, chain[1:]))
    print("   (idle gaps on the path: %.1f us)" % waits)


if __name__ == "__main__":
    main()
