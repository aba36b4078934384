#!/usr/bin/env python
"""Per-launch global-memory FOOTPRINT of one training step at the benchmark size (no GPU): the engine's real host code run on the
recording CUDA runtime (tests/hostsim/), every launch decoded into the words it reads / writes / accumulates into, and per
launch the number of DISTINCT bytes it touches - the DRAM traffic of a launch whose every word moves exactly once.  Summed per
kernel and next to the measured serialised times of the committed ncu launch list this is the HBM floor of each kernel family
(bytes / measured HBM bandwidth); measured DRAM traffic above it is re-reading, below it L2 hits on a predecessor's output.

usage: python tools/footprint_model.py [preset] [batch]   ->  text table on stdout, profiles/r2_footprint.json"""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tests", "hostsim"))
sys.path.insert(0, os.path.join(REPO, "wave-u-net_b200"))
sys.path.insert(0, REPO)
import schedule  # noqa: E402


def words_of(acc):
    """distinct-word intervals [(lo, hi)] in units of 4 bytes (absolute addresses / 4) of one access."""
    out = []
    if acc[0] == "F":
        _, _, addr, nbytes = acc
        out.append((addr // 4, (addr + nbytes + 3) // 4))
    else:
        _, _, base, batch, bstride, rlo, rhi, rstride, C = acc
        for b in range(batch):
            w0 = base // 4 + b * bstride + rlo * rstride
            if rstride == C:
                out.append((w0, w0 + (rhi - rlo) * C))
            else:
                starts = w0 + np.arange(rhi - rlo, dtype=np.int64) * rstride
                out.extend(zip(starts.tolist(), (starts + C).tolist()))
    return out


def union_len(intervals):
    if not intervals:
        return 0
    a = np.array(sorted(intervals), dtype=np.int64)
    total, cur_lo, cur_hi = 0, a[0, 0], a[0, 1]
    for lo, hi in a[1:]:
        if lo > cur_hi:
            total += cur_hi - cur_lo
            cur_lo, cur_hi = lo, hi
        elif hi > cur_hi:
            cur_hi = hi
    return int(total + cur_hi - cur_lo)


def main():
    import bench
    import Config
    preset = sys.argv[1] if len(sys.argv) > 1 else "baseline_stereo"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    cfg = Config.build_config([preset], experiment_id=0)["model_config"]
    meta, ops = schedule.trace("train", [preset], {}, batch, cfg["num_frames"])
    launches = [o for o in ops if o[0] == "L"]
    adam = [i for i, o in enumerate(launches) if "adam_advance_kernel" in o[2]]
    step = launches[adam[0] + 1:adam[1] + 1]                      # the second (steady-state) step
    peaks = bench.load_peaks()
    fam = {}
    rows = []
    for o in step:
        name = o[2].split("(")[0].replace("void_", "").replace("wun::", "")
        rd = union_len([w for a in o[3] if a[1] == "R" for w in words_of(a)]) * 4
        wr = union_len([w for a in o[3] if a[1] in ("W", "A") for w in words_of(a)]) * 4
        both = union_len([w for a in o[3] for w in words_of(a)]) * 4
        rows.append({"kernel": name, "read_bytes": rd, "written_bytes": wr, "distinct_bytes": both})
        f = fam.setdefault(name, {"launches": 0, "read_bytes": 0, "written_bytes": 0, "distinct_bytes": 0})
        f["launches"] += 1
        f["read_bytes"] += rd; f["written_bytes"] += wr; f["distinct_bytes"] += both
    total = sum(f["distinct_bytes"] for f in fam.values())
    print("# python tools/footprint_model.py %s %d   (host only; kernel-source sha %s)" % (preset, batch, bench.kernel_source_hash()))
    print("# one steady-state training step: %d launches, %.2f GB of distinct bytes summed over the launches = %.0f us at %.0f GB/s"
          % (len(step), total / 1e9, total / (peaks["hbm_gbs"] * 1e9) * 1e6, peaks["hbm_gbs"]))
    print("%-42s %4s %10s %10s %10s %9s" % ("kernel", "n", "read MB", "written MB", "distinct MB", "HBM us"))
    for name, f in sorted(fam.items(), key=lambda kv: -kv[1]["distinct_bytes"]):
        print("%-42s %4d %10.1f %10.1f %10.1f %9.1f" % (name, f["launches"], f["read_bytes"] / 1e6, f["written_bytes"] / 1e6,
                                                        f["distinct_bytes"] / 1e6, f["distinct_bytes"] / (peaks["hbm_gbs"] * 1e9) * 1e6))
    out = {"source_hash": bench.kernel_source_hash(), "preset": preset, "batch": batch, "hbm_gbs": peaks["hbm_gbs"],
           "step_distinct_bytes": total, "families": fam,
           "note": "distinct bytes each launch touches (reads + writes + accumulations, every word once), from the engine's real launch "
                   "parameters decoded on the host (tests/hostsim); a model of the minimum DRAM traffic, not a measurement"}
    with open(os.path.join(REPO, "profiles", "r2_footprint.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
