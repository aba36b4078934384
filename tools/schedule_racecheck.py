#!/usr/bin/env python
"""Host-side racecheck of the multi-stream schedule at the BENCHMARK sizes (no GPU; ~1-2 min and ~12 GB of RAM per training case).

The unit tests (tests/test_stream_schedule.py) run mid-size nets; this tool runs the exact configurations bench.py times -
M4 baseline_stereo batch 16, M5 full batch 16, M6 full_multi_instrument at the per-GPU batches of the 8-/2-GPU runs, M1 batch 16,
and the Predict window batch - through the engine's real host code on the recording CUDA runtime and prints, per case, the
launches per stream and the races / out-of-bounds accesses / uninitialised reads found.  Output of the last run: profiles/r2_schedule_racecheck.txt."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "hostsim"))
import schedule  # noqa: E402

CASES = [
    ("M4 baseline_stereo, batch 16, full window: 2 steps of fwd+bwd, bucketed all-reduce, Adam", "train_dp", ["baseline_stereo"], 16),
    ("M5 full (learned upsampling), batch 16", "train_dp", ["full"], 16),
    ("M6 full_multi_instrument, batch 4 (global 32 over 8 GPUs)", "train_dp", ["full_multi_instrument"], 4),
    ("M6 full_multi_instrument, batch 16 (global 32 over 2 GPUs)", "train_dp", ["full_multi_instrument"], 16),
    ("M6 full_multi_instrument, batch 8 (global 32 over 4 GPUs)", "train_dp", ["full_multi_instrument"], 8),
    ("M6 full_multi_instrument, batch 32 (global 32 on one GPU; 4 GB workspace)", "train_dp", ["full_multi_instrument"], 32),
    ("M1 baseline (same padding, mono), batch 16", "train_dp", ["baseline"], 16),
    ("Predict: full_44KHz, 16 windows per batch, two forward calls", "infer", ["full_44KHz"], 16),
]


def main():
    only = sys.argv[1:]
    total = 0
    for title, scenario, named, batch in CASES:
        if only and not any(o in title for o in only):
            continue
        if schedule.PKG not in sys.path:
            sys.path.insert(0, schedule.PKG)
        import Config
        cfg = Config.build_config(named, experiment_id=0)["model_config"]
        t0 = time.time()
        meta, ops = schedule.trace(scenario, named, {}, batch, cfg["num_frames"])
        violations, stats = schedule.check(meta, ops)
        total += len(violations)
        names = {16: "caller", 32: "comm"}
        streams = ", ".join("%s %d" % (names.get(h, "internal#%d" % k), stats["per_stream"][k]) for h, k in stats["streams"].items())
        kinds = {}
        for o in ops:
            if o[0] == "L":
                k = o[2].split("(")[0].replace("void_", "").replace("wun::", "")
                kinds[k] = kinds.get(k, 0) + 1
        print("%s\n   workspace %.2f GB, %d launches on %d streams (%s), %d event records, %d stream waits: %d race(s) / out-of-bounds / uninitialised reads, joined back into the caller's stream: %s, %.0f s"
              % (title, meta["regions"]["ws" if scenario != "infer" else "ws_infer"][1] / 1e9, stats["launches"], len(stats["per_stream"]),
                 streams, sum(1 for o in ops if o[0] == "E"), sum(1 for o in ops if o[0] == "S"), len(violations), stats["joined_into_caller"], time.time() - t0))
        print("   kernels: " + ", ".join("%s x%d" % kv for kv in sorted(kinds.items())))
        for v in violations:
            print("   RACE: %(name)s [%(mode)s %(region)s, %(words)d words] is not ordered after %(other_name)s" % v)
        sys.stdout.flush()
    print("total races: %d" % total)
    return 1 if total else 0


if __name__ == "__main__":
    sys.exit(main())
