"""Runs the engine's REAL host code against the recording CUDA runtime (tests/hostsim/fake_cudart.cpp) and prints the trace.

Started by tests/hostsim/schedule.py as a subprocess with WUN_LIB pointing at the `--cudart none` build of the engine, so that
`import wun` binds the simulated library.  Device pointers are made-up addresses in disjoint regions (never dereferenced: no
kernel runs); the caller's stream and the data-parallel communication stream are made-up handles as well.

usage: run_trace.py <scenario> <batch> <num_frames> <json overrides> <preset> [<preset> ...]
scenarios: train (two consecutive steps: forward_backward + device Adam), train_dp (the same with the bucketed gradient
all-reduce of wun/parallel.py on a communication stream), infer (two forward calls), train_out (training step that also
returns the source estimates, then an inference call on the same handle), train_validate (training steps alternating with a
validation forward at another batch size)."""
import ctypes
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
for p in (REPO, os.path.join(REPO, "wave-u-net_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

REGION_SHIFT = 36
REGION_BASE = 1 << 40
REGIONS = ["ws", "params", "grads", "mix", "targets", "outputs", "loss", "adam_m", "adam_v", "adam_state", "ws_infer"]
MAIN, COMM = 0x10, 0x20


def addr(region):
    return REGION_BASE + (REGIONS.index(region) << REGION_SHIFT)


def main():
    scenario, batch, nf = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    overrides, named = json.loads(sys.argv[4]), sys.argv[5:]
    import Config
    import wun
    assert "libwun_sim" in wun.LIB_PATH, wun.LIB_PATH
    fake = ctypes.CDLL(os.path.join(os.path.dirname(wun.LIB_PATH), "libfakecudart.so"))
    fake.fakecuda_trace.restype = ctypes.c_longlong
    fake.fakecuda_trace.argtypes = [ctypes.c_char_p, ctypes.c_longlong]
    fake.fakecuda_user_line.argtypes = [ctypes.c_char_p]

    cfg = Config.build_config(named, overrides, experiment_id=0)["model_config"]
    wcfg = wun.config_from_model_config(cfg)
    t_in, t_out = wun.get_padding(wcfg, nf)
    eng = wun.Engine(wcfg, input_frames=t_in)
    lib, h = wun.lib, eng._h
    K, C = len(cfg["source_names"]), (1 if cfg["mono_downmix"] else 2)
    ws_bytes = eng.workspace_bytes(batch, True)
    batch_inf = batch + 1 if scenario in ("train_validate", "train_two_batches") else batch      # validation / the other batch size
    ws_inf = eng.workspace_bytes(batch_inf, scenario == "train_two_batches")
    sizes = {"ws": ws_bytes, "params": 4 * eng.param_numel, "grads": 4 * eng.param_numel, "mix": 4 * batch_inf * t_in * C,
             "targets": 4 * K * batch_inf * t_out * C, "outputs": 4 * K * batch_inf * t_out * C, "loss": 4, "adam_m": 4 * eng.param_numel,
             "adam_v": 4 * eng.param_numel, "adam_state": 12, "ws_infer": ws_inf}
    VP = ctypes.c_void_p

    def fwd_bwd(outputs):
        wun.check(lib.wun_forward_backward(h, VP(addr("params")), VP(addr("mix")), VP(addr("targets")), batch,
                                           VP(addr("outputs")) if outputs else None, VP(addr("loss")), VP(addr("grads")), 1.0,
                                           VP(addr("ws")), ws_bytes, VP(MAIN)))

    def adam():
        wun.check(lib.wun_adam_step_device(h, VP(addr("params")), VP(addr("grads")), VP(addr("adam_m")), VP(addr("adam_v")),
                                           VP(addr("adam_state")), 1e-4, 0.9, 0.999, 1e-8, VP(MAIN)))

    def user(line):
        fake.fakecuda_user_line(line.encode())

    if scenario in ("train", "train_out"):
        for _ in range(2):
            fwd_bwd(scenario == "train_out")
            adam()
        if scenario == "train_out":
            wun.check(lib.wun_forward(h, VP(addr("params")), VP(addr("mix")), batch, 0, VP(addr("outputs")), VP(addr("ws_infer")),
                                      ws_inf, VP(MAIN)))
    elif scenario == "train_dp":
        from wun import parallel                      # (imports torch: only this scenario pays for it)
        firsts = parallel.bucket_offsets(eng.param_table, 4)
        eng.set_grad_buckets(firsts)
        his = [eng.param_numel] + firsts[:-1]
        done = 0x900000
        for _ in range(2):
            fwd_bwd(False)
            for k, (lo, hi) in enumerate(zip(firsts, his)):              # wun/parallel.py BucketedAllReduce.run
                wun.check(lib.wun_stream_wait_grad_bucket(h, k, VP(COMM)))
                a, n = addr("grads") + 4 * lo, 4 * (hi - lo)
                user("L %d nccl_all_reduce_bucket%d F:R:%d:%d F:W:%d:%d" % (COMM, k, a, n, a, n))
            user("E %d %d" % (done, COMM))
            user("S %d %d" % (MAIN, done))
            adam()
    elif scenario == "train_validate":
        # Training.optimise's pattern (Training.py:123-150): training steps, a validation forward at another batch size on its own
        # workspace, training again - one handle, the internal streams and events shared by all of it
        for _ in range(2):
            fwd_bwd(False)
            adam()
            wun.check(lib.wun_forward(h, VP(addr("params")), VP(addr("mix")), batch_inf, 0, VP(addr("outputs")), VP(addr("ws_infer")),
                                      ws_inf, VP(MAIN)))
    elif scenario == "train_two_batches":
        # an epoch whose last batch is smaller / larger: training steps at two batch sizes on one handle, each with its own
        # workspace (the facade caches one per batch size); the planner's choices, pack arena and split arena differ per batch
        for b, ws_name, nbytes in ((batch, "ws", ws_bytes), (batch_inf, "ws_infer", ws_inf), (batch, "ws", ws_bytes)):
            wun.check(lib.wun_forward_backward(h, VP(addr("params")), VP(addr("mix")), VP(addr("targets")), b, None, VP(addr("loss")),
                                               VP(addr("grads")), 1.0, VP(addr(ws_name)), nbytes, VP(MAIN)))
            adam()
    elif scenario == "infer":
        for _ in range(2):
            wun.check(lib.wun_forward(h, VP(addr("params")), VP(addr("mix")), batch, 0, VP(addr("outputs")), VP(addr("ws_infer")),
                                      ws_inf, VP(MAIN)))
    else:
        raise SystemExit("unknown scenario %r" % scenario)

    n = fake.fakecuda_trace(None, 0)
    buf = ctypes.create_string_buffer(int(n))
    fake.fakecuda_trace(buf, n)
    sys.stdout.write(json.dumps({"regions": {r: [addr(r), sizes[r]] for r in REGIONS}, "main": MAIN, "comm": COMM,
                                 "launches_reported": int(eng.launches(True))}) + "\n")
    sys.stdout.write(buf.value.decode())


if __name__ == "__main__":
    main()
