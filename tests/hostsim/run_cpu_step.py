"""Two whole training steps (+ an inference call) of the engine's REAL host code, executed on the CPU by the reference kernels of
tests/hostsim/cpu_kernels.cpp (the recording runtime in execute mode), compared with the oracle.  TEST INFRASTRUCTURE.

Started by tests/test_cpu_device.py as a subprocess with WUN_LIB pointing at the `--cudart none` build of the engine.  The
"device" buffers are numpy arrays; the library sees their addresses.  Prints one JSON object: the relative errors of the loss, of
every gradient tensor (worst and which), of the source estimates, and of the parameters after two TF-form Adam steps.

usage: run_cpu_step.py <batch> <num_frames> <grad_scale> <json overrides> <preset> [<preset> ...]"""
import ctypes
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
for p in (REPO, os.path.join(REPO, "wave-u-net_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def aligned(n_floats):
    raw = np.zeros(n_floats + 64, np.float32)
    shift = (-raw.ctypes.data % 256) // 4
    return raw[shift:shift + n_floats]


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / max(np.linalg.norm(np.asarray(b, np.float64)), 1e-30))


def main():
    batch, nf, grad_scale = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
    overrides, named = json.loads(sys.argv[4]), sys.argv[5:]
    import Config
    import wun
    from oracle import wave_unet_oracle as O
    assert "libwun_sim" in wun.LIB_PATH, wun.LIB_PATH
    fake = ctypes.CDLL(os.path.join(os.path.dirname(wun.LIB_PATH), "libfakecudart.so"))
    fake.fakecuda_trace.restype = ctypes.c_longlong
    fake.fakecuda_trace.argtypes = [ctypes.c_char_p, ctypes.c_longlong]
    fake.fakecuda_set_execute(1)

    cfg = Config.build_config(named, overrides, experiment_id=0)["model_config"]
    try:
        t_in, t_out = O.get_padding(cfg, nf)
        probe = O.forward_np(cfg, O.init_params(cfg, seed=1), np.zeros((1, t_in, O.num_channels(cfg)), np.float32), False)
        real_out = int(next(iter(probe.values())).shape[1])
        if real_out != t_out:
            # reference quirk: get_padding's formula and the graph disagree for some filter combinations
            # (UnetAudioSeparator.py:69-76).  The engine's plan walks the graph: its T_out must be the graph's.
            eng_q = wun.Engine(wun.config_from_model_config(cfg), input_frames=t_in)
            print(json.dumps({"formula_quirk": True, "formula_t_out": int(t_out), "graph_t_out": real_out, "engine_t_out": int(eng_q.T_out)}))
            return
    except AssertionError as ex:                              # infeasible shapes (UnetAudioSeparator.py:55, :121, Utils.py:114-117):
        engine_error = None                                   # the engine has to refuse them too, with the same exception type
        try:
            wcfg = wun.config_from_model_config(cfg)
            e_in, _ = wun.get_padding(wcfg, nf)
            wun.Engine(wcfg, input_frames=e_in)
        except AssertionError as ex2:
            engine_error = "AssertionError: %s" % str(ex2)[:120]
        print(json.dumps({"infeasible": True, "oracle_error": str(ex)[:120], "engine_error": engine_error}))
        return
    params = O.init_params(cfg, seed=11)
    rng = np.random.default_rng(12)
    for k in params:
        if k.endswith("/bias") or "interp" in k:
            params[k] = rng.uniform(-0.3, 0.3, size=params[k].shape).astype(np.float32)
    mix, targets = O.synthetic_batch(cfg, batch, t_in, t_out, seed=13)
    shard = os.environ.get("HOSTSIM_SHARD")                 # "lo:hi": this process is one data-parallel rank holding examples [lo, hi)
    if shard:
        lo, hi = (int(x) for x in shard.split(":"))
        mix, targets = mix[lo:hi], {k: v[lo:hi] for k, v in targets.items()}
        batch = hi - lo
    names = list(cfg["source_names"])
    K, C = len(names), O.num_channels(cfg)

    eng = wun.Engine(wun.config_from_model_config(cfg), input_frames=t_in)
    lib, h = wun.lib, eng._h
    n = eng.param_numel
    par, grads, m, v = aligned(n), aligned(n), aligned(n), aligned(n)
    for pname, shape, off, numel in eng.param_table:
        par[off:off + numel] = np.asarray(params[pname], np.float32).reshape(-1)
    ws_bytes = eng.workspace_bytes(batch, True)
    ws = aligned(ws_bytes // 4 + 64)
    ws_inf_bytes = eng.workspace_bytes(batch, False)
    ws_inf = aligned(ws_inf_bytes // 4 + 64)
    mix_d = aligned(mix.size); mix_d[:] = mix.reshape(-1)
    tg = np.stack([targets[s] for s in names]).astype(np.float32)
    tg_d = aligned(tg.size); tg_d[:] = tg.reshape(-1)
    out_d = aligned(tg.size)
    loss = aligned(4)
    state = aligned(4); state[:3] = [0.9, 0.999, 0.0]
    VP = ctypes.c_void_p
    P = lambda a: VP(a.ctypes.data)      # noqa: E731
    MAIN = VP(0x10)

    res = {"unknown": []}
    lr = 1e-3
    o_par = {k: np.asarray(p, np.float32).copy() for k, p in params.items()}
    o_m = {k: np.zeros_like(p) for k, p in o_par.items()}
    o_v = {k: np.zeros_like(p) for k, p in o_par.items()}
    for step in (1, 2):
        wun.check(lib.wun_forward_backward(h, P(par), P(mix_d), P(tg_d), batch, P(out_d), P(loss), P(grads), grad_scale, P(ws), ws_bytes, MAIN))
        loss_o, outs_o, grads_o = O.forward_backward(cfg, o_par, mix, targets)
        worst, which = 0.0, None
        for pname, shape, off, numel in eng.param_table:
            e = rel(grads[off:off + numel], np.asarray(grads_o[pname]).reshape(-1) * grad_scale)
            if e > worst:
                worst, which = e, pname
        if step == 1 and os.environ.get("HOSTSIM_DUMP_GRADS"):
            np.save(os.environ["HOSTSIM_DUMP_GRADS"], grads[:n].copy())
        res["step%d" % step] = {
            "loss": float(loss[0]), "loss_oracle": float(loss_o), "loss_rel": abs(float(loss[0]) - loss_o) / abs(loss_o),
            "grad_worst_rel": worst, "grad_worst_tensor": which,
            "outputs_rel": max(rel(out_d.reshape(K, batch, t_out, C)[k], outs_o[s]) for k, s in enumerate(names))}
        wun.check(lib.wun_adam_step_device(h, P(par), P(grads), P(m), P(v), P(state), lr, 0.9, 0.999, 1e-8, MAIN))
        for k in o_par:        # the oracle's own trajectory: its gradients (scaled like the engine's), TF-form Adam
            o_par[k], o_m[k], o_v[k] = O.adam_update(o_par[k], np.asarray(grads_o[k], np.float32) * np.float32(grad_scale), o_m[k], o_v[k], step, lr)
        res["step%d" % step]["params_rel"] = max(rel(par[off:off + numel], o_par[pname].reshape(-1)) for pname, shape, off, numel in eng.param_table)
    res["adam_state"] = [float(x) for x in state[:3]]
    # inference on the same handle (test-time clip), with the updated parameters
    wun.check(lib.wun_forward(h, P(par), P(mix_d), batch, 0, P(out_d), P(ws_inf), ws_inf_bytes, MAIN))
    want = O.forward_np(cfg, o_par, mix, False)
    res["infer_outputs_rel"] = max(rel(out_d.reshape(K, batch, t_out, C)[k], want[s]) for k, s in enumerate(names))
    # get_output(training=True) without a loss: no test-time clip (UnetAudioSeparator.py:131-136, Utils.py:89-92)
    wun.check(lib.wun_forward(h, P(par), P(mix_d), batch, 1, P(out_d), P(ws_inf), ws_inf_bytes, MAIN))
    want = O.forward_np(cfg, o_par, mix, True)
    res["train_mode_outputs_rel"] = max(rel(out_d.reshape(K, batch, t_out, C)[k], want[s]) for k, s in enumerate(names))

    # window gather / scatter of predict_track (Evaluate.py:125-139): hop = T_out, the last window shifted back, plain overwrite
    n_frames = 3 * t_out - 7
    starts = []
    for pos in range(0, n_frames, t_out):
        starts.append(n_frames - t_out if pos + t_out > n_frames else pos)
    nw = len(starts)
    padded = rng.standard_normal((n_frames + (t_in - t_out), C)).astype(np.float32)
    st = np.asarray(starts, np.int64)
    win = aligned(nw * t_in * C)
    wun.check(lib.wun_gather_windows(h, P(padded), padded.shape[0], P(st), nw, P(win), MAIN))
    res["gather_exact"] = bool(np.array_equal(win.reshape(nw, t_in, C), np.stack([padded[s0:s0 + t_in] for s0 in starts])))
    outs = rng.standard_normal((K, nw, t_out, C)).astype(np.float32)
    preds = aligned(K * n_frames * C)
    wun.check(lib.wun_scatter_windows(h, P(outs), P(st), nw, P(preds), n_frames, MAIN))
    want_p = np.zeros((K, n_frames, C), np.float32)
    for w, s0 in enumerate(starts):
        want_p[:, s0:s0 + t_out] = outs[:, w]
    res["scatter_exact"] = bool(np.array_equal(preds.reshape(K, n_frames, C), want_p))

    # the whole predict_track pipeline through the C-ABI (Evaluate.py:82-145 as wave-u-net_b200/Evaluate.py drives it): pad, gather
    # the windows, batched forward at test time, scatter - against the oracle's window-by-window restatement
    if os.environ.get("HOSTSIM_PREDICT"):
        n_audio = 2 * t_out + t_out // 3
        audio = (rng.uniform(-1.0, 1.0, size=(n_audio, C)) * 0.5).astype(np.float32)
        pad = (t_in - t_out) // 2
        padded_a = np.ascontiguousarray(np.pad(audio, [(pad, pad), (0, 0)]))
        starts_a = []
        for pos in range(0, n_audio, t_out):
            starts_a.append(n_audio - t_out if pos + t_out > n_audio else pos)
        nwa = len(starts_a)
        sta = np.asarray(starts_a, np.int64)
        win_a = aligned(nwa * t_in * C)
        wun.check(lib.wun_gather_windows(h, P(padded_a), padded_a.shape[0], P(sta), nwa, P(win_a), MAIN))
        ws_p_bytes = eng.workspace_bytes(nwa, False)
        ws_p = aligned(ws_p_bytes // 4 + 64)
        outs_a = aligned(K * nwa * t_out * C)
        wun.check(lib.wun_forward(h, P(par), P(win_a), nwa, 0, P(outs_a), P(ws_p), ws_p_bytes, MAIN))
        preds_a = aligned(K * n_audio * C)
        wun.check(lib.wun_scatter_windows(h, P(outs_a), P(sta), nwa, P(preds_a), n_audio, MAIN))
        want_a = O.predict_track(cfg, o_par, audio, t_in, t_out)
        res["predict_rel"] = max(rel(preds_a.reshape(K, n_audio, C)[k], want_a[s]) for k, s in enumerate(names))
        res["predict_windows"] = nwa

    nbytes = fake.fakecuda_trace(None, 0)
    buf = ctypes.create_string_buffer(int(nbytes))
    fake.fakecuda_trace(buf, nbytes)
    kinds = {}
    for ln in buf.value.decode().splitlines():
        if ln.startswith("L "):
            name = ln.split(" ")[2]
            if name.startswith("UNKNOWN:"):
                res["unknown"].append(name)
            k = name.split("(")[0].replace("void_", "").replace("wun::", "").replace("UNKNOWN:", "")
            kinds[k] = kinds.get(k, 0) + 1
    res["kernels"] = kinds
    print(json.dumps(res))


if __name__ == "__main__":
    main()
