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
    poison = bool(os.environ.get("HOSTSIM_POISON"))          # the footprint "poison test" of fake_cudart.cpp (execute mode 2)
    fake.fakecuda_set_execute(2 if poison else 1)
    fake.fakecuda_register_buffer.argtypes = [ctypes.c_void_p, ctypes.c_longlong]

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
    if poison:
        for a in (par, grads, m, v, ws, ws_inf, mix_d, tg_d, out_d, loss, state):
            fake.fakecuda_register_buffer(P(a), a.size * 4)

    res = {"unknown": [], "footprint_miss": []}
    lr = 1e-3
    def engine_params():
        return {pname: par[off:off + numel].reshape(np.asarray(params[pname]).shape).copy() for pname, shape, off, numel in eng.param_table}

    for step in (1, 2):
        # the oracle at the parameters the ENGINE holds at this step: the two differ by arithmetic only (a comparison along two
        # separate Adam trajectories would also measure how ill-conditioned some gradients are - the mean residual of a bias)
        o_par = engine_params()
        wun.check(lib.wun_forward_backward(h, P(par), P(mix_d), P(tg_d), batch, P(out_d), P(loss), P(grads), grad_scale, P(ws), ws_bytes, MAIN))
        loss_o, outs_o, grads_o = O.forward_backward(cfg, o_par, mix, targets)
        if os.environ.get("HOSTSIM_GRAD_REF", "f64") == "f64":
            # gradients against float64 autograd through the oracle: the reference routines accumulate in double, and a few
            # gradients (the mean residual behind an output bias, interp_*) are sums with heavy cancellation in fp32
            import torch
            tp = O._as_torch(o_par, torch.float64, True)
            outs64 = O.forward(cfg, tp, torch.as_tensor(mix).to(torch.float64), True)
            l64 = O.mse_loss(cfg, outs64, {k: torch.as_tensor(vv).to(torch.float64) for k, vv in targets.items()})
            grads_o = {k: g.numpy() for k, g in zip(tp, torch.autograd.grad(l64, list(tp.values())))}
        worst, which = 0.0, None
        for pname, shape, off, numel in eng.param_table:
            e = rel(grads[off:off + numel], np.asarray(grads_o[pname]).reshape(-1) * grad_scale)
            if e > worst:
                worst, which = e, pname
        explained = None
        if worst >= 1e-4 and os.environ.get("HOSTSIM_F64"):
            # diagnostic: the same gradients from the oracle in float64 - is the difference the fp32 oracle's own rounding?
            import torch
            tp = O._as_torch(o_par, torch.float64, True)
            outs64 = O.forward(cfg, tp, torch.as_tensor(mix).to(torch.float64), True)
            l64 = O.mse_loss(cfg, outs64, {k: torch.as_tensor(vv).to(torch.float64) for k, vv in targets.items()})
            g64 = dict(zip(tp, torch.autograd.grad(l64, list(tp.values()))))
            off_w, num_w = [(off, numel) for pname, shape, off, numel in eng.param_table if pname == which][0]
            res["f64_step%d" % step] = {"tensor": which, "engine_vs_f64": rel(grads[off_w:off_w + num_w] / grad_scale, g64[which].numpy().reshape(-1)),
                                        "oracle32_vs_f64": rel(np.asarray(grads_o[which]).reshape(-1), g64[which].numpy().reshape(-1))}
        if worst >= 1e-4:
            # a pre-activation within rounding noise of zero takes the other LeakyReLU slope in one of the two computations
            # (double accumulation here, fp32 in the oracle): accept only if the oracle reproduces these gradients once the
            # slopes of (at most 4) such elements are flipped - the same proof the GPU tests and smoke() use
            got = {pname: grads[off:off + numel].reshape(np.asarray(grads_o[pname]).shape) / np.float32(grad_scale)
                   for pname, shape, off, numel in eng.param_table}
            flips, w_after = O.explain_gradient_mismatch(cfg, o_par, mix, targets, got, tol=1e-4)
            explained = {"flips": None if flips is None else len(flips), "worst_after": float(w_after)}
        if step == 1 and os.environ.get("HOSTSIM_DUMP_GRADS"):
            np.save(os.environ["HOSTSIM_DUMP_GRADS"], grads[:n].copy())
        res["step%d" % step] = {
            "loss": float(loss[0]), "loss_oracle": float(loss_o), "loss_rel": abs(float(loss[0]) - loss_o) / abs(loss_o),
            "grad_worst_rel": worst, "grad_worst_tensor": which, "grad_explained": explained,
            "outputs_rel": max(rel(out_d.reshape(K, batch, t_out, C)[k], outs_o[s]) for k, s in enumerate(names))}
        # TF-form Adam (Training.py:77) from the engine's own state and gradient: what the update must be
        g_now, m_prev, v_prev, p_prev = grads[:n].copy(), m[:n].copy(), v[:n].copy(), par[:n].copy()
        wun.check(lib.wun_adam_step_device(h, P(par), P(grads), P(m), P(v), P(state), lr, 0.9, 0.999, 1e-8, MAIN))
        p_want, m_want, v_want = O.adam_update(p_prev, g_now, m_prev, v_prev, step, lr)
        res["step%d" % step]["params_rel"] = rel(par[:n], p_want)
        # (the slots agree to ~1e-5 only: the kernels - like TF's ApplyAdam - form 1 - beta2 in float32, the oracle in double)
        res["step%d" % step]["adam_slots_rel"] = max(rel(m[:n], m_want), rel(v[:n], v_want))
    o_par = engine_params()
    if os.environ.get("HOSTSIM_OTHER_BATCH") and batch > 1:
        # the same handle at another batch size (the last batch of an epoch) and back: its own workspace, other planner choices
        import torch
        worst_other = 0.0
        for nb in (batch - 1, batch):
            wsb = eng.workspace_bytes(nb, True)
            ws2 = aligned(wsb // 4 + 64)
            mix2 = aligned(nb * t_in * C); mix2[:] = mix[:nb].reshape(-1)
            tg2 = aligned(K * nb * t_out * C); tg2[:] = tg[:, :nb].reshape(-1)
            wun.check(lib.wun_forward_backward(h, P(par), P(mix2), P(tg2), nb, None, P(loss), P(grads), 1.0, P(ws2), wsb, MAIN))
            tp = O._as_torch(o_par, torch.float64, True)
            outs64 = O.forward(cfg, tp, torch.as_tensor(mix[:nb]).to(torch.float64), True)
            l64 = O.mse_loss(cfg, outs64, {k: torch.as_tensor(vv[:nb]).to(torch.float64) for k, vv in targets.items()})
            g64 = {k: g.numpy() for k, g in zip(tp, torch.autograd.grad(l64, list(tp.values())))}
            worst_other = max([worst_other, abs(float(loss[0]) - float(l64)) / abs(float(l64))] +
                              [rel(grads[off:off + numel], g64[pname].reshape(-1)) for pname, shape, off, numel in eng.param_table])
        res["other_batch_worst_rel"] = worst_other
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
    if os.environ.get("HOSTSIM_DUMP_TRACE"):
        open(os.environ["HOSTSIM_DUMP_TRACE"], "w").write(buf.value.decode())
    kinds = {}
    for ln in buf.value.decode().splitlines():
        if ln.startswith("L "):
            name = ln.split(" ")[2]
            if name.startswith("UNKNOWN:"):
                res["unknown"].append(name)
            if name.startswith("FOOTPRINT-MISS"):
                res["footprint_miss"].append(name[:160])
                name = name.split(":", 1)[1]
            k = name.split("(")[0].replace("void_", "").replace("wun::", "").replace("UNKNOWN:", "")
            kinds[k] = kinds.get(k, 0) + 1
    res["kernels"] = kinds
    print(json.dumps(res))


if __name__ == "__main__":
    main()
