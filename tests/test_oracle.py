"""CPU tests that pin the oracle (oracle/wave_unet_oracle.py).

1. against fixtures produced by the reference's own code (tests/golden/make_golden.py),
2. hand-computed micro cases that make each restated TF leaf semantic observable,
3. fp64 finite differences of the oracle's backward,
4. an independent numpy im2col formulation of the conv.
"""
import glob
import json
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import wave_unet_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if not p.endswith("feeder.npz"))


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = json.loads(str(z["cfg_json"]))
    params = OrderedDict((n, z["param/" + n]) for n in z["param_names"])
    targets = OrderedDict((s, z["target/" + s]) for s in cfg["source_names"])
    return z, cfg, params, targets


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def test_fixtures_present():
    assert len(CASES) >= 6


def test_get_padding_matches_reference():
    rows = json.load(open(os.path.join(GOLDEN, "padding.json")))
    import Config
    n_ok = 0
    for r in rows:
        if "preset" in r:
            cfg = Config.build_config([r["preset"]], experiment_id=0)["model_config"]
        else:
            cfg = Config.build_config(["baseline_context"], r["overrides"], experiment_id=0)["model_config"]
        if "error" in r:
            with pytest.raises(AssertionError):
                O.get_padding(cfg, r["num_frames"])
        else:
            t_in, t_out = O.get_padding(cfg, r["num_frames"])
            assert [t_in, t_out] == [r["in_shape"][1], r["out_shape"][1]], r
            assert O.num_channels(cfg) == r["in_shape"][2]
            n_ok += 1
    assert n_ok > 80


def test_baseline_json_shapes():
    """BASELINE.json configs: M4/M5/M6 -> 147443 in / 16389 out."""
    import Config
    for p in ["baseline_stereo", "full", "full_multi_instrument", "full_44KHz"]:
        cfg = Config.build_config([p], experiment_id=0)["model_config"]
        assert O.get_padding(cfg, cfg["num_frames"]) == (147443, 16389)
    cfg = Config.build_config(["baseline"], experiment_id=0)["model_config"]
    assert O.get_padding(cfg, 16384) == (16384, 16384)


def test_param_counts():
    """10 263 390 (M4) / 10 265 550 (M5) / 10 263 498 (M6) / 10 263 028 (M1)."""
    import Config
    want = {"baseline_stereo": 10263390, "full": 10265550, "full_multi_instrument": 10263498,
            "baseline": 10263028}
    for p, n in want.items():
        cfg = Config.build_config([p], experiment_id=0)["model_config"]
        tot = sum(int(np.prod(s)) for _, s in O.param_table(cfg))
        assert tot == n, (p, tot)


@pytest.mark.parametrize("name", CASES)
def test_oracle_fp64_matches_reference_graph(name):
    z, cfg, params, targets = load_case(name)
    assert [n for n, _ in O.param_table(cfg)] == list(params.keys())      # TF creation order + names
    for n, s in O.param_table(cfg):
        assert tuple(params[n].shape) == tuple(s)
    loss, outs, grads = O.forward_backward(cfg, params, z["mix"], targets, dtype=torch.float64)
    assert abs(loss - float(z["loss"])) <= 1e-12 * max(1, abs(float(z["loss"])))
    for s in cfg["source_names"]:
        assert list(outs.keys()) == cfg["source_names"]
        np.testing.assert_allclose(outs[s], z["out_train/" + s], rtol=0, atol=1e-12)
    for n in params:
        g = z["grad/" + n]
        assert rel_l2(grads[n], g) < 1e-10, n
    test_outs = O.forward_np(cfg, params, z["mix"], False, dtype=torch.float64)
    for s in cfg["source_names"]:
        np.testing.assert_allclose(test_outs[s], z["out_test/" + s], rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", CASES)
def test_oracle_fp32_noise_floor(name):
    z, cfg, params, targets = load_case(name)
    loss, outs, grads = O.forward_backward(cfg, params, z["mix"], targets, dtype=torch.float32)
    for s in cfg["source_names"]:
        assert rel_l2(outs[s], z["out_train/" + s]) < 2e-6
    for n in params:
        assert rel_l2(grads[n], z["grad/" + n]) < 2e-5, n


# ---------------------------------------------------------------------------------------------------
# hand-computed micro cases (expected values written out literally)
# ---------------------------------------------------------------------------------------------------
def t64(a):
    return torch.tensor(a, dtype=torch.float64)


def test_micro_conv_valid_and_same():
    x = t64([[[1.], [2.], [3.], [4.], [5.]]])                 # [1,5,1]
    k = t64([[[1.]], [[0.]], [[-1.]]])                        # [3,1,1]: y[t] = x[t] - x[t+2]
    b = t64([0.5])
    y = O.conv1d(x, k, b, "valid")
    assert y[0, :, 0].tolist() == [-1.5, -1.5, -1.5]
    y = O.conv1d(x, k, b, "same")                             # pad 1/1
    assert y[0, :, 0].tolist() == [0 - 2 + .5, 1 - 3 + .5, 2 - 4 + .5, 3 - 5 + .5, 4 - 0 + .5]
    # even kernel: TF SAME pads (k-1)//2 left, rest right -> k=2: 0 left, 1 right
    k2 = t64([[[1.]], [[10.]]])
    y = O.conv1d(x, k2, t64([0.]), "same")
    assert y[0, :, 0].tolist() == [21., 32., 43., 54., 5.]


def test_micro_conv_is_cross_correlation_multi_channel():
    x = t64([[[1., 2.], [3., 4.], [5., 6.]]])                 # [1,3,2]
    k = torch.zeros(2, 2, 3, dtype=torch.float64)
    k[0, 0, 0] = 1.0        # out0 += x[t,0]
    k[1, 1, 0] = 2.0        # out0 += 2*x[t+1,1]
    k[0, 1, 1] = -1.0       # out1 = -x[t,1]
    k[1, 0, 2] = 1.0        # out2 = x[t+1,0]
    y = O.conv1d(x, k, t64([0., 0., 100.]), "valid")
    assert y[0].tolist() == [[1 + 8., -2., 103.], [3 + 12., -4., 105.]]


def test_micro_leaky_relu_and_subgradient():
    x = torch.tensor([-2.0, 0.0, 3.0], dtype=torch.float64, requires_grad=True)
    y = O.leaky_relu(x)
    assert y.tolist() == [-0.4, 0.0, 3.0]
    y.sum().backward()
    assert x.grad.tolist() == [0.2, 0.2, 1.0]                 # slope 0.2 AT zero (MaximumGrad)


def test_micro_upsample_linear():
    x = t64([[[0.], [2.], [8.]]])
    assert O.upsample_linear(x, True)[0, :, 0].tolist() == [0., 1., 2., 5., 8.]          # 2N-1
    assert O.upsample_linear(x, False)[0, :, 0].tolist() == [0., 1., 2., 5., 8., 8.]      # 2N, clamp


def test_upsample_linear_context_vs_an_independent_align_corners_resize():
    """A second, independent implementation of the align_corners=True bilinear resize the context model uses
    (UnetAudioSeparator.py:115: tf.image.resize_bilinear(..., [1, 2N-1], align_corners=True)): torch's own
    F.interpolate(mode="linear", align_corners=True) - sample positions i * (N-1)/(2N-2) = i/2, the definition both libraries
    document.  The oracle, the golden fixtures' stand-in (tests/golden/tf_shim.py) and torch must agree."""
    import sys
    sys.path.insert(0, GOLDEN)
    import tf_shim
    rng = np.random.default_rng(11)
    for n, c in ((2, 1), (9, 3), (64, 5), (517, 2)):
        x = torch.as_tensor(rng.standard_normal((2, n, c)))
        want = torch.nn.functional.interpolate(x.transpose(1, 2), size=2 * n - 1, mode="linear", align_corners=True).transpose(1, 2)
        np.testing.assert_allclose(O.upsample_linear(x, True).numpy(), want.numpy(), rtol=0, atol=1e-14)
        shim = tf_shim._resize_bilinear(tf_shim.T(x.unsqueeze(1)), [1, 2 * n - 1], align_corners=True).t[:, 0]
        np.testing.assert_allclose(shim.numpy(), want.numpy(), rtol=0, atol=1e-14)


def test_micro_upsample_learned():
    x = t64([[[1., 10.], [3., 30.], [5., 50.]]])
    var = t64([0.0, np.log(3.0)])                             # sigmoid -> 0.5, 0.75
    v = O.upsample_learned(x, var, "valid")
    np.testing.assert_allclose(v[0].numpy(), [[1, 10], [2, 15], [3, 30], [4, 35], [5, 50]], atol=1e-12)
    s = O.upsample_learned(x, var, "same")                    # x[N] = 0 on the right
    np.testing.assert_allclose(s[0].numpy(), [[1, 10], [2, 15], [3, 30], [4, 35], [5, 50], [2.5, 37.5]],
                               atol=1e-12)


def test_micro_crop():
    x = torch.arange(10, dtype=torch.float64).reshape(1, 10, 1)
    assert O.crop(x, 4)[0, :, 0].tolist() == [3., 4., 5., 6.]
    assert O.crop(x, 5)[0, :, 0].tolist() == [2., 3., 4., 5., 6.]   # odd diff: extra frame off the END
    assert O.crop(x, 10) is x


def test_micro_difference_output_and_clip():
    """L=1 net, all-zero conv kernels: accompaniment = tanh(b), vocals = crop(mix) - tanh(b)."""
    import Config
    cfg = Config.build_config(["baseline_context"], dict(num_layers=1, num_initial_filters=1,
                                                         filter_size=3, merge_filter_size=3,
                                                         input_filter_size=3), experiment_id=0)["model_config"]
    t_in, t_out = O.get_padding(cfg, 2)
    params = OrderedDict((n, np.zeros(s, np.float64)) for n, s in O.param_table(cfg))
    last = [n for n in params if n.endswith("/bias")][-1]
    params[last][:] = 0.5
    mix = np.linspace(-2, 2, t_in).reshape(1, t_in, 1)
    out = O.forward_np(cfg, params, mix, True, dtype=torch.float64)
    c = (t_in - t_out) // 2
    np.testing.assert_allclose(out["accompaniment"], np.tanh(0.5) * np.ones((1, t_out, 1)), atol=1e-15)
    np.testing.assert_allclose(out["vocals"], mix[:, c:c + t_out] - np.tanh(0.5), atol=1e-15)
    out_t = O.forward_np(cfg, params, mix, False, dtype=torch.float64)
    np.testing.assert_allclose(out_t["vocals"], np.clip(mix[:, c:c + t_out] - np.tanh(0.5), -1, 1), atol=1e-15)


def test_micro_adam_tf_form():
    p, g = np.array([1.0]), np.array([0.5])
    p1, m1, v1 = O.adam_update(p, g, np.zeros(1), np.zeros(1), 1, 1e-4)
    # t=1: m=.05, v=2.5e-4*... lr_t = 1e-4*sqrt(1-.999)/(1-.9); p -= lr_t*m/(sqrt(v)+1e-8)
    m, v = 0.05, 0.001 * 0.25
    lr_t = 1e-4 * np.sqrt(1 - 0.999) / (1 - 0.9)
    assert abs(m1[0] - m) < 1e-15 and abs(v1[0] - v) < 1e-15
    assert abs(p1[0] - (1.0 - lr_t * m / (np.sqrt(v) + 1e-8))) < 1e-15


def test_adam_tf_form_vs_torch_adam_through_the_epsilon_identity():
    """Independent check of the TF-form Adam (tf.train.AdamOptimizer, Training.py:77): torch.optim.Adam's update
    lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps) equals TF's lr*sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps_tf) exactly when
    eps_tf = eps * sqrt(1-b2^t).  Five steps with that per-step epsilon must reproduce torch's trajectory; with the SAME epsilon
    the two differ (the epsilon placement is what the restatement has to get right)."""
    rng = np.random.default_rng(3)
    p0 = rng.standard_normal(64)
    grads = [rng.standard_normal(64) * 10.0 ** rng.integers(-9, 1, size=64) for _ in range(5)]      # tiny gradients make eps matter
    eps = 1e-8
    tp = torch.nn.Parameter(torch.as_tensor(p0.copy()))
    opt = torch.optim.Adam([tp], lr=1e-3, betas=(0.9, 0.999), eps=eps)
    p, m, v = p0.copy(), np.zeros(64), np.zeros(64)
    q, mq, vq = p0.copy(), np.zeros(64), np.zeros(64)
    for t, g in enumerate(grads, 1):
        tp.grad = torch.as_tensor(g.copy())
        opt.step()
        p, m, v = O.adam_update(p, g, m, v, t, 1e-3, eps=eps * np.sqrt(1 - 0.999 ** t))
        q, mq, vq = O.adam_update(q, g, mq, vq, t, 1e-3, eps=eps)
    np.testing.assert_allclose(p, tp.detach().numpy(), rtol=1e-12, atol=1e-15)
    assert np.abs(q - tp.detach().numpy()).max() > 1e-6


def test_conv_vs_numpy_im2col():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 37, 5))
    k = rng.standard_normal((7, 5, 3))
    b = rng.standard_normal(3)
    y = O.conv1d(t64(x), t64(k), t64(b), "valid").numpy()
    To = 37 - 6
    cols = np.stack([x[:, t:t + 7, :].reshape(2, -1) for t in range(To)], axis=1)     # [B,To,k*Cin]
    ref = cols @ k.reshape(-1, 3) + b
    np.testing.assert_allclose(y, ref, atol=1e-12)


def test_finite_difference_gradients():
    """fp64 central differences of the oracle loss wrt a few parameter entries of each tensor."""
    z, cfg, params, targets = load_case("ctx_learned_diff_multi")
    p64 = OrderedDict((k, v.astype(np.float64)) for k, v in params.items())
    _, _, grads = O.forward_backward(cfg, p64, z["mix"], targets, dtype=torch.float64)
    rng = np.random.default_rng(1)

    def loss_of(pp):
        with torch.no_grad():
            tp = O._as_torch(pp, torch.float64, False)
            outs = O.forward(cfg, tp, torch.tensor(z["mix"], dtype=torch.float64), True)
            tg = {k: torch.tensor(v, dtype=torch.float64) for k, v in targets.items()}
            return float(O.mse_loss(cfg, outs, tg))

    for name in p64:
        flat = p64[name].reshape(-1)
        for idx in rng.choice(flat.size, size=min(2, flat.size), replace=False):
            old = flat[idx]
            h = 1e-6
            flat[idx] = old + h
            lp = loss_of(p64)
            flat[idx] = old - h
            lm = loss_of(p64)
            flat[idx] = old
            fd = (lp - lm) / (2 * h)
            an = grads[name].reshape(-1)[idx]
            assert abs(fd - an) <= 1e-6 * max(1.0, abs(an)) + 1e-9, (name, idx, fd, an)


def test_predict_track_tiling():
    z, cfg, params, _ = load_case("ctx_linear_diff_stereo")
    T_in, T_out = int(z["T_in"]), int(z["T_out"])
    rng = np.random.default_rng(3)
    n = 8 * T_out + 7                                   # forces a shifted last window
    audio = rng.uniform(-1, 1, size=(n, 2)).astype(np.float32)
    preds = O.predict_track(cfg, params, audio, T_in, T_out)
    pad = (T_in - T_out) // 2
    padded = np.pad(audio, [(pad, pad), (0, 0)])
    # first window and shifted last window reproduce single-window forward outputs
    w0 = O.forward_np(cfg, params, padded[None, 0:T_in], False)
    wl = O.forward_np(cfg, params, padded[None, n - T_out:n - T_out + T_in], False)
    for s in cfg["source_names"]:
        assert preds[s].shape == audio.shape
        np.testing.assert_array_equal(preds[s][:T_out], w0[s][0])
        np.testing.assert_array_equal(preds[s][n - T_out:], wl[s][0])
    # short input gets padded to T_in and cut back
    short = audio[:T_in // 2]
    ps = O.predict_track(cfg, params, short, T_in, T_out)
    assert ps["vocals"].shape == short.shape


def test_mask_flip_analysis_explains_exactly_the_flipped_elements():
    """The checker used by __graft_entry__.smoke(): a gradient set computed with the LeakyReLU slope of a near-zero
    pre-activation inverted (what a not-bit-identical forward does) is recognised as such; a scaled tensor is not."""
    import Config
    cfg = Config.build_config(["baseline_stereo"], dict(num_layers=4), experiment_id=0)["model_config"]
    t_in, t_out = O.get_padding(cfg, 64)
    params = O.init_params(cfg, seed=1337)
    mix, targets = O.synthetic_batch(cfg, 2, t_in, t_out, seed=1)
    near = O.near_zero_preactivations(cfg, params, mix, 1e-4)
    assert near and near == sorted(near) and near[0][0] < 1e-6          # this seed has one within fp32 noise of zero
    _, _, g0 = O.forward_backward(cfg, params, mix, targets)
    flip = (near[0][1], near[0][2])
    _, _, g1 = O.forward_backward_with_flips(cfg, params, mix, targets, [flip])
    first = list(g0)[0]
    e = np.linalg.norm(g1[first] - g0[first]) / np.linalg.norm(g0[first])
    assert 1e-3 < e < 5e-2                                               # one element moves the first-layer kernel gradient by 6.6e-3
    assert O.explain_gradient_mismatch(cfg, params, mix, targets, g0) == ((), 0.0)
    flips, w = O.explain_gradient_mismatch(cfg, params, mix, targets, g1)
    assert flips == (flip,) and w < 1e-6
    bad = {k: v.copy() for k, v in g0.items()}
    bad[first] *= 1.01
    flips, w = O.explain_gradient_mismatch(cfg, params, mix, targets, bad)
    assert flips is None and w > 5e-3
    # the instrumentation leaves the oracle untouched
    _, _, g2 = O.forward_backward(cfg, params, mix, targets)
    assert all(np.array_equal(g2[k], g0[k]) for k in g0)


def test_backward_with_dictated_slopes_and_live_rows():
    """forward_backward_with_masks (the GPU gradient tests' proof step): with the oracle's OWN signs at the rows the
    engine keeps (tests/helpers.live_rows) the gradients are unchanged and no slope differs; with one sign inverted the
    report names it and the gradients move."""
    import Config
    from helpers import live_rows
    for named, ov, nf in ((["full"], dict(num_layers=3, num_initial_filters=6), 40),
                          (["baseline"], dict(num_layers=3, num_initial_filters=6), 64)):
        cfg = Config.build_config(named, ov, experiment_id=0)["model_config"]
        t_in, t_out = O.get_padding(cfg, nf)
        params = O.init_params(cfg, seed=3)
        mix, targets = O.synthetic_batch(cfg, 2, t_in, t_out, seed=4)
        with torch.no_grad():
            _, inter = O.forward(cfg, O._as_torch(params, torch.float32, False), torch.from_numpy(mix), True,
                                 return_intermediates=True)
        L = cfg["num_layers"]
        tensors = [inter["down%d" % i] for i in range(L)] + [inter["bottleneck"]] + [inter["up%d" % i] for i in range(L)]
        rows = live_rows(cfg, t_in)
        assert sorted(rows) == list(range(2 * L + 1))
        masks = {}
        for idx, views in rows.items():
            assert all(r.max() < tensors[idx].shape[1] for _, r in views if len(r))
            if idx >= L:
                assert len(views[0][1]) == tensors[idx].shape[1]
            masks[idx] = [(r, (tensors[idx][:, r, :] > 0).numpy()) for _, r in views]
        loss0, _, g0 = O.forward_backward(cfg, params, mix, targets)
        loss1, _, g1, rep = O.forward_backward_with_masks(cfg, params, mix, targets, masks)
        assert loss0 == loss1 and all(r[1] == 0 for r in rep)
        for n in g0:
            np.testing.assert_array_equal(g0[n], g1[n])
        masks[L][0][1][0, 0, 0] ^= True
        _, _, g2, rep = O.forward_backward_with_masks(cfg, params, mix, targets, masks)
        assert [r for r in rep if r[0] == L][0][1] == 1          # (later layers see the changed forward value too)
        assert any(np.abs(g2[n] - g0[n]).max() > 0 for n in g0)
