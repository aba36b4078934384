"""CPU ORACLE for the Wave-U-Net forward/backward hot path.  TEST INFRASTRUCTURE ONLY.

This file is a restatement, on torch-CPU tensors (fp32 or fp64), of the arithmetic the reference
performs on the path named by BASELINE.json:

    Models/UnetAudioSeparator.py:15-144   (ctor keys, get_padding, get_output)
    Models/InterpolationLayer.py:4-40     (learned interpolation)
    Models/OutputLayer.py:5-23            (independent / difference outputs)
    Utils.py:11-24, 79-92, 104-123        (crop_and_concat, LeakyReLU, AudioClip, crop)
    Training.py:50-77                     (MSE loss, Adam)
    Evaluate.py:82-145                    (predict_track tiling)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
it; the product (wave-u-net_b200/) never does.

PINNING STATUS.  The reference has no tests, golden vectors or fixtures for this path and its
arithmetic lives in tensorflow==1.8.0 (requirements.txt:3), which cannot be installed here.  What
pins this oracle instead:
  * get_padding is checked against the reference's OWN get_padding, imported unmodified from
    /root/reference (it only needs numpy) - tests/golden/make_golden.py, tests/golden/*.npz.
  * the graph wiring (layer order, crop offsets, concat order, gather interleave, output-layer
    algebra) is checked against the reference's OWN get_output, imported unmodified and executed over
    a small eager "tensorflow" stand-in (tests/golden/tf_shim.py) whose leaf ops follow TF-1.8 op
    definitions; outputs AND autograd gradients are committed as fixtures.
  * the TF leaf-op semantics themselves (SAME-padding split, legacy resize_bilinear, Maximum
    sub-gradient, Adam epsilon placement) are restated from the TF-1.8 op definitions; they are made
    observable by hand-computed micro cases in tests/test_oracle.py.
So: wiring pinned by the reference's code, leaf kernels restated => "parity pinned at graph level,
leaf ops unpinned".

Tensor layout everywhere: channels-last [B, T, C] like the reference (UnetAudioSeparator.py:88).
Conv kernels: [k, C_in, C_out] (tf.layers.conv1d kernel layout).  Parameters are an ordered dict in
TF variable-creation order with TF auto-names under scope "separator".
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

LEAK = 0.2  # Utils.py:79


# ----------------------------------------------------------------------------------------------
# shapes
# ----------------------------------------------------------------------------------------------
def get_padding(cfg, num_frames):
    """(T_in, T_out) for a desired output length.  Follows UnetAudioSeparator.py:34-83."""
    if not cfg["context"]:
        return int(num_frames), int(num_frames)
    L = cfg["num_layers"]
    fs, mfs = cfg["filter_size"], cfg["merge_filter_size"]
    ifs, ofs = cfg["input_filter_size"], cfg["output_filter_size"]
    rem = float(num_frames)
    rem = rem - ofs + 1                                  # :46
    for _ in range(L):                                   # :49-51
        rem = rem + mfs - 1
        rem = (rem + 1.0) / 2.0
    x = int(math.ceil(rem))                              # :54
    assert x >= 2                                        # :55
    out = x
    inp = x + fs - 1                                     # :62
    for i in range(L):                                   # :65-73
        out = 2 * out - 1
        out = out - mfs + 1
        inp = 2 * inp - 1
        inp = inp + (fs if i < L - 1 else ifs) - 1
    out = out - ofs + 1                                  # :76
    return int(inp), int(out)


def num_channels(cfg):
    return 1 if cfg["mono_downmix"] else 2               # :31


def n_output_convs(cfg):
    K = len(cfg["source_names"])
    if cfg["output_type"] == "direct":
        return K
    if cfg["output_type"] == "difference":
        return K - 1
    raise NotImplementedError


def param_table(cfg):
    """[(name, shape)] in TF creation order (get_output, :92-142)."""
    L, F0 = cfg["num_layers"], cfg["num_initial_filters"]
    fs, mfs, ofs = cfg["filter_size"], cfg["merge_filter_size"], cfg["output_filter_size"]
    C = num_channels(cfg)
    tab = []
    n = [0]

    def conv(k, cin, cout):
        sfx = "" if n[0] == 0 else "_%d" % n[0]
        tab.append(("separator/conv1d%s/kernel" % sfx, (k, cin, cout)))
        tab.append(("separator/conv1d%s/bias" % sfx, (cout,)))
        n[0] += 1

    cin = C
    for i in range(L):                                   # :97-100
        conv(fs, cin, F0 * (i + 1))
        cin = F0 * (i + 1)
    conv(fs, cin, F0 * (L + 1))                          # :102
    cur = F0 * (L + 1)
    for i in range(L):                                   # :107-125
        if cfg["upsampling"] == "learned":
            tab.append(("separator/interp_%d" % i, (cur,)))
        skip = F0 * (L - i)
        conv(mfs, skip + cur, F0 * (L - i))
        cur = F0 * (L - i)
    for _ in range(n_output_convs(cfg)):                 # OutputLayer.py:8,15
        conv(ofs, C + cur, C)
    return tab


def init_params(cfg, seed=1337, dtype=np.float32):
    """Glorot-uniform kernels, zero biases (tf.layers.conv1d defaults), glorot interp vars."""
    rng = np.random.default_rng(seed)
    p = OrderedDict()
    for name, shape in param_table(cfg):
        if name.endswith("/bias"):
            p[name] = np.zeros(shape, dtype)
        elif name.endswith("/kernel"):
            k, cin, cout = shape
            lim = math.sqrt(6.0 / (k * cin + k * cout))
            p[name] = rng.uniform(-lim, lim, size=shape).astype(dtype)
        else:  # interp_<level>, shape [F]: fan_in = fan_out = F
            lim = math.sqrt(6.0 / (2 * shape[0]))
            p[name] = rng.uniform(-lim, lim, size=shape).astype(dtype)
    return p


# ----------------------------------------------------------------------------------------------
# leaf ops (TF-1.8 semantics, restated)
# ----------------------------------------------------------------------------------------------
def leaky_relu(x):
    """tf.maximum(0.2*x, x) (Utils.py:79-80).  TF's MaximumGrad routes the gradient to the FIRST
    argument where 0.2*x >= x, i.e. slope 0.2 at x == 0 - identical to torch's leaky_relu."""
    return F.leaky_relu(x, LEAK)


def audio_clip(x, training):
    """Utils.py:82-92"""
    return x if training else torch.clamp(x, -1.0, 1.0)


def conv1d(x, kernel, bias, padding):
    """tf.layers.conv1d, stride 1: cross-correlation of [B,T,Cin] with [k,Cin,Cout] (+bias).
    'same': total pad k-1, left = (k-1)//2, remainder on the right (TF SAME rule)."""
    k = kernel.shape[0]
    xt = x.transpose(1, 2)                               # [B,Cin,T]
    if padding == "same":
        left = (k - 1) // 2
        xt = F.pad(xt, (left, k - 1 - left))
    w = kernel.permute(2, 1, 0)                          # [Cout,Cin,k]
    return F.conv1d(xt, w, bias).transpose(1, 2)


def crop(x, target_len):
    """Centre crop on the time axis (Utils.py:104-123); odd difference drops the extra frame at
    the end."""
    diff = x.shape[1] - target_len
    assert diff >= 0
    if diff == 0:
        return x
    start = diff // 2
    end = diff - start
    return x[:, start:x.shape[1] - end, :]


def upsample_linear(x, context):
    """UnetAudioSeparator.py:114-117 via tf.image.resize_bilinear on a height-1 image.
    context: align_corners=True to 2N-1 -> src = dst*0.5 exactly.
    else   : legacy (no half-pixel) resize to 2N -> src = dst*0.5, upper index clamped to N-1.
    TF evaluates lerp as  left + (right-left)*frac."""
    N = x.shape[1]
    if context:
        mid = x[:, :-1] + (x[:, 1:] - x[:, :-1]) * 0.5
        out = x.new_empty(x.shape[0], 2 * N - 1, x.shape[2])
        out[:, 0::2] = x
        out[:, 1::2] = mid
        return out
    right = torch.cat([x[:, 1:], x[:, -1:]], dim=1)
    mid = x + (right - x) * 0.5
    out = x.new_empty(x.shape[0], 2 * N, x.shape[2])
    out[:, 0::2] = x
    out[:, 1::2] = mid
    return out


def upsample_learned(x, var, padding):
    """InterpolationLayer.py:4-40.  w = sigmoid(var); the [1,2,F,F] filter [diag(w); diag(1-w)]
    gives mid[s] = w*x[s] + (1-w)*x[s+1].  'valid': N-1 mids.  'same': TF SAME pads a width-2 filter
    with 0 on the left and 1 on the right -> x[N] = 0, N mids.  Interleave starts with x[0]."""
    w = torch.sigmoid(var)
    cw = 1.0 - w
    N = x.shape[1]
    if padding == "valid":
        mid = x[:, :-1] * w + x[:, 1:] * cw
        out = x.new_empty(x.shape[0], 2 * N - 1, x.shape[2])
    else:
        nxt = torch.cat([x[:, 1:], torch.zeros_like(x[:, :1])], dim=1)
        mid = x * w + nxt * cw
        out = x.new_empty(x.shape[0], 2 * N, x.shape[2])
    out[:, 0::2] = x
    out[:, 1::2] = mid
    return out


# ----------------------------------------------------------------------------------------------
# the network
# ----------------------------------------------------------------------------------------------
def _as_torch(params, dtype, requires_grad):
    out = OrderedDict()
    for k, v in params.items():
        t = torch.as_tensor(np.asarray(v)).to(dtype).clone()
        t.requires_grad_(requires_grad)
        out[k] = t
    return out


def forward(cfg, params, mix, training, return_intermediates=False):
    """get_output (UnetAudioSeparator.py:85-144).  `params`: ordered dict of torch tensors,
    `mix`: torch [B,T_in,C].  Returns OrderedDict name -> [B,T_out,C]."""
    L = cfg["num_layers"]
    padding = "valid" if cfg["context"] else "same"
    names = list(params.keys())
    it = iter(names)

    def next_conv():
        kn = next(it)
        bn = next(it)
        assert kn.endswith("/kernel") and bn.endswith("/bias"), (kn, bn)
        return params[kn], params[bn]

    inter = OrderedDict()
    enc = []
    cur = mix
    for i in range(L):                                                   # :97-100
        k, b = next_conv()
        cur = leaky_relu(conv1d(cur, k, b, padding))
        enc.append(cur)
        inter["down%d" % i] = cur
        cur = cur[:, ::2, :]
    k, b = next_conv()
    cur = leaky_relu(conv1d(cur, k, b, padding))                         # :102
    inter["bottleneck"] = cur
    for i in range(L):                                                   # :107-125
        if cfg["upsampling"] == "learned":
            var = params[next(it)]
            cur = upsample_learned(cur, var, padding)
        else:
            cur = upsample_linear(cur, cfg["context"])
        skip = enc[-i - 1]
        assert skip.shape[1] == cur.shape[1] or cfg["context"]           # :121
        cur = torch.cat([crop(skip, cur.shape[1]), cur], dim=2)          # Utils.py:23-24
        k, b = next_conv()
        cur = leaky_relu(conv1d(cur, k, b, padding))
        inter["up%d" % i] = cur
    cur = torch.cat([crop(mix, cur.shape[1]), cur], dim=2)               # :127

    act = cfg["output_activation"]
    if act == "tanh":
        out_act = torch.tanh
    elif act == "linear":
        out_act = lambda t: audio_clip(t, training)                      # :133-134
    else:
        raise NotImplementedError

    srcs = cfg["source_names"]
    outputs = OrderedDict()
    if cfg["output_type"] == "direct":                                   # OutputLayer.py:5-9
        for name in srcs:
            k, b = next_conv()
            outputs[name] = out_act(conv1d(cur, k, b, padding))
    elif cfg["output_type"] == "difference":                             # OutputLayer.py:11-23
        total = 0
        for name in srcs[:-1]:
            k, b = next_conv()
            o = out_act(conv1d(cur, k, b, padding))
            outputs[name] = o
            total = total + o
        last = crop(mix, total.shape[1]) - total
        outputs[srcs[-1]] = audio_clip(last, training)
    else:
        raise NotImplementedError
    if return_intermediates:
        return outputs, inter
    return outputs


def mse_loss(cfg, outputs, targets):
    """Training.py:50-63: sum_k mean((real_k - est_k)^2) / K."""
    loss = 0
    for name in cfg["source_names"]:
        loss = loss + torch.mean((targets[name] - outputs[name]) ** 2)
    return loss / float(len(cfg["source_names"]))


def forward_backward(cfg, params_np, mix_np, targets_np, dtype=torch.float32, training=True):
    """One forward + loss + backward.  Returns (loss, outputs(np), grads(OrderedDict np))."""
    params = _as_torch(params_np, dtype, True)
    mix = torch.as_tensor(mix_np).to(dtype)
    targets = {k: torch.as_tensor(v).to(dtype) for k, v in targets_np.items()}
    outputs = forward(cfg, params, mix, training)
    loss = mse_loss(cfg, outputs, targets)
    loss.backward()
    grads = OrderedDict((k, v.grad.detach().numpy().copy()) for k, v in params.items())
    outs = OrderedDict((k, v.detach().numpy().copy()) for k, v in outputs.items())
    return float(loss.detach()), outs, grads


# ---- LeakyReLU mask-flip analysis -------------------------------------------------------------------------------
# LeakyReLU's derivative jumps at 0, so a pre-activation that lies within the forward rounding noise of zero can get the
# other slope in an implementation that is not bit-identical to this one; one such element moves whole gradient
# tensors of a small net by ~1e-2 although every kernel is exact (measured: tools/grad_diag.py).  The two
# helpers below let a checker PROVE that a gradient mismatch is exactly that: list the pre-activations near zero, redo
# the backward with some of their slopes flipped, and compare again.
_FLIP_STATE = {"active": False, "call": 0, "flips": (), "record": None, "tol": 0.0, "masks": None, "mask_report": None}


def _leaky_relu_instrumented(x):
    st = _FLIP_STATE
    idx = st["call"]
    st["call"] += 1
    if st["record"] is not None:
        with torch.no_grad():
            rms = x.pow(2).mean().sqrt().clamp_min(1e-30)
            rel = (x.abs() / rms).reshape(-1)
            for flat in torch.nonzero(rel < st["tol"]).reshape(-1).tolist():
                st["record"].append((float(rel[flat]), idx, int(flat)))
    if st["masks"] is not None:
        # slopes dictated by another implementation at the rows it keeps (forward_backward_with_masks)
        mask = x.detach() > 0
        for rows, given in st["masks"].get(idx, ()):
            rows = torch.as_tensor(rows, dtype=torch.long)
            given = torch.as_tensor(given, dtype=torch.bool)
            own = mask[:, rows, :]
            diff = own != given
            n = int(diff.sum())
            worst = 0.0
            if n:
                rms = float(x.detach().pow(2).mean().sqrt().clamp_min(1e-30))
                worst = float((x.detach()[:, rows, :].abs() * diff).max()) / rms
            st["mask_report"].append((idx, n, worst))
            mask[:, rows, :] = given
        return torch.where(mask, x, LEAK * x)
    mine = [f for c, f in st["flips"] if c == idx]
    if not mine:
        return F.leaky_relu(x, LEAK)
    mask = (x.detach() > 0).reshape(-1).clone()
    for flat in mine:
        mask[flat] = ~mask[flat]
    return torch.where(mask.reshape(x.shape), x, LEAK * x)


def _run_instrumented(fn, flips=(), record=None, tol=0.0, masks=None, mask_report=None):
    global leaky_relu
    saved = leaky_relu
    _FLIP_STATE.update(active=True, call=0, flips=tuple(flips), record=record, tol=tol, masks=masks, mask_report=mask_report)
    leaky_relu = _leaky_relu_instrumented
    try:
        return fn()
    finally:
        leaky_relu = saved
        _FLIP_STATE.update(active=False, call=0, flips=(), record=None, tol=0.0, masks=None, mask_report=None)


def near_zero_preactivations(cfg, params_np, mix_np, tol=1e-4, dtype=torch.float64):
    """[(|x| / rms of its layer, LeakyReLU call index, flat element index)] of every pre-activation with |x| < tol*rms,
    smallest first.  Call index = order of the LeakyReLUs in forward(): down0..down(L-1), bottleneck, up0..up(L-1)."""
    rec = []
    with torch.no_grad():
        params = _as_torch(params_np, dtype, False)
        _run_instrumented(lambda: forward(cfg, params, torch.as_tensor(mix_np).to(dtype), True), record=rec, tol=tol)
    return sorted(rec)


def forward_backward_with_flips(cfg, params_np, mix_np, targets_np, flips, dtype=torch.float32):
    """forward_backward with the LeakyReLU slope of the listed (call index, flat index) elements inverted."""
    return _run_instrumented(lambda: forward_backward(cfg, params_np, mix_np, targets_np, dtype), flips=flips)


def forward_backward_with_masks(cfg, params_np, mix_np, targets_np, masks, dtype=torch.float32):
    """forward_backward where the LeakyReLU slope at given rows is DICTATED: masks = {call index: [(rows, bool[B,len(rows),C])]}
    (True = slope 1).  Used to prove that an implementation's gradients are exact for ITS OWN activation signs: the caller
    passes sign(saved activation) of every row the implementation keeps (rows it never computes get no gradient anyway).
    Returns (loss, outputs, grads, report) with report = [(call index, #elements whose slope differs from the oracle's own,
    largest |pre-activation| / layer rms among those)] - a faithful implementation differs only where the pre-activation
    is within forward rounding noise of zero."""
    report = []
    loss, outs, grads = _run_instrumented(lambda: forward_backward(cfg, params_np, mix_np, targets_np, dtype),
                                          masks=masks, mask_report=report)
    return loss, outs, grads, report


def explain_gradient_mismatch(cfg, params_np, mix_np, targets_np, grads_got, tol=1e-3, margin=1e-4, max_candidates=4):
    """Is `grads_got` (name -> array) the exact gradient for SOME assignment of slopes to the (at most max_candidates)
    pre-activations within margin*rms of zero?  Returns (flips, worst per-tensor rel-L2) of the best assignment, or
    (None, worst error without flips) if none brings every tensor within tol."""
    import itertools
    cands = [(c, f) for _, c, f in near_zero_preactivations(cfg, params_np, mix_np, margin)][:max_candidates]

    def worst(grads):
        return max(float(np.linalg.norm(np.asarray(grads_got[n], np.float64) - grads[n]) / max(np.linalg.norm(grads[n]), 1e-30))
                   for n in grads)

    base = worst(forward_backward(cfg, params_np, mix_np, targets_np)[2])
    if base <= tol:
        return (), base
    for k in range(1, len(cands) + 1):
        for subset in itertools.combinations(cands, k):
            w = worst(forward_backward_with_flips(cfg, params_np, mix_np, targets_np, subset)[2])
            if w <= tol:
                return tuple(subset), w
    return None, base


def forward_np(cfg, params_np, mix_np, training, dtype=torch.float32):
    with torch.no_grad():
        params = _as_torch(params_np, dtype, False)
        outs = forward(cfg, params, torch.as_tensor(mix_np).to(dtype), training)
    return OrderedDict((k, v.numpy().copy()) for k, v in outs.items())


def adam_update(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer (Training.py:77) update for one tensor, `step` = t >= 1 after increment.
    TF form: lr_t = lr*sqrt(1-b2^t)/(1-b1^t);  p -= lr_t * m / (sqrt(v) + eps)   (eps OUTSIDE sqrt,
    not bias-corrected - differs from torch.optim.Adam)."""
    p = np.asarray(p)
    dt = p.dtype
    g = np.asarray(g, dtype=dt)
    m2 = (beta1 * m + (1 - beta1) * g).astype(dt)
    v2 = (beta2 * v + (1 - beta2) * g * g).astype(dt)
    lr_t = dt.type(lr * math.sqrt(1 - beta2 ** step) / (1 - beta1 ** step))
    p2 = (p - lr_t * m2 / (np.sqrt(v2) + dt.type(eps))).astype(dt)
    return p2, m2, v2


# ----------------------------------------------------------------------------------------------
# synthetic batches (Utils.py:26-42, Datasets.py:207)
# ----------------------------------------------------------------------------------------------
def synthetic_batch(cfg, batch, T_in, T_out, seed=1337):
    """sources s_k ~ U(-1,1)/K iid; gain g_k ~ U(0.7,1) per example (Utils.py:33);
    mix = sum_k g_k s_k (Utils.py:35); targets centre-cropped by (T_in-T_out)//2 (Utils.py:38-42)."""
    rng = np.random.default_rng(seed)
    C = num_channels(cfg)
    srcs = cfg["source_names"]
    K = len(srcs)
    mix = np.zeros((batch, T_in, C), np.float32)
    targets = OrderedDict()
    cropf = (T_in - T_out) // 2
    for name in srcs:
        s = rng.uniform(-1.0, 1.0, size=(batch, T_in, C)).astype(np.float32) / np.float32(K)
        g = rng.uniform(0.7, 1.0, size=(batch, 1, 1)).astype(np.float32)
        s = s * g
        mix += s
        targets[name] = np.ascontiguousarray(s[:, cropf:T_in - cropf, :] if cropf > 0 else s)
    return mix, targets


# ----------------------------------------------------------------------------------------------
# windowed inference (Evaluate.py:82-145), resampling excluded (librosa; out of scope)
# ----------------------------------------------------------------------------------------------
def predict_track(cfg, params_np, mix_audio, T_in, T_out, dtype=torch.float32):
    """mix_audio: [n_frames, n_channels] already at expected_sr.  Batch-1 windows, hop T_out, last
    window shifted to the end, plain overwrite."""
    assert mix_audio.ndim == 2
    if cfg["mono_downmix"]:
        mix_audio = np.mean(mix_audio, axis=1, keepdims=True)            # :98-99
    elif mix_audio.shape[1] == 1:
        mix_audio = np.tile(mix_audio, [1, 2])                           # :101-102
    if mix_audio.shape[0] < T_in:                                        # :107-111
        extra = T_in - mix_audio.shape[0]
        mix_audio = np.pad(mix_audio, [(0, extra), (0, 0)], mode="constant")
    else:
        extra = 0
    n = mix_audio.shape[0]
    preds = OrderedDict((k, np.zeros(mix_audio.shape, np.float32)) for k in cfg["source_names"])
    pad = (T_in - T_out) // 2                                            # :121
    padded = np.pad(mix_audio, [(pad, pad), (0, 0)], mode="constant")
    for pos in range(0, n, T_out):                                       # :125-139
        if pos + T_out > n:
            pos = n - T_out
        part = padded[pos:pos + T_in, :][None].astype(np.float32)
        outs = forward_np(cfg, params_np, part, False, dtype)
        for k in cfg["source_names"]:
            preds[k][pos:pos + T_out] = outs[k][0]
    if extra > 0:
        preds = OrderedDict((k, v[:-extra, :]) for k, v in preds.items())
    return preds
