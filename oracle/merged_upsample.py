"""TEST INFRASTRUCTURE (like everything under oracle/): the algebra behind the round-2 plan of folding the x2 upsampling of
the up blocks (UnetAudioSeparator.py:109-118, InterpolationLayer.py:19-39) into the merge convolution's weights.

The up block convolves the interleaved sequence  U[2s] = x[s],  U[2s+1] = a*x[s] + (1-a)*x[s+1]   (a = 0.5 for linear
upsampling, a = sigmoid(var)[c] per input channel for the learned layer) with k taps:

    y[u] = sum_j  W_j . U[u + j - pad_left]

Splitting the output rows by parity, u = 2m + p, every U row is a fixed combination of x[m + s], so

    y[2m + p] = sum_s  W'_{p,s} . x[m + s]          with   W'_{p,s}[c, n] = sum_j coef_{p,s,j}[c] * W_j[c, n]

and coef in {1, a_c, 1 - a_c}.  For k = 5 that is 3 (p = 0) and 4 (p = 1) taps on x instead of 5 + 5 taps on copy / mid
planes: fewer MMAs, no interpolated tensor, and the dgrad of the merged conv IS the gradient w.r.t. x (no g_ue / g_mid /
upsample-backward pass).  The gradients of the original parameters follow by the chain rule from dW':

    dW_j[c, n]  = sum_{p,s} coef_{p,s,j}[c] * dW'_{p,s}[c, n]
    da[c]       = sum_{p,s,j,n} dcoef_{p,s,j}/da * W_j[c, n] * dW'_{p,s}[c, n]        (dcoef/da = +1 for a, -1 for 1-a)
    dvar[c]     = da[c] * a_c * (1 - a_c)

tests/test_merged_upsample.py checks all of it against the oracle's upsample + conv1d and autograd (valid / context mode)."""
from collections import OrderedDict

ONE, A, B = 0, 1, 2          # coefficient kinds: 1, a, 1 - a


def merged_taps(k, parity, pad_left=0):
    """OrderedDict shift -> [(j, kind)] : which original taps feed the merged tap that reads x[m + shift] for output rows
    2m + parity."""
    taps = {}
    for j in range(k):
        e = parity + j - pad_left
        if e % 2 == 0:
            taps.setdefault(e // 2, []).append((j, ONE))
        else:
            s = (e - 1) // 2
            taps.setdefault(s, []).append((j, A))          # mid[s] = a*x[s] + (1-a)*x[s+1]
            taps.setdefault(s + 1, []).append((j, B))
    return OrderedDict(sorted(taps.items()))


def coef(kind, a):
    return 1.0 if kind == ONE else (a if kind == A else 1.0 - a)


def merged_weights(W, a, parity, pad_left=0):
    """W: [k, C, N] (torch or numpy), a: scalar or [C] -> OrderedDict shift -> [C, N] merged weight."""
    out = OrderedDict()
    for s, srcs in merged_taps(W.shape[0], parity, pad_left).items():
        acc = 0
        for j, kind in srcs:
            c = coef(kind, a)
            acc = acc + (W[j] * (c[:, None] if hasattr(c, "shape") and len(getattr(c, "shape", ())) == 1 else c))
        out[s] = acc
    return out


def original_gradients(dWm, W, a, k, pad_left=0):
    """dWm: {parity: {shift: dL/dW'_{p,s} [C, N]}} -> (dW [k, C, N], da [C] or scalar 0-d) by the chain rule above."""
    dW = [0 for _ in range(k)]
    da = 0
    for p in (0, 1):
        for s, srcs in merged_taps(k, p, pad_left).items():
            g = dWm[p][s]
            for j, kind in srcs:
                c = coef(kind, a)
                dW[j] = dW[j] + g * (c[:, None] if hasattr(c, "shape") and len(getattr(c, "shape", ())) == 1 else c)
                if kind == A:
                    da = da + (W[j] * g).sum(-1)
                elif kind == B:
                    da = da - (W[j] * g).sum(-1)
    return dW, da
