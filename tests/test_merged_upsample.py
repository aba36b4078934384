"""The algebra of oracle/merged_upsample.py (round-2 plan: fold the x2 upsampling of the up blocks into the merge conv's
weights) against the oracle's upsample + conv1d and torch autograd.  CPU, fp64."""
import numpy as np
import pytest
import torch

from oracle import merged_upsample as MU
from oracle import wave_unet_oracle as O


def test_tap_table_for_the_merge_filter():
    # k = 5 (merge_filter_size): 3 taps for even output rows, 4 for odd ones - instead of 5 + 5 on copy / mid planes
    t0, t1 = MU.merged_taps(5, 0), MU.merged_taps(5, 1)
    assert list(t0) == [0, 1, 2] and list(t1) == [0, 1, 2, 3]
    assert t0[0] == [(0, MU.ONE), (1, MU.A)] and t0[1] == [(1, MU.B), (2, MU.ONE), (3, MU.A)] and t0[2] == [(3, MU.B), (4, MU.ONE)]
    assert t1[0] == [(0, MU.A)] and t1[3] == [(4, MU.B)]
    # every original tap is used with total coefficient 1 per parity
    for p in (0, 1):
        tot = {}
        for s, srcs in MU.merged_taps(5, p).items():
            for j, kind in srcs:
                tot[j] = tot.get(j, 0.0) + MU.coef(kind, 0.3)
        assert all(abs(v - 1.0) < 1e-12 for v in tot.values()) and sorted(tot) == [0, 1, 2, 3, 4]


@pytest.mark.parametrize("learned", [False, True])
@pytest.mark.parametrize("k", [1, 3, 5, 7])
def test_merged_conv_equals_upsample_then_conv_forward_and_backward(learned, k):
    torch.manual_seed(k * 2 + learned)
    B, N, C, No = 2, 11, 6, 4
    x = torch.randn(B, N, C, dtype=torch.float64, requires_grad=True)
    W = torch.randn(k, C, No, dtype=torch.float64, requires_grad=True)
    bias = torch.zeros(No, dtype=torch.float64)
    var = torch.randn(C, dtype=torch.float64, requires_grad=True)
    # ---- reference: interleave, then a valid k-tap conv (UnetAudioSeparator.py:109-125 without the skip half) ----
    up = O.upsample_learned(x, var, "valid") if learned else O.upsample_linear(x, True)
    y = O.conv1d(up, W, bias, "valid")                               # [B, 2N-1-k+1, No]
    V = y.shape[1]
    gy = torch.randn_like(y)
    (y * gy).sum().backward()
    gx_ref, gW_ref = x.grad.clone(), W.grad.clone()
    gvar_ref = var.grad.clone() if learned else None
    # ---- merged formulation: per output parity a short conv over x itself ----
    a = torch.sigmoid(var.detach()) if learned else 0.5
    x2 = x.detach().clone().requires_grad_(True)
    Wm = {p: {s: w.detach().clone().requires_grad_(True) for s, w in MU.merged_weights(W.detach(), a, p).items()} for p in (0, 1)}
    ym = torch.zeros_like(y)
    for p in (0, 1):
        rows = (V + 1 - p) // 2                                       # output rows 2m + p < V
        for s, w in Wm[p].items():
            ym[:, p::2] = ym[:, p::2] + torch.einsum("bmc,cn->bmn", x2[:, s:s + rows], w)
    assert torch.allclose(ym, y.detach(), rtol=1e-12, atol=1e-12)
    (ym * gy).sum().backward()
    assert torch.allclose(x2.grad, gx_ref, rtol=1e-11, atol=1e-12)    # the merged conv's dgrad IS the gradient w.r.t. x
    dWm = {p: {s: w.grad for s, w in Wm[p].items()} for p in (0, 1)}
    dW, da = MU.original_gradients(dWm, W.detach(), a, k)
    assert torch.allclose(torch.stack(dW), gW_ref, rtol=1e-11, atol=1e-12)
    if learned:
        assert torch.allclose(da * a * (1 - a), gvar_ref, rtol=1e-10, atol=1e-12)


def test_merged_tap_count_of_the_m4_up_path():
    """MMAs of the upsampled half of every up block: 7 merged taps instead of 10 (copy + mid planes, both parities)."""
    n_now = sum(len(range(5)) for _ in (0, 1))
    n_merged = sum(len(MU.merged_taps(5, p)) for p in (0, 1))
    assert (n_now, n_merged) == (10, 7)
