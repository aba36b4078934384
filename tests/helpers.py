"""Shared by the CPU and GPU parity tests: which rows of each activation the engine keeps ("live" rows)."""
import numpy as np


def live_rows(cfg, t_in):
    """{LeakyReLU call index in the oracle's forward(): [(engine tensor name, full-rate row indices)]}.
    Call order: down0..down(L-1), bottleneck, up0..up(L-1) (UnetAudioSeparator.py:97-125).  The engine keeps, per down
    block, the decimated rows ([:, ::2, :], :100) as dec<i> and the odd rows inside the centre-cropped skip window
    (Utils.py:104-123) as odd<i>; everything else of the full-rate conv output is never computed."""
    L, fs, mfs, ctx = cfg["num_layers"], cfg["filter_size"], cfg["merge_filter_size"], cfg["context"]
    full, cur = [], int(t_in)
    for i in range(L):
        cur = cur - (fs - 1) if ctx else cur
        full.append(cur)
        cur = (cur + 1) // 2
    ups = [cur - (fs - 1) if ctx else cur]
    for i in range(L):
        u = 2 * ups[-1] - 1 if ctx else 2 * ups[-1]
        ups.append(u - (mfs - 1) if ctx else u)
    out = {}
    for i in range(L):
        U = 2 * ups[L - 1 - i] - 1 if ctx else 2 * ups[L - 1 - i]
        cs = (full[i] - U) // 2
        odd_pos = np.asarray([a for a in range(cs, cs + U) if a % 2 == 1], np.int64)
        out[i] = [("dec%d" % i, np.arange(0, full[i], 2)), ("odd%d" % i, odd_pos)]
    out[L] = [("z", np.arange(ups[0]))]
    for i in range(L):
        out[L + 1 + i] = [("up%d" % i, np.arange(ups[i + 1]))]
    return out
