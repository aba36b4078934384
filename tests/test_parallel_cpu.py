"""world_size-2 gloo tests (CPU) of the data-parallel host logic (wave-u-net_b200/wun/parallel.py).

The engine itself needs a GPU, so the per-rank "gradient" here comes from the CPU oracle on the rank's shard: the test
proves that shard + 1/world pre-scale + ONE all-reduce(sum) of the flat buffer reproduces the full-batch gradient of
Training.py:50-63, and that the inference window partition + all-gather reassembles every window in order."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "wave-u-net_b200"))
    import Config
    from oracle import wave_unet_oracle as O
    from wun import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        cfg = Config.build_config(["baseline_stereo"], dict(num_layers=3, num_initial_filters=8), experiment_id=0)["model_config"]
        t_in, t_out = O.get_padding(cfg, 40)
        params = O.init_params(cfg, seed=5)
        B = 5                                           # uneven split: ranks get 3 and 2 examples
        mix, targets = O.synthetic_batch(cfg, B, t_in, t_out, seed=6)
        names = list(params.keys())
        # replicas must start identical
        flat_p = torch.from_numpy(np.concatenate([params[n].ravel() for n in names]) + (rank * 1.0)).float()
        parallel.broadcast_parameters(flat_p, src=0)
        assert torch.equal(flat_p, torch.from_numpy(np.concatenate([params[n].ravel() for n in names])).float())
        lo, hi = parallel.shard_range(B, rank, world)
        assert parallel.shard_batch(torch.from_numpy(mix), rank, world).shape[0] == hi - lo
        _, _, g_local = O.forward_backward(cfg, params, mix[lo:hi], {k: v[lo:hi] for k, v in targets.items()},
                                           dtype=torch.float64)
        scale = parallel.grad_scale_for(hi - lo, B)
        flat = torch.from_numpy(np.concatenate([g_local[n].ravel() for n in names]) * scale)
        parallel.allreduce_gradients(flat)
        _, _, g_full = O.forward_backward(cfg, params, mix, targets, dtype=torch.float64)
        want = np.concatenate([g_full[n].ravel() for n in names])
        err = np.linalg.norm(flat.numpy() - want) / np.linalg.norm(want)
        # timing reduction + window gather
        slow = parallel.max_over_ranks(10.0 + rank)
        n_win = 7
        wlo, whi = parallel.shard_range(n_win, rank, world)
        local = torch.arange(wlo, whi, dtype=torch.float32).reshape(1, -1, 1, 1).repeat(2, 1, 3, 2)
        full = parallel.gather_window_predictions(local, n_win)
        ok_gather = bool(torch.equal(full[0, :, 0, 0], torch.arange(n_win, dtype=torch.float32))) and full.shape == (2, n_win, 3, 2)
        if rank == 0:
            out.put((err, slow, ok_gather, (lo, hi)))
    finally:
        dist.destroy_process_group()


def test_dp_gradient_allreduce_and_window_gather_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    err, slow, ok_gather, rng0 = q.get(timeout=5)
    assert err < 1e-12, err
    assert slow == 11.0
    assert ok_gather
    assert rng0 == (0, 3)


def test_shard_range_properties():
    sys.path.insert(0, os.path.join(REPO, "wave-u-net_b200"))
    from wun import parallel
    for n in (0, 1, 7, 16, 485):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_window_starts_match_reference_loop():
    sys.path.insert(0, os.path.join(REPO, "wave-u-net_b200"))
    import Evaluate
    # Evaluate.py:125-128: hop T_out, last window shifted to end exactly at n_frames
    assert Evaluate.window_starts(100, 30) == [0, 30, 60, 70]
    assert Evaluate.window_starts(90, 30) == [0, 30, 60]
    assert Evaluate.window_starts(7938000, 16389)[-1] == 7938000 - 16389
    assert len(Evaluate.window_starts(7938000, 16389)) == 485
