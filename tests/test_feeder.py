"""Training-batch feeder: oracle vs the reference's own random_amplify / crop_sample (golden fixture),
properties of the hash stream, and - on the GPU - bit-exact parity of wun_feed_batch with the oracle."""
import os

import numpy as np
import pytest

import Config
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
from oracle import feeder_oracle as F


def test_oracle_matches_reference_random_amplify_and_crop():
    z = np.load(os.path.join(GOLDEN, "feeder.npz"))
    for ci in range(3):
        names = [str(n) for n in z["c%d/names" % ci]]
        gains, t_out = z["c%d/gains" % ci], int(z["c%d/t_out" % ci])
        snips = [z["c%d/in/%s" % (ci, n)] for n in names]
        mix, targets = F.amplify_and_crop(snips, None, gains, t_out, True)
        assert np.array_equal(mix, z["c%d/out/mix" % ci])                      # bit-exact: one multiply, sum in source order
        for n, t in zip(names, targets):
            assert np.array_equal(t, z["c%d/out/%s" % (ci, n)]), (ci, n)
        rec_mix = sum(snips).astype(np.float32)
        mix2, targets2 = F.amplify_and_crop(snips, rec_mix, gains, t_out, False)
        assert np.array_equal(mix2, z["c%d/out_noaug/mix" % ci])
        for n, t in zip(names, targets2):
            assert np.array_equal(t, z["c%d/out_noaug/%s" % (ci, n)]), (ci, n)


def test_hash_stream_ranges_and_determinism():
    lengths = np.array([500, 91, 4000], np.int64)
    seen_tracks, starts = set(), []
    for step in range(40):
        for b in range(8):
            tr, st, g = F.choose(1234, step, b, 3, lengths, 90, 4, True)
            assert (tr, st, tuple(g)) == tuple(x if not isinstance(x, np.ndarray) else tuple(x)
                                               for x in F.choose(1234, step, b, 3, lengths, 90, 4, True))
            assert 0 <= tr < 3 and 0 <= st < lengths[tr] - 90                    # Datasets.py:18: maxval exclusive
            assert np.all(g >= np.float32(0.7)) and np.all(g < np.float32(1.0))   # Utils.py:33
            seen_tracks.add(tr); starts.append(st)
    assert seen_tracks == {0, 1, 2} and len(set(starts)) > 50
    # another seed / step / example changes the draw
    base = F.choose(1234, 3, 2, 3, lengths, 90, 4, True)
    assert any(F.choose(s, st, b, 3, lengths, 90, 4, True)[:2] != base[:2] for s, st, b in ((1235, 3, 2), (1234, 4, 2), (1234, 3, 3)))
    # no augmentation: unit gains
    assert np.all(F.choose(1, 0, 0, 3, lengths, 90, 4, False)[2] == 1.0)


def test_gain_statistics():
    g = np.array([F.choose(7, s, b, 1, np.array([1000]), 10, 2, True)[2] for s in range(200) for b in range(16)])
    assert abs(float(g.mean()) - 0.85) < 0.01 and float(g.min()) < 0.71 and float(g.max()) > 0.99


def _problem(preset, overrides, n_tracks=5, seed=3):
    from wun.feeder import synthetic_tracks, build_pool
    import wun
    cfg = Config.build_config([preset], overrides, experiment_id=0)["model_config"]
    eng = wun.Engine(wun.config_from_model_config(cfg), num_frames=cfg["num_frames"])
    tracks = synthetic_tracks(cfg["source_names"], cfg["num_channels"], n_tracks, (eng.T_in + 1, eng.T_in + 700), seed=seed)
    return cfg, eng, tracks, build_pool(tracks, cfg["source_names"])


def test_pool_layout():
    cfg, eng, tracks, (pool, offsets, lengths) = _problem("baseline_stereo", dict(num_layers=3, num_frames=64))
    K = len(cfg["source_names"])
    assert pool.shape == (K + 1, int(lengths.sum()), cfg["num_channels"]) and offsets[0] == 0
    for i, t in enumerate(tracks):
        assert np.array_equal(pool[K, offsets[i]:offsets[i] + lengths[i]], t["mix"])
        assert np.array_equal(pool[0, offsets[i]:offsets[i] + lengths[i]], t[cfg["source_names"][0]])


def test_device_feeder_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("has GPU")
    from wun.feeder import DeviceFeeder
    cfg, eng, tracks, _ = _problem("baseline_stereo", dict(num_layers=3, num_frames=64))
    with pytest.raises(RuntimeError, match="no host fallback"):
        DeviceFeeder(eng, tracks, cfg["source_names"], 4)


@pytest.mark.gpu
@pytest.mark.parametrize("preset,overrides,aug", [
    ("baseline_stereo", dict(num_layers=3, num_frames=64), True),
    ("baseline_stereo", dict(num_layers=3, num_frames=64), False),
    ("full_multi_instrument", dict(num_layers=4, num_frames=200), True),
    ("baseline", dict(num_layers=2, num_frames=128), True),                   # mono, 'same' padding: T_out == T_in, no crop
])
def test_feed_batch_bit_exact_vs_oracle(preset, overrides, aug):
    import torch
    from wun.feeder import DeviceFeeder
    cfg, eng, tracks, (pool, offsets, lengths) = _problem(preset, overrides)
    B = 6
    fd = DeviceFeeder(eng, tracks, cfg["source_names"], B, augmentation=aug, seed=99, record_choice=True)
    for step in range(3):                                                      # the device counter advances by itself
        mix_d, tg_d = fd.next_batch()
        torch.cuda.synchronize()
        mix_o, tg_o, chosen_o = F.feed_batch(pool, offsets, lengths, B, eng.T_in, eng.T_out, aug, 99, step)
        assert np.array_equal(fd.chosen.cpu().numpy(), chosen_o)
        assert np.array_equal(mix_d.cpu().numpy(), mix_o)
        assert np.array_equal(tg_d.cpu().numpy(), tg_o)
    assert int(fd.step_state.item()) == 3


@pytest.mark.gpu
def test_feeder_inside_cuda_graph_draws_a_new_batch_per_replay():
    import torch
    from wun.feeder import DeviceFeeder
    cfg, eng, tracks, (pool, offsets, lengths) = _problem("baseline_stereo", dict(num_layers=3, num_frames=64))
    fd = DeviceFeeder(eng, tracks, cfg["source_names"], 4, augmentation=True, seed=5)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        fd.next_batch()                                                        # step 0 eagerly
        stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            fd.next_batch()
        for step in (1, 2, 3):
            g.replay()
            stream.synchronize()
            mix_o, tg_o, _ = F.feed_batch(pool, offsets, lengths, 4, eng.T_in, eng.T_out, True, 5, step)
            assert np.array_equal(fd.mix.cpu().numpy(), mix_o) and np.array_equal(fd.targets.cpu().numpy(), tg_o)


@pytest.mark.gpu
def test_training_on_device_fed_batches_reduces_the_loss():
    """Training.train with the device feeder as batch source: the loss of a small net on a FIXED validation batch drops."""
    import torch
    import Training
    from Models.UnetAudioSeparator import UnetAudioSeparator
    cfg = Config.build_config(["baseline_stereo"], dict(num_layers=3, num_initial_filters=8, num_frames=64, batch_size=4,
                                                        epoch_it=60, init_sup_sep_lr=1e-3), experiment_id=0)["model_config"]
    sep = UnetAudioSeparator(cfg)
    l0 = None
    path, sep = Training.train(cfg, "feedtest", sep=sep, log_every=0, feeder="device")
    assert sep.global_step == 60 and sep.last_feeder is not None and int(sep.last_feeder.step_state.item()) == 60
    # same tracks, fresh batch from another seed as validation: compare an untrained replica with the trained one
    from wun.feeder import DeviceFeeder
    fd = DeviceFeeder(sep.engine(input_frames=sep.last_feeder.eng.T_in), sep.last_feeder_tracks, cfg["source_names"], 4,
                      augmentation=False, seed=777)
    mix, tg = fd.next_batch()
    fresh = UnetAudioSeparator(cfg)
    fresh._ensure_params(fresh.engine(input_frames=fd.eng.T_in), mix.device, create=True)
    l_fresh = float(fresh.loss_and_gradients(mix, tg).item())
    l_trained = float(sep.loss_and_gradients(mix, tg).item())
    assert l_trained < 0.8 * l_fresh, (l_trained, l_fresh)
