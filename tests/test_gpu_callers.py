"""GPU tests of the callers either side of the hot path: windowed inference (Evaluate.predict_track) and the
training entry point (Training.train), both against the oracle / reference semantics."""
import os

import numpy as np
import pytest
import torch

import Config
import Evaluate
import TFCheckpoint
import Training
from Models.UnetAudioSeparator import UnetAudioSeparator
from oracle import wave_unet_oracle as O

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("preset,ov", [("baseline_stereo", dict(num_layers=4)),
                                        ("full", dict(num_layers=3, num_initial_filters=16)),
                                        ("baseline", dict(num_layers=4, num_frames=256))])
def test_predict_track_matches_oracle(preset, ov):
    cfg = Config.build_config([preset], ov, experiment_id=0)["model_config"]
    if cfg["context"]:
        cfg["num_frames"] = 200
    t_in, t_out = O.get_padding(cfg, cfg["num_frames"])
    params = O.init_params(cfg, seed=3)
    sep = UnetAudioSeparator(cfg)
    sep.load_variables(params, input_frames=t_in)
    rng = np.random.default_rng(4)
    n = 5 * t_out + 13                                   # last window gets shifted back (Evaluate.py:127-128)
    audio = rng.uniform(-1, 1, size=(n, 2)).astype(np.float32)
    got = Evaluate.predict_track(cfg, sep, audio, batch_windows=3)
    want = O.predict_track(cfg, params, audio, t_in, t_out)
    assert list(got.keys()) == cfg["source_names"]
    for s in cfg["source_names"]:
        assert got[s].shape == want[s].shape
        assert rel_l2(got[s], want[s]) <= 1e-4, (s, rel_l2(got[s], want[s]))
    # short input: zero-extended to T_in and cut back (Evaluate.py:107-111, :142-143)
    short = audio[: t_in // 3]
    got = Evaluate.predict_track(cfg, sep, short)
    want = O.predict_track(cfg, params, short, t_in, t_out)
    for s in cfg["source_names"]:
        assert got[s].shape == want[s].shape == (short.shape[0], want[s].shape[1])
        assert rel_l2(got[s], want[s]) <= 1e-4


def test_mono_input_is_tiled_for_stereo_model():
    cfg = Config.build_config(["baseline_stereo"], dict(num_layers=3, num_frames=100), experiment_id=0)["model_config"]
    t_in, t_out = O.get_padding(cfg, 100)
    params = O.init_params(cfg, seed=8)
    sep = UnetAudioSeparator(cfg)
    sep.load_variables(params, input_frames=t_in)
    mono = np.random.default_rng(1).uniform(-1, 1, size=(2 * t_out + 5, 1)).astype(np.float32)
    got = Evaluate.predict_track(cfg, sep, mono)
    want = O.predict_track(cfg, params, mono, t_in, t_out)
    for s in cfg["source_names"]:
        assert got[s].shape[1] == 2 and rel_l2(got[s], want[s]) <= 1e-4


def test_training_entry_point_and_checkpoint(tmp_path):
    cfg = Config.build_config(["baseline_stereo"], dict(num_layers=3, num_frames=64, batch_size=4, epoch_it=12,
                                                        init_sup_sep_lr=1e-3, model_base_dir=str(tmp_path)),
                              experiment_id=42)
    mc = cfg["model_config"]
    path, sep = Training.train(mc, 42, log_every=0)
    # the reference's Saver layout: <base>/<id>/<id>-<global_step>.{index,data-00000-of-00001} + the `checkpoint` state file
    assert path == os.path.join(str(tmp_path), "42", "42-12") and sep.global_step == 12
    assert os.path.exists(path + ".index") and os.path.exists(path + ".data-00000-of-00001")
    assert TFCheckpoint.latest_checkpoint(os.path.dirname(path)) == path
    ck = TFCheckpoint.read_checkpoint(path)
    assert int(ck["global_step"]) == 12 and ck["global_step"].dtype == np.int64
    names = [n for n, _, _, _ in sep.param_table()]
    assert all(n in ck for n in names)
    assert all("separator_solver/%s/Adam" % n in ck and "separator_solver/%s/Adam_1" % n in ck for n in names)
    assert len(ck) == 3 * len(names) + 3
    # resume: a fresh separator restored from the checkpoint continues from the same variables / Adam slots
    sep2 = UnetAudioSeparator(mc)
    t_in = int(sep.get_padding(np.array([4, 64, 0]))[0][1])
    Training.load_checkpoint(path, sep2, t_in)
    assert torch.equal(sep2.params.cpu(), sep.params.cpu()) and sep2.global_step == 12
    assert torch.equal(sep2.adam_m.cpu(), sep.adam_m.cpu()) and torch.equal(sep2.adam_v.cpu(), sep.adam_v.cpu())
    # the round-1 .npz container still round-trips
    legacy = Training.save_checkpoint(str(tmp_path / "legacy.npz"), sep)
    sep3 = UnetAudioSeparator(mc)
    Training.load_checkpoint(legacy, sep3, t_in)
    assert torch.equal(sep3.params.cpu(), sep.params.cpu()) and torch.equal(sep3.adam_v.cpu(), sep.adam_v.cpu())
    # Predict.py's loader (Evaluate.py:55-57: variables only) gives the same separation as the live separator
    mix = np.random.default_rng(2).uniform(-0.5, 0.5, size=(300, 2)).astype(np.float32)
    want = Evaluate.predict_track(mc, sep, mix)
    got = Evaluate.produce_source_estimates(mc, path, mix)
    for k in mc["source_names"]:
        assert np.array_equal(got[k], want[k])
    v = Training.validation_loss(mc, sep, batches=2)
    assert np.isfinite(v) and v > 0


def test_cli_parsing_drives_training(tmp_path):
    cfg, extras = Config.parse_command_line(["with", "cfg.baseline_context", "cfg.model_config.num_layers=2",
                                             "cfg.model_config.epoch_it=3", "cfg.model_config.batch_size=2",
                                             "cfg.model_config.num_frames=40",
                                             "cfg.model_config.model_base_dir=%s" % tmp_path, "max_epochs=1"])
    assert extras == {"max_epochs": 1}
    path, sep = Training.train(cfg["model_config"], cfg["experiment_id"], log_every=0)
    assert sep.global_step == 3
