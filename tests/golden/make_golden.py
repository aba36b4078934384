"""Generate tests/golden/*.npz and padding.json by executing the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

What runs from the reference, verbatim: Models/UnetAudioSeparator.py (ctor, get_padding, get_output),
Models/InterpolationLayer.py, Models/OutputLayer.py, Utils.py (crop, crop_and_concat, LeakyReLU,
AudioClip).  `tensorflow` is replaced by tests/golden/tf_shim.py (eager torch-CPU, fp64).  The loss is
the two lines Training.py:62-63 restated here (Training.py itself needs sacred + a dataset).
Gradients come from torch autograd through the graph the reference code built.

Fixture content (per case): cfg (json), mix [B,T_in,C] f32, targets, params in TF creation order (f32
values, stored exactly), outputs for training=True and training=False (f64), loss (f64), grads (f64).
"""
import json
import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, HERE)

import tf_shim  # noqa: E402

tf_shim.install()
sys.path.insert(0, REF)
import Models.UnetAudioSeparator as RefSep  # noqa: E402  (the reference, unmodified)

import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location("b200_Config", os.path.join(REPO, "wave-u-net_b200", "Config.py"))
Config = importlib.util.module_from_spec(_spec)   # ours: only used to build the model_config dict
_spec.loader.exec_module(Config)

CASES = {
    # name: (named presets, overrides, batch, num_frames)
    "ctx_linear_diff_stereo": (["baseline_stereo"], dict(num_layers=3, num_initial_filters=8), 2, 40),
    "ctx_learned_diff_multi": (["full_multi_instrument"],
                               dict(num_layers=3, num_initial_filters=8, upsampling="learned"), 2, 33),
    "ctx_learned_mono_direct": (["baseline"], dict(num_layers=2, num_initial_filters=8, context=True,
                                                   upsampling="learned", filter_size=7,
                                                   input_filter_size=7, merge_filter_size=3), 2, 21),
    "same_linear_direct_mono": (["baseline"], dict(num_layers=3, num_initial_filters=8), 2, 64),
    "same_learned_diff_linear": (["baseline_diff"], dict(num_layers=3, num_initial_filters=8,
                                                         upsampling="learned", mono_downmix=False,
                                                         output_activation="linear"), 2, 64),
    "ctx_linear_direct_linearact": (["baseline_context"], dict(num_layers=2, num_initial_filters=8,
                                                               output_type="direct",
                                                               output_activation="linear",
                                                               output_filter_size=3), 1, 30),
}


def run_case(name, named, overrides, batch, num_frames):
    cfg = Config.build_config(named, overrides, experiment_id=0)["model_config"]
    sep = RefSep.UnetAudioSeparator(cfg)
    in_shape, out_shape = sep.get_padding(np.array([batch, num_frames, 0]))
    T_in, T_out, C = int(in_shape[1]), int(out_shape[1]), int(in_shape[2])
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    srcs = cfg["source_names"]
    K = len(srcs)
    mix = np.zeros((batch, T_in, C), np.float32)
    targets = {}
    cropf = (T_in - T_out) // 2
    for s in srcs:
        a = (rng.uniform(-1, 1, size=(batch, T_in, C)) * 1.5 / K).astype(np.float32)  # some |mix|>1
        mix += a
        targets[s] = np.ascontiguousarray(a[:, cropf:T_in - cropf] if cropf else a)

    tf_shim.STATE.reset(seed=1337)
    x = tf_shim.T(torch.tensor(mix, dtype=torch.float64))
    outs_train = sep.get_output(x, True, False, reuse=False)
    # non-zero biases make the bias path observable: perturb, then rebuild (variables are reused)
    for vname, v in tf_shim.STATE.vars.items():
        with torch.no_grad():
            v.copy_(torch.tensor(v.detach().numpy().astype(np.float32).astype(np.float64)))
            if vname.endswith("/bias"):
                v.copy_(torch.tensor(rng.uniform(-0.05, 0.05, size=tuple(v.shape)).astype(np.float32)
                                     .astype(np.float64)))
    tf_shim.STATE.counters = {}
    outs_train = sep.get_output(x, True, False, reuse=True)
    tf_shim.STATE.counters = {}
    outs_test = sep.get_output(x, False, False, reuse=True)

    # Training.py:50-63
    loss = 0
    for s in srcs:
        real = torch.tensor(targets[s], dtype=torch.float64)
        loss = loss + torch.mean((real - outs_train[s].t) ** 2)
    loss = loss / float(cfg["num_sources"])
    loss.backward()

    rec = {"cfg_json": np.array(json.dumps(cfg)), "mix": mix, "loss": np.float64(loss.item()),
           "T_in": np.int64(T_in), "T_out": np.int64(T_out),
           "param_names": np.array(list(tf_shim.STATE.vars.keys()))}
    for s in srcs:
        rec["target/" + s] = targets[s]
        rec["out_train/" + s] = outs_train[s].t.detach().numpy()
        rec["out_test/" + s] = outs_test[s].t.detach().numpy()
    for vname, v in tf_shim.STATE.vars.items():
        rec["param/" + vname] = v.detach().numpy().astype(np.float32)
        rec["grad/" + vname] = v.grad.detach().numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print("%-32s T_in=%d T_out=%d params=%d loss=%.6g" % (name, T_in, T_out, len(tf_shim.STATE.vars),
                                                       loss.item()))


def padding_table():
    rows = []
    presets = ["baseline", "baseline_context", "baseline_stereo", "full", "full_44KHz",
               "full_multi_instrument", "baseline_context_smallfilter_deep", "baseline_comparison"]
    for p in presets:
        cfg = Config.build_config([p], experiment_id=0)["model_config"]
        sep = RefSep.UnetAudioSeparator(cfg)
        for nf in [cfg["num_frames"], 1, 2, 100, 1000, 16384, 16389, 16390, 44100, 98560, 200000]:
            try:
                i, o = sep.get_padding(np.array([cfg["batch_size"], nf, 0]))
                rows.append(dict(preset=p, num_frames=int(nf), in_shape=[int(v) for v in i],
                                 out_shape=[int(v) for v in o]))
            except AssertionError:
                rows.append(dict(preset=p, num_frames=int(nf), error="AssertionError"))
    # small / odd configs
    for L, fs, mfs, ifs, ofs in [(1, 3, 3, 3, 1), (2, 15, 5, 15, 1), (3, 15, 5, 15, 1), (3, 5, 1, 5, 3),
                                 (2, 7, 3, 7, 1), (4, 9, 5, 11, 3)]:
        cfg = Config.build_config(["baseline_context"],
                                  dict(num_layers=L, filter_size=fs, merge_filter_size=mfs,
                                       input_filter_size=ifs, output_filter_size=ofs),
                                  experiment_id=0)["model_config"]
        sep = RefSep.UnetAudioSeparator(cfg)
        for nf in [1, 2, 3, 7, 16, 33, 40, 100, 1001]:
            try:
                i, o = sep.get_padding(np.array([3, nf, 0]))
                rows.append(dict(overrides=dict(num_layers=L, filter_size=fs, merge_filter_size=mfs,
                                                input_filter_size=ifs, output_filter_size=ofs),
                                 num_frames=int(nf), in_shape=[int(v) for v in i],
                                 out_shape=[int(v) for v in o]))
            except AssertionError:
                rows.append(dict(overrides=dict(num_layers=L, filter_size=fs, merge_filter_size=mfs,
                                                input_filter_size=ifs, output_filter_size=ofs),
                                 num_frames=int(nf), error="AssertionError"))
    with open(os.path.join(HERE, "padding.json"), "w") as f:
        json.dump(rows, f, indent=0)
    print("padding.json: %d rows" % len(rows))


def feeder_case():
    """Utils.random_amplify + Utils.crop_sample (the reference's own functions, fp32 here: the feeder is byte movement plus
    one multiply and one sum per sample, and parity is bit-exact) on fixed snippets and fixed gains."""
    import collections
    import Utils as RefUtils                                   # /root/reference/Utils.py, unmodified
    rng = np.random.default_rng(4242)
    rec = {}
    for ci, (names, C, t_in, t_out) in enumerate([(["accompaniment", "vocals"], 2, 40, 12),
                                                  (["bass", "drums", "other", "vocals"], 1, 33, 33),
                                                  (["bass", "drums", "other", "vocals"], 2, 57, 21)]):
        tf_shim.STATE.reset(dtype=torch.float32)
        gains = rng.uniform(0.7, 1.0, size=len(names)).astype(np.float32)
        snips = collections.OrderedDict((n, rng.uniform(-1, 1, size=(t_in, C)).astype(np.float32)) for n in names)
        sample = collections.OrderedDict((n, tf_shim.T(torch.tensor(v))) for n, v in snips.items())
        sample["mix"] = tf_shim.T(torch.zeros((t_in, C)))      # overwritten by random_amplify
        tf_shim.STATE.uniform_queue = [float(g) for g in gains]
        out = RefUtils.random_amplify(sample)
        out = RefUtils.crop_sample(out, (t_in - t_out) // 2)
        rec["c%d/names" % ci] = np.array(names)
        rec["c%d/gains" % ci] = gains
        rec["c%d/t_out" % ci] = np.int64(t_out)
        for n in names:
            rec["c%d/in/%s" % (ci, n)] = snips[n]
            rec["c%d/out/%s" % (ci, n)] = out[n].t.numpy()
        rec["c%d/out/mix" % ci] = out["mix"].t.numpy()
        # without augmentation only crop_sample runs (Datasets.py:204-208)
        sample2 = collections.OrderedDict((n, tf_shim.T(torch.tensor(v))) for n, v in snips.items())
        sample2["mix"] = tf_shim.T(torch.tensor(sum(snips.values()).astype(np.float32)))
        out2 = RefUtils.crop_sample(sample2, (t_in - t_out) // 2)
        for n in names:
            rec["c%d/out_noaug/%s" % (ci, n)] = out2[n].t.numpy()
        rec["c%d/out_noaug/mix" % ci] = out2["mix"].t.numpy()
    np.savez_compressed(os.path.join(HERE, "feeder.npz"), **rec)
    print("feeder.npz: 3 cases (reference Utils.random_amplify / crop_sample)")


if __name__ == "__main__":
    feeder_case()
    padding_table()
    for name, (named, ov, b, nf) in CASES.items():
        run_case(name, named, ov, b, nf)
