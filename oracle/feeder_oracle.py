"""TEST INFRASTRUCTURE ONLY - CPU restatement of the training-batch feeder (csrc/kernels_feed.cu).

What it restates from the reference (per training example):
  take_random_snippets  /root/reference/Datasets.py:16-19   start ~ U{0 .. length - T_in - 1}; slice [start, start + T_in)
  take_snippets_at_pos  /root/reference/Datasets.py:27-34   sample[key][pos:pos + T_in, :] for every source and the mix
  random_amplify        /root/reference/Utils.py:26-36      source *= U(0.7, 1.0); mix = add_n(sources) in dict order
  crop_sample           /root/reference/Utils.py:38-42      sources keep [crop:-crop], crop = (T_in - T_out) // 2
Pinned by tests/golden/feeder.npz, produced by tests/golden/make_golden.py running the reference's own random_amplify /
crop_sample (Utils.py, unmodified) on fixed gains.  The random numbers themselves are this framework's counter-based
hash (the reference uses TF's stateful RNG ops, which cannot be reproduced): the same arithmetic as kernels_feed.cu, so
device and oracle agree bit for bit.  Only tests/ and bench.py's checker legs may import this module.
"""
import numpy as np

M64 = (1 << 64) - 1


def mix64(z):
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def feed_rand(seed, step, example, slot):
    key = ((step << 32) & M64) ^ ((example << 8) & M64) ^ slot
    return mix64((seed & M64) ^ mix64((key + 0x9E3779B97F4A7C15) & M64))


def choose(seed, step, example, n_tracks, lengths, t_in, num_sources, augmentation):
    """(track, start, gains[K]) of one example - slots 0, 1, 2 + k of the hash stream."""
    track = int(feed_rand(seed, step, example, 0) % n_tracks)
    span = int(lengths[track]) - t_in
    start = int(feed_rand(seed, step, example, 1) % span) if span > 0 else 0
    gains = np.ones(num_sources, np.float32)
    if augmentation:
        for k in range(num_sources):
            u = np.float32(feed_rand(seed, step, example, 2 + k) >> 40) * np.float32(1.0 / 16777216.0)
            gains[k] = np.float32(np.float32(u * np.float32(0.3)) + np.float32(0.7))       # rand * (1.0 - 0.7) + 0.7
    return track, start, gains


def amplify_and_crop(snippets, mix_snippet, gains, t_out, augmentation):
    """snippets: [K][T_in, C] float32.  Returns (mix [T_in, C], targets [K][T_out, C]) - Utils.py:26-42."""
    t_in = snippets[0].shape[0]
    crop = (t_in - t_out) // 2
    if augmentation:
        scaled = [np.float32(g) * s.astype(np.float32) for g, s in zip(gains, snippets)]
        mix = scaled[0].copy()
        for s in scaled[1:]:
            mix = (mix + s).astype(np.float32)                       # tf.add_n accumulates in list order
    else:
        scaled = [s.astype(np.float32) for s in snippets]
        mix = mix_snippet.astype(np.float32)
    targets = [s[crop:t_in - crop] if crop > 0 else s for s in scaled]
    targets = [t[:t_out] for t in targets]                            # odd T_in - T_out: the kernel keeps exactly T_out frames
    return mix, targets


def feed_batch(pool, offsets, lengths, batch, t_in, t_out, augmentation, seed, step):
    """pool [K + 1][total, C].  Returns mix [B, T_in, C], targets [K, B, T_out, C], chosen [B, 2]."""
    K, C = pool.shape[0] - 1, pool.shape[2]
    mix = np.zeros((batch, t_in, C), np.float32)
    targets = np.zeros((K, batch, t_out, C), np.float32)
    chosen = np.zeros((batch, 2), np.int64)
    for b in range(batch):
        track, start, gains = choose(seed, step, b, len(offsets), lengths, t_in, K, augmentation)
        f0 = int(offsets[track]) + start
        snips = [pool[k, f0:f0 + t_in] for k in range(K)]
        m, tg = amplify_and_crop(snips, pool[K, f0:f0 + t_in], gains, t_out, augmentation)
        mix[b] = m
        for k in range(K):
            targets[k, b] = tg[k]
        chosen[b] = (track, start)
    return mix, targets, chosen
