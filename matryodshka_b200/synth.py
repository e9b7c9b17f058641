"""Seeded synthetic inputs and weights for tests and bench.

There is no network for datasets or checkpoints, so every measurement and
parity test runs on synthetic ODS pairs of the named shape and random-init
weights of the reference architecture.  Seed default 8964 is the reference's
own ``random_seed`` (test.py:62).
"""
from __future__ import annotations

import numpy as np

from .nets import layer_shapes

DEFAULT_SEED = 8964
DEFAULT_BASELINE = 0.032  # export.py:225


def band_limited_images(B, H, W, seed=DEFAULT_SEED, n_waves=8, noise=0.02):
    """[B, H, W, 3] float32 in [0, 1]: a sum of low-frequency 2-D sinusoids,
    periodic in x (an ERP image wraps horizontally), scaled to [0.05, 0.95], plus
    U(-noise, noise)."""
    rng = np.random.default_rng(seed)
    yy = np.linspace(0.0, 1.0, H, dtype=np.float64)[:, None]
    xx = (np.arange(W, dtype=np.float64) / W)[None, :]
    out = np.zeros((B, H, W, 3), np.float64)
    for b in range(B):
        for c in range(3):
            img = np.zeros((H, W), np.float64)
            for _ in range(n_waves):
                fx = rng.integers(0, 7)
                fy = rng.uniform(0.0, 5.0)
                ph = rng.uniform(0, 2 * np.pi)
                amp = rng.uniform(0.3, 1.0)
                img += amp * np.sin(2 * np.pi * (fx * xx + fy * yy) + ph)
            img = (img - img.min()) / (img.max() - img.min() + 1e-12)
            out[b, ..., c] = 0.05 + 0.9 * img
    out += rng.uniform(-noise, noise, out.shape)
    return np.clip(out, 0.0, 1.0).astype(np.float32)


def white_noise_images(B, H, W, seed=DEFAULT_SEED):
    """Stress variant: U(0, 1) white noise."""
    rng = np.random.default_rng(seed)
    return rng.uniform(0.0, 1.0, (B, H, W, 3)).astype(np.float32)


def ods_pair(B, H, W, seed=DEFAULT_SEED, kind="band"):
    """(ref, src) images: src is ref shifted by a few pixels plus its own detail,
    so the two PSVs correlate the way a stereo pair does."""
    make = band_limited_images if kind == "band" else white_noise_images
    ref = make(B, H, W, seed)
    other = make(B, H, W, seed + 1)
    src = np.clip(0.8 * np.roll(ref, 3, axis=2) + 0.2 * other, 0.0, 1.0).astype(np.float32)
    return ref, src


def identity_poses(B):
    return np.tile(np.eye(4, dtype=np.float32)[None], (B, 1, 1))


def intrinsics(B, baseline=DEFAULT_BASELINE):
    """data_loader.py:160: intrinsics[b, 0, 0] carries the ODS baseline."""
    k = np.array([[baseline, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    return np.tile(k[None], (B, 1, 1))


def target_positions(B, seed=DEFAULT_SEED, scale=0.05):
    """tgt_pos [B, 3] ~ U(-scale, scale)^3 (data_loader.py:175 'tgt_pose')."""
    rng = np.random.default_rng(seed + 17)
    return rng.uniform(-scale, scale, (B, 3)).astype(np.float32)


def net_weights(num_inputs, num_outputs, ngf=64, seed=DEFAULT_SEED, coord=True):
    """Random-init weights keyed by the TF checkpoint variable names:
    conv/deconv ~ N(0, 2/fan_in), gamma ~ U(0.5, 1.5), beta ~ N(0, 0.1^2),
    head bias ~ N(0, 0.1^2)."""
    rng = np.random.default_rng(seed + 101)
    out = {}
    for name, shp in layer_shapes(num_inputs, num_outputs, ngf, coord).items():
        if name.endswith("/weights"):
            if len(shp) == 4 and name.split("/")[1] in ("conv6_1", "conv7_1", "conv8_1"):
                fan_in = shp[0] * shp[1] * shp[3] / 4.0  # stride-2 deconv: 4 of 16 taps hit each output
            else:
                fan_in = shp[0] * shp[1] * shp[2]
            out[name] = rng.normal(0.0, np.sqrt(2.0 / fan_in), shp).astype(np.float32)
        elif name.endswith("gamma"):
            out[name] = rng.uniform(0.5, 1.5, shp).astype(np.float32)
        else:
            out[name] = rng.normal(0.0, 0.1, shp).astype(np.float32)
    return out
