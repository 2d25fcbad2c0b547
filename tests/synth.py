"""Seeded synthetic Cityscapes-shaped inputs, numpy so they are identical on every box.

RGB uniform[-1,1] (folder.py:187-190 normalisation), seg = one-hot float of a blocky 20-class label map
(folder.py:193-200), masks = sigmoid of a smooth field (nets/SubNets.py:257-258), flows = smooth N(0,1)
field on a coarse lattice upsampled bilinearly and scaled to sigma_px pixels, in normalised units.
"""
from __future__ import annotations

import numpy as np


def _upsample(coarse: np.ndarray, H: int, W: int) -> np.ndarray:
    """bilinear upsample [..., h, w] -> [..., H, W] (align-corners lattice)."""
    h, w = coarse.shape[-2:]
    ys = np.linspace(0, h - 1, H)
    xs = np.linspace(0, w - 1, W)
    y0 = np.clip(np.floor(ys).astype(int), 0, max(h - 2, 0))
    x0 = np.clip(np.floor(xs).astype(int), 0, max(w - 2, 0))
    y1 = np.minimum(y0 + 1, h - 1)
    x1 = np.minimum(x0 + 1, w - 1)
    ty = (ys - y0)[:, None]
    tx = (xs - x0)[None, :]
    a = coarse[..., y0[:, None], x0[None, :]]
    b = coarse[..., y0[:, None], x1[None, :]]
    c = coarse[..., y1[:, None], x0[None, :]]
    d = coarse[..., y1[:, None], x1[None, :]]
    return (a * (1 - tx) + b * tx) * (1 - ty) + (c * (1 - tx) + d * tx) * ty


def smooth_field(rng, shape, H, W, cell=16):
    return _upsample(rng.standard_normal(shape + (H // cell + 2, W // cell + 2)), H, W)


def rgb(seed, N, H, W, C=3):
    return np.random.default_rng(seed).uniform(-1, 1, (N, C, H, W)).astype(np.float32)


def seg(seed, N, H, W, classes=20, block=8):
    rng = np.random.default_rng(seed)
    lab = rng.integers(0, classes, (N, (H + block - 1) // block, (W + block - 1) // block))
    lab = np.repeat(np.repeat(lab, block, 1), block, 2)[:, :H, :W]
    return (lab[:, None] == np.arange(classes)[None, :, None, None]).astype(np.float32)


def mask(seed, N, H, W, T=None):
    rng = np.random.default_rng(seed)
    shape = (N,) if T is None else (N, T)
    f = smooth_field(rng, shape, H, W)
    m = 1.0 / (1.0 + np.exp(-f))
    return (m[:, None] if T is None else m).astype(np.float32)  # [N,1,H,W] or [N,T,H,W]


def flow(seed, N, H, W, sigma_px=8.0, T=None, oob_frac=0.01):
    """[N,2,H,W] or [N,2,T,H,W], normalised units (2.0 = full width/height); >= oob_frac of the pixels are
    pushed out of the image to exercise the validity bits."""
    rng = np.random.default_rng(seed)
    shape = (N, 2) if T is None else (N, 2, T)
    f = smooth_field(rng, shape, H, W)
    f[:, 0] *= 2.0 * sigma_px / max(W, 1)
    f[:, 1] *= 2.0 * sigma_px / max(H, 1)
    if oob_frac > 0:
        kick = rng.random(f.shape[:1] + f.shape[2:]) < oob_frac
        sign = rng.choice([-1.0, 1.0], size=kick.shape)
        f[:, 0] = np.where(kick, f[:, 0] + 2.5 * sign, f[:, 0])
    return f.astype(np.float32)


def adversarial_flow(seed, N, H, W):
    """iid uniform +-0.5: quarter-image random gather/scatter, the worst case."""
    return np.random.default_rng(seed).uniform(-0.5, 0.5, (N, 2, H, W)).astype(np.float32)


def grad(seed, shape):
    return np.random.default_rng(seed).standard_normal(shape).astype(np.float32)
