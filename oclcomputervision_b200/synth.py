"""Synthetic inputs for the RAISR path.

The pretrained table (``filter.p``, /root/reference/super_resolution/download-pre-trained-weights.txt:1)
needs a download, so benchmarks and tests use a random-init table of the reference shape
``(n_angle, n_strength, n_coherence, scale*scale, 121)`` (raisr.py:77-78) and smooth-noise frames
whose gradient statistics reach every hash bucket.
"""
from __future__ import annotations

import numpy as np

FLEN = 11


def random_filters(scale: int = 2, n_angle: int = 24, n_strength: int = 3, n_coherence: int = 3,
                   seed: int = 1234) -> np.ndarray:
    """Identity-plus-noise taps, each bucket distinct, rows normalised to sum 1."""
    rng = np.random.default_rng(seed)
    f = rng.normal(0.0, 0.02, (n_angle, n_strength, n_coherence, scale * scale, FLEN * FLEN))
    f[..., (FLEN * FLEN) // 2] += 1.0
    f /= f.sum(-1, keepdims=True)
    return np.ascontiguousarray(f.astype(np.float32))


def _blur_axis(a: np.ndarray, k: np.ndarray, axis: int) -> np.ndarray:
    r = len(k) // 2
    pad = [(0, 0), (0, 0)]
    pad[axis] = (r, r)
    p = np.pad(a, pad, mode="reflect")
    out = np.zeros_like(a)
    n = a.shape[axis]
    for i, w in enumerate(k):
        sl = [slice(None), slice(None)]
        sl[axis] = slice(i, i + n)
        out += w * p[tuple(sl)]
    return out


def synthetic_frame(h: int, w: int, seed: int = 1000, sigma: float = 4.0) -> np.ndarray:
    """u8 luma frame: Gaussian-blurred white noise under a cubic left-to-right envelope, so
    strength bins 0/1/2 are all populated (plain noise only reaches the top strength bin)."""
    rng = np.random.default_rng(seed)
    n = rng.standard_normal((h, w), dtype=np.float32)
    r = int(np.ceil(4 * sigma))
    x = np.arange(-r, r + 1, dtype=np.float32)
    k = np.exp(-x * x / (2 * sigma * sigma)).astype(np.float32)
    k /= k.sum()
    n = _blur_axis(_blur_axis(n, k, 1), k, 0)
    n /= np.abs(n).max()
    env = np.linspace(0.0, 1.0, w, dtype=np.float32) ** 3
    img = 0.5 + 0.5 * n * env[None, :]
    return np.clip(np.rint(img * 255.0), 0, 255).astype(np.uint8)


def synthetic_batch(n_frames: int, h: int, w: int, pool: int = 8, seed: int = 1000) -> np.ndarray:
    """(n_frames, h, w) u8; a pool of distinct frames tiled to bound host time."""
    pool = max(1, min(pool, n_frames))
    frames = np.stack([synthetic_frame(h, w, seed + k) for k in range(pool)])
    reps = (n_frames + pool - 1) // pool
    return np.ascontiguousarray(np.tile(frames, (reps, 1, 1))[:n_frames])
