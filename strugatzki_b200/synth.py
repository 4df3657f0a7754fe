"""Synthetic MFCC+loudness feature files -- numpy twin of the device
generator `k_db_synth` (csrc/db.cuh).  Both use an integer hash, an EXACT integer sum of 8
consecutive 24-bit values and one float scale, so host and device produce bit-identical floats:

    x[c][t] = mu[c] + sigma[c] * g,   g = float(sum_{k<8} u24(seed, stream, c, t+k) - 67108860) * SCALE

Channel 0 (loudness) is clamped to >= floor0.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

BASE_SEED = 20240611
SCALE = np.frombuffer(np.uint32(0x339CC471).tobytes(), np.float32)[0]  # float32(sqrt(1.5) / 2**24)

_A = np.uint64(0xD1B54A32D192ED03)
_B = np.uint64(0x9E3779B97F4A7C15)
_C = np.uint64(0xBF58476D1CE4E5B9)
_D = np.uint64(0x94D049BB133111EB)


def u24(seed: int, stream: int, c: int, t: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + np.uint64(stream) * _A + np.uint64(c) * _B
             + t.astype(np.uint64) * _C)
        z ^= z >> np.uint64(30)
        z *= _C
        z ^= z >> np.uint64(27)
        z *= _D
        z ^= z >> np.uint64(31)
    return (z >> np.uint64(40)).astype(np.int64)


def unit_noise(seed: int, stream: int, c: int, n: int, t0: int = 0) -> np.ndarray:
    """g(seed, stream, c, t) for t in [t0, t0+n): zero mean, unit variance, 8-frame triangular correlation."""
    u = u24(seed, stream, c, np.arange(t0, t0 + n + 7, dtype=np.uint64))
    cs = np.concatenate(([0], np.cumsum(u)))
    s = cs[8:8 + n] - cs[0:n]
    return (s - 67108860).astype(np.int32).astype(np.float32) * SCALE


def default_profile(num_ch: int = 14):
    """(mu, sigma, floor0, norm[numCh][2]) of SURVEY.md section 8d."""
    c = np.arange(num_ch, dtype=np.float32)
    mu = (np.float32(0.35) + np.float32(0.03) * c).astype(np.float32)
    sigma = np.full(num_ch, 0.10, np.float32)
    mu[0] = 0.5
    sigma[0] = 0.12
    norm = np.stack([np.float32(-0.10) + np.float32(0.01) * c, np.float32(1.10) - np.float32(0.005) * c], 1)
    norm = norm.astype(np.float32)
    norm[0] = (0.05, 1.20)
    return mu, sigma, np.float32(1e-3), norm


def synth_file(seed: int, stream: int, n_frames: int, mu: np.ndarray, sigma: np.ndarray,
               floor0: float) -> np.ndarray:
    """Raw (un-normalised) feature frames [n_frames][numCh], float32."""
    num_ch = len(mu)
    out = np.empty((n_frames, num_ch), np.float32)
    for c in range(num_ch):
        g = unit_noise(seed, stream, c, n_frames)
        x = np.float32(mu[c]) + np.float32(sigma[c]) * g
        if c == 0:
            x = np.maximum(x, np.float32(floor0))
        out[:, c] = x
    return out


def plant(window: np.ndarray, seed: int, stream: int, amp: float = 0.02, floor0: float = 1e-3) -> np.ndarray:
    """A noisy copy of `window` ([W][numCh] raw frames): window + amp * g'."""
    w = np.array(window, np.float32, copy=True)
    for c in range(w.shape[1]):
        w[:, c] = w[:, c] + np.float32(amp) * unit_noise(seed ^ 0x5EED, stream, c, w.shape[0])
    w[:, 0] = np.maximum(w[:, 0], np.float32(floor0))
    return w


def regime_file(seed: int, stream: int, n_frames: int, num_ch: int = 14, n_regimes: int = 26):
    """Segmentation test file: the per-channel mean profile switches abruptly at random frames."""
    mu, sigma, floor0, _ = default_profile(num_ch)
    base = synth_file(seed, stream, n_frames, np.zeros(num_ch, np.float32), sigma, -1e9)
    rng = np.random.default_rng(seed + 7919 * stream)
    margin = min(400, n_frames // 8)
    cuts = np.sort(rng.choice(np.arange(margin, n_frames - margin), n_regimes - 1, replace=False))
    bounds = np.concatenate(([0], cuts, [n_frames]))
    out = np.empty_like(base)
    for k in range(len(bounds) - 1):
        prof = (mu + rng.uniform(-0.12, 0.12, num_ch).astype(np.float32)).astype(np.float32)
        out[bounds[k]:bounds[k + 1]] = base[bounds[k]:bounds[k + 1]] + prof
    out[:, 0] = np.maximum(out[:, 0], np.float32(floor0))
    return out, cuts
