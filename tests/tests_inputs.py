"""Deterministic synthetic inputs shared by tests and golden generation (numpy/scipy only)."""
import numpy as np
from scipy.ndimage import gaussian_filter


def synth_volume(shape, seed):
    """SURVEY 8(d) recipe: min-max of a sigma-weighted sum of Gaussian-smoothed uniform noise."""
    rng = np.random.default_rng(seed)
    n = rng.random(shape)
    v = sum(s * gaussian_filter(n, s) for s in (1.5, 4.0, 8.0))
    return ((v - v.min()) / (v.max() - v.min())).astype(np.float32)


def smooth_flow(shape, seed, magnitude=2.0, sigma=8.0):
    """Smooth random displacement field (Z,Y,X,3) float32 with max |component| ~ magnitude."""
    rng = np.random.default_rng(seed)
    f = np.stack([gaussian_filter(rng.standard_normal(shape), sigma) for _ in range(3)], -1)
    f *= magnitude / np.abs(f).max()
    return f.astype(np.float32)
