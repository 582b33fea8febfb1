"""Torch-RNG-independent deterministic parameters for parity fixtures.

TEST INFRASTRUCTURE - see oracle/__init__.py.  Golden vectors must not depend on
the torch version's default-init RNG stream, so fixtures fill every parameter
from numpy's PCG64 keyed by (seed, position in state_dict order): weights and
biases uniform in +-1/sqrt(fan_in) (the scale nn.Linear's default init uses),
LayerNorm gains 1 + small noise and shifts small noise so that the affine part
of the norm is actually exercised.
"""
from __future__ import annotations

import numpy as np
import torch


def fill_deterministic(model: torch.nn.Module, seed: int = 0) -> None:
    with torch.no_grad():
        for i, (name, p) in enumerate(model.state_dict().items()):
            rng = np.random.default_rng([seed, i])
            if p.dim() == 2:
                bound = 1.0 / np.sqrt(p.shape[1])
                v = rng.uniform(-bound, bound, size=tuple(p.shape))
            else:
                is_norm_gain = name.endswith("model.5.weight") or name.endswith("norm.weight")
                if is_norm_gain:
                    v = 1.0 + 0.1 * rng.standard_normal(tuple(p.shape))
                else:
                    v = 0.05 * rng.standard_normal(tuple(p.shape))
            p.copy_(torch.from_numpy(v.astype(np.float32)))


def fill_parameters(model: torch.nn.Module, seed: int = 0) -> None:
    """Deterministic values for ``named_parameters`` only (buffers such as BatchNorm's counters stay): matrices
    uniform in +-1/sqrt(fan_in), vectors 0.5 + 0.1 N(0, 1)."""
    with torch.no_grad():
        for i, (_, p) in enumerate(model.named_parameters()):
            g = np.random.default_rng([seed, i])
            v = g.uniform(-1.0, 1.0, tuple(p.shape)) / np.sqrt(p.shape[-1]) if p.dim() == 2 else 0.5 + 0.1 * g.standard_normal(tuple(p.shape))
            p.copy_(torch.from_numpy(v.astype(np.float32)))


def synthetic_images(batch: int, resize: int, seed: int = 0) -> np.ndarray:
    """BASELINE.md section 4 inputs: uint8 [B, r, r, 3], i.i.d. uniform 0..255."""
    return np.random.default_rng(seed).integers(0, 256, size=(batch, resize, resize, 3), dtype=np.uint8)


def voronoi_labels(H: int, W: int, n_seeds: int, seed: int = 0) -> np.ndarray:
    """Synthetic superpixel label map: nearest of
    ``n_seeds`` random sites; int64 [H, W] with labels 0..n_seeds-1 (a label may
    be absent, which exercises the rank-among-unique renumbering)."""
    rng = np.random.default_rng(seed)
    sites = rng.random((n_seeds, 2)) * np.array([H, W])
    rr, cc = np.meshgrid(np.arange(H) + 0.5, np.arange(W) + 0.5, indexing="ij")
    d = (rr[..., None] - sites[:, 0]) ** 2 + (cc[..., None] - sites[:, 1]) ** 2
    return np.argmin(d, axis=-1).astype(np.int64)
