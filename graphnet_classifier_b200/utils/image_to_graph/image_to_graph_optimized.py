"""Pixel-graph builder with the reference's function names and return types
(reference utils/image_to_graph/image_to_graph_optimized.py), computed by the
``gnc_build_pixel_graph_u8`` kernel.

These are the single-image, host-visible entry points: numpy in the reference's
dtypes (uint8 ``x``, int64 ``pos``, int64 ``edge_index``), so existing callers keep
working.  The device-resident batched form is ``build_pixel_graphs`` (batched.py).
PIL decode / resize stay on the host (out of the hot path).
"""
from __future__ import annotations

from functools import lru_cache

import numpy as np
import torch
from PIL import Image

from .batched import build_pixel_graphs


def _grid_edges_device(H: int, W: int, diagonals: bool) -> np.ndarray:
    dummy = torch.zeros(1, H, W, 3, dtype=torch.uint8, device="cuda")
    gb = build_pixel_graphs(dummy, diagonals=diagonals, use_cache=False)
    # the reference returns the transposed view of an [E, 2] array; values are what matter
    return gb.edge_index.cpu().numpy()


def create_grid_edges_optimized(H, W, diagonals=False):
    """int64 ``[2, E]`` directed grid edges: horizontal, vertical, then the two diagonal
    families, each row-major (reference image_to_graph_optimized.py:7-39)."""
    return _grid_edges_device(int(H), int(W), bool(diagonals))


@lru_cache(maxsize=128)
def get_cached_edge_index(resize_value, diagonals):
    """Same memoisation as the reference (:42-47): one array object per (size, diagonals)."""
    return create_grid_edges_optimized(resize_value, resize_value, diagonals)


def _load_rgb(image_or_path, resize_value):
    image = Image.open(image_or_path) if isinstance(image_or_path, str) else image_or_path
    return np.asarray(image.convert("RGB").resize((resize_value, resize_value)))


def image_to_graph_pixel_optimized(image_or_path, resize_value=128, diagonals=False, use_cache=True):
    """(x uint8 [N, 3], pos int64 [N, 2], edge_index int64 [2, E]) for one image
    (reference :50-87)."""
    tab = _load_rgb(image_or_path, resize_value)
    gb = build_pixel_graphs(torch.from_numpy(np.ascontiguousarray(tab)), diagonals=diagonals, use_cache=False)
    x = gb.x.to(torch.uint8).cpu().numpy()
    pos = gb.pos.to(torch.int64).cpu().numpy()
    if use_cache:
        edge_index = get_cached_edge_index(resize_value, diagonals)
    else:
        edge_index = gb.edge_index.cpu().numpy()
    return x, pos, edge_index
