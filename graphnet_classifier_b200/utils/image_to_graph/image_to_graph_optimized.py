"""Pixel-graph builder with the reference's function names and return types
(reference utils/image_to_graph/image_to_graph_optimized.py), computed by the
``gnc_build_pixel_graph_u8`` kernel.

These are the single-image, host-visible entry points: numpy in the reference's
dtypes (uint8 ``x``, int64 ``pos``, int64 ``edge_index``), so existing callers keep
working.  The device-resident batched form is ``build_pixel_graphs`` (batched.py).
The file decode (``Image.open`` + ``convert('RGB')``) stays on the host; the resize to the working
resolution runs on the device, bit-identical to PIL's (``ops.resize_bicubic``).
"""
from __future__ import annotations

from functools import lru_cache

import numpy as np
import torch
from PIL import Image

from ... import ops
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


def load_rgb_device(image_or_path, resize_value) -> torch.Tensor:
    """``image.convert('RGB').resize((r, r))`` (reference :65-70) as ``uint8 [r, r, 3]`` on the device:
    decoded on the host, resized by the Pillow-exact device kernel."""
    image = Image.open(image_or_path) if isinstance(image_or_path, str) else image_or_path
    rgb = torch.from_numpy(np.array(image.convert("RGB"), dtype=np.uint8))
    dev = torch.device("cuda", torch.cuda.current_device())
    return ops.resize_bicubic(rgb.to(dev, non_blocking=True), int(resize_value), int(resize_value))


def _load_rgb(image_or_path, resize_value):
    return load_rgb_device(image_or_path, resize_value).cpu().numpy()


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
