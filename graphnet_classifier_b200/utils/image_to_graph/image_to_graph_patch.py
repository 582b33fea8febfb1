"""Patch-graph builder (reference utils/image_to_graph/image_to_graph_patch.py) on the
``gnc_build_patch_graph_u8`` kernel."""
from __future__ import annotations

import numpy as np
import torch

from .batched import build_patch_graphs
from .image_to_graph_optimized import _load_rgb


def image_to_graph_patch(image_or_path, resize_value=128, patch_size=8):
    """(x float [n, 3] tile means 0..255, pos int64 [n, 2] tile centres, edge_index
    int64 [2, E]) - reference :6-54.  ``x`` is float32 here (the reference returns
    float64 lists that its loader casts to float32; values are identical after that cast)."""
    tab = _load_rgb(image_or_path, resize_value)
    gb = build_patch_graphs(torch.from_numpy(np.ascontiguousarray(tab)), patch_size=patch_size, use_cache=False)
    return gb.x.cpu().numpy(), gb.pos.to(torch.int64).cpu().numpy(), gb.edge_index.cpu().numpy()
