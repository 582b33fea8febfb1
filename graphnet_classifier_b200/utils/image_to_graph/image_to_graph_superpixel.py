"""Superpixel-graph builder (reference utils/image_to_graph/image_to_graph_superpixel.py).

Two stages: (1) SLIC segmentation -> integer label map; (2) label map -> node
features, centroids and 4-connected adjacency.  Stage 2 runs in the
``gnc_build_superpixel_graph`` kernel and is pinned bit-exactly (indices) against the
reference.  Stage 1 is scikit-image's ``slic`` in the reference - un-vendored, unpinned
and not installed here; ``segments=`` lets the caller supply a label map, otherwise
``slic_labels`` (slic.py, a from-the-paper device SLIC, parity unpinned) is used.
"""
from __future__ import annotations

import numpy as np
import torch

from .batched import build_superpixel_graphs
from .image_to_graph_optimized import _load_rgb


def image_to_graph_superpixel(image_or_path, resize_value=128, n_segments=100, compactness=10, segments=None):
    """(x [S, 3] mean RGB in 0..1, pos [S, 2] centroid (row, col), edge_index int64
    [2, E]) - reference :8-73."""
    tab = np.ascontiguousarray(_load_rgb(image_or_path, resize_value))
    if segments is None:
        from .slic import slic_labels
        segments = slic_labels(torch.from_numpy(tab).cuda(), n_segments=n_segments, compactness=compactness)[0]
    seg = torch.as_tensor(np.asarray(segments) if not isinstance(segments, torch.Tensor) else segments)
    n_nodes, x, pos, n_edges, edges = build_superpixel_graphs(torch.from_numpy(tab), seg)
    S, E = int(n_nodes[0].item()), int(n_edges[0].item())
    xs = x[0, :S].cpu().numpy()
    ps = pos[0, :S].cpu().numpy()
    if E == 0:
        return xs, ps, np.empty((2, 0))           # reference :70-71 (float64, no edges)
    return xs, ps, edges[0, :, :E].cpu().numpy()
