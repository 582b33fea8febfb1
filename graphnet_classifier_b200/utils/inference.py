"""``gnn_inference`` with the reference's signature and return value (reference
utils/inference.py:32-71): build the 3-block GraphNet + head for ``resize_value``, load a
checkpoint, classify one image, return ``(logits [1, classes], probabilities [1, classes])``.
Same kernels as the batched path; the debug prints of the reference are kept behind ``verbose``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from ..models.GNN import CombinedModel, GraphNet
from .image_to_graph.batched import build_pixel_graphs
from .image_to_graph.image_to_graph_optimized import load_rgb_device


def gnn_inference(image_path, weights_path: str = "weights/GNN/best_model_epoch2.pth", resize_value: int = 64,
                  verbose: bool = False):
    graph_net = GraphNet(num_local_features=3, space_dim=2, out_channels=1, n_blocks=3)
    model = CombinedModel(graph_net=graph_net, num_nodes=resize_value * resize_value, classes=2)
    state = torch.load(weights_path, map_location="cpu")
    model.load_state_dict(state)
    model = model.cuda().eval()
    with torch.no_grad():
        gb = build_pixel_graphs(load_rgb_device(image_path, resize_value))
        if verbose:
            print(f"Input x shape: {gb.x.shape}, x range: [{gb.x.min():.3f}, {gb.x.max():.3f}]")
            print(f"Input pos shape: {gb.pos.shape}, pos range: [{gb.pos.min():.3f}, {gb.pos.max():.3f}]")
            print(f"Input edge_index shape: {gb.edge_index.shape}")
        logits = model(gb.as_tuple())
        if logits.dim() == 1:
            logits = logits.unsqueeze(0)
        probabilities = F.softmax(logits, dim=-1)
    return logits, probabilities
