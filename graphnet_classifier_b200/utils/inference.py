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


def mlp_inference(image_path, weights="weights/MLP/final_model_.pth", resize_value=128):
    """The MLP baseline on one image (reference utils/inference.py:16-29): ``Image.open(p).resize((r, r)).convert('RGB')``
    - resize BEFORE the conversion, raw 0..255 pixels, flattened HWC - through ``MLP``; returns (logits, probabilities)."""
    import numpy as np
    from PIL import Image
    from .. import ops
    from ..models.MLP import MLP
    model = MLP(in_dim=resize_value * resize_value * 3, out_dim=2)
    model.load_state_dict(torch.load(weights, map_location="cpu"))
    model = model.cuda().eval()
    image = Image.open(image_path) if isinstance(image_path, str) else image_path
    with torch.no_grad():
        if image.mode == "RGB":      # the resize runs on the device, bit-identical to PIL's (conversion is then a no-op)
            pixels = ops.resize_bicubic(torch.from_numpy(np.array(image, dtype=np.uint8)).cuda(), resize_value, resize_value)
        else:                        # other modes resize in their own mode (palette images: NEAREST) - PIL's decode-side rule
            pixels = torch.from_numpy(np.array(image.resize((resize_value, resize_value)).convert("RGB"), dtype=np.uint8)).cuda()
        logits = model(pixels.reshape(1, -1).float())
        probabilities = F.softmax(logits, dim=1)
    return logits, probabilities


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
