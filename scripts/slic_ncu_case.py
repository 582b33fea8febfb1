#!/usr/bin/env python
"""One SLIC call (k-means + connectivity post-pass) on 296 config-4 images for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphnet_classifier_b200.utils.image_to_graph.slic import slic_labels
B, r, dev = int(sys.argv[1]) if len(sys.argv) > 1 else 296, 256, "cuda"
g = torch.Generator(device=dev).manual_seed(0)
low = torch.rand(B, 3, 8, 8, device=dev, generator=g)
imgs = torch.nn.functional.interpolate(low, size=(r, r), mode="bilinear", align_corners=False)
imgs = (imgs + 0.05 * torch.randn(B, 3, r, r, device=dev, generator=g)).clamp(0, 1).mul(255).byte().permute(0, 2, 3, 1).contiguous()
for _ in range(2):
    lab = slic_labels(imgs)
torch.cuda.synchronize()
print("ok", int(lab.max()))
