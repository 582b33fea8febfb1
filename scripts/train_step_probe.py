#!/usr/bin/env python
"""One training step at resize 128 on a small batch (for ncu launch lists / kernel shares)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphnet_classifier_b200 import ops, build
build.build()
from graphnet_classifier_b200.models.GNN import CombinedModel, GraphNet
from graphnet_classifier_b200.pipeline import GraphClassifierPipeline
from graphnet_classifier_b200.utils.distributed import GradBucket
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
r = 128
torch.manual_seed(0)
model = CombinedModel(GraphNet(num_local_features=3, space_dim=2, out_channels=1, n_blocks=3), num_nodes=r * r).cuda()
pipe = GraphClassifierPipeline(model, resize_value=r)
imgs = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (B, r, r, 3), dtype=np.uint8)).cuda()
labels = torch.from_numpy(np.random.default_rng(1).integers(0, 2, B)).cuda()
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
bucket = GradBucket(model.parameters())
for i in range(3):
    ops.PROFILE = ops.KernelProfile() if i == 2 else None
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); loss = pipe.train_step(imgs, labels, opt, grad_bucket=bucket); e.record(); torch.cuda.synchronize()
    print(f"step {i}: {s.elapsed_time(e):.2f} ms loss {loss.item():.4f}")
prof = ops.PROFILE.summary(); ops.PROFILE = None
tot = sum(d["ms"] for d in prof.values())
for k, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k:18s} {d['ms']:8.3f} ms  {100 * d['ms'] / tot:5.1f}%  x{d['calls']}")
print("  sum of profiled libgnc launches", round(tot, 2), "ms")
