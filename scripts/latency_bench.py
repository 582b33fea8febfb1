#!/usr/bin/env python
"""Single-image / small-batch latency of the public inference call: eager launches
vs the CUDA-graph replay, pinned host image in, logits back on the host."""
import sys, os, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphnet_classifier_b200 import build
build.build()
from graphnet_classifier_b200.models.GNN import CombinedModel, GraphNet
from graphnet_classifier_b200.pipeline import GraphClassifierPipeline
r = int(sys.argv[1]) if len(sys.argv) > 1 else 128
torch.manual_seed(0)
model = CombinedModel(GraphNet(num_local_features=3, space_dim=2, out_channels=1, n_blocks=3), num_nodes=r * r).cuda().eval()
pipe = GraphClassifierPipeline(model, resize_value=r)
rng = np.random.default_rng(0)
for B in (1, 4, 16):
    imgs = [torch.from_numpy(rng.integers(0, 256, (B, r, r, 3), dtype=np.uint8)).pin_memory() for _ in range(8)]
    res = {}
    for name, fn in (("eager", pipe.infer), ("graphed", pipe.infer_graphed)):
        for i in range(5):
            fn(imgs[i % 8]).cpu()
        torch.cuda.synchronize()
        ts = []
        for i in range(60):
            t0 = time.perf_counter()
            out = fn(imgs[i % 8]).cpu()
            ts.append((time.perf_counter() - t0) * 1e3)
        res[name] = (statistics.median(ts), min(ts), out)
    same = torch.equal(pipe.infer(imgs[3]).cpu(), pipe.infer_graphed(imgs[3]).cpu())
    print(f"resize {r} batch {B:3d}: eager {res['eager'][0]:7.3f} ms (min {res['eager'][1]:.3f})   graphed {res['graphed'][0]:7.3f} ms "
          f"(min {res['graphed'][1]:.3f})   identical logits: {same}")
