#!/usr/bin/env python
"""Files -> logits (utils.staging.infer_files) for chunk sizes of the decode / inference pipeline, device JPEG decode."""
import os, sys, time, tempfile, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from PIL import Image
from graphnet_classifier_b200 import build
build.build()
from graphnet_classifier_b200.models.GNN import CombinedModel, GraphNet
from graphnet_classifier_b200.pipeline import GraphClassifierPipeline
from graphnet_classifier_b200.utils.staging import DecodePool, infer_files
B, r = 512, 128
torch.manual_seed(0)
model = CombinedModel(GraphNet(num_local_features=3, space_dim=2, out_channels=1, n_blocks=3), num_nodes=r * r).cuda()
pipe = GraphClassifierPipeline(model, resize_value=r)
rng = np.random.default_rng(0)
d = tempfile.mkdtemp()
try:
    paths = []
    for i in range(B):
        low = rng.integers(0, 256, (375 // 16 + 2, 500 // 16 + 2, 3), dtype=np.uint8)
        p = os.path.join(d, f"{i}.jpg"); Image.fromarray(low).resize((500, 375), Image.BICUBIC).save(p, quality=90); paths.append(p)
    with DecodePool() as pool:
        for chunk in (512, 256, 128, 64):
            infer_files(pipe, paths, pool=pool, chunk=chunk).cpu()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(3): infer_files(pipe, paths, pool=pool, chunk=chunk).cpu()
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
            print(f"chunk {chunk:4d}: {dt * 1e3:7.2f} ms per {B} files -> {B / dt:,.0f} graphs/s")
finally:
    shutil.rmtree(d, ignore_errors=True)
