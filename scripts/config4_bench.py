#!/usr/bin/env python
"""BASELINE config 4 end to end: superpixel graphs at resize 256, batch 1024 - device SLIC (10 iterations + connectivity
post-pass), label map -> per-image graphs -> one block-diagonal batch + CSR (build_superpixel_batch), GraphNet + pad /
truncate readout + head -> logits.  Per-stage CUDA-event times, L2 flushed between repetitions."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphnet_classifier_b200 import ops, build
build.build()
from graphnet_classifier_b200.models.GNN import CombinedModel, GraphNet
from graphnet_classifier_b200.utils.image_to_graph.batched import build_superpixel_batch
from graphnet_classifier_b200.utils.image_to_graph.slic import slic_labels
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
r = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = "cuda"
torch.manual_seed(0)
net = CombinedModel(GraphNet(num_local_features=3, space_dim=2, out_channels=1, n_blocks=3), num_nodes=r // 2, classes=2).cuda().eval()
# smooth random images (a few low-frequency blobs + noise) so that SLIC has structure to follow
g = torch.Generator(device=dev).manual_seed(0)
low = torch.rand(B, 3, 8, 8, device=dev, generator=g)
imgs = torch.nn.functional.interpolate(low, size=(r, r), mode="bilinear", align_corners=False)
imgs = (imgs + 0.05 * torch.randn(B, 3, r, r, device=dev, generator=g)).clamp(0, 1).mul(255).byte().permute(0, 2, 3, 1).contiguous()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def stage_times():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    flush.zero_()
    ev[0].record()
    labels = slic_labels(imgs, n_segments=100, compactness=10.0)
    ev[1].record()
    gb = build_superpixel_batch(imgs, labels=labels, max_nodes=128)
    ev[2].record()
    with torch.no_grad():
        out = net(gb.as_tuple())
    ev[3].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(3)], int(gb.x.shape[0]), int(gb.edge_index.shape[1]), out


for _ in range(2):
    stage_times()
runs = [stage_times() for _ in range(5)]
ts = [sorted(rn[0][i] for rn in runs)[2] for i in range(3)]
_, N, E, out = runs[-1]
names = ["SLIC (10 iterations + connectivity)", "label maps -> block-diagonal batch + CSR", "GraphNet + readout + head"]
print(f"config 4: {B} images, resize {r}: {N} nodes, {E} edges in the batch ({N / B:.1f} nodes, {E / B:.1f} edges per image)")
for n, t in zip(names, ts):
    print(f"  {n:42s} {t:8.3f} ms")
tot = sum(ts)
print(f"  {'total':42s} {tot:8.3f} ms  -> {B / tot * 1e3:,.0f} images/s   (logits {tuple(out.shape)}, finite: {bool(torch.isfinite(out).all())})")
