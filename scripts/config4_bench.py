#!/usr/bin/env python
"""BASELINE config 4 end to end: superpixel graphs at resize 256, batch 1024 - device SLIC, label map ->
graph (features, centroids, ordered adjacency), compaction into one block-diagonal batch, CSR build, GraphNet
forward (node outputs; the reference's dense head needs a fixed node count).  Per-stage CUDA-event
times, L2 flushed between repetitions."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphnet_classifier_b200 import ops, build
build.build()
from graphnet_classifier_b200.models.GNN import GraphNet
from graphnet_classifier_b200.utils.image_to_graph.batched import build_superpixel_graphs
from graphnet_classifier_b200.utils.image_to_graph.slic import slic_labels
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
r = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = "cuda"
torch.manual_seed(0)
net = GraphNet(num_local_features=3, space_dim=2, out_channels=1, n_blocks=3).cuda().eval()
# smooth random images (a few low-frequency blobs + noise) so that SLIC has structure to follow
g = torch.Generator(device=dev).manual_seed(0)
low = torch.rand(B, 3, 8, 8, device=dev, generator=g)
imgs = torch.nn.functional.interpolate(low, size=(r, r), mode="bilinear", align_corners=False)
imgs = (imgs + 0.05 * torch.randn(B, 3, r, r, device=dev, generator=g)).clamp(0, 1).mul(255).byte().permute(0, 2, 3, 1).contiguous()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def stage_times():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    flush.zero_()
    ev[0].record()
    labels = slic_labels(imgs, n_segments=100, compactness=10.0)
    ev[1].record()
    n_nodes, x, pos, n_edges, edges = build_superpixel_graphs(imgs, labels, max_nodes=128)
    ev[2].record()
    # compaction: block-diagonal batch of the per-image graphs
    S_max = x.shape[1]
    node_off = torch.cumsum(n_nodes, 0) - n_nodes
    nmask = torch.arange(S_max, device=dev)[None, :] < n_nodes[:, None]
    xb, pb = x[nmask], pos[nmask]
    E_max = edges.shape[2]
    emask = torch.arange(E_max, device=dev)[None, :] < n_edges[:, None]
    eb = (edges + node_off[:, None, None].long()).permute(1, 0, 2)[:, emask]
    ev[3].record()
    graph = ops.GraphIndex.from_edge_index(eb, xb.shape[0])
    ops.attach_graph(eb, graph)
    ev[4].record()
    with torch.no_grad():
        out = net(xb, pb, eb)
    ev[5].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(5)], int(xb.shape[0]), int(eb.shape[1]), out


for _ in range(2):
    stage_times()
runs = [stage_times() for _ in range(5)]
ts = [sorted(rn[0][i] for rn in runs)[2] for i in range(5)]
_, N, E, out = runs[-1]
names = ["SLIC (10 iterations)", "label map -> graph", "compaction (torch indexing)", "CSR build", "GraphNet forward"]
print(f"config 4: {B} images, resize {r}: {N} nodes, {E} edges in the batch ({N / B:.1f} nodes, {E / B:.1f} edges per image)")
for n, t in zip(names, ts):
    print(f"  {n:30s} {t:8.3f} ms")
tot = sum(ts)
print(f"  {'total':30s} {tot:8.3f} ms  -> {B / tot * 1e3:,.0f} images/s   (finite outputs: {bool(torch.isfinite(out).all())})")
