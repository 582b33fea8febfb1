#!/usr/bin/env python
"""BASELINE config 4 (graph construction): pixel-grid build at resize 256 x batch 1024 and the
label-map -> superpixel-graph stage (100-seed Voronoi labels), plus resize 128 x 512.
Bytes per SURVEY.md 8(d): per image 3N in, 12N (x) out, + topology once per shape."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphnet_classifier_b200 import build
from graphnet_classifier_b200.utils.image_to_graph.batched import build_pixel_graphs, build_superpixel_graphs, build_patch_graphs
build.build()
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return sum(ts) / len(ts)
print("| case | images | ms | images/s | algorithmic GB | GB/s | frac of 6547.8 |")
print("|---|---|---|---|---|---|---|")
for r, B in ((128, 512), (256, 1024)):
    imgs = torch.randint(0, 256, (B, r, r, 3), dtype=torch.uint8, device=dev)
    N, E = r * r, 2 * r * (r - 1)
    for name, cache, nbytes in (("pixel build, topology cached (steady state)", True, B * N * 15.0),
                                ("pixel build, full (x, pos, edge_index, int32 ends, 2 CSRs)", False,
                                 B * (N * 15.0 + N * 8 + E * 16 + E * 8 + 2 * (4 * (N + 1) + 4 * E)))):
        ms = timeit(lambda: build_pixel_graphs(imgs, use_cache=cache))
        print(f"| r={r} {name} | {B} | {ms:.3f} | {B/ms*1e3:.0f} | {nbytes/1e9:.3f} | {nbytes/ms/1e6:.0f} | {nbytes/ms/1e6/6547.8:.3f} |", flush=True)
    ms = timeit(lambda: build_patch_graphs(imgs, patch_size=8, use_cache=True))
    print(f"| r={r} patch(8) build, topology cached | {B} | {ms:.3f} | {B/ms*1e3:.0f} | {B*N*3/1e9:.3f} | {B*N*3/ms/1e6:.0f} | {B*N*3/ms/1e6/6547.8:.3f} |", flush=True)
    # superpixel: label map -> graph (S = 100 Voronoi cells per image)
    rng = np.random.default_rng(0)
    sites = torch.tensor(rng.random((B, 100, 2)) * r, device=dev, dtype=torch.float32)
    yy, xx = torch.meshgrid(torch.arange(r, device=dev) + 0.5, torch.arange(r, device=dev) + 0.5, indexing="ij")
    labs = torch.empty(B, r, r, dtype=torch.int32, device=dev)
    for b0 in range(0, B, 64):
        s = sites[b0:b0 + 64]
        d = (yy[None, :, :, None] - s[:, None, None, :, 0]) ** 2 + (xx[None, :, :, None] - s[:, None, None, :, 1]) ** 2
        labs[b0:b0 + 64] = d.argmin(-1).int()
    ms = timeit(lambda: build_superpixel_graphs(imgs, labs, max_nodes=100))
    nb = B * N * (3 + 4.0)
    print(f"| r={r} superpixel label-map -> graph (S=100) | {B} | {ms:.3f} | {B/ms*1e3:.0f} | {nb/1e9:.3f} | {nb/ms/1e6:.0f} | {nb/ms/1e6/6547.8:.3f} |", flush=True)
