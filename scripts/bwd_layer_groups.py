#!/usr/bin/env python
"""Fused backward-layer kernel against the dynamic range of dZ: rows of equal magnitude (every 4 blocks share a scale
group) vs per-GRAPH magnitudes (blocks of one graph share a scale, a new group per graph boundary) vs per-ROW magnitudes
over several decades (almost every block opens its own group: one accumulator drain per block)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphnet_classifier_b200 import ops, build
build.build()
B = 104
E = 2 * 128 * 127
M = B * E
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
base = torch.randn(M, 128, device=dev, generator=g) * 1e-4
X = torch.relu(torch.randn(M, 128, device=dev, generator=g))
W = torch.randn(128, 128, device=dev, generator=g) / 11
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
cases = {
    "uniform rows": base,
    "per-graph magnitudes (3 decades)": base * torch.pow(10.0, -3 * torch.rand(B, 1, 1, device=dev, generator=g)).expand(B, E, 1).reshape(M, 1),
    "per-row magnitudes (1 decade)": base * torch.pow(10.0, -1 * torch.rand(M, 1, device=dev, generator=g)),
    "per-row magnitudes (4 decades)": base * torch.pow(10.0, -4 * torch.rand(M, 1, device=dev, generator=g)),
    "per-32-row-block magnitudes (4 decades)": base * torch.pow(10.0, -4 * torch.rand(M // 32, 1, 1, device=dev, generator=g)).expand(M // 32, 32, 1).reshape(M, 1),
}
for name, dZ in cases.items():
    dZ = dZ.contiguous()
    for _ in range(2):
        ops.tc_bwd_layer(dZ, X, W, mask=True, want_db=True)
    ts = []
    for _ in range(5):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ops.tc_bwd_layer(dZ, X, W, mask=True, want_db=True); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    t = sorted(ts)[2]
    print(f"{name:45s} {t:7.3f} ms  {3 * M * 512 / t / 1e6:6.0f} GB/s")
