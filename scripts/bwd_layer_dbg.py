#!/usr/bin/env python
"""Timing experiments on the fused backward-layer kernel with stages switched off (GNC_BWD_DBG bit mask:
1 no dX MMAs, 2 no dW MMAs, 4 no dW drain reads, 8 no dX stores, 16 no conversion).  Results are invalid with a mask."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphnet_classifier_b200 import ops, build
build.build()
B = 104
M = B * 2 * 128 * 127
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
dZ = torch.randn(M, 128, device=dev, generator=g) * 1e-4
X = torch.relu(torch.randn(M, 128, device=dev, generator=g))
W = torch.randn(128, 128, device=dev, generator=g) / 11
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
addend_mode = "addend" in sys.argv
ad = torch.randn(M, 128, device=dev, generator=g) * 1e-4 if addend_mode else None
masks = [int(a) for a in sys.argv[1:] if a != "addend"] or [0, 1, 2, 3, 4, 8, 16, 7, 15, 31]
for mask in masks:
    os.environ["GNC_BWD_DBG"] = str(mask)
    ts = []
    for rep in range(5):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ops.tc_bwd_layer(dZ, X, W, mask=not addend_mode, addend=ad, want_db=True); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    print(f"dbg mask {mask & 255:2d} prefetch-ahead {(mask >> 8) - 1 if mask >> 8 else 'default'}: {sorted(ts)[2]:.3f} ms")
