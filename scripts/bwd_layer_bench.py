#!/usr/bin/env python
"""Micro-benchmark of the fused backward-layer kernel (csrc/tc_bwd.cu) against the round-1 pair (tc_linear data
gradient with the mask epilogue + tc_wgrad) on the edge rows of one training micro-batch (CUDA events, L2 flushed,
variants interleaved, medians)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphnet_classifier_b200 import ops, build
build.build()

B = int(sys.argv[1]) if len(sys.argv) > 1 else 104
M = B * 2 * 128 * 127
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
dZ = torch.randn(M, 128, device=dev, generator=g) * 1e-4
X = torch.relu(torch.randn(M, 128, device=dev, generator=g))
W = torch.randn(128, 128, device=dev, generator=g) / 11
ad = torch.randn(M, 128, device=dev, generator=g) * 1e-4
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def pair():
    dW, db = ops.tc_wgrad(dZ, X, want_db=True)
    return ops.tc_linear(dZ, W, transpose_w=True, mask=X), dW


def pair_addend():
    dW, db = ops.tc_wgrad(dZ, X, want_db=True)
    return ops.tc_linear(dZ, W, transpose_w=True, addend=ad), dW


def fused():
    return ops.tc_bwd_layer(dZ, X, W, mask=True, want_db=True)


def fused_addend():
    return ops.tc_bwd_layer(dZ, X, W, addend=ad, want_db=True)


variants = [("pair (tc_wgrad + tc_linear mask)", pair, 5), ("fused mask", fused, 3),
            ("pair (tc_wgrad + tc_linear addend)", pair_addend, 6), ("fused addend", fused_addend, 4)]
times = {n: [] for n, _, _ in variants}
for name, fn, _ in variants:
    fn()
torch.cuda.synchronize()
for rep in range(7):
    for name, fn, _ in variants:
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        times[name].append(s.elapsed_time(e))
print(f"rows M = {M} ({B} graphs of resize 128), one [M,128] fp32 tensor = {M * 512 / 1e9:.2f} GB")
for name, fn, rows in variants:
    t = sorted(times[name])[len(times[name]) // 2]
    gb = rows * M * 512 / 1e9
    print(f"{name:38s} {t:8.3f} ms   {gb / t * 1e3:7.0f} GB/s algorithmic ({rows} rows of traffic per input row)")
