#!/usr/bin/env python
"""Stage times of the device SLIC (colour conversion + k-means iterations, connectivity post-pass) on BASELINE config 4's
images (resize 256), default forms against the streaming forms."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphnet_classifier_b200 import build, _lib
build.build()
from graphnet_classifier_b200.utils.image_to_graph.slic import slic_labels, enforce_connectivity
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
r = 256
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
low = torch.rand(B, 3, 8, 8, device=dev, generator=g)
imgs = torch.nn.functional.interpolate(low, size=(r, r), mode="bilinear", align_corners=False)
imgs = (imgs + 0.05 * torch.randn(B, 3, r, r, device=dev, generator=g)).clamp(0, 1).mul(255).byte().permute(0, 2, 3, 1).contiguous()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lib = _lib.load()


def t(fn, n=5):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[n // 2]


raw = slic_labels(imgs, enforce_connectivity_=False)
K = int(lib.gnc_slic_num_centers(r, r, 100))
ms = int(0.5 * r * r / K)
print(f"{B} images {r}x{r}, {K} centres")
for name, run, stream in (("default (one CTA per image)", 8, 0), ("streaming", -8, 1)):
    lib.gnc_debug_slic_run_length(run); lib.gnc_debug_slic_connect_streaming(stream)
    for iters in (0, 1, 10):
        print(f"  {name:30s} k-means, {iters:2d} iterations: {t(lambda: slic_labels(imgs, max_num_iter=iters, enforce_connectivity_=False)):8.3f} ms")
    print(f"  {name:30s} connectivity post-pass:    {t(lambda: enforce_connectivity(raw, ms)):8.3f} ms")
lib.gnc_debug_slic_run_length(8); lib.gnc_debug_slic_connect_streaming(0)
