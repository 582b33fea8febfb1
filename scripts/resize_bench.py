#!/usr/bin/env python
"""Input staging: Pillow-exact bicubic resize on the device vs PIL on one host core.
Algorithmic bytes per image: 3*H*W read + 3*r*r written."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from PIL import Image
from graphnet_classifier_b200 import ops, build
build.build()
peak = 6547.8
try:
    import json
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
print("| source | target | batch | device ms | images/s | algorithmic GB/s | % of HBM peak | PIL 1 core images/s | bit-exact |")
print("|---|---|---|---|---|---|---|---|---|")
for (H, W, r, B) in [(375, 500, 128, 512), (256, 256, 128, 1024), (1080, 1920, 256, 64), (3000, 4000, 128, 16), (64, 64, 128, 2048)]:
    rng = np.random.default_rng(0)
    imgs = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    dev = torch.from_numpy(imgs).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        out = ops.resize_bicubic(dev, r, r)
    ts = []
    for _ in range(7):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); out = ops.resize_bicubic(dev, r, r); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = float(np.median(ts))
    n_cpu = min(B, 16)
    t0 = time.perf_counter()
    refs = [np.asarray(Image.fromarray(imgs[i]).resize((r, r))) for i in range(n_cpu)]
    cpu = n_cpu / (time.perf_counter() - t0)
    exact = all(np.array_equal(out[i].cpu().numpy(), refs[i]) for i in range(n_cpu))
    nbytes = 3.0 * B * (H * W + r * r)
    gbs = nbytes / ms / 1e6
    print(f"| {H}x{W} | {r} | {B} | {ms:.3f} | {B / ms * 1e3:,.0f} | {gbs:,.0f} | {100 * gbs / peak:.1f} | {cpu:,.0f} | {exact} |")
