#!/usr/bin/env python
"""A/B of library variants (graphnet_classifier_b200/variants/libgnc_<name>.so) on the SLIC k-means kernel: config-4 images,
10 iterations, interleaved rounds; prints the median time and a checksum of the labels (variants must agree bit for bit).
    python scripts/slic_ab.py base,rows8 [rounds]"""
import os, subprocess, sys, statistics, collections
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, root)
    import torch
    from graphnet_classifier_b200.utils.image_to_graph.slic import slic_labels
    B, r, dev = 1024, 256, "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    low = torch.rand(B, 3, 8, 8, device=dev, generator=g)
    imgs = torch.nn.functional.interpolate(low, size=(r, r), mode="bilinear", align_corners=False)
    imgs = (imgs + 0.05 * torch.randn(B, 3, r, r, device=dev, generator=g)).clamp(0, 1).mul(255).byte().permute(0, 2, 3, 1).contiguous()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    fn = lambda: slic_labels(imgs, max_num_iter=10, enforce_connectivity_=False)
    lab = fn(); fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    w = torch.arange(lab.numel(), device=dev, dtype=torch.int64) % 1000003
    print(sorted(ts)[2], int((lab.reshape(-1).long() * w).sum()))
    sys.exit(0)
names = sys.argv[1].split(",")
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
res, sums = collections.defaultdict(list), {}
for _ in range(rounds):
    for n in names:
        lib, _, minb = n.partition("@")            # "name@3": GNC_SLIC_MINB=3 (CTAs per SM the kernel is compiled for)
        env = dict(os.environ, GNC_LIB=os.path.join(root, "graphnet_classifier_b200", "variants", f"libgnc_{lib}.so"))
        if minb:
            env["GNC_SLIC_MINB"] = minb
        out = subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True, text=True)
        if out.returncode != 0:
            print(n, "FAILED", out.stderr[-600:]); continue
        t, c = out.stdout.split()[-2:]
        res[n].append(float(t)); sums[n] = c
for n in names:
    if res[n]:
        print(f"{n:12s} k-means 10 iterations, 1024 images: median {statistics.median(res[n]):7.3f} ms (min {min(res[n]):7.3f})  labels checksum {sums[n]}")
