#!/usr/bin/env python
"""Small invocations of the round-2 kernels for `compute-sanitizer --tool memcheck` (one tool per call, small shapes):
SLIC image kernel + streaming form, connectivity (both forms), superpixel graph (both kernels), device JPEG decode,
deferred backward reduction, segment readout."""
import io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from PIL import Image
from graphnet_classifier_b200 import _lib, ops
from graphnet_classifier_b200.utils.image_to_graph.slic import slic_labels, enforce_connectivity
from graphnet_classifier_b200.utils.image_to_graph.batched import build_superpixel_graphs, build_superpixel_batch
from graphnet_classifier_b200.utils import jpeg as gjpeg
lib = _lib.load()
rng = np.random.default_rng(0)
for (B, H, W, S) in ((3, 64, 64, 16), (2, 40, 72, 12), (1, 256, 256, 100), (2, 50, 61, 9)):
    img = torch.from_numpy(rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)).cuda()
    for run in (8, -8):
        lib.gnc_debug_slic_run_length(run)
        raw = slic_labels(img, n_segments=S, enforce_connectivity_=False)
    lib.gnc_debug_slic_run_length(8)
    for streaming in (0, 1):
        lib.gnc_debug_slic_connect_streaming(streaming)
        lab = enforce_connectivity(raw, 6)
    lib.gnc_debug_slic_connect_streaming(0)
    build_superpixel_graphs(img, lab)
    gb = build_superpixel_batch(img, labels=lab)
torch.cuda.synchronize()
datas = []
for h, w, kw in ((33, 47, dict(quality=75)), (17, 23, dict(quality=95, subsampling=0)), (40, 50, dict(quality=60, subsampling=1)),
                 (50, 70, dict(quality=80, restart_marker_blocks=3)), (1, 1, dict(quality=90)), (64, 64, dict(quality=90))):
    buf = io.BytesIO()
    Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)).save(buf, format="JPEG", **kw)
    datas.append(buf.getvalue())
out = gjpeg.decode_batch(datas)
torch.cuda.synchronize()
for d, t in zip(datas, out):
    assert np.array_equal(t.cpu().numpy(), np.asarray(Image.open(io.BytesIO(d)).convert("RGB")))
g = torch.Generator(device="cuda").manual_seed(0)
with ops.DeferredBwdReduce():
    for M in (5, 1000, 148 * 32 + 7):
        ops.tc_bwd_layer(torch.randn(M, 128, device="cuda", generator=g), torch.relu(torch.randn(M, 128, device="cuda", generator=g)),
                         torch.randn(128, 128, device="cuda", generator=g) / 11, mask=True, want_db=True)
torch.cuda.synchronize()
print("sanitize case ok")
