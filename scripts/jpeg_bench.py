#!/usr/bin/env python
"""Device JPEG decoder (csrc/jpeg.cu) alone: B synthetic 375 x 500 quality-90 files, bytes in host memory -> RGB pixels
in HBM; checked against Pillow on the first files; CUDA events, median of 5."""
import io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from PIL import Image
from graphnet_classifier_b200 import build, ops
build.build()
from graphnet_classifier_b200.utils import jpeg as gjpeg
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rng = np.random.default_rng(0)
datas = []
for i in range(B):
    low = rng.integers(0, 256, (375 // 16 + 2, 500 // 16 + 2, 3), dtype=np.uint8)
    buf = io.BytesIO()
    Image.fromarray(low).resize((500, 375), Image.BICUBIC).save(buf, format="JPEG", quality=90)
    datas.append(buf.getvalue())
st = {}
out = gjpeg.decode_batch(datas, staging=st)
for d, t in list(zip(datas, out))[:8]:
    assert np.array_equal(t.cpu().numpy(), np.asarray(Image.open(io.BytesIO(d)).convert("RGB")))
ts = []
for _ in range(5):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); gjpeg.decode_batch(datas, staging=st); e.record(); torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
ms = sorted(ts)[2]
ops.PROFILE = ops.KernelProfile()
for _ in range(3):
    gjpeg.decode_batch(datas, staging=st)
prof = ops.PROFILE.summary()["jpeg_decode"]
ops.PROFILE = None
print(f"device kernels alone (memset + entropy + IDCT + upsampling/colour): {prof['ms'] / prof['calls']:.2f} ms per call "
      f"-> {B / (prof['ms'] / prof['calls']) * 1e3:,.0f} images/s; the rest of the call is host work (descriptor table, copies into pinned memory)")
print(f"{B} files of {sum(map(len, datas)) / B / 1e3:.1f} KB: {ms:.2f} ms -> {B / ms * 1e3:,.0f} images/s (bit-identical to Pillow on the checked files)")
