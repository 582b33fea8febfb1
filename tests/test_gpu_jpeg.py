"""Device JPEG decoder (csrc/jpeg.cu through utils/jpeg.py) against Pillow - the decoder behind the reference's
``Image.open(path).convert('RGB')`` (utils/dataloader.py:34, image_to_graph_optimized.py:65-68) - bit for bit: the
reference's shipped JPEGs (tests/golden/jpeg_files.npz: file bytes + the pixels the unmodified reference builder saw),
generated files over sizes that are not multiples of the MCU, the three chroma layouts, greyscale, optimised Huffman
tables, restart intervals, low / high quality; files outside the decoder's scope come back as None and take the host
path with the same pixels."""
import io
import os

import numpy as np
import pytest
import torch
from PIL import Image

from oracle.jpeg import decode_baseline

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jpeg_files.npz")


def _jpeg(rng, h, w, gray=False, noise=12, **kw):
    low = rng.integers(0, 256, (h // 6 + 2, w // 6 + 2, 3), dtype=np.uint8)
    arr = np.asarray(Image.fromarray(low).resize((w, h), Image.BICUBIC)).astype(int)
    arr = np.clip(arr + rng.integers(-noise, noise + 1, (h, w, 3)), 0, 255).astype(np.uint8)
    im = Image.fromarray(arr)
    if gray:
        im = im.convert("L")
    buf = io.BytesIO()
    im.save(buf, format="JPEG", **kw)
    return buf.getvalue()


CASES = [(32, 32, dict(quality=90)), (33, 47, dict(quality=75)), (17, 23, dict(quality=95, subsampling=0)),
         (40, 50, dict(quality=60, subsampling=1)), (64, 48, dict(quality=85, optimize=True)),
         (31, 65, dict(quality=90, gray=True)), (50, 70, dict(quality=80, restart_marker_blocks=3)),
         (57, 91, dict(quality=80, restart_marker_rows=1, subsampling=1)), (20, 3, dict(quality=90)), (5, 4, dict(quality=90)),
         (100, 150, dict(quality=30)), (16, 16, dict(quality=100, subsampling=2)), (8, 8, dict(quality=50, subsampling=0)),
         (1, 1, dict(quality=90)), (375, 500, dict(quality=90)), (256, 256, dict(quality=97, noise=60)),
         (129, 255, dict(quality=88, subsampling=1, optimize=True)),
         # scans of more than 4 KB take the parallel entropy decoder (32 segments per image, self-synchronisation)
         (300, 420, dict(quality=85, subsampling=0)), (333, 517, dict(quality=70, subsampling=1)),
         (400, 300, dict(quality=92, gray=True)), (512, 512, dict(quality=99, noise=90)), (240, 320, dict(quality=95, optimize=True, noise=40)),
         (600, 800, dict(quality=25)), (97, 1201, dict(quality=90, subsampling=2))]


def test_shipped_jpegs_equal_reference_pixels(libgnc):
    from graphnet_classifier_b200.utils.jpeg import decode_batch
    g = np.load(GOLDEN)
    names = sorted(k[:-6] for k in g.files if k.endswith("_bytes"))
    datas = [g[n + "_bytes"].tobytes() for n in names]
    out = decode_batch(datas)
    for n, t in zip(names, out):
        assert t is not None and t.dtype == torch.uint8
        assert np.array_equal(t.cpu().numpy(), g[n + "_rgb"]), n


@pytest.mark.parametrize("sequential", [0, 1])
def test_generated_files_equal_pillow(libgnc, sequential):
    """Both forms of the entropy stage: the parallel one (default for scans of at least 4 KB without restart intervals)
    and the sequential one (debug switch; always used for short scans and restart intervals)."""
    from graphnet_classifier_b200.utils.jpeg import decode_batch
    rng = np.random.default_rng(7)
    datas = [_jpeg(rng, h, w, **dict(kw)) for h, w, kw in CASES]
    libgnc.gnc_debug_jpeg_sequential(sequential)
    try:
        out = decode_batch(datas)                      # one batch: images of different sizes and layouts side by side
        torch.cuda.synchronize()
    finally:
        libgnc.gnc_debug_jpeg_sequential(0)
    for (h, w, kw), data, t in zip(CASES, datas, out):
        ref = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
        assert t is not None, (h, w, kw)
        got = t.cpu().numpy()
        assert got.shape == ref.shape and np.array_equal(got, ref), (h, w, kw, int(np.abs(got.astype(int) - ref).max()))
        if h * w <= 6000:
            assert np.array_equal(decode_baseline(data), ref)


def test_unsupported_files_fall_back_to_the_host(libgnc, tmp_path):
    from graphnet_classifier_b200.utils import jpeg as gjpeg
    from graphnet_classifier_b200.utils.staging import DecodePool
    rng = np.random.default_rng(9)
    arr = rng.integers(0, 256, (40, 56, 3), dtype=np.uint8)
    files = {"a.jpg": dict(quality=90), "p.jpg": dict(quality=90, progressive=True), "b.png": {}, "c.jpg": dict(quality=70, subsampling=0)}
    paths = []
    for name, kw in files.items():
        Image.fromarray(arr).save(tmp_path / name, **kw)
        paths.append(str(tmp_path / name))
    cmyk = tmp_path / "k.jpg"
    Image.fromarray(arr).convert("CMYK").save(cmyk, quality=90)
    paths.append(str(cmyk))
    infos = [gjpeg.parse(open(p, "rb").read()) for p in paths]
    assert [i is not None for i in infos] == [True, False, False, True, False]
    want = np.stack([np.array(Image.open(p).convert("RGB").resize((24, 24))) for p in paths])
    with DecodePool(workers=2, device_jpeg=True) as pool:
        got = pool.stage(paths, 24)
        assert pool.stats == {"device_jpeg": 2, "host_decoded": 3}
        assert np.array_equal(got.cpu().numpy(), want)
        got2 = torch.cat(list(pool.batches(paths * 3, 24, chunk=4)))
        assert np.array_equal(got2.cpu().numpy(), np.concatenate([want] * 3))
    with DecodePool(workers=2, device_jpeg=False) as pool:      # everything on the host processes: same pixels
        assert np.array_equal(pool.stage(paths, 24).cpu().numpy(), want)
