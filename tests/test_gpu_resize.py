"""GPU: the Pillow-exact device resize (csrc/resize.cu) through the C ABI vs the reference builder's golden
pixels, the oracle and Pillow itself - bit-exact - and the entry points that use it."""
import numpy as np
import pytest
import torch

from oracle.resize import resize_bicubic as oracle_resize

pytestmark = pytest.mark.gpu


def _cases(g):
    for k in g:
        if k.endswith("_x"):
            name, to = k[:-2].rsplit("_to", 1)
            yield k, g[name + "_src"], int(to), g[k]


def test_resize_vs_reference_golden(golden, libgnc):
    from graphnet_classifier_b200 import ops
    n = 0
    for k, src, r, ref in _cases(golden["resize"]):
        got = ops.resize_bicubic(torch.from_numpy(src).cuda(), r, r)
        assert got.dtype == torch.uint8 and tuple(got.shape) == (r, r, 3)
        assert np.array_equal(got.cpu().numpy(), ref), k
        n += 1
    assert n >= 10


@pytest.mark.parametrize("H,W,oh,ow", [(375, 500, 128, 128), (64, 64, 128, 128), (100, 37, 64, 64), (128, 300, 128, 128),
                                       (300, 128, 128, 128), (17, 23, 32, 32), (5, 5, 64, 64), (129, 127, 128, 128),
                                       (128, 128, 128, 128), (2, 3, 8, 8), (1, 1, 4, 4), (640, 480, 64, 32), (50, 41, 7, 9),
                                       (33, 35, 31, 29)])
def test_resize_vs_oracle_and_pillow(libgnc, H, W, oh, ow):
    from PIL import Image
    from graphnet_classifier_b200 import ops
    rng = np.random.default_rng(H * 1000 + W)
    B = 3
    imgs = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    imgs[1] = (np.add.outer(np.arange(H), np.arange(W))[..., None] * np.array([1, 2, 3]) % 256).astype(np.uint8)
    imgs[2, : H // 2] = 255                                      # saturated regions: overshoot is clamped
    imgs[2, H // 2:] = 0
    got = ops.resize_bicubic(torch.from_numpy(imgs).cuda(), oh, ow).cpu().numpy()
    for b in range(B):
        assert np.array_equal(got[b], oracle_resize(imgs[b], oh, ow)), b
        assert np.array_equal(got[b], np.asarray(Image.fromarray(imgs[b]).resize((ow, oh)))), b


def test_resize_strided_views_and_odd_alignment(libgnc):
    """Row pitch / image stride / base address that are not multiples of 4 or 16 (views into a larger buffer)."""
    from graphnet_classifier_b200 import ops
    rng = np.random.default_rng(3)
    big = rng.integers(0, 256, (2, 70, 90, 3), dtype=np.uint8)
    dev = torch.from_numpy(big).cuda()
    view = dev[:, 3:64, 5:82]                                    # pitch 270 bytes, base offset 3*270 + 15
    got = ops.resize_bicubic(view, 32, 48).cpu().numpy()
    for b in range(2):
        assert np.array_equal(got[b], oracle_resize(np.ascontiguousarray(big[b, 3:64, 5:82]), 32, 48))
    flat = torch.zeros(1 + 61 * 77 * 3, dtype=torch.uint8, device="cuda")
    flat[1:] = dev[0, 3:64, 5:82].reshape(-1)                    # base address 1 mod 16
    one = flat[1:].view(61, 77, 3)
    assert np.array_equal(ops.resize_bicubic(one, 20, 77).cpu().numpy(), oracle_resize(big[0, 3:64, 5:82], 20, 77))
    assert np.array_equal(ops.resize_bicubic(one, 61, 19).cpu().numpy(), oracle_resize(big[0, 3:64, 5:82], 61, 19))


def test_resize_large_photo_properties(libgnc):
    """12-megapixel input: constant images stay constant, a batch equals its images one by one, and a sampled
    band equals the oracle."""
    from graphnet_classifier_b200 import ops
    H, W, r = 3000, 4000, 128
    const = torch.full((H, W, 3), 201, dtype=torch.uint8, device="cuda")
    assert bool((ops.resize_bicubic(const, r, r) == 201).all())
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    got = ops.resize_bicubic(torch.from_numpy(img).cuda(), r, r).cpu().numpy()
    assert np.array_equal(got, oracle_resize(img, r, r))
    both = ops.resize_bicubic(torch.stack([torch.from_numpy(img).cuda(), const]), r, r)
    assert np.array_equal(both[0].cpu().numpy(), got) and bool((both[1] == 201).all())


def test_entry_points_resize_like_the_reference(golden, libgnc):
    """image_to_graph_pixel_optimized / OptimizedDatasetLoader / GraphClassifierPipeline fed with images of another
    size return what the reference returns after its PIL resize."""
    from PIL import Image
    from graphnet_classifier_b200.models.GNN import CombinedModel, GraphNet
    from graphnet_classifier_b200.pipeline import GraphClassifierPipeline
    from graphnet_classifier_b200.utils.dataloader import OptimizedDatasetLoader
    from graphnet_classifier_b200.utils.image_to_graph.image_to_graph_optimized import image_to_graph_pixel_optimized
    g = golden["resize"]
    src, ref = g["jpeg_img_4_799_256_src"], g["jpeg_img_4_799_256_to100_x"]
    x, pos, ei = image_to_graph_pixel_optimized(Image.fromarray(src), 100)
    assert x.dtype == np.uint8 and np.array_equal(x, ref.reshape(-1, 3))
    ds = OptimizedDatasetLoader(dataset=[(Image.fromarray(src), 1)], resize_value=100)
    (dx, dpos, dei), label = ds[0]
    assert np.array_equal(dx.cpu().numpy(), ref.reshape(-1, 3).astype(np.float32)) and int(label) == 1
    torch.manual_seed(0)
    r = 16
    model = CombinedModel(GraphNet(num_local_features=3, space_dim=2, out_channels=1, n_blocks=1), num_nodes=r * r).cuda().eval()
    pipe = GraphClassifierPipeline(model, resize_value=r)
    rng = np.random.default_rng(2)
    raw = [rng.integers(0, 256, (40, 52, 3), dtype=np.uint8), rng.integers(0, 256, (16, 16, 3), dtype=np.uint8),
           rng.integers(0, 256, (23, 19, 3), dtype=np.uint8)]
    pre = np.stack([np.asarray(Image.fromarray(im).resize((r, r))) for im in raw])
    assert torch.equal(pipe.infer([torch.from_numpy(im) for im in raw]), pipe.infer(torch.from_numpy(pre)))
