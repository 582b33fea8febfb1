"""CPU: the oracle restatement reproduces the committed golden vectors, which were
produced by the UNMODIFIED reference (oracle/make_golden.py)."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import gnn as ognn
from oracle import graph_build as ogb
from oracle.weights import fill_deterministic


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def test_grid_edges_bit_exact(golden):
    g = golden["grid_edges"]
    n = 0
    for k, ref in g.items():
        if k.startswith("grid_"):
            hw, d = k[len("grid_"):].split("_d")
            H, W = map(int, hw.split("x"))
            out = ogb.grid_edges(H, W, bool(int(d)))
            assert out.dtype == np.int64 and np.array_equal(out, ref), k
            n += 1
        elif k.startswith("gridsha_"):
            r, d = k[len("gridsha_"):].split("_d")
            assert np.array_equal(_sha(ogb.grid_edges(int(r), int(r), bool(int(d)))), ref), k
            n += 1
    assert n >= 20


def test_pixel_patch_superpixel_builders(golden):
    b = golden["builders"]
    for k in [k for k in b if k.startswith("pixel_") and k.endswith("_img")]:
        tag = k[:-4]
        diag = tag.endswith("d1")
        x, pos, ei = ogb.pixel_graph(b[k], diag)
        assert np.array_equal(x, b[tag + "_x"]) and x.dtype == np.uint8
        assert np.array_equal(pos, b[tag + "_pos"]) and np.array_equal(ei, b[tag + "_ei"])
    for k in [k for k in b if k.startswith("patch_") and k.endswith("_img")]:
        tag = k[:-4]
        p = int(tag.split("_")[-1])
        x, pos, ei = ogb.patch_graph(b[k], p)
        assert np.array_equal(x, b[tag + "_x"]) and np.array_equal(pos, b[tag + "_pos"])
        assert np.array_equal(ei, b[tag + "_ei"])
    for k in [k for k in b if k.startswith("superpixel_") and k.endswith("_img")]:
        tag = k[:-4]
        x, pos, ei = ogb.superpixel_graph_from_labels(b[k], b[tag + "_labels"])
        assert np.array_equal(x, b[tag + "_x"]) and np.array_equal(pos, b[tag + "_pos"])
        assert np.array_equal(ei, b[tag + "_ei"]) and ei.dtype == np.int64
    for k in [k for k in b if k.startswith("jpeg_") and k.endswith("_img")]:
        x, _, _ = ogb.pixel_graph(b[k])
        assert np.array_equal(_sha(x), b[k[:-4] + "_xsha"])


def test_superpixel_no_edges_is_float_empty():
    img = np.zeros((4, 4, 3), np.uint8)
    _, _, ei = ogb.superpixel_graph_from_labels(img, np.zeros((4, 4), np.int64))
    assert ei.shape == (2, 0) and ei.dtype == np.float64     # reference superpixel.py:70-71


def test_state_dict_contract(golden):
    m = golden["model"]
    om = ognn.build_reference_config_model(8, seed=0)
    assert list(om.state_dict().keys()) == list(m["state_dict_keys"])
    assert [str(tuple(v.shape)) for v in om.state_dict().values()] == list(m["state_dict_shapes_r8"])
    om32 = ognn.build_reference_config_model(32, seed=0)
    assert list(om32.state_dict().keys()) == list(m["checkpoint_keys"])
    assert [str(tuple(v.shape)) for v in om32.state_dict().values()] == list(m["checkpoint_shapes"])
    assert len(m["checkpoint_keys"]) == 76


@pytest.mark.parametrize("r,diag", [(8, False), (8, True), (12, False)])
def test_model_logits_loss_grads(golden, r, diag):
    m = golden["model"]
    tag = f"model_r{r}_d{int(diag)}"
    om = ognn.build_reference_config_model(r, seed=None)
    fill_deterministic(om, seed=7)
    imgs = m[tag + "_imgs"]
    for b in range(imgs.shape[0]):
        inp = ogb.to_model_inputs(*ogb.pixel_graph(imgs[b], diag))
        with torch.no_grad():
            out = om(inp)
        np.testing.assert_allclose(out.numpy(), m[tag + "_logits"][b], rtol=2e-6, atol=1e-7)
    inp = ogb.to_model_inputs(*ogb.pixel_graph(imgs[0], diag))
    loss = torch.nn.functional.cross_entropy(om(inp), torch.tensor(int(m[tag + "_label"])))
    assert abs(loss.item() - float(m[tag + "_loss"])) < 2e-6
    loss.backward()
    for name, p in om.named_parameters():
        samp = m[f"{tag}_grad_{name}_samp"]
        f = p.grad.flatten()
        stride = max(1, -(-f.numel() // 512))
        got = f[::stride].numpy()
        scale = float(m[f"{tag}_grad_{name}_sn"][1]) / np.sqrt(f.numel()) + 1e-30
        assert np.max(np.abs(got - samp)) <= 1e-5 * max(scale, np.max(np.abs(samp))), name


def test_scatter_sum_golden(golden):
    m = golden["model"]
    out = ognn.scatter_sum(torch.from_numpy(m["scatter_src"]), torch.from_numpy(m["scatter_idx"]))
    assert np.array_equal(out.numpy(), m["scatter_out"])
    with pytest.raises(NotImplementedError):
        ognn.scatter_sum(torch.zeros(3, 2), torch.zeros(3, dtype=torch.long), dim=1)


def test_batched_equals_per_sample(golden):
    m = golden["model"]
    om = ognn.build_reference_config_model(8, seed=None)
    fill_deterministic(om, seed=7)
    imgs = m["model_r8_d0_imgs"]
    gs = [ogb.pixel_graph(im) for im in imgs]
    bx, bp, be = ogb.batch_graphs([g[0] for g in gs], [g[1] for g in gs], [g[2] for g in gs])
    with torch.no_grad():
        out = om(ogb.to_model_inputs(bx, bp, be))
    np.testing.assert_allclose(out.numpy(), m["model_r8_d0_logits"], rtol=2e-6, atol=1e-7)


def _resize_cases(g):
    for k in g:
        if k.endswith("_x"):
            name, to = k[:-2].rsplit("_to", 1)
            yield k, g[name + "_src"], int(to), g[k]


def test_resize_oracle_vs_reference_golden_and_pillow(golden):
    """oracle/resize.py == the reference builder's resized pixels (golden, produced through
    image_to_graph_pixel_optimized) == Pillow itself on further shapes."""
    from PIL import Image
    from oracle.resize import precompute_coeffs, resize_bicubic
    n = 0
    for k, src, r, ref in _resize_cases(golden["resize"]):
        assert np.array_equal(resize_bicubic(src, r, r), ref), k
        n += 1
    assert n >= 10
    rng = np.random.default_rng(5)
    for H, W, oh, ow in [(375, 500, 128, 128), (64, 64, 128, 128), (128, 300, 128, 128), (300, 128, 128, 128),
                         (128, 128, 128, 128), (1, 1, 4, 4), (2, 3, 8, 8), (640, 480, 64, 32), (50, 40, 7, 9)]:
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        assert np.array_equal(resize_bicubic(img, oh, ow), np.asarray(Image.fromarray(img).resize((ow, oh)))), (H, W, oh, ow)
    b, kk = precompute_coeffs(500, 128)
    assert kk.shape == (128, 17) and int(kk.sum(1).min()) > (1 << 22) - 20 and int(kk.sum(1).max()) < (1 << 22) + 20


def test_resize_coefficient_tables_through_the_c_abi(libgnc):
    """Host-only entry point: Pillow's per-axis tables from libgnc == the oracle's, bit for bit."""
    from oracle.resize import precompute_coeffs
    for i, o in [(500, 128), (375, 128), (64, 128), (37, 64), (1000, 256), (5, 64), (129, 128), (127, 128), (3, 8), (1, 4),
                 (640, 32), (4000, 128), (128, 256), (333, 77)]:
        ks = libgnc.gnc_resize_bicubic_ksize(i, o)
        ob, ok = precompute_coeffs(i, o)
        assert ks == ok.shape[1]
        b = np.zeros((o, 2), np.int32)
        k = np.full((o, ks), -7, np.int32)
        assert libgnc.gnc_resize_bicubic_coeffs(i, o, b.ctypes.data, k.ctypes.data) == 0
        assert np.array_equal(b, ob) and np.array_equal(k, ok), (i, o)
    assert libgnc.gnc_resize_bicubic_ksize(0, 5) == 0
    assert libgnc.gnc_resize_bicubic_coeffs(5, 5, None, None) == 1


def test_oracle_reproduces_reference_on_shipped_checkpoint():
    """tests/golden/checkpoint.npz holds the shipped checkpoint's tensors and what the UNMODIFIED reference computes
    with them on two shipped JPEGs (oracle/make_golden.py:gen_checkpoint); the oracle must reproduce both, and the
    large-activation case (hidden activations beyond the fp16 two-piece domain of the chained kernels)."""
    import os
    import numpy as np
    import torch
    from oracle import gnn as ognn
    from oracle.weights import fill_deterministic
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "checkpoint.npz"))
    om = ognn.build_reference_config_model(32, seed=0)
    om.load_state_dict({str(k): torch.from_numpy(g[f"w{i:02d}"]) for i, k in enumerate(g["keys"])}, strict=True)
    om.eval()
    from oracle import graph_build as ogb
    for tag in ("chihuahua", "muffin"):
        x, pos, ei = ogb.pixel_graph(g[f"{tag}_pixels"])
        with torch.no_grad():
            logits = om(ogb.to_model_inputs(x, pos, ei))
        np.testing.assert_allclose(logits.numpy(), g[f"{tag}_logits"], rtol=2e-6, atol=1e-7)
    cfg = dict(num_local_features=3, space_dim=2, out_channels=1, n_blocks=3)
    om = ognn.OracleCombinedModel(ognn.OracleGraphNet(**cfg), num_nodes=256, classes=2)
    fill_deterministic(om, seed=int(g["big_seed"]))
    with torch.no_grad():
        om.graph_net.node_encoder.model[0].weight.mul_(float(g["big_scale"]))
        for b in range(2):
            out = om(ogb.to_model_inputs(*ogb.pixel_graph(g["big_imgs"][b])))
            np.testing.assert_allclose(out.numpy(), g["big_logits"][b], rtol=2e-6, atol=1e-7)


# ---- JPEG decode: the oracle restatement of libjpeg against Pillow and the reference's shipped files (CPU) ----
def test_jpeg_oracle_equals_pillow_and_shipped_files():
    import io
    from PIL import Image
    from oracle.jpeg import decode_baseline
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jpeg_files.npz"))
    for k in [k for k in g.files if k.endswith("_bytes")]:
        data = g[k].tobytes()
        assert np.array_equal(decode_baseline(data), g[k[:-6] + "_rgb"]), k
        assert np.array_equal(np.asarray(Image.open(io.BytesIO(data)).convert("RGB")), g[k[:-6] + "_rgb"]), k
    rng = np.random.default_rng(3)
    for h, w, kw in ((33, 47, dict(quality=75)), (17, 23, dict(quality=95, subsampling=0)), (40, 50, dict(quality=60, subsampling=1)),
                     (31, 65, dict(quality=90)), (50, 70, dict(quality=80, restart_marker_blocks=3)), (20, 3, dict(quality=90)),
                     (64, 48, dict(quality=85, optimize=True))):
        arr = np.asarray(Image.fromarray(rng.integers(0, 256, (h // 5 + 2, w // 5 + 2, 3), dtype=np.uint8)).resize((w, h)))
        buf = io.BytesIO()
        Image.fromarray(arr).save(buf, format="JPEG", **kw)
        ref = np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGB"))
        assert np.array_equal(decode_baseline(buf.getvalue()), ref), (h, w, kw)


def test_jpeg_host_parser_matches_oracle_parser(libgnc):
    """gnc_jpeg_parse (host code of csrc/jpeg.cu, no GPU needed): geometry, tables and scan position against the
    oracle's own marker parser; files outside the decoder's scope are reported as unsupported."""
    import ctypes
    import io
    from PIL import Image
    from graphnet_classifier_b200 import _lib
    from oracle.jpeg import parse
    rng = np.random.default_rng(5)
    arr = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    for kw in (dict(quality=80), dict(quality=92, subsampling=0), dict(quality=50, subsampling=1, optimize=True),
               dict(quality=85, restart_marker_blocks=2)):
        buf = io.BytesIO()
        Image.fromarray(arr).save(buf, format="JPEG", **kw)
        data = buf.getvalue()
        info = _lib.GncJpegImage()
        assert libgnc.gnc_jpeg_parse(data, len(data), ctypes.byref(info)) == 0
        (H, W, comps), quant, huff, dri, tabs, pos = parse(data)
        assert (info.height, info.width, info.ncomp) == (H, W, len(comps)) and info.restart_interval == dri
        assert [info.hsamp[c] for c in range(3)] == [c[1] for c in comps] and [info.vsamp[c] for c in range(3)] == [c[2] for c in comps]
        assert info.scan_offset == pos and data[pos + info.scan_bytes:pos + info.scan_bytes + 2] == b"\xff\xd9"
        for c in range(3):
            assert np.array_equal(np.array(info.quant[info.qtab[c]][:]), quant[comps[c][3]])
            # every code of the oracle's canonical table decodes to the same symbol through the lookahead / maxcode tables
            for (tc, th), slot in (((0, tabs[c][1]), info.dc_tab[c]), ((1, tabs[c][2]), 4 + info.ac_tab[c])):
                t = info.huff[slot]
                for (length, code), sym in huff[(tc, th)].items():
                    if length <= 9:
                        e = t.look[code << (9 - length)]
                        assert (e >> 8, e & 255) == (length, sym)
                    else:
                        assert code <= t.maxcode[length] and t.vals[code + t.valoff[length]] == sym
        assert info.n_blocks == info.mcu_x * info.mcu_y * sum(info.hsamp[c] * info.vsamp[c] for c in range(3))
    for bad in (dict(progressive=True), None):
        buf = io.BytesIO()
        if bad is None:
            Image.fromarray(arr).save(buf, format="PNG")
        else:
            Image.fromarray(arr).save(buf, format="JPEG", **bad)
        info = _lib.GncJpegImage()
        assert libgnc.gnc_jpeg_parse(buf.getvalue(), len(buf.getvalue()), ctypes.byref(info)) == _lib.GNC_JPEG_UNSUPPORTED


def test_jpeg_batch_pack_host(libgnc):
    """gnc_jpeg_pack (host): one call parses a batch on several threads, drops the files the device decoder does not cover
    and lays descriptors and bytes out with all offsets set - checked against per-file gnc_jpeg_parse."""
    import ctypes
    import io
    from PIL import Image
    from graphnet_classifier_b200 import _lib
    rng = np.random.default_rng(8)
    datas = []
    for h, w, fmt in [(33, 47, "JPEG"), (20, 20, "PNG"), (64, 48, "JPEG"), (100, 90, "JPEG")] * 4:
        buf = io.BytesIO()
        Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)).save(buf, format=fmt)
        datas.append(buf.getvalue())
    datas.append(b"")
    n = len(datas)
    stream = np.zeros(sum(map(len, datas)), np.uint8)
    infos, index, totals = (_lib.GncJpegImage * n)(), (ctypes.c_int32 * n)(), (ctypes.c_int64 * 5)()
    ptrs, sizes = (ctypes.c_char_p * n)(*datas), (ctypes.c_int64 * n)(*[len(d) for d in datas])
    cast = lambda a: ctypes.cast(a, ctypes.c_void_p)
    assert libgnc.gnc_jpeg_pack(cast(ptrs), cast(sizes), n, 4, stream.ctypes.data, stream.size, cast(infos), cast(index), cast(totals)) == 0
    m = int(totals[0])
    assert [index[j] for j in range(m)] == [i for i in range(n - 1) if i % 4 != 1]
    off = blocks = plane = pixels = 0
    for j in range(m):
        d = datas[index[j]]
        ref = _lib.GncJpegImage()
        assert libgnc.gnc_jpeg_parse(d, len(d), ctypes.byref(ref)) == 0
        assert bytes(stream[off:off + len(d)]) == d and infos[j].scan_offset == ref.scan_offset + off
        assert (infos[j].block_offset, infos[j].coef_offset, infos[j].plane_offset, infos[j].pixel_offset) == (blocks, 64 * blocks, plane, pixels)
        off, blocks, plane, pixels = off + len(d), blocks + ref.n_blocks, plane + ref.plane_bytes, pixels + ref.width * ref.height
    assert list(totals) == [m, off, blocks, plane, pixels]
