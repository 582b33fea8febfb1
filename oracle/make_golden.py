"""Generate tests/golden/*.npz from the UNMODIFIED reference, and pin the oracle.

TEST INFRASTRUCTURE - see oracle/__init__.py.  Run in the build container (where
/root/reference is mounted):

    python -m oracle.make_golden

For every case the reference's own function is executed (behind the stand-ins in
oracle/reference_loader.py), the oracle restatement is executed on the same
inputs, the two are compared (bit-exact for integer / builder outputs, <= 2e-6
relative for the fp32 model), and only then is the REFERENCE output written as
the golden vector.  The script exits non-zero if any comparison fails.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch
from PIL import Image

from . import gnn as ognn
from . import graph_build as ogb
from . import reference_loader as rl
from .weights import fill_deterministic, fill_parameters, synthetic_images, voronoi_labels

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sha(a: np.ndarray) -> str:
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def _check(cond, what):
    if not cond:
        print("MISMATCH:", what)
        sys.exit(1)


def grad_digest(t: torch.Tensor, max_samples: int = 512):
    f = t.detach().double().flatten()
    stride = max(1, -(-f.numel() // max_samples))
    return np.array([f.sum().item(), f.norm().item()]), f[::stride].float().numpy(), stride


def gen_grids(ref):
    out = {}
    shapes = [(1, 1), (1, 4), (5, 1), (2, 2), (3, 5), (4, 4), (8, 8), (7, 16), (32, 32)]
    for (H, W) in shapes:
        for diag in (False, True):
            r = ref.optimized.create_grid_edges_optimized(H, W, diag)
            o = ogb.grid_edges(H, W, diag)
            _check(r.dtype == np.int64 and o.dtype == np.int64, f"grid dtype {H}x{W}")
            _check(np.array_equal(r, o), f"grid edges {H}x{W} diag={diag}")
            out[f"grid_{H}x{W}_d{int(diag)}"] = np.ascontiguousarray(r)
    # large grids (BASELINE resize shapes): checksum only
    for r_ in (64, 128, 256):
        for diag in (False, True):
            r = np.ascontiguousarray(ref.optimized.get_cached_edge_index(r_, diag))
            o = ogb.grid_edges(r_, r_, diag)
            _check(np.array_equal(r, o), f"grid edges {r_} diag={diag}")
            out[f"gridsha_{r_}_d{int(diag)}"] = np.frombuffer(bytes.fromhex(sha(r)), dtype=np.uint8)
    np.savez_compressed(os.path.join(GOLDEN, "grid_edges.npz"), **out)
    print("grid_edges.npz:", len(out), "entries")


def gen_builders(ref):
    out = {}
    # synthetic images through the reference's pixel / patch builders
    for r_, diag in ((8, False), (8, True), (32, False)):
        img = synthetic_images(1, r_, seed=100 + r_)[0]
        x, pos, ei = ref.optimized.image_to_graph_pixel_optimized(Image.fromarray(img), r_, diag, use_cache=False)
        ox, opos, oei = ogb.pixel_graph(img, diag)
        _check(x.dtype == np.uint8 and np.array_equal(x, ox), f"pixel x r={r_}")
        _check(np.array_equal(pos, opos) and pos.dtype == opos.dtype, f"pixel pos r={r_}")
        _check(np.array_equal(ei, oei), f"pixel ei r={r_}")
        k = f"pixel_{r_}_d{int(diag)}"
        out[k + "_img"], out[k + "_x"], out[k + "_pos"], out[k + "_ei"] = img, x, pos, np.ascontiguousarray(ei)
    for r_, p in ((32, 8), (20, 8), (16, 4), (8, 8)):
        img = synthetic_images(1, r_, seed=200 + r_)[0]
        x, pos, ei = ref.patch.image_to_graph_patch(Image.fromarray(img), r_, p)
        ox, opos, oei = ogb.patch_graph(img, p)
        x, pos = np.asarray(x, dtype=np.float64).reshape(-1, 3), np.asarray(pos, dtype=np.int64).reshape(-1, 2)
        _check(np.array_equal(x, ox), f"patch x r={r_} p={p}")
        _check(np.array_equal(pos, opos), f"patch pos r={r_} p={p}")
        _check(np.array_equal(ei, oei), f"patch ei r={r_} p={p}")
        k = f"patch_{r_}_{p}"
        out[k + "_img"], out[k + "_x"], out[k + "_pos"], out[k + "_ei"] = img, x, pos, np.ascontiguousarray(ei)
    # label map -> superpixel graph through the reference's own function body
    for r_, S, seed in ((16, 5, 1), (32, 12, 2), (48, 30, 3)):
        img = synthetic_images(1, r_, seed=300 + r_)[0]
        labels = voronoi_labels(r_, r_, S, seed)
        if seed == 2:                      # make the label set non-contiguous: rank renumbering
            labels = labels * 3 + 7
        rl.set_next_slic_labels(labels)
        x, pos, ei = ref.superpixel.image_to_graph_superpixel(Image.fromarray(img), r_, n_segments=S)
        ox, opos, oei = ogb.superpixel_graph_from_labels(img, labels)
        x, pos = np.asarray(x, dtype=np.float64), np.asarray(pos, dtype=np.float64)
        _check(np.array_equal(x, ox), f"superpixel x r={r_}")
        _check(np.array_equal(pos, opos), f"superpixel pos r={r_}")
        _check(np.array_equal(ei, oei) and ei.dtype == oei.dtype, f"superpixel ei r={r_}")
        k = f"superpixel_{r_}"
        out[k + "_img"], out[k + "_labels"] = img, labels
        out[k + "_x"], out[k + "_pos"], out[k + "_ei"] = x, pos, np.ascontiguousarray(ei)
    # a one-segment map: no edges -> float64 [2, 0] (image_to_graph_superpixel.py:70-71)
    img = synthetic_images(1, 8, seed=399)[0]
    labels = np.zeros((8, 8), dtype=np.int64)
    rl.set_next_slic_labels(labels)
    x, pos, ei = ref.superpixel.image_to_graph_superpixel(Image.fromarray(img), 8, n_segments=1)
    ox, opos, oei = ogb.superpixel_graph_from_labels(img, labels)
    _check(ei.shape == (2, 0) and oei.shape == (2, 0) and ei.dtype == oei.dtype, "superpixel empty edges")
    _check(np.array_equal(np.asarray(x), ox) and np.array_equal(np.asarray(pos), opos), "superpixel single")
    # the shipped JPEGs (real inputs) at 32 and 64: decoded pixels + reference outputs
    for rel, r_ in (("static/chihuahua/img_4_799_32.jpg", 32), ("static/muffin/img_4_880_32.jpg", 32),
                    ("static/chihuahua/img_4_799_64.jpg", 64)):
        path = os.path.join(rl.REFERENCE_ROOT, rel)
        x, pos, ei = ref.optimized.image_to_graph_pixel_optimized(path, r_)
        img = np.array(Image.open(path).convert("RGB").resize((r_, r_)))
        ox, opos, oei = ogb.pixel_graph(img)
        _check(np.array_equal(x, ox) and np.array_equal(pos, opos) and np.array_equal(ei, oei), f"jpeg {rel}")
        k = "jpeg_" + os.path.basename(rel).replace(".jpg", "")
        out[k + "_img"] = img
        out[k + "_xsha"] = np.frombuffer(bytes.fromhex(sha(x)), dtype=np.uint8)
    np.savez_compressed(os.path.join(GOLDEN, "builders.npz"), **out)
    print("builders.npz:", len(out), "entries")


def _ref_model(ref, r_, n_blocks=3, classes=2, seed=None, num_nodes=None):
    if seed is not None:
        torch.manual_seed(seed)
    gn = ref.GNN.GraphNet(num_local_features=3, space_dim=2, out_channels=1, n_blocks=n_blocks)
    return ref.GNN.CombinedModel(graph_net=gn, num_nodes=num_nodes or r_ * r_, classes=classes)


def gen_model(ref):
    out = {}
    # 1) construction parity: same seed -> same tensors, same key order
    rm = _ref_model(ref, 8, seed=0)
    om = ognn.build_reference_config_model(8, seed=0)
    rs, os_ = rm.state_dict(), om.state_dict()
    _check(list(rs.keys()) == list(os_.keys()), "state_dict key order")
    for k in rs:
        _check(torch.equal(rs[k], os_[k]), f"seeded init {k}")
    out["state_dict_keys"] = np.array(list(rs.keys()))
    out["state_dict_shapes_r8"] = np.array([str(tuple(v.shape)) for v in rs.values()])
    # 2) the shipped checkpoint loads into the oracle tree
    ck = torch.load(os.path.join(rl.REFERENCE_ROOT, "weights/GNN/dim32_3block/best_model_epoch5.pth"),
                    map_location="cpu")
    om32 = ognn.build_reference_config_model(32, seed=0)
    missing = om32.load_state_dict(ck, strict=True)
    _check(len(missing.missing_keys) == 0 and len(missing.unexpected_keys) == 0, "checkpoint keys")
    out["checkpoint_keys"] = np.array(list(ck.keys()))
    out["checkpoint_shapes"] = np.array([str(tuple(v.shape)) for v in ck.values()])

    # 3) forward / loss / gradients on deterministic weights
    crit = torch.nn.CrossEntropyLoss()
    for r_, diag, label in ((8, False, 1), (8, True, 0), (12, False, 1)):
        rm = _ref_model(ref, r_)
        fill_deterministic(rm, seed=7)
        om = ognn.build_reference_config_model(r_, seed=None)
        om.load_state_dict(rm.state_dict())
        imgs = synthetic_images(3, r_, seed=r_)
        tag = f"model_r{r_}_d{int(diag)}"
        logits_all = []
        for b in range(3):
            x, pos, ei = ref.optimized.image_to_graph_pixel_optimized(Image.fromarray(imgs[b]), r_, diag, False)
            tx, tpos, tei = ogb.to_model_inputs(x, pos, ei)
            with torch.no_grad():
                lr = rm((tx, tpos, tei))
                lo = om((tx, tpos, tei))
            _check(torch.allclose(lr, lo, rtol=2e-6, atol=1e-7), f"{tag} logits image {b}: {lr} vs {lo}")
            logits_all.append(lr.numpy())
        out[tag + "_imgs"] = imgs
        out[tag + "_logits"] = np.stack(logits_all)
        # gradients on image 0
        x, pos, ei = ref.optimized.image_to_graph_pixel_optimized(Image.fromarray(imgs[0]), r_, diag, False)
        tx, tpos, tei = ogb.to_model_inputs(x, pos, ei)
        lab = torch.tensor(label, dtype=torch.long)
        for m in (rm, om):
            m.zero_grad()
            crit(m((tx, tpos, tei)), lab).backward()
        loss_r = crit(rm((tx, tpos, tei)), lab).item()
        out[tag + "_label"] = np.array(label)
        out[tag + "_loss"] = np.array(loss_r, dtype=np.float64)
        for (k, pr), (_, po) in zip(rm.named_parameters(), om.named_parameters()):
            rel = (pr.grad - po.grad).norm() / (pr.grad.norm() + 1e-30)
            _check(rel < 5e-6, f"{tag} grad {k} rel {rel}")
            sn, samp, stride = grad_digest(pr.grad)
            out[f"{tag}_grad_{k}_sn"] = sn
            out[f"{tag}_grad_{k}_samp"] = samp
        # batched == per-sample (block-diagonal batching is our extension; pin it
        # against the reference's own GraphNet so the extension is semantics-preserving)
        xs, ps, es = [], [], []
        for b in range(3):
            x, pos, ei = ref.optimized.image_to_graph_pixel_optimized(Image.fromarray(imgs[b]), r_, diag, False)
            xs.append(x), ps.append(pos), es.append(ei)
        bx, bpos, bei = ogb.batch_graphs(xs, ps, es)
        tx, tpos, tei = ogb.to_model_inputs(bx, bpos, bei)
        with torch.no_grad():
            y = rm.graph_net(tx, tpos, tei).reshape(3, -1)
            lb = rm.classifier(y)
            lo = om((tx, tpos, tei))
        _check(torch.allclose(lb, torch.from_numpy(np.stack(logits_all)), rtol=2e-6, atol=1e-7), f"{tag} batched ref")
        _check(torch.allclose(lb, lo, rtol=2e-6, atol=1e-7), f"{tag} batched oracle")
    # 4) survey sanity values: seed-0 default init, r=32
    rm = _ref_model(ref, 32, seed=0)
    imgs = np.stack([np.random.default_rng(0).integers(0, 256, (32, 32, 3), dtype=np.uint8)])
    x, pos, ei = ref.optimized.image_to_graph_pixel_optimized(Image.fromarray(imgs[0]), 32)
    with torch.no_grad():
        l0 = rm(ogb.to_model_inputs(x, pos, ei))
    out["seed0_r32_logits_img0"] = l0.numpy()
    print("seed-0 r=32 logits image 0:", l0.numpy(), "(survey probe: [-0.0638899, -0.1402419])")
    # 5) scatter_sum fallback semantics
    g = torch.Generator().manual_seed(3)
    src = torch.randn(40, 6, generator=g)
    idx = torch.randint(0, 9, (40,), generator=g)
    r = ref.GNN.scatter_sum(src, idx, dim=0)
    o = ognn.scatter_sum(src, idx, dim=0)
    _check(torch.equal(r, o), "scatter_sum")
    out["scatter_src"], out["scatter_idx"], out["scatter_out"] = src.numpy(), idx.numpy(), r.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "model.npz"), **out)
    print("model.npz:", len(out), "entries")


def gen_resize(ref):
    """Input staging: ``image.convert('RGB').resize((r, r))`` inside the reference's own builder
    (image_to_graph_optimized.py:65-70; Pillow BICUBIC) on the shipped JPEGs at sizes other than the file's and on
    random non-square images.  The builder's ``x`` IS the resized image; oracle/resize.py must reproduce it bit for bit."""
    from oracle.resize import resize_bicubic
    out = {}
    cases = [("static/chihuahua/img_4_799_256.jpg", 128), ("static/chihuahua/img_4_799_256.jpg", 100),
             ("static/muffin/img_0_187_128.jpg", 256), ("static/muffin/img_0_187_256.jpg", 64),
             ("static/muffin/img_4_880_32.jpg", 48)]
    srcs = {}
    for rel, r_ in cases:
        path = os.path.join(rl.REFERENCE_ROOT, rel)
        x, _, _ = ref.optimized.image_to_graph_pixel_optimized(path, r_)
        src = np.array(Image.open(path).convert("RGB"))
        _check(x.dtype == np.uint8 and np.array_equal(resize_bicubic(src, r_, r_).reshape(-1, 3), x), f"resize {rel} -> {r_}")
        name = os.path.basename(rel).replace(".jpg", "")
        srcs[name] = src
        out[f"jpeg_{name}_to{r_}_x"] = x.reshape(r_, r_, 3)
    for name, src in srcs.items():
        out[f"jpeg_{name}_src"] = src
    rng = np.random.default_rng(77)
    for i, (H, W, r_) in enumerate([(75, 100, 32), (33, 17, 24), (64, 200, 64), (200, 64, 64), (9, 9, 40), (301, 203, 50)]):
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        if i % 2:                       # smooth content with saturated regions: exercises the clamp
            yy, xx = np.mgrid[0:H, 0:W]
            img = np.stack([(yy * 255 // max(H - 1, 1)), (xx * 255 // max(W - 1, 1)), ((yy // 3 + xx // 3) % 2) * 255], -1).astype(np.uint8)
        x, _, _ = ref.optimized.image_to_graph_pixel_optimized(Image.fromarray(img), r_)
        _check(np.array_equal(resize_bicubic(img, r_, r_).reshape(-1, 3), x), f"resize random {H}x{W} -> {r_}")
        out[f"rand{i}_src"], out[f"rand{i}_to{r_}_x"] = img, x.reshape(r_, r_, 3)
    np.savez_compressed(os.path.join(GOLDEN, "resize.npz"), **out)
    print("resize.npz:", len(out), "entries")


def gen_mlp(ref):
    """The MLP baseline (reference models/MLP.py:5-47 as main.py:21-29 / utils/inference.py:16-29 use it): logits,
    cross-entropy loss and every gradient of the UNMODIFIED reference class on a ToTensor-style batch and on a raw
    0..255 flattened image, deterministic weights."""
    out = {}
    for tag, in_shape, kw in (("b4", (4, 3, 6, 6), dict(hidden_layers=2)), ("raw", (1, 3 * 8 * 8), dict(hidden_layers=3)),
                              ("bn", (5, 3, 4, 4), dict(hidden_layers=1, norm_type="BatchNorm1d", activation="Tanh"))):
        in_dim = int(np.prod(in_shape[1:]))
        m = ref.MLP.MLP(in_dim=in_dim, out_dim=2, **kw)
        fill_parameters(m, seed=31)
        rng = np.random.default_rng(in_dim)
        x = rng.random(in_shape, dtype=np.float32) if tag != "raw" else rng.integers(0, 256, in_shape).astype(np.float32)
        labels = rng.integers(0, 2, in_shape[0])
        logits = m(torch.from_numpy(x))
        loss = torch.nn.functional.cross_entropy(logits, torch.from_numpy(labels))
        loss.backward()
        out[f"{tag}_x"], out[f"{tag}_labels"] = x, labels
        out[f"{tag}_logits"], out[f"{tag}_loss"] = logits.detach().numpy(), np.float64(loss.item())
        out[f"{tag}_keys"] = np.array(list(m.state_dict().keys()))
        for name, p in m.named_parameters():
            out[f"{tag}_grad_{name}"] = p.grad.numpy().copy()
    np.savez_compressed(os.path.join(GOLDEN, "mlp.npz"), **out)
    print("mlp.npz:", len(out), "entries")


def gen_checkpoint(ref):
    """The shipped checkpoint (weights/GNN/dim32_3block/best_model_epoch5.pth, resize 32) through the UNMODIFIED reference
    on two shipped JPEGs: logits, the GraphNet node outputs and the largest hidden activation (386-427 with raw 0..255
    pixels: the fp16 two-piece kernels' domain is 4094).  The 76 tensors travel with the fixture so that the GPU tests can
    load them without /root/reference.  Second case: deterministic weights whose first node-encoder layer is scaled by
    60 - hidden activations beyond that domain, where the product path must fall back to its 3xTF32 engine."""
    out = {}
    ck = torch.load(os.path.join(rl.REFERENCE_ROOT, "weights/GNN/dim32_3block/best_model_epoch5.pth"), map_location="cpu")
    rm = _ref_model(ref, 32)
    rm.load_state_dict(ck)
    rm.eval()
    out["keys"] = np.array(list(ck.keys()))
    for i, (k, v) in enumerate(ck.items()):
        out[f"w{i:02d}"] = v.numpy()
    acts = []
    hook = rm.graph_net.node_encoder.model[1].register_forward_hook(lambda m, a, o: acts.append(float(o.abs().max())))
    for tag, rel in (("chihuahua", "static/chihuahua/img_4_799_32.jpg"), ("muffin", "static/muffin/img_4_880_32.jpg")):
        path = os.path.join(rl.REFERENCE_ROOT, rel)
        x, pos, ei = ref.optimized.image_to_graph_pixel_optimized(path, 32)
        tx, tpos, tei = ogb.to_model_inputs(x, pos, ei)
        with torch.no_grad():
            out[f"{tag}_logits"] = rm((tx, tpos, tei)).numpy()
            out[f"{tag}_nodes"] = rm.graph_net(tx, tpos, tei).numpy()
        out[f"{tag}_pixels"] = x.reshape(32, 32, 3)
    hook.remove()
    out["max_hidden_activation"] = np.array(max(acts))
    print("shipped checkpoint: largest node-encoder hidden activation", max(acts))
    # activations outside the fp16 two-piece domain
    r_ = 16
    rm = _ref_model(ref, r_)
    fill_deterministic(rm, seed=13)
    with torch.no_grad():
        rm.graph_net.node_encoder.model[0].weight.mul_(60.0)
    acts = []
    hook = rm.graph_net.node_encoder.model[1].register_forward_hook(lambda m, a, o: acts.append(float(o.abs().max())))
    imgs = synthetic_images(2, r_, seed=5)
    logits, nodes = [], []
    for b in range(2):
        x, pos, ei = ref.optimized.image_to_graph_pixel_optimized(Image.fromarray(imgs[b]), r_)
        tx, tpos, tei = ogb.to_model_inputs(x, pos, ei)
        with torch.no_grad():
            logits.append(rm((tx, tpos, tei)).numpy())
            nodes.append(rm.graph_net(tx, tpos, tei).numpy())
    hook.remove()
    _check(max(acts) > 4094, f"the large-activation case must leave the fp16 domain (max {max(acts)})")
    out["big_imgs"], out["big_logits"], out["big_nodes"] = imgs, np.stack(logits), np.stack(nodes)
    out["big_max_hidden_activation"] = np.array(max(acts))
    out["big_scale"], out["big_seed"] = np.array(60.0), np.array(13)
    np.savez_compressed(os.path.join(GOLDEN, "checkpoint.npz"), **out)
    print("checkpoint.npz:", len(out), "entries; large-activation case max", max(acts))


def gen_jpeg(ref):
    """The shipped JPEG files themselves (bytes) with the pixels the unmodified reference builder is fed from them:
    ``image_to_graph_pixel_optimized(path, r)`` with ``r`` = the file's own size returns the decoded RGB pixels as its
    node features (reference image_to_graph_optimized.py:65-72: open, convert('RGB'), resize to the same size is the
    identity, reshape).  oracle/jpeg.py must reproduce them bit for bit; the device decoder is tested against both."""
    from oracle.jpeg import decode_baseline
    out = {}
    for rel, r_ in (("static/chihuahua/img_4_799_32.jpg", 32), ("static/muffin/img_4_880_32.jpg", 32),
                    ("static/chihuahua/img_4_799_64.jpg", 64), ("static/muffin/img_4_880_64.jpg", 64),
                    ("static/muffin/img_0_187_128.jpg", 128), ("static/chihuahua/img_4_799_256.jpg", 256)):
        path = os.path.join(rl.REFERENCE_ROOT, rel)
        data = open(path, "rb").read()
        x, _, _ = ref.optimized.image_to_graph_pixel_optimized(path, r_)
        px = np.asarray(x, dtype=np.uint8).reshape(r_, r_, 3)
        _check(np.array_equal(px, np.array(Image.open(path).convert("RGB"))), f"reference builder pixels {rel}")
        _check(np.array_equal(decode_baseline(data), px), f"oracle JPEG decode {rel}")
        k = os.path.basename(rel).replace(".jpg", "")
        out[k + "_bytes"] = np.frombuffer(data, dtype=np.uint8)
        out[k + "_rgb"] = px
    np.savez_compressed(os.path.join(GOLDEN, "jpeg_files.npz"), **out)
    print("jpeg_files.npz:", len(out), "entries,", sum(v.nbytes for k, v in out.items() if k.endswith("_bytes")), "file bytes")


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    ref = rl.load_reference()
    torch.set_num_threads(4)
    if "--only-resize" in sys.argv:
        gen_resize(ref)
        return
    if "--only-mlp" in sys.argv:
        gen_mlp(ref)
        return
    if "--only-checkpoint" in sys.argv:
        gen_checkpoint(ref)
        return
    if "--only-jpeg" in sys.argv:
        gen_jpeg(ref)
        return
    gen_grids(ref)
    gen_builders(ref)
    gen_model(ref)
    gen_resize(ref)
    gen_mlp(ref)
    gen_checkpoint(ref)
    gen_jpeg(ref)
    print("all reference-vs-oracle comparisons passed; golden vectors written to", GOLDEN)


if __name__ == "__main__":
    main()
