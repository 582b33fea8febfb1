"""GPU: OptimizedDatasetLoader / train_GNN over a synthetic ImageFolder (reference
utils/dataloader.py:10-53, main.py:32-76): item contract, dtypes, parity with the oracle
builders on the PIL-resized pixels, unknown-method error, and a one-epoch train_GNN run."""
import os

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import graph_build as ogb

pytestmark = pytest.mark.gpu


@pytest.fixture()
def image_folder(tmp_path):
    rng = np.random.default_rng(0)
    for cls in ("chihuahua", "muffin"):
        d = tmp_path / "dataset" / cls
        d.mkdir(parents=True)
        for i in range(3):
            Image.fromarray(rng.integers(0, 256, (40, 52, 3), dtype=np.uint8)).save(d / f"img_{i}.png")
    return str(tmp_path / "dataset")


def test_item_contract_matches_reference_loader(libgnc, image_folder):
    from graphnet_classifier_b200.utils.dataloader import OptimizedDatasetLoader
    r = 16
    for method, kw in (("pixel", dict(diagonals=True)), ("patch", dict(patch_size=4))):
        ds = OptimizedDatasetLoader(image_folder, resize_value=r, method=method, **kw)
        assert len(ds) == 6 and ds.dataset.classes == ["chihuahua", "muffin"]
        (x, pos, ei), label = ds[4]
        assert x.is_cuda and x.dtype == torch.float32 and pos.dtype == torch.float32
        assert ei.dtype == torch.int64 and label.dtype == torch.int64 and label.dim() == 0 and int(label) == 1
        img, _ = ds.dataset[4]
        tab = np.array(img.convert("RGB").resize((r, r)))           # what the reference builder sees
        if method == "pixel":
            ox, opos, oei = ogb.pixel_graph(tab, True)
        else:
            ox, opos, oei = ogb.patch_graph(tab, 4)
        assert np.array_equal(x.cpu().numpy(), np.asarray(ox, dtype=np.float32))
        assert np.array_equal(pos.cpu().numpy(), np.asarray(opos, dtype=np.float32))
        assert np.array_equal(ei.cpu().numpy(), oei)
    with pytest.raises(ValueError, match="Unknown method"):
        OptimizedDatasetLoader(image_folder, resize_value=r, method="voxel")[0]


def test_train_gnn_entry_point(libgnc, image_folder, tmp_path):
    from graphnet_classifier_b200.main import train_GNN
    out = str(tmp_path / "weights" / "GNN")
    best = train_GNN(epochs=1, resize_value=8, max_samples=4, output_path=out, dataset_path=image_folder)
    assert np.isfinite(best)
    names = os.listdir(out)
    assert "final_model.pth" in names and "best_model_epoch1.pth" in names
    sd = torch.load(os.path.join(out, "final_model.pth"), map_location="cpu")
    assert len(sd) == 76 and tuple(sd["classifier.fc1.weight"].shape) == (128, 64)


def test_gnn_inference_helper(libgnc, tmp_path):
    """reference utils/inference.py:32-71: checkpoint -> single-image logits / probabilities."""
    from graphnet_classifier_b200.utils.inference import gnn_inference
    from oracle import gnn as ognn
    r = 16
    om = ognn.build_reference_config_model(r, seed=3)
    ck = str(tmp_path / "ck.pth")
    torch.save(om.state_dict(), ck)                       # a checkpoint in the reference's format
    img = np.random.default_rng(5).integers(0, 256, (r, r, 3), dtype=np.uint8)
    path = str(tmp_path / "img.png")
    Image.fromarray(img).save(path)
    logits, probs = gnn_inference(path, ck, resize_value=r)
    assert logits.shape == (1, 2) and probs.shape == (1, 2) and abs(float(probs.sum()) - 1.0) < 1e-6
    with torch.no_grad():
        exp = om(ogb.to_model_inputs(*ogb.pixel_graph(img)))
    np.testing.assert_allclose(logits[0].cpu().numpy(), exp.numpy(), rtol=1e-5, atol=1e-7)


def test_train_gnn_batched_one_step_matches_oracle(libgnc, tmp_path):
    """main.train_GNN_batched, single process: a dataset of B unresized PIL images with batch_size = B is one Adam step
    on the mean cross entropy of the B graphs - compared with the oracle doing exactly that on PIL-resized pixels."""
    import numpy as np
    from PIL import Image
    from graphnet_classifier_b200.main import train_GNN_batched
    from oracle import gnn as ognn
    from oracle import graph_build as ogb
    r, B = 8, 4
    rng = np.random.default_rng(17)
    photos = [rng.integers(0, 256, (20 + 3 * i, 31 - 2 * i, 3), dtype=np.uint8) for i in range(B)]
    labels = [0, 1, 1, 0]
    data = [(Image.fromarray(p), l) for p, l in zip(photos, labels)]
    torch.manual_seed(0)
    best, gm = train_GNN_batched(epochs=1, resize_value=r, batch_size=B, output_path=str(tmp_path), dataset=data, shuffle=False)
    # oracle: same init (seed 0, same construction order), PIL resize, one graph per forward, mean loss, one Adam step
    om = ognn.build_reference_config_model(r, seed=0)
    opt = torch.optim.Adam(om.parameters(), lr=1e-3)
    loss = sum(torch.nn.functional.cross_entropy(
        om(ogb.to_model_inputs(*ogb.pixel_graph(np.asarray(Image.fromarray(p).resize((r, r)))))), torch.tensor(l))
        for p, l in zip(photos, labels)) / B
    opt.zero_grad(), loss.backward(), opt.step()
    assert abs(best - loss.item()) < 1e-5 * max(1.0, loss.item())
    for (name, p), (_, po) in zip(gm.named_parameters(), om.named_parameters()):
        assert float((p.detach().cpu() - po.detach()).abs().max()) < 2e-3 * 1.01, name     # first Adam step: |update| <= lr
    assert sorted(f for f in os.listdir(tmp_path) if f.endswith(".pth")) == ["best_model_epoch1.pth", "final_model.pth"]


def test_decode_pool_equals_reference_pixels(libgnc, tmp_path):
    """utils/staging.DecodePool (threaded decode into pinned buffers, double-buffered copies, device resize) hands the
    graph builder the very pixels the reference's ``Image.open(path).convert('RGB').resize((r, r))`` produces - JPEG and
    PNG files of several sizes (one of them already r x r, one greyscale), in file order, chunked or not."""
    from graphnet_classifier_b200.utils.staging import DecodePool
    rng = np.random.default_rng(4)
    r = 24
    paths = []
    shapes = [(40, 52), (24, 24), (40, 52), (64, 30), (24, 24), (375, 500), (40, 52), (33, 47), (64, 30)]
    for i, (h, w) in enumerate(shapes):
        low = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2, 3), dtype=np.uint8)
        img = Image.fromarray(low).resize((w, h), Image.BILINEAR)
        if i == 7:
            img = img.convert("L")
        path = tmp_path / (f"f{i}.jpg" if i % 3 else f"f{i}.png")
        img.save(path, quality=90) if path.suffix == ".jpg" else img.save(path)
        paths.append(str(path))
    want = np.stack([np.array(Image.open(p).convert("RGB").resize((r, r))) for p in paths])
    for device_jpeg in (True, False):
        with DecodePool(workers=3, device_jpeg=device_jpeg) as pool:
            got = pool.stage(paths, r)
            assert got.is_cuda and got.dtype == torch.uint8 and tuple(got.shape) == (len(paths), r, r, 3)
            assert np.array_equal(got.cpu().numpy(), want)
            chunks = list(pool.batches(paths, r, chunk=4))
            assert [c.shape[0] for c in chunks] == [4, 4, 1]
            assert np.array_equal(torch.cat(chunks).cpu().numpy(), want)
            again = torch.cat(list(pool.batches(paths, r, chunk=2))).cpu().numpy()     # staging buffers reused many times
            assert np.array_equal(again, want)


def test_infer_files_equals_infer(libgnc, tmp_path):
    from graphnet_classifier_b200.models.GNN import CombinedModel, GraphNet
    from graphnet_classifier_b200.pipeline import GraphClassifierPipeline
    from graphnet_classifier_b200.utils.staging import infer_files
    rng = np.random.default_rng(5)
    r = 16
    paths = []
    for i in range(7):
        p = tmp_path / f"g{i}.jpg"
        Image.fromarray(rng.integers(0, 256, (30 + 3 * (i % 2), 41, 3), dtype=np.uint8)).save(p, quality=85)
        paths.append(str(p))
    torch.manual_seed(0)
    model = CombinedModel(GraphNet(num_local_features=3, space_dim=2, out_channels=1, n_blocks=3), num_nodes=r * r).cuda()
    pipe = GraphClassifierPipeline(model, resize_value=r)
    got = infer_files(pipe, paths, chunk=3)
    px = np.stack([np.array(Image.open(p).convert("RGB").resize((r, r))) for p in paths])
    want = pipe.infer(torch.from_numpy(px))
    assert torch.equal(got, want)
