"""GPU: GraphNet / CombinedModel end to end through the reference-shaped API vs the
golden vectors (reference output) and the oracle: logits, loss and every parameter
gradient within 1e-5 relative; batching; checkpoint exchange; the train loop."""
import os

import numpy as np
import pytest
import torch

from oracle import gnn as ognn
from oracle import graph_build as ogb
from oracle.weights import fill_deterministic, synthetic_images

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _models(r, n_blocks=3, seed=7, **gn_kw):
    from graphnet_classifier_b200.models.GNN import CombinedModel, GraphNet
    kw = dict(num_local_features=3, space_dim=2, out_channels=1, n_blocks=n_blocks)
    kw.update(gn_kw)
    om = ognn.OracleCombinedModel(ognn.OracleGraphNet(**kw), num_nodes=r * r, classes=2)
    fill_deterministic(om, seed=seed)
    gm = CombinedModel(GraphNet(**kw), num_nodes=r * r, classes=2)
    gm.load_state_dict(om.state_dict())
    return om, gm.cuda()


def _inputs(img, diag):
    return ogb.to_model_inputs(*ogb.pixel_graph(img, diag))


@pytest.mark.parametrize("r,diag", [(8, False), (8, True), (12, False)])
def test_logits_loss_grads_vs_golden_reference(golden, libgnc, r, diag):
    m = golden["model"]
    tag = f"model_r{r}_d{int(diag)}"
    om, gm = _models(r)
    imgs = m[tag + "_imgs"]
    for b in range(imgs.shape[0]):
        x, pos, ei = _inputs(imgs[b], diag)
        # the reference hands over a non-contiguous edge_index (transposed view): keep that
        out = gm((x.cuda(), pos.cuda(), ei.cuda()))
        assert out.shape == (2,)                                        # 1-D logits (Q11)
        np.testing.assert_allclose(out.detach().cpu().numpy(), m[tag + "_logits"][b], rtol=RTOL, atol=1e-7)
    x, pos, ei = _inputs(imgs[0], diag)
    label = torch.tensor(int(m[tag + "_label"]))
    loss = torch.nn.functional.cross_entropy(gm((x.cuda(), pos.cuda(), ei.cuda())), label.cuda())
    assert abs(loss.item() - float(m[tag + "_loss"])) < 1e-5 * max(1.0, float(m[tag + "_loss"]))
    loss.backward()
    torch.nn.functional.cross_entropy(om((x, pos, ei)), label).backward()
    for (name, p), (_, po) in zip(gm.named_parameters(), om.named_parameters()):
        assert _rel(p.grad, po.grad) < RTOL, name                      # full tensor vs oracle
        samp = m[f"{tag}_grad_{name}_samp"]                             # sampled entries vs reference golden
        f = p.grad.flatten().cpu()
        stride = max(1, -(-f.numel() // 512))
        scale = float(m[f"{tag}_grad_{name}_sn"][1]) / np.sqrt(f.numel()) + 1e-30
        assert np.max(np.abs(f[::stride].numpy() - samp)) <= 10 * RTOL * max(scale, np.max(np.abs(samp))), name


def test_batched_equals_per_sample_and_builder_path(golden, libgnc):
    from graphnet_classifier_b200.utils.image_to_graph.batched import build_pixel_graphs
    m = golden["model"]
    om, gm = _models(8)
    imgs = m["model_r8_d0_imgs"]
    gb = build_pixel_graphs(torch.from_numpy(imgs))
    out = gm(gb.as_tuple())                                            # [B, classes]
    assert out.shape == (3, 2)
    np.testing.assert_allclose(out.detach().cpu().numpy(), m["model_r8_d0_logits"], rtol=RTOL, atol=1e-7)
    out2 = gm(gb.x, gb.pos, gb.edge_index)                             # positional form
    assert torch.equal(out, out2)
    # edge_index without the attached CSR (generic csr_build path) gives the same bits
    out3 = gm(gb.x, gb.pos, gb.edge_index.clone())
    assert torch.equal(out, out3)


def test_resize32_seed0_matches_survey_probe_and_checkpoint_roundtrip(golden, libgnc, tmp_path):
    from graphnet_classifier_b200.models.GNN import CombinedModel, GraphNet
    torch.manual_seed(0)
    gm = CombinedModel(GraphNet(num_local_features=3, space_dim=2, out_channels=1, n_blocks=3), num_nodes=1024,
                       classes=2)
    om = ognn.build_reference_config_model(32, seed=0)
    for k, v in om.state_dict().items():
        assert torch.equal(v, gm.state_dict()[k]), k
    gm = gm.cuda()
    img = np.random.default_rng(0).integers(0, 256, (32, 32, 3), dtype=np.uint8)
    x, pos, ei = _inputs(img, False)
    out = gm((x.cuda(), pos.cuda(), ei.cuda()))
    np.testing.assert_allclose(out.detach().cpu().numpy(), golden["model"]["seed0_r32_logits_img0"], rtol=RTOL)
    # checkpoint written by us loads into the reference-shaped oracle tree and vice versa
    path = os.path.join(tmp_path, "ck.pth")
    torch.save(gm.state_dict(), path)
    sd = torch.load(path, map_location="cpu")
    assert list(sd.keys()) == list(golden["model"]["checkpoint_keys"])
    om2 = ognn.build_reference_config_model(32, seed=1)
    om2.load_state_dict(sd)
    with torch.no_grad():
        np.testing.assert_allclose(om2((x, pos, ei)).numpy(), out.detach().cpu().numpy(), rtol=RTOL)


@pytest.mark.parametrize("kw", [
    dict(n_blocks=2, out_dim_node=64, out_dim_edge=32, hidden_dim_processor_edge=96),
    dict(n_blocks=1, activation="Tanh", hidden_layers_processor_node=3),
    dict(n_blocks=1, norm_type=None, out_channels=2),
])
def test_non_default_configurations(libgnc, kw):
    from graphnet_classifier_b200.models.GNN import CombinedModel, GraphNet
    base = dict(num_local_features=3, space_dim=2, out_channels=1)
    base.update(kw)
    r = 6
    ogn = ognn.OracleGraphNet(**base)
    om = ognn.OracleCombinedModel(ogn, num_nodes=r * r, classes=3)
    fill_deterministic(om, seed=3)
    gm = CombinedModel(GraphNet(**base), num_nodes=r * r, classes=3)
    gm.load_state_dict(om.state_dict())
    gm = gm.cuda()
    x, pos, ei = _inputs(synthetic_images(1, r, 5)[0], True)
    label = torch.tensor(2)
    lo = torch.nn.functional.cross_entropy(om((x, pos, ei)), label)
    lg = torch.nn.functional.cross_entropy(gm((x.cuda(), pos.cuda(), ei.cuda())), label.cuda())
    assert abs(lo.item() - lg.item()) < RTOL * max(1.0, abs(lo.item()))
    lo.backward(), lg.backward()
    for (name, p), (_, po) in zip(gm.named_parameters(), om.named_parameters()):
        assert _rel(p.grad, po.grad) < 5 * RTOL, name


def test_pipeline_infer_and_train_step_match_oracle(libgnc):
    from graphnet_classifier_b200.pipeline import GraphClassifierPipeline
    r, B = 10, 7
    om, gm = _models(r, n_blocks=2)
    imgs = synthetic_images(B, r, seed=21)
    labels = np.random.default_rng(1).integers(0, 2, B)
    pipe = GraphClassifierPipeline(gm, resize_value=r, micro_batch=3, train_micro_batch=2)
    got = pipe.infer(torch.from_numpy(imgs))                           # 3 micro-batches: 3 + 3 + 1
    with torch.no_grad():
        exp = torch.stack([om(_inputs(im, False)) for im in imgs])
    np.testing.assert_allclose(got.cpu().numpy(), exp.numpy(), rtol=RTOL, atol=1e-7)
    # one training step: mean CE over the batch, gradient accumulation over micro-batches, Adam
    opt_g = torch.optim.Adam(gm.parameters(), lr=1e-3)
    opt_o = torch.optim.Adam(om.parameters(), lr=1e-3)
    loss_g = pipe.train_step(torch.from_numpy(imgs), torch.from_numpy(labels), opt_g)
    opt_o.zero_grad()
    loss_o = sum(torch.nn.functional.cross_entropy(om(_inputs(im, False)), torch.tensor(int(l))) for im, l in
                 zip(imgs, labels)) / B
    loss_o.backward()
    assert abs(loss_g.item() - loss_o.item()) < RTOL * max(1.0, loss_o.item())
    for (name, p), (_, po) in zip(gm.named_parameters(), om.named_parameters()):
        assert _rel(p.grad, po.grad) < RTOL, name
    opt_o.step()
    # Adam's first step is lr*sign(g)-like, so parameters are compared with an absolute bar
    for (name, p), (_, po) in zip(gm.named_parameters(), om.named_parameters()):
        assert float((p.detach().cpu() - po.detach()).abs().max()) < 2e-3 * 1.01, name


def test_reference_train_loop_one_graph_per_step(libgnc, tmp_path):
    # utils/train_model.train with a batch_size=1 style iterable, like main.py:60
    from graphnet_classifier_b200.utils.image_to_graph.batched import build_pixel_graphs
    from graphnet_classifier_b200.utils.train_model import train
    r = 8
    om, gm = _models(r, n_blocks=1)
    imgs = synthetic_images(4, r, seed=9)
    labels = [0, 1, 1, 0]
    data = [(build_pixel_graphs(torch.from_numpy(im), use_cache=True).as_tuple(), torch.tensor(l)) for im, l in
            zip(imgs, labels)]
    best = train(gm, data, epochs=2, patience=5, output_path=str(tmp_path))
    # same loop on the oracle model (reference semantics: Adam 1e-3, CE, one step per item)
    opt = torch.optim.Adam(om.parameters(), lr=1e-3)
    losses = []
    for _ in range(2):
        tot = 0.0
        for im, l in zip(imgs, labels):
            loss = torch.nn.functional.cross_entropy(om(_inputs(im, False)), torch.tensor(l))
            opt.zero_grad(), loss.backward(), opt.step()
            tot += loss.item()
        losses.append(tot / 4)
    assert abs(best - min(losses)) < 5e-3
    files = sorted(os.listdir(tmp_path))
    assert "final_model.pth" in files and "best_model_epoch1.pth" in files
    log = [f for f in files if f.startswith("training_logs_")]
    assert len(log) == 1
    text = open(os.path.join(tmp_path, log[0])).read()
    assert "Epochs: 2, Patience: 5" in text and "Epoch 2/2, avg_loss=" in text and "Best loss achieved:" in text


def test_train_loop_cuda_graph_equals_eager(libgnc, tmp_path):
    """utils/train_model.train replays one captured step per item (forward, CE, backward, Adam).  The replay is
    bit-identical to eager steps with the same (capturable) Adam, also when a second sample layout shows up
    mid-training; against torch's host-scalar Adam flavour the two update rules differ by rounding (1e-7 on the
    first step), which a dozen sign-like early Adam steps amplify."""
    from graphnet_classifier_b200.utils.image_to_graph.batched import build_pixel_graphs
    from graphnet_classifier_b200.utils.train_model import _GraphedStep, train
    r = 8
    imgs = synthetic_images(6, r, seed=13)
    labels = [0, 1, 1, 0, 1, 0]
    data = [(build_pixel_graphs(torch.from_numpy(im), use_cache=True).as_tuple(), torch.tensor(l)) for im, l in zip(imgs, labels)]
    # items 4 and 5 on another topology object (diagonals): a second layout, captured with a non-empty Adam state
    for i in (4, 5):
        data[i] = (build_pixel_graphs(torch.from_numpy(imgs[i]), diagonals=True, use_cache=True).as_tuple(), torch.tensor(labels[i]))
    crit = torch.nn.CrossEntropyLoss()
    finals = []
    for graphed in (True, False):
        _, gm = _models(r, n_blocks=1)
        opt = torch.optim.Adam(gm.parameters(), lr=1e-3, capturable=True)
        steps, losses = {}, []
        for _ in range(2):
            for sample, label in data:
                label = label.cuda()
                if graphed:
                    key = _GraphedStep.layout(sample, label)
                    assert key is not None
                    if key not in steps:
                        steps[key] = _GraphedStep(gm, opt, crit, sample, label)
                        steps[key].capture()
                    losses.append(steps[key].run(sample, label).item())
                else:
                    loss = crit(gm(sample), label)
                    opt.zero_grad(), loss.backward(), opt.step()
                    losses.append(loss.item())
        if graphed:
            assert len(steps) == 2
        finals.append((losses, [p.detach().clone() for p in gm.parameters()]))
    assert finals[0][0] == finals[1][0]                                  # every step's loss, bit for bit
    for a, b in zip(finals[0][1], finals[1][1]):
        assert torch.equal(a, b)
    # the public entry point, graph on (default) vs off (host-scalar Adam)
    results = []
    for use_graph in (True, False):
        _, gm = _models(r, n_blocks=1)
        best = train(gm, data, epochs=2, patience=5, output_path=str(tmp_path / f"g{int(use_graph)}"), cuda_graph=use_graph)
        results.append((best, [p.detach().clone() for p in gm.parameters()]))
    (b1, p1), (b0, p0) = results
    assert abs(b1 - b0) < 1e-5
    for a, c in zip(p1, finals[0][1]):
        assert torch.equal(a, c)                                         # train() == the hand-driven captured steps
    # against the host-scalar Adam only the aggregate is compared: elements whose true gradient is zero carry
    # rounding-noise gradients, which Adam normalises to +-lr steps of either sign in either flavour
    num = sum(float((a - b).pow(2).sum()) for a, b in zip(p1, p0)) ** 0.5
    den = sum(float(b.pow(2).sum()) for b in p0) ** 0.5
    assert num / den < 1e-3, num / den


def test_infer_graphed_equals_infer(libgnc):
    """CUDA-graph replay of the small-batch call returns the eager call's logits, for
    changing inputs and after a weight update."""
    import numpy as np
    from graphnet_classifier_b200.models.GNN import CombinedModel, GraphNet
    from graphnet_classifier_b200.pipeline import GraphClassifierPipeline
    torch.manual_seed(3)
    r = 32
    model = CombinedModel(GraphNet(num_local_features=3, space_dim=2, out_channels=1, n_blocks=3), num_nodes=r * r).cuda().eval()
    pipe = GraphClassifierPipeline(model, resize_value=r)
    rng = np.random.default_rng(5)
    for B in (1, 3):
        for _ in range(3):
            img = torch.from_numpy(rng.integers(0, 256, (B, r, r, 3), dtype=np.uint8))
            assert torch.equal(pipe.infer_graphed(img), pipe.infer(img))
    with torch.no_grad():
        for prm in model.parameters():
            prm.mul_(1.01)
    img = torch.from_numpy(rng.integers(0, 256, (1, r, r, 3), dtype=np.uint8))
    assert torch.equal(pipe.infer_graphed(img), pipe.infer(img))


def test_shipped_checkpoint_and_large_activation_fallback(libgnc):
    """ADVICE r1 / VERDICT r1 item 3.  (1) The reference's shipped checkpoint (weights/GNN/dim32_3block/
    best_model_epoch5.pth; its tensors travel in tests/golden/checkpoint.npz) on two shipped JPEGs: logits and node
    outputs of the unmodified reference, reproduced by the chained inference kernels - hidden activations reach 427 with
    raw 0..255 pixels, inside the fp16 two-piece domain (4094).  (2) Weights whose hidden activations reach 17 838: the
    chained kernels return non-finite rows there, the guard notices and the forward is evaluated on the 3xTF32 engine."""
    import os
    from graphnet_classifier_b200 import ops
    from graphnet_classifier_b200.models.GNN import CombinedModel, GraphNet
    from graphnet_classifier_b200.utils.image_to_graph.batched import build_pixel_graphs
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "checkpoint.npz"))
    cfg = dict(num_local_features=3, space_dim=2, out_channels=1, n_blocks=3)
    gm = CombinedModel(GraphNet(**cfg), num_nodes=32 * 32, classes=2)
    sd = {str(k): torch.from_numpy(g[f"w{i:02d}"]) for i, k in enumerate(g["keys"])}
    gm.load_state_dict(sd, strict=True)
    gm = gm.cuda().eval()
    assert 300 < float(g["max_hidden_activation"]) < 4094
    before = ops.CHAIN_GUARD_EVENTS
    for tag in ("chihuahua", "muffin"):
        gb = build_pixel_graphs(torch.from_numpy(g[f"{tag}_pixels"])[None])
        with torch.no_grad():
            logits = gm(gb.as_tuple())
            nodes = gm.graph_net(*gb.as_tuple())
        np.testing.assert_allclose(logits.cpu().numpy(), g[f"{tag}_logits"], rtol=1e-5, atol=1e-6)
        err = float((nodes.cpu() - torch.from_numpy(g[f"{tag}_nodes"])).abs().max() / np.abs(g[f"{tag}_nodes"]).max())
        assert err < 1e-5, err
    assert ops.CHAIN_GUARD_EVENTS == before                   # the chained kernels handled it: no fallback taken
    # activations outside the domain
    from oracle.weights import fill_deterministic
    r = 16
    om = ognn.OracleCombinedModel(ognn.OracleGraphNet(**cfg), num_nodes=r * r, classes=2)
    fill_deterministic(om, seed=int(g["big_seed"]))
    with torch.no_grad():
        om.graph_net.node_encoder.model[0].weight.mul_(float(g["big_scale"]))
    gm = CombinedModel(GraphNet(**cfg), num_nodes=r * r, classes=2)
    gm.load_state_dict(om.state_dict())
    gm = gm.cuda().eval()
    assert float(g["big_max_hidden_activation"]) > 4094
    gb = build_pixel_graphs(torch.from_numpy(g["big_imgs"]))
    with torch.no_grad():
        raw = gm.graph_net._forward_tc(gb.x, gb.pos, gb.graph)
        assert not bool(torch.isfinite(raw).all())            # loud, not silently wrong
        logits = gm(gb.as_tuple())
        nodes = gm.graph_net(*gb.as_tuple())
    assert ops.CHAIN_GUARD_EVENTS == before + 2
    np.testing.assert_allclose(logits.cpu().numpy(), g["big_logits"], rtol=1e-5, atol=1e-6)
    ref_nodes = torch.from_numpy(g["big_nodes"]).reshape(-1, 1)
    assert float((nodes.cpu() - ref_nodes).abs().max() / ref_nodes.abs().max()) < 1e-5
