"""BASELINE config 4 end to end on supplied label maps: superpixel graphs with DIFFERENT node
counts in one block-diagonal batch, GraphNet, the pad / truncate readout and the reference's head - logits, loss and
gradients against the oracle (per-graph reference GraphNet + oracle.gnn.padded_readout_logits)."""
import numpy as np
import pytest
import torch

from oracle import gnn as ognn
from oracle import graph_build as ogb
from oracle.weights import fill_deterministic

pytestmark = pytest.mark.gpu


def _voronoi_labels(rng, H, W, k):
    seeds = rng.random((k, 2)) * [H, W]
    yy, xx = np.mgrid[0:H, 0:W]
    d = (yy[..., None] - seeds[:, 0]) ** 2 + (xx[..., None] - seeds[:, 1]) ** 2
    return d.argmin(-1).astype(np.int32)


@pytest.mark.parametrize("num_nodes", [20, 40])
def test_variable_node_count_batch_logits_and_gradients(num_nodes):
    from graphnet_classifier_b200.models.GNN import CombinedModel, GraphNet
    from graphnet_classifier_b200.utils.image_to_graph.batched import build_superpixel_batch
    r, B = 48, 5
    rng = np.random.default_rng(3)
    imgs = rng.integers(0, 256, (B, r, r, 3), dtype=np.uint8)
    ks = [17, 33, 40, 25, 52]                       # fewer than, equal to and more than num_nodes
    labels = np.stack([_voronoi_labels(rng, r, r, k) for k in ks])
    cfg = dict(num_local_features=3, space_dim=2, out_channels=1, n_blocks=3)
    om = ognn.OracleCombinedModel(ognn.OracleGraphNet(**cfg), num_nodes=num_nodes, classes=2)
    fill_deterministic(om, seed=17)
    gm = CombinedModel(GraphNet(**cfg), num_nodes=num_nodes, classes=2)
    gm.load_state_dict(om.state_dict())
    gm = gm.cuda()
    gb = build_superpixel_batch(torch.from_numpy(imgs), labels=torch.from_numpy(labels), max_nodes=64)
    counts = (gb.graph.node_ptr[1:] - gb.graph.node_ptr[:-1]).cpu().tolist()
    graphs = [ogb.to_model_inputs(*ogb.superpixel_graph_from_labels(imgs[b], labels[b])) for b in range(B)]
    assert counts == [g[0].shape[0] for g in graphs] and len(set(counts)) > 1
    lab = torch.tensor([0, 1, 1, 0, 1])
    logits = gm(gb.as_tuple())
    assert logits.shape == (B, 2)
    loss = torch.nn.functional.cross_entropy(logits, lab.cuda())
    loss.backward()
    exp = ognn.padded_readout_logits(om, graphs)
    lo = torch.nn.functional.cross_entropy(exp, lab)
    lo.backward()
    np.testing.assert_allclose(logits.detach().cpu().numpy(), exp.detach().numpy(), rtol=1e-5, atol=1e-6)
    assert abs(loss.item() - lo.item()) < 1e-5 * max(1.0, lo.item())
    worst = 0.0
    for (name, p), (_, po) in zip(gm.named_parameters(), om.named_parameters()):
        rel = float((p.grad.cpu().double() - po.grad.double()).norm() / po.grad.double().norm().clamp_min(1e-30))
        worst = max(worst, rel)
        assert rel < 2e-5, (name, rel)
    print(f"variable-N readout num_nodes={num_nodes}: node counts {counts}, worst gradient rel-L2 {worst:.2e}")
    # one graph whose count equals num_nodes: the readout IS the reference's flatten (models/GNN.py:339)
    if num_nodes in counts:
        b = counts.index(num_nodes)
        with torch.no_grad():
            ref_b = om(graphs[b])
        np.testing.assert_allclose(logits[b].detach().cpu().numpy(), ref_b.numpy(), rtol=1e-5, atol=1e-6)


def test_segment_readout_pad_truncate_and_backward():
    from graphnet_classifier_b200 import ops
    ptr = torch.tensor([0, 3, 3, 10, 14], dtype=torch.int32, device="cuda")      # counts 3, 0, 7, 4
    y = torch.arange(14, dtype=torch.float32, device="cuda").reshape(14, 1).requires_grad_()
    out = ops.segment_readout(y, ptr, 5)
    exp = torch.tensor([[0, 1, 2, 0, 0], [0, 0, 0, 0, 0], [3, 4, 5, 6, 7], [10, 11, 12, 13, 0]], dtype=torch.float32)
    assert torch.equal(out.cpu(), exp)
    w = torch.arange(20, dtype=torch.float32, device="cuda").reshape(4, 5) + 1
    (out * w).sum().backward()
    g = torch.tensor([1, 2, 3, 11, 12, 13, 14, 15, 0, 0, 16, 17, 18, 19], dtype=torch.float32).reshape(14, 1)
    assert torch.equal(y.grad.cpu(), g)
