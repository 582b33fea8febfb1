"""Which tensors miss the 1e-5 gradient bar at resize 64 / 128, on which engine, and how the float32 oracle itself
compares with the float64 oracle there (ReLU-kink flips show as noise >> 1e-6).  python tests/diagnostics/grad_r64_probe.py [r] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import gnn as ognn, graph_build as ogb
from oracle.weights import fill_deterministic, synthetic_images
from graphnet_classifier_b200 import ops, tc_train
from graphnet_classifier_b200.models.GNN import CombinedModel, GraphNet
from graphnet_classifier_b200.utils.image_to_graph.batched import build_pixel_graphs

r = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 11
rel = lambda a, b: float((a.double().cpu() - b.double().cpu()).norm() / b.double().norm().clamp_min(1e-300))
cfg = dict(num_local_features=3, space_dim=2, out_channels=1, n_blocks=3)
om = ognn.OracleCombinedModel(ognn.OracleGraphNet(**cfg), num_nodes=r * r, classes=2)
fill_deterministic(om, seed=seed)
imgs = synthetic_images(B, r, seed=3 * r)
labels = torch.tensor([i % 2 for i in range(B)])
lo = sum(torch.nn.functional.cross_entropy(om(ogb.to_model_inputs(*ogb.pixel_graph(im, False))), l) for im, l in zip(imgs, labels)) / B
lo.backward()
om64 = ognn.OracleCombinedModel(ognn.OracleGraphNet(**cfg), num_nodes=r * r, classes=2).double()
om64.load_state_dict(om.state_dict())
for mlp in [m for m in om64.modules() if isinstance(m, ognn.OracleMLP)]:
    mlp.forward = (lambda x, mlp=mlp: mlp.model(x.reshape(x.shape[0], -1)))
l64 = sum(torch.nn.functional.cross_entropy(
    om64(tuple(t.double() if t.is_floating_point() else t for t in ogb.to_model_inputs(*ogb.pixel_graph(im, False)))), l)
    for im, l in zip(imgs, labels)) / B
l64.backward()
res = {}
for eng in ("tc", "tc-pair", "fp32"):
    ops.ENGINE = eng.split("-")[0]
    tc_train.BWD = "pair" if eng == "tc-pair" else "fused"
    gm = CombinedModel(GraphNet(**cfg), num_nodes=r * r, classes=2)
    gm.load_state_dict(om.state_dict())
    gm = gm.cuda()
    gb = build_pixel_graphs(torch.from_numpy(imgs))
    loss = torch.nn.functional.cross_entropy(gm(gb.as_tuple()).reshape(B, -1), labels.cuda())
    loss.backward()
    res[eng] = {n: p.grad.detach().cpu() for n, p in gm.named_parameters()}
    print(f"{eng}: loss {loss.item():.8f} oracle32 {lo.item():.8f} oracle64 {l64.item():.8f}")
print(f"{'tensor':78s} {'noise32v64':>10s} " + " ".join(f"{e:>9s}" for e in res))
for (n, po), (_, p64) in zip(om.named_parameters(), om64.named_parameters()):
    noise = rel(po.grad, p64.grad)
    errs = [min(rel(res[e][n], po.grad), rel(res[e][n], p64.grad)) for e in res]
    flag = " <<<" if max(errs) > 1e-5 or noise > 1e-5 else ""
    print(f"{n:78s} {noise:10.2e} " + " ".join(f"{x:9.2e}" for x in errs) + flag)
