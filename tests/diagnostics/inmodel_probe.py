"""In-model check of every tensor-core training op against torch on the SAME inputs (failing smoke config)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from graphnet_classifier_b200 import ops, build
build.build()
from graphnet_classifier_b200.models.GNN import CombinedModel, GraphNet
from graphnet_classifier_b200.pipeline import GraphClassifierPipeline
from oracle import gnn as ognn
from oracle.weights import fill_deterministic, synthetic_images
r, B = 16, 4
cfg = dict(num_local_features=3, space_dim=2, out_channels=1, n_blocks=3)
om = ognn.OracleCombinedModel(ognn.OracleGraphNet(**cfg), num_nodes=r * r, classes=2)
fill_deterministic(om, seed=1)
gm = CombinedModel(GraphNet(**cfg), num_nodes=r * r, classes=2)
gm.load_state_dict(om.state_dict()); gm = gm.cuda()
imgs = synthetic_images(B, r, seed=3)
labels = np.array([i % 2 for i in range(B)])
orig_wgrad, orig_lin = ops.tc_wgrad, ops.tc_linear
log = []
def wgrad(dZ, X, out=None, accumulate=False, want_db=False):
    res = orig_wgrad(dZ, X, out=out, accumulate=accumulate, want_db=want_db)
    dW = res[0] if want_db else res
    ref = dZ.double().t() @ X.double()
    err = float((dW.double() - ref).norm() / (ref.norm() + 1e-30))
    log.append(("wgrad", tuple(dZ.shape), dZ.is_contiguous(), X.is_contiguous(), tuple(X.stride()), err))
    return res
def lin(A, W, **kw):
    out = orig_lin(A, W, **kw)
    if kw.get("mask") is not None and kw.get("transpose_w"):
        ref = (A.double() @ W.double()) * (kw["mask"] > 0)
        err = float((out.double() - ref).norm() / (ref.norm() + 1e-30))
        log.append(("masked_dgrad", tuple(A.shape), A.is_contiguous(), kw["mask"].is_contiguous(), tuple(kw["mask"].stride()), err))
    return out
ops.tc_wgrad, ops.tc_linear = wgrad, lin
pipe = GraphClassifierPipeline(gm, resize_value=r)
loss = pipe.forward_backward(torch.from_numpy(imgs), torch.from_numpy(labels))
torch.cuda.synchronize()
bad = [l for l in log if l[-1] > 1e-5]
print(len(log), "ops logged;", len(bad), "with error > 1e-5")
for l in log[:6] + bad[:10]:
    print(l)
ops.tc_wgrad, ops.tc_linear = orig_wgrad, orig_lin
from oracle import graph_build as ogb
for im, lab in zip(imgs, labels):
    out = om(ogb.to_model_inputs(*ogb.pixel_graph(im)))
    (torch.nn.functional.cross_entropy(out, torch.tensor(int(lab))) / B).backward()
g = dict(gm.named_parameters())["graph_net.node_decoder.model.0.weight"].grad.cpu().double()
o = dict(om.named_parameters())["graph_net.node_decoder.model.0.weight"].grad.double()
d = g - o
print("dec0.weight grad: |g|", float(g.norm()), "|o|", float(o.norm()), "|d|", float(d.norm()), "ratio g.o/o.o", float((g * o).sum() / (o * o).sum()))
rows = d.norm(dim=1) / (o.norm(dim=1) + 1e-30)
cols = d.norm(dim=0) / (o.norm(dim=0) + 1e-30)
print("per-row rel err: max", float(rows.max()), "median", float(rows.median()), " rows > 1e-3:", int((rows > 1e-3).sum()))
print("per-col rel err: max", float(cols.max()), "median", float(cols.median()), " cols > 1e-3:", int((cols > 1e-3).sum()))
gb_ = dict(gm.named_parameters())["graph_net.node_decoder.model.0.bias"].grad.cpu().double()
ob_ = dict(om.named_parameters())["graph_net.node_decoder.model.0.bias"].grad.double()
print("bias diff top5:", torch.topk((gb_ - ob_).abs(), 5))
