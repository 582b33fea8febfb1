"""Diagnostic: per-step parameter difference between the captured training step and eager steps (same Adam flavour or not)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from graphnet_classifier_b200 import build
build.build()
from graphnet_classifier_b200.models.GNN import CombinedModel, GraphNet
from graphnet_classifier_b200.utils.image_to_graph.batched import build_pixel_graphs
from graphnet_classifier_b200.utils.train_model import _GraphedStep
from oracle.weights import fill_deterministic, synthetic_images
from oracle import gnn as ognn

r = 8
kw = dict(num_local_features=3, space_dim=2, out_channels=1, n_blocks=1)
imgs = synthetic_images(6, r, seed=13)
labels = [0, 1, 1, 0, 1, 0]
data = [(build_pixel_graphs(torch.from_numpy(im), use_cache=True).as_tuple(), torch.tensor(l).cuda()) for im, l in zip(imgs, labels)]


def fresh():
    om = ognn.OracleCombinedModel(ognn.OracleGraphNet(**kw), num_nodes=r * r, classes=2)
    fill_deterministic(om, seed=7)
    gm = CombinedModel(GraphNet(**kw), num_nodes=r * r, classes=2)
    gm.load_state_dict(om.state_dict())
    return gm.cuda()


crit = torch.nn.CrossEntropyLoss()
runs = {}
for mode in ("graph", "eager-capturable", "eager-plain"):
    gm = fresh()
    opt = torch.optim.Adam(gm.parameters(), lr=1e-3, capturable=(mode != "eager-plain"))
    hist, gst = [], None
    for ep in range(2):
        for sample, label in data:
            if mode == "graph":
                if gst is None:
                    gst = _GraphedStep(gm, opt, crit, sample, label)
                    gst.capture()
                loss = gst.run(sample, label)
            else:
                loss = crit(gm(sample), label)
                opt.zero_grad(); loss.backward(); opt.step()
            torch.cuda.synchronize()
            hist.append((loss.item(), [p.detach().clone() for p in gm.parameters()], [p.grad.detach().clone() for p in gm.parameters()]))
    runs[mode] = hist
names = [n for n, _ in fresh().named_parameters()]
for other in ("eager-capturable", "eager-plain"):
    print("graph vs", other)
    for i, (a, b) in enumerate(zip(runs["graph"], runs[other])):
        dp = [float((x - y).abs().max()) for x, y in zip(a[1], b[1])]
        dg = [float((x - y).abs().max() / (y.abs().max() + 1e-30)) for x, y in zip(a[2], b[2])]
        k = int(np.argmax(dp))
        print(f"  step {i}: loss diff {abs(a[0] - b[0]):.2e}  max param diff {max(dp):.2e} ({names[k]})  max rel grad diff {max(dg):.2e} ({names[int(np.argmax(dg))]})")
