#!/usr/bin/env python
"""The reference's own loop (utils/train_model.py:8-81: one optimizer step per graph, main.py:60 batch_size=1) on the
device: eager launches vs the captured step, next to the oracle port on the host cores.  BASELINE configs[0] shape
(resize 64) and the resize-128 shape."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphnet_classifier_b200 import build
build.build()
from graphnet_classifier_b200.models.GNN import CombinedModel, GraphNet
from graphnet_classifier_b200.utils.image_to_graph.batched import build_pixel_graphs
from graphnet_classifier_b200.utils.train_model import train
from oracle import gnn as ognn, graph_build as ogb

n_items = int(sys.argv[1]) if len(sys.argv) > 1 else 64
print("| resize | items | eager ms/graph | CUDA-graph ms/graph | oracle port (host, %d threads) ms/graph |" % (os.cpu_count() or 1))
print("|---|---|---|---|---|")
for r in (64, 128):
    rng = np.random.default_rng(0)
    imgs = rng.integers(0, 256, (n_items, r, r, 3), dtype=np.uint8)
    labels = rng.integers(0, 2, n_items)
    data = [(build_pixel_graphs(torch.from_numpy(im), use_cache=True).as_tuple(), torch.tensor(int(l))) for im, l in zip(imgs, labels)]
    res = {}

    class Timed:                       # per-epoch wall time, measured around the loop's own iteration
        def __init__(self, items):
            self.items, self.durs = items, []

        def __iter__(self):
            t0 = time.perf_counter()
            for it in self.items:
                yield it
            torch.cuda.synchronize()
            self.durs.append(time.perf_counter() - t0)

    for use_graph in (False, True):
        torch.manual_seed(0)
        model = CombinedModel(GraphNet(num_local_features=3, space_dim=2, out_channels=1, n_blocks=3), num_nodes=r * r).cuda()
        timed = Timed(data)
        with tempfile.TemporaryDirectory() as d:
            sys.stdout = open(os.devnull, "w")
            try:
                train(model, timed, epochs=3, output_path=d, cuda_graph=use_graph)     # epoch 1 holds warm-up / capture
            finally:
                sys.stdout = sys.__stdout__
        res[use_graph] = min(timed.durs[1:]) / n_items * 1e3
    torch.set_num_threads(os.cpu_count() or 1)
    om = ognn.build_reference_config_model(r, seed=0)
    opt = torch.optim.Adam(om.parameters(), lr=1e-3)
    n_cpu = 6
    t0 = time.perf_counter()
    for im, l in zip(imgs[:n_cpu], labels[:n_cpu]):
        loss = torch.nn.functional.cross_entropy(om(ogb.to_model_inputs(*ogb.pixel_graph(im))), torch.tensor(int(l)))
        opt.zero_grad(); loss.backward(); opt.step(); loss.item()
    cpu = (time.perf_counter() - t0) / n_cpu * 1e3
    print(f"| {r} | {n_items} | {res[False]:.2f} | {res[True]:.2f} | {cpu:.0f} |")
