#!/usr/bin/env python
"""Small invocations of the kernels added at the end of round 2 for `compute-sanitizer --tool memcheck`: the chained
training forward with its stash (edge and node forms, ragged last tile), the node form with the aggregating loader
(in-degrees 0 - 2, rows past M), the paired aggregation, one training step and one inference step of the model."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphnet_classifier_b200 import ops
from graphnet_classifier_b200.models.GNN import CombinedModel, GraphNet
from graphnet_classifier_b200.pipeline import GraphClassifierPipeline

g = torch.Generator().manual_seed(0)
mk = lambda *s: torch.randn(*s, generator=g).cuda()
layers = [(mk(128, 128) / 11, mk(128) * 0.1) for _ in range(3)]
gamma, beta = torch.ones(128).cuda(), torch.zeros(128).cuda()
for M in (1, 255, 256 * 3 + 17):
    A = mk(M, 128)
    R = 50
    P, Q, T = mk(R, 128), mk(R, 128), mk(M, 128)
    i0 = torch.randint(0, R, (M,), generator=g).int().cuda(); i1 = torch.randint(0, R, (M,), generator=g).int().cuda()
    for kw in (dict(gather0=(P, i0), gather1=(Q, i1)), dict(gather0=(T, None))):
        st = [torch.empty(M, 128).cuda() for _ in range(3)] + [torch.empty(M).cuda(), torch.empty(M).cuda()]
        ops.tc_mlp_chain(A, layers, gamma=gamma, beta=beta, residual=A, stash=tuple(st), **kw)
    deg = torch.randint(0, 3, (M,), generator=g)
    deg[0] = 2                                  # at least one edge (the model takes this form only for graphs with edges)
    E = max(int(deg.sum()), 1)
    rowptr = torch.zeros(M + 1, dtype=torch.int32); rowptr[1:] = torch.cumsum(deg, 0).int()
    eid = torch.randperm(E, generator=g)[:int(deg.sum())].int()
    e = mk(E, 128); h = mk(M, 128); V0 = mk(128, 256) / 16
    ops.tc_mlp_chain(e, [(V0[:, 128:256], layers[0][1]), layers[1], layers[2]], operand2=(h, V0[:, 0:128]), gamma=gamma, beta=beta,
                     residual=h, agg=(rowptr.cuda(), eid.cuda()))
    N = max(M // 2, 1)
    ei = torch.stack([torch.randint(0, N, (M,), generator=g), torch.randint(0, N, (M,), generator=g)]).cuda()
    gi = ops.GraphIndex.from_edge_index(ei, N)
    ops._agg_pair_raw(gi.src_rowptr, gi.src_eid, gi.dst_rowptr, gi.dst_eid, A, N)
torch.cuda.synchronize()
r, B = 16, 3
model = CombinedModel(GraphNet(num_local_features=3, space_dim=2, out_channels=1, n_blocks=3), num_nodes=r * r, classes=2).cuda()
pipe = GraphClassifierPipeline(model, resize_value=r)
imgs = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (B, r, r, 3), dtype=np.uint8))
pipe.forward_backward(imgs, torch.tensor([0, 1, 1]))
pipe.infer(imgs)
torch.cuda.synchronize()
print("sanitize_case2 ok")
