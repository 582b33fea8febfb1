#!/usr/bin/env python
"""One edge-MLP-shaped chained launch (gathered addends, LayerNorm, residual) for ncu: B graphs of resize 128."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphnet_classifier_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
r = 128; dev = "cuda"
N, E = B * r * r, B * 2 * r * (r - 1)
g = torch.Generator(device=dev).manual_seed(0)
mk = lambda *s: torch.randn(*s, device=dev, generator=g)
layers = [(mk(128, 128) / 11, mk(128) * 0.1) for _ in range(3)]
gamma, beta = torch.ones(128, device=dev), torch.zeros(128, device=dev)
v = torch.arange(r * r, device=dev).view(r, r)
src1 = torch.cat([v[:, :-1].reshape(-1), v[:-1, :].reshape(-1)]); dst1 = torch.cat([v[:, 1:].reshape(-1), v[1:, :].reshape(-1)])
off = (torch.arange(B, device=dev) * r * r).view(B, 1)
src = (src1.view(1, -1) + off).reshape(-1).int(); dst = (dst1.view(1, -1) + off).reshape(-1).int()
e = mk(E, 128); P = mk(N, 128); Q = mk(N, 128); out = torch.empty(E, 128, device=dev)
for _ in range(3):
    ops.tc_mlp_chain(e, layers, gather0=(P, src), gather1=(Q, dst), gamma=gamma, beta=beta, residual=e, out=out)
torch.cuda.synchronize()
print("ok", float(out[0, 0]))
# second case (argv[2] == "node"): the node processor with the aggregation folded into its loader (tc_chain2n_kernel, AGG)
if len(sys.argv) > 2 and sys.argv[2] == "node":
    from graphnet_classifier_b200.ops import GraphIndex
    g_idx = GraphIndex.from_edge_index(torch.stack([src.long(), dst.long()]), N)
    h = mk(N, 128); V0 = mk(128, 256) / 16; c0 = mk(128) * 0.1; outn = torch.empty(N, 128, device=dev)
    for _ in range(3):
        ops.tc_mlp_chain(e, [(V0[:, 128:256], c0), layers[1], layers[2]], operand2=(h, V0[:, 0:128]), gamma=gamma, beta=beta,
                         residual=h, out=outn, agg=(g_idx.dst_rowptr, g_idx.dst_eid))
    torch.cuda.synchronize()
    print("ok node", float(outn[0, 0]))
# third case (argv[2] == "stash"): the training forward of the edge MLP (chained launch that also writes a1, a2, z, mean, rstd)
if len(sys.argv) > 2 and sys.argv[2] == "stash":
    st = [torch.empty(E, 128, device=dev) for _ in range(3)] + [torch.empty(E, device=dev), torch.empty(E, device=dev)]
    for _ in range(3):
        ops.tc_mlp_chain(e, layers, gather0=(P, src), gather1=(Q, dst), gamma=gamma, beta=beta, residual=e, out=out, stash=tuple(st))
    torch.cuda.synchronize()
    print("ok stash", float(st[0][0, 0]))
