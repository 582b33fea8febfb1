#!/usr/bin/env python
"""Soak of the kernels added at the end of round 2: many launches on fresh random inputs, every result compared bit for bit
with the form it replaces (timing-dependent faults - a barrier or prefetch race - would show as rare mismatches).
    python scripts/soak_new_kernels.py [iterations] [graphs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphnet_classifier_b200 import ops

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
B = int(sys.argv[2]) if len(sys.argv) > 2 else 48
r, dev = 128, "cuda"
N, E = B * r * r, B * 2 * r * (r - 1)
g = torch.Generator(device=dev).manual_seed(1)
mk = lambda *s: torch.randn(*s, device=dev, generator=g)
v = torch.arange(r * r, device=dev).view(r, r)
src1 = torch.cat([v[:, :-1].reshape(-1), v[:-1, :].reshape(-1)]); dst1 = torch.cat([v[:, 1:].reshape(-1), v[1:, :].reshape(-1)])
off = (torch.arange(B, device=dev) * r * r).view(B, 1)
src = (src1.view(1, -1) + off).reshape(-1).int(); dst = (dst1.view(1, -1) + off).reshape(-1).int()
gi = ops.GraphIndex.from_edge_index(torch.stack([src.long(), dst.long()]), N)
gamma, beta = torch.rand(128, device=dev, generator=g) + 0.5, mk(128) * 0.1
bad = {"agg": 0, "stash_out": 0, "stash_a": 0, "pair": 0}
for it in range(iters):
    layers = [(mk(128, 128) / 11, mk(128) * 0.1) for _ in range(3)]
    V0 = mk(128, 256) / 16
    e, h, P, Q = mk(E, 128), mk(N, 128), mk(N, 128), mk(N, 128)
    # node processor: aggregation folded into the launch against aggregation kernel + launch
    ln = [(V0[:, 128:256], layers[0][1]), layers[1], layers[2]]
    ref = ops.tc_mlp_chain(ops._agg_raw(gi.dst_rowptr, gi.dst_eid, e, N), ln, operand2=(h, V0[:, 0:128]), gamma=gamma, beta=beta, residual=h)
    got = ops.tc_mlp_chain(e, ln, operand2=(h, V0[:, 0:128]), gamma=gamma, beta=beta, residual=h, agg=(gi.dst_rowptr, gi.dst_eid))
    bad["agg"] += int(not torch.equal(ref, got))
    # edge processor: chained launch with the stash against the plain launch and the per-layer engine's activations
    st = [torch.empty(E, 128, device=dev) for _ in range(3)] + [torch.empty(E, device=dev), torch.empty(E, device=dev)]
    plain = ops.tc_mlp_chain(e, layers, gather0=(P, src), gather1=(Q, dst), gamma=gamma, beta=beta, residual=e)
    out = ops.tc_mlp_chain(e, layers, gather0=(P, src), gather1=(Q, dst), gamma=gamma, beta=beta, residual=e, stash=tuple(st))
    bad["stash_out"] += int(not torch.equal(plain, out))
    a1 = ops.tc_linear(e, layers[0][0], bias=layers[0][1], gather0=(P, src), gather1=(Q, dst), relu=True)
    bad["stash_a"] += int(float((st[0] - a1).abs().max()) > 1e-4 * float(a1.abs().max()))
    # paired aggregation against two single ones
    pa, pb = ops._agg_pair_raw(gi.src_rowptr, gi.src_eid, gi.dst_rowptr, gi.dst_eid, e, N)
    bad["pair"] += int(not (torch.equal(pa, ops._agg_raw(gi.src_rowptr, gi.src_eid, e, N)) and torch.equal(pb, ops._agg_raw(gi.dst_rowptr, gi.dst_eid, e, N))))
torch.cuda.synchronize()
print(f"soak: {iters} iterations at {B} graphs of resize {r} ({E} edge rows): mismatches {bad}")
sys.exit(1 if any(bad.values()) else 0)
