#!/usr/bin/env python
"""Micro-benchmark of the chained-MLP kernel (csrc/tc_chain.cu) against the per-layer tensor-core
engine on the shapes of one GraphNet block at BASELINE configs[1] (CUDA events, L2 flushed)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphnet_classifier_b200 import ops, build
build.build()


def timeit(fn, n=4):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return min(ts), sum(ts) / len(ts)


B = int(sys.argv[1]) if len(sys.argv) > 1 else 512          # graphs of resize 128
r = 128
N, E = B * r * r, B * 2 * r * (r - 1)
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
mk = lambda *s: torch.randn(*s, device=dev, generator=g)
layers = [(mk(128, 128) / 11, mk(128) * 0.1) for _ in range(3)]
gamma, beta = torch.ones(128, device=dev), torch.zeros(128, device=dev)
# grid topology in the reference's edge order (horizontal then vertical, per graph)
v = torch.arange(r * r, device=dev).view(r, r)
src1 = torch.cat([v[:, :-1].reshape(-1), v[:-1, :].reshape(-1)])
dst1 = torch.cat([v[:, 1:].reshape(-1), v[1:, :].reshape(-1)])
off = (torch.arange(B, device=dev) * r * r).view(B, 1)
src = (src1.view(1, -1) + off).reshape(-1).int()
dst = (dst1.view(1, -1) + off).reshape(-1).int()
e = mk(E, 128); P = mk(N, 128); Q = mk(N, 128)
out = torch.empty(E, 128, device=dev)


def per_layer_edge():
    a1 = ops.tc_linear(e, layers[0][0], bias=layers[0][1], gather0=(P, src), gather1=(Q, dst), relu=True)
    a2 = ops.tc_linear(a1, layers[1][0], bias=layers[1][1], relu=True)
    return ops.tc_linear(a2, layers[2][0], bias=layers[2][1], gamma=gamma, beta=beta, residual=e, out=out)


def chain_edge():
    return ops.tc_mlp_chain(e, layers, gather0=(P, src), gather1=(Q, dst), gamma=gamma, beta=beta, residual=e, out=out)


h = mk(N, 128); T = mk(N, 128); agg = mk(N, 128); outn = torch.empty(N, 128, device=dev)


def per_layer_node():
    a1 = ops.tc_linear(agg, layers[0][0], bias=layers[0][1], addend=T, relu=True)
    a2 = ops.tc_linear(a1, layers[1][0], bias=layers[1][1], relu=True)
    return ops.tc_linear(a2, layers[2][0], bias=layers[2][1], gamma=gamma, beta=beta, residual=h, out=outn)


g_idx = ops.GraphIndex.from_edge_index(torch.stack([src.long(), dst.long()]), N)
cases = {
    "edge MLP per-layer (3 launches)": (per_layer_edge, E, 3, 4.0 * 128 * (2 * E + 2 * N)),
    "edge MLP chained": (chain_edge, E, 3, 4.0 * 128 * (2 * E + 2 * N)),
    "edge chain, no addends": (lambda: ops.tc_mlp_chain(e, layers, gamma=gamma, beta=beta, residual=e, out=out), E, 3, 4.0 * 128 * 2 * E),
    "edge chain, no addends/LN/res": (lambda: ops.tc_mlp_chain(e, layers, out=out), E, 3, 4.0 * 128 * 2 * E),
    "edge chain 2 layers, LN+res": (lambda: ops.tc_mlp_chain(e, layers[:2], gamma=gamma, beta=beta, residual=e, out=out), E, 2, 4.0 * 128 * 2 * E),
    "node MLP per-layer (3 launches)": (per_layer_node, N, 3, 4.0 * 128 * 4 * N),
    "node MLP chained": (lambda: ops.tc_mlp_chain(agg, layers, gather0=(T, None), gamma=gamma, beta=beta, residual=h, out=outn), N, 3, 4.0 * 128 * 4 * N),
    "node MLP chained, two-operand L0": (lambda: ops.tc_mlp_chain(agg, layers, operand2=(h, layers[0][0]), gamma=gamma, beta=beta, residual=h, out=outn), N, 4, 4.0 * 128 * 3 * N),
    "node MLP, two-operand L0 + aggregating loader": (lambda: ops.tc_mlp_chain(e, layers, operand2=(h, layers[0][0]), gamma=gamma, beta=beta, residual=h, out=outn, agg=(g_idx.dst_rowptr, g_idx.dst_eid)), N, 4, 4.0 * 128 * (E + 2 * N)),
    "P,Q,T products (multi, 3 sets)": (lambda: ops.tc_linear_multi(h, [layers[0][0], layers[1][0], layers[2][0]]), N, 3, 4.0 * 128 * 4 * N),
    "P,Q products (multi, 2 sets)": (lambda: ops.tc_linear_multi(h, [layers[0][0], layers[1][0]]), N, 2, 4.0 * 128 * 3 * N),
    "encoder tail (2 layers, LN)": (lambda: ops.tc_mlp_chain(h, layers[:2], gamma=gamma, beta=beta, out=outn), N, 2, 4.0 * 128 * 2 * N),
}
only = sys.argv[2] if len(sys.argv) > 2 else None
if os.environ.get("CHAIN_NO_PREFETCH"):
    from graphnet_classifier_b200 import _lib
    _lib.load().gnc_debug_chain_trace(None, -1)
# the GPU runs under a software power cap: what ran just before shifts the clocks, so the cases are
# interleaved and repeated, and the median of the per-round best is reported
import statistics
rounds = 3
res = {k: [] for k in cases}
for _ in range(rounds):
    for name, (fn, rows, nl, nbytes) in cases.items():
        if only and only not in name:
            continue
        res[name].append(timeit(fn, n=3)[0])
for name, (fn, rows, nl, nbytes) in cases.items():
    if not res[name]:
        continue
    best = statistics.median(res[name])
    tf = 2.0 * rows * 128 * 128 * nl * 3 / best / 1e9      # executed fp16 tensor flops (3 MMAs per product)
    print(f"{name:34s} rows={rows}  median {best:8.3f} ms  (min {min(res[name]):7.3f})   {nbytes/best/1e6:7.0f} GB/s algorithmic   {tf:7.1f} fp16-TFLOP/s executed")
