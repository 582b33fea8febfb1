#!/usr/bin/env python
"""BASELINE config 5: aggregation micro-benchmark sweep - D in {32..512}, E in {1M, 10M, 100M}
(where E*D*4 fits), topology (i) batched pixel grid in reference edge order and (ii) uniform
random destinations with N = E/2 - gnc_agg_csr_sum_f32 vs torch index_add_ on the same GPU.
Prints a markdown table; bytes per SURVEY.md 8(d): 4*(E*D + E + (N+1) + N*D)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphnet_classifier_b200 import build, ops
from graphnet_classifier_b200.ops import GraphIndex
from graphnet_classifier_b200.utils.image_to_graph.batched import build_pixel_graphs
build.build()
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists("MEASURED_PEAKS.json") else 6547.8
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return sum(ts) / len(ts)

def graph(kind, E):
    if kind == "grid":
        per = 2 * 128 * 127
        B = max(1, round(E / per))
        return build_pixel_graphs(torch.zeros(B, 128, 128, 3, dtype=torch.uint8, device=dev), use_cache=False).graph
    N = E // 2
    g = torch.Generator(device=dev).manual_seed(0)
    ei = torch.stack([torch.randint(0, N, (E,), device=dev, generator=g), torch.randint(0, N, (E,), device=dev, generator=g)])
    return GraphIndex.from_edge_index(ei, N, validate=False)

print("| topology | E | N | D | ours ms | ours GB/s | frac of measured HBM peak | torch index_add_ ms | speed-up | bit-exact vs index_add_ |")
print("|---|---|---|---|---|---|---|---|---|---|")
for kind in ("grid", "random"):
    for E_t in (1_000_000, 10_000_000, 100_000_000):
        g = graph(kind, E_t)
        E, N = g.num_edges, g.num_nodes
        for D in (32, 64, 128, 256, 512):
            if E * D * 4 > 60e9:
                print(f"| {kind} | {E} | {N} | {D} | - | - | - | - | - | skipped: E*D*4 = {E*D*4/1e9:.0f} GB does not fit one GPU with the baseline's buffers |")
                continue
            src = torch.randn(E, D, device=dev)
            ms = timeit(lambda: ops.aggregate(src, g))
            nbytes = 4.0 * (E * D + E + (N + 1) + N * D)
            idx = g.dst.long()
            out = torch.zeros(N, D, device=dev)
            ms_t = timeit(lambda: out.zero_().index_add_(0, idx, src), n=3)
            exact = "n/a (atomics: order varies)"
            if E <= 12_000_000 and D <= 128:
                ref = torch.zeros(N, D).index_add_(0, idx.cpu(), src.cpu())
                exact = str(bool(torch.equal(ops.aggregate(src, g).cpu(), ref)))
            print(f"| {kind} | {E} | {N} | {D} | {ms:.3f} | {nbytes/ms/1e6:.0f} | {nbytes/ms/1e6/PEAK:.3f} | {ms_t:.3f} | {ms_t/ms:.1f}x | {exact} |", flush=True)
            del src, out
        del g
        torch.cuda.empty_cache()
