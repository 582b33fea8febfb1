#!/usr/bin/env python
"""BASELINE config 5 at several GPUs: the E = 100M rows of the aggregation sweep that do not fit one GPU
(D = 256, 512), sharded by destination range - rank k owns a contiguous node range and the edges that target it, so
the forward sum needs no exchange.  Launch with torchrun; device-timed, max over ranks.
Bytes per SURVEY.md 8(d): 4*(E*D + E + (N+1) + N*D) per rank."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from graphnet_classifier_b200 import build, ops
from graphnet_classifier_b200.ops import GraphIndex
from graphnet_classifier_b200.utils.image_to_graph.batched import build_pixel_graphs

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
if rank == 0:
    build.build()
if world > 1:
    dist.barrier()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAK = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(root, "MEASURED_PEAKS.json")) else 6547.8
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
E_total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t.item()))
    return sum(ts) / len(ts)


def shard_graph(kind, E):
    if kind == "grid":
        B = max(1, round(E / (2 * 128 * 127)))                      # whole images per rank: node range = its images
        return build_pixel_graphs(torch.zeros(B, 128, 128, 3, dtype=torch.uint8, device=dev), use_cache=False).graph
    N = E // 2
    g = torch.Generator(device=dev).manual_seed(rank)
    ei = torch.stack([torch.randint(0, N, (E,), device=dev, generator=g), torch.randint(0, N, (E,), device=dev, generator=g)])
    return GraphIndex.from_edge_index(ei, N, validate=False)           # destinations local to the rank's node range


if rank == 0:
    print(f"| topology | GPUs | E total | E per GPU | D | ms (max over ranks) | aggregate GB/s | frac of {world} x measured HBM peak | torch index_add_ ms (max over ranks) | speed-up | mass conservation |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
for kind in ("grid", "random"):
    g = shard_graph(kind, E_total // world)
    E, N = g.num_edges, g.num_nodes
    for D in (128, 256, 512):
        if 4.0 * D * (E + N) > 140e9:                             # messages + output must fit one GPU's 180 GB
            if rank == 0:
                print(f"| {kind} | {world} | {E * world} | {E} | {D} | - | - | - | skipped: {E * D * 4 / 1e9:.0f} GB per GPU |")
            continue
        src = torch.empty(E, D, device=dev)
        for lo in range(0, E, 1 << 22):                              # filled in pieces: randn's temporaries stay small
            src[lo:lo + (1 << 22)].normal_()
        out = torch.empty(N, D, device=dev)
        ms = timeit(lambda: ops._agg_raw(g.dst_rowptr, g.dst_eid, src, N, out=out))
        # size-independent property: every message lands in exactly one row (column sums agree in fp64)
        a = torch.zeros(D, dtype=torch.float64, device=dev)
        b = torch.zeros(D, dtype=torch.float64, device=dev)
        for lo in range(0, N, 1 << 22):
            a += out[lo:lo + (1 << 22)].sum(0, dtype=torch.float64)
        for lo in range(0, E, 1 << 22):
            b += src[lo:lo + (1 << 22)].sum(0, dtype=torch.float64)
        ok = torch.tensor([float(((a - b).abs().max() / (b.abs().max() + 1e-30)) < 1e-6)], device=dev)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        nbytes = 4.0 * (E * D + E + (N + 1) + N * D)
        idx = g.dst.long()
        ms_t = timeit(lambda: out.zero_().index_add_(0, idx, src), n=3)      # the baseline the sweep is quoted against
        del idx
        if rank == 0:
            gbs = world * nbytes / ms / 1e6
            print(f"| {kind} | {world} | {E * world} | {E} | {D} | {ms:.3f} | {gbs:.0f} | {gbs / (world * PEAK):.3f} | {ms_t:.3f} | {ms_t / ms:.1f}x | {bool(ok.item())} |", flush=True)
        del src, out
        torch.cuda.empty_cache()
    del g
    torch.cuda.empty_cache()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
