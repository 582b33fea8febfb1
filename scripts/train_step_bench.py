#!/usr/bin/env python
"""A/B timing of one training step (graph build + forward + CE + backward + Adam, BASELINE configs[2] per-GPU shape)
under the backward schedules of tc_train.py; prints ms per step, launches per step and the per-kernel shares.
    python scripts/train_step_bench.py [graphs] [modes...]      modes: fused (chained forward with stash), fused-layer (per-layer
    forward), pair (round-1 kernel pair, per-layer forward)"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from graphnet_classifier_b200 import ops, build, tc_train, _lib
from graphnet_classifier_b200.models.GNN import CombinedModel, GraphNet
from graphnet_classifier_b200.pipeline import GraphClassifierPipeline
build.build()

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
modes = sys.argv[2:] or ["fused", "fused-layer", "pair"]
r = 128
torch.manual_seed(0)
model = CombinedModel(GraphNet(num_local_features=3, space_dim=2, out_channels=1, n_blocks=3), num_nodes=r * r, classes=2).cuda()
pipe = GraphClassifierPipeline(model, resize_value=r, train_micro_batch=int(os.environ.get("TRAIN_MB", "0")) or None)
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
rng = np.random.default_rng(0)
img = torch.from_numpy(rng.integers(0, 256, (B, r, r, 3), dtype=np.uint8)).cuda()
lab = torch.from_numpy(rng.integers(0, 2, B)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for mode in modes:
    tc_train.BWD = mode.split("-")[0]
    tc_train.FWD = "layer" if (mode.endswith("-layer") or mode == "pair") else "chain"
    for _ in range(2):
        pipe.train_step(img, lab, opt)
    torch.cuda.synchronize()
    ts = []
    for rep in range(3):
        flush.zero_()
        _lib.reset_launch_count()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); loss = pipe.train_step(img, lab, opt); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    launches = _lib.launch_count()
    ops.PROFILE = ops.KernelProfile()
    pipe.train_step(img, lab, opt)
    prof = ops.PROFILE.summary()
    ops.PROFILE = None
    tot = sum(v["ms"] for v in prof.values())
    print(f"mode {mode}: {sorted(ts)[1]:.1f} ms per step of {B} graphs ({B / sorted(ts)[1] * 1e3:.0f} graphs/s), micro-batch "
          f"{pipe.train_micro_batch}, {launches} libgnc launches, loss {float(loss):.6f}, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
    for name, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:10]:
        print(f"    {name:24s} {v['ms']:8.2f} ms {v['calls']:4d} calls  {v['bytes'] / max(v['ms'], 1e-9) / 1e6:7.0f} GB/s algorithmic")
    print(f"    (profiled kernel time {tot:.1f} ms)")
