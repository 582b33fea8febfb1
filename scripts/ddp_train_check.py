#!/usr/bin/env python
"""torchrun check of main.train_GNN_batched over NCCL: every rank trains on its shard of each global batch, gradients meet
in one all-reduce, all ranks end with identical parameters, and the first epoch's mean loss / the parameters after the
first step equal the oracle's single-process step on the whole batch (tests/ is single-GPU; this runs by hand:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/ddp_train_check.py)."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
from PIL import Image
from graphnet_classifier_b200 import build
rank = int(os.environ.get("RANK", 0))
if rank == 0:
    build.build()
from graphnet_classifier_b200.main import train_GNN_batched
from oracle import gnn as ognn, graph_build as ogb          # checker only

r, B = 16, 8
rng = np.random.default_rng(3)
photos = [rng.integers(0, 256, (24 + i, 40 - i, 3), dtype=np.uint8) for i in range(B)]
labels = [int(v) for v in rng.integers(0, 2, B)]
data = [(Image.fromarray(p), l) for p, l in zip(photos, labels)]
torch.manual_seed(0)
with tempfile.TemporaryDirectory() as d:
    best, model = train_GNN_batched(epochs=1, resize_value=r, batch_size=B, output_path=d, dataset=data, shuffle=False)
world = dist.get_world_size() if dist.is_initialized() else 1
flat = torch.cat([p.detach().flatten() for p in model.parameters()])
if world > 1:
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
else:
    same = True
if rank == 0:
    om = ognn.build_reference_config_model(r, seed=0)
    opt = torch.optim.Adam(om.parameters(), lr=1e-3)
    per = [torch.nn.functional.cross_entropy(
        om(ogb.to_model_inputs(*ogb.pixel_graph(np.asarray(Image.fromarray(p).resize((r, r)))))), torch.tensor(l))
        for p, l in zip(photos, labels)]
    loss = sum(per) / B
    local = sum(per[:B // world]).item() / (B // world)       # train() logs the rank's own shard loss (rank 0: first shard)
    opt.zero_grad(); loss.backward(); opt.step()
    ref = torch.cat([p.detach().flatten() for p in om.parameters()])
    dmax = float((flat.cpu() - ref).abs().max())
    print(f"world_size {world}: parameters identical on all ranks: {same}; rank-0 shard loss {best:.7f} vs oracle {local:.7f} "
          f"(diff {abs(best - local):.1e}); max |param - oracle param| after the Adam step on the global batch {dmax:.2e} (lr = 1e-3)",
          flush=True)
    assert same and abs(best - local) < 1e-5 * max(1.0, local) and dmax < 2e-3 * 1.01
if dist.is_initialized():
    dist.barrier()
    dist.destroy_process_group()
