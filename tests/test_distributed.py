"""CPU, world_size 2 over gloo: the data-parallel plumbing (graph sharding + flat
gradient bucket + all-reduce) reproduces single-process gradients on the whole batch."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graphnet_classifier_b200.utils.distributed import GradBucket, broadcast_parameters, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model(seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 2))


def _worker(rank, world, port, n_items, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(123)
        X, y = torch.randn(n_items, 5), torch.randint(0, 2, (n_items,))
        model = _model(seed=rank)                 # ranks start different; broadcast fixes that
        broadcast_parameters(model, src=0)
        bucket = GradBucket(model.parameters())
        lo, hi = shard_range(n_items, rank, world)
        bucket.zero()
        # every rank scales by the GLOBAL batch, then the bucket is summed (average=False)
        loss = torch.nn.functional.cross_entropy(model(X[lo:hi]), y[lo:hi], reduction="sum") / n_items
        loss.backward()
        bucket.all_reduce(average=False)
        torch.save({"flat": bucket.flat.clone(), "params": [p.detach().clone() for p in model.parameters()]},
                   os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_two_rank_allreduce_equals_single_process(tmp_path):
    n_items, world = 11, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_items, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (torch.load(tmp_path / f"rank{r}.pt") for r in range(world))
    assert torch.equal(r0["flat"], r1["flat"])                         # identical after the all-reduce
    for a, b in zip(r0["params"], r1["params"]):
        assert torch.equal(a, b)                                       # broadcast worked
    torch.manual_seed(123)
    X, y = torch.randn(n_items, 5), torch.randint(0, 2, (n_items,))
    model = _model(seed=0)
    torch.nn.functional.cross_entropy(model(X), y, reduction="mean").backward()
    ref = torch.cat([p.grad.flatten() for p in model.parameters()])
    assert torch.allclose(r0["flat"], ref, rtol=1e-5, atol=1e-7)
