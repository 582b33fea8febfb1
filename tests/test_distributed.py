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


class _ShardLossModel(torch.nn.Module):
    """A model whose loss does not depend on its parameter: the per-shard epoch losses are scripted by the data."""

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(2))

    def forward(self, sample):
        return sample + 0.0 * self.w


def _train_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from graphnet_classifier_b200.utils.train_model import train

        class Data:
            """One item per epoch; rank 0's shard loss keeps falling, rank 1's rises: on their OWN losses rank 1 would
            stop after ``patience`` epochs while rank 0 went on - and block in the next all-reduce."""
            def __init__(self):
                self.epoch = 0

            def __iter__(self):
                e = self.epoch
                self.epoch += 1
                margin = (4.0 - 0.5 * e) if rank == 0 else (-1.0 + 0.7 * e)     # loss = softplus(margin)
                yield torch.tensor([[0.0, margin]]), torch.tensor([0])

        model = _ShardLossModel()
        syncs = []

        def grad_sync():
            t = torch.zeros(1)
            dist.all_reduce(t)                     # a rank that left the loop early would hang its peers here
            syncs.append(1)

        best = train(model, Data(), epochs=12, patience=2, output_path=os.path.join(out_dir, "w"), grad_sync=grad_sync,
                     cuda_graph=False)
        torch.save({"best": best, "epochs": len(syncs)}, os.path.join(out_dir, f"train_rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_two_rank_train_takes_global_early_stopping_decision(tmp_path):
    """best-model / early-stopping decisions come from the loss averaged over all ranks, so every rank runs
    the same number of epochs (reference utils/train_model.py:57-69 semantics on the global loss)."""
    world = 2
    mp.spawn(_train_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (torch.load(tmp_path / f"train_rank{r}.pt") for r in range(world))
    assert r0["epochs"] == r1["epochs"]
    assert r0["best"] == r1["best"]
    assert r0["epochs"] < 12                       # the rising global loss did stop the run early
    import glob
    assert len(glob.glob(str(tmp_path / "w" / "training_logs_*.txt"))) == 1      # rank 0 alone writes files
