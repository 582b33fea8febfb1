"""CPU: host-side logic that needs no kernel - the train loop's file/log contract
(reference utils/train_model.py), sharding, micro-batch arithmetic, num_nodes rule."""
import os

import torch

from graphnet_classifier_b200.main import num_nodes_for
from graphnet_classifier_b200.pipeline import GraphClassifierPipeline
from graphnet_classifier_b200.utils.distributed import GradBucket, shard_range
from graphnet_classifier_b200.utils.train_model import train


class _Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.fc = torch.nn.Linear(6, 2)

    def forward(self, sample):
        return self.fc(sample)


def test_train_loop_files_log_and_early_stopping(tmp_path):
    torch.manual_seed(0)
    data = [(torch.randn(6), torch.tensor(i % 2)) for i in range(6)]
    model = _Tiny()
    best = train(model, data, epochs=3, patience=1, output_path=str(tmp_path / "w" / "run"))
    out = tmp_path / "w" / "run"
    names = sorted(os.listdir(out))
    assert "final_model.pth" in names and "best_model_epoch1.pth" in names
    logs = [n for n in names if n.startswith("training_logs_") and n.endswith(".txt")]
    assert len(logs) == 1
    lines = open(out / logs[0]).read().splitlines()
    # header (utils/train_model.py:27-30), per-epoch pairs (:53-54), footer (:78-81)
    assert lines[0].startswith("Training started at: ") and lines[1] == "Epochs: 3, Patience: 1"
    assert lines[2].startswith("Output path: ") and lines[3] == "-" * 50
    assert lines[4].startswith("Epoch 1/3, avg_loss=") and lines[5].startswith("Epoch 1/3, needed ") and lines[5].endswith(" minutes")
    assert lines[-3].startswith("Training completed at: ") and lines[-2].startswith("Best loss achieved: ")
    assert lines[-1].startswith("Final model saved: ")
    assert best < 10
    sd = torch.load(out / "final_model.pth")
    assert set(sd) == {"fc.weight", "fc.bias"}
    # resume from a checkpoint (start_weights, :14-15)
    m2 = _Tiny()
    with torch.no_grad():
        m2.fc.weight.zero_()
    train(m2, data[:1], epochs=1, output_path=str(tmp_path / "w2"), start_weights=str(out / "final_model.pth"))


def test_early_stopping_uses_training_loss(tmp_path):
    # constant-loss model: epoch 1 sets the best, epochs 2.. do not improve -> stops after `patience` more
    class Const(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(2))

        def forward(self, s):
            return self.p * 0 + torch.tensor([0.3, 0.3])

    data = [(torch.zeros(1), torch.tensor(0))] * 2
    train(Const(), data, epochs=10, patience=2, output_path=str(tmp_path))
    text = open([tmp_path / n for n in os.listdir(tmp_path) if n.startswith("training_logs")][0]).read()
    assert "Epoch 3/10" in text and "Epoch 4/10" not in text


def test_grad_sync_hook_runs_between_backward_and_step(tmp_path):
    calls = []
    model = _Tiny()

    def hook():
        calls.append(float(model.fc.weight.grad.abs().sum()))

    train(model, [(torch.randn(6), torch.tensor(1))] * 3, epochs=1, output_path=str(tmp_path), grad_sync=hook)
    assert len(calls) == 3 and all(c > 0 for c in calls)


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 512, 4096, 4099):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_num_nodes_rule_matches_reference_main():
    assert num_nodes_for("pixel", 64) == 4096
    assert num_nodes_for("superpixel", 128) == 64          # reference's guess (Q7), reproduced
    assert num_nodes_for("patch", 128) == 256
    assert num_nodes_for("other", 16) == 256


def test_even_chunking():
    f = GraphClassifierPipeline._even_chunk
    assert f(512, 512) == 512 and f(512, 104) == 103 and f(7, 3) == 3 and f(1, 64) == 1
    for total in (1, 5, 512, 1000):
        for lim in (1, 3, 104, 2048):
            c = f(total, lim)
            assert 1 <= c <= max(lim, 1) and -(-total // c) == -(-total // min(lim, total))


def test_grad_bucket_views_and_rebind():
    m = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    b = GradBucket(m.parameters())
    assert b.flat.numel() == sum(p.numel() for p in m.parameters())
    m(torch.randn(5, 4)).sum().backward()
    flat_copy = torch.cat([p.grad.flatten() for p in m.parameters()])
    assert torch.equal(flat_copy, b.flat) and float(b.flat.abs().sum()) > 0
    b.zero()
    assert all(float(p.grad.abs().sum()) == 0 for p in m.parameters())
    for p in m.parameters():
        p.grad = None
    m(torch.randn(5, 4)).sum().backward()
    b.rebind()
    assert torch.equal(torch.cat([p.grad.flatten() for p in m.parameters()]), b.flat)
    b.all_reduce()       # no process group: no-op


def test_core_param_list_layout_and_graphed_step_layout():
    """Host logic of the training schedule: the parameter order the hand-scheduled backward indexes into
    (tc_train.GraphNetCoreFn: 12 + 16 * n_blocks + 6 tensors) and the eligibility rule of the captured step."""
    from graphnet_classifier_b200 import tc_train
    from graphnet_classifier_b200.models.GNN import GraphNet
    from graphnet_classifier_b200.utils.train_model import _GraphedStep
    for nb in (1, 3):
        gn = GraphNet(num_local_features=3, space_dim=2, out_channels=1, n_blocks=nb)
        ps = tc_train.core_param_list(gn)
        assert len(ps) == 12 + 16 * nb + 6
        for k in range(nb):
            pe = 12 + 16 * k
            assert tuple(ps[pe].shape) == (128, 384) and ps[pe] is gn.graph_processor.blocks[k].edge_model.edge_processor.model[0].weight
            assert tuple(ps[pe + 8].shape) == (128, 256) and ps[pe + 8] is gn.graph_processor.blocks[k].node_model.node_processor.model[0].weight
            assert ps[pe + 6] is gn.graph_processor.blocks[k].edge_model.edge_processor.model[5].weight      # LayerNorm gamma
        assert ps[-6] is gn.node_decoder.model[0].weight and ps[-3] is gn.node_decoder.model[2].bias
        assert ps[-2] is gn.node_decoder.model[4].weight and ps[-1] is gn.node_decoder.model[4].bias
        assert len({id(p) for p in ps}) == len(ps)
        # everything except the encoders' first layers (the decoder's Linear(128, 1) is part of the core)
        rest = {id(p) for p in gn.parameters()} - {id(p) for p in ps}
        assert len(rest) == 4
    # host tensors, or index tensors without an attached topology, never take the captured path
    x, pos, ei = torch.zeros(4, 3), torch.zeros(4, 2), torch.zeros(2, 3, dtype=torch.long)
    assert _GraphedStep.layout((x, pos, ei), torch.tensor(1)) is None
    assert _GraphedStep.layout(x, 1) is None


def test_batched_shard_loader_plan():
    """main.BatchedShardLoader: all ranks shuffle alike, shards of a global batch are disjoint, equal-sized and cover it
    (truncated to a multiple of the world size), epochs reshuffle."""
    import pytest
    from graphnet_classifier_b200.main import BatchedShardLoader
    data = list(range(23))
    world = 4
    loaders = [BatchedShardLoader(data, batch_size=10, resize_value=8, rank=r, world_size=world, seed=5) for r in range(world)]
    for epoch in (0, 1):
        plans = [ld.plan(epoch) for ld in loaders]
        assert all(len(p) == 2 for p in plans)                 # batches of 10, 10, (3 -> truncated to 0)
        seen = []
        for b in range(2):
            shards = [p[b] for p in plans]
            assert [len(s) for s in shards] == [2, 2, 2, 2]   # 10 -> 8 graphs, 2 per rank
            flat = [i for s in shards for i in s]
            assert len(set(flat)) == len(flat)
            seen += flat
        assert len(set(seen)) == 16
    assert loaders[0].plan(0) != loaders[0].plan(1)
    one = BatchedShardLoader(data, batch_size=10, resize_value=8, shuffle=False)
    assert one.plan(0) == [list(range(0, 10)), list(range(10, 20)), [20, 21, 22]] and len(one) == 3
    with pytest.raises(ValueError):
        BatchedShardLoader(data, 10, 8, method="superpixel")
    with pytest.raises(ValueError):
        BatchedShardLoader(data, 2, 8, world_size=4)
