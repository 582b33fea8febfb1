"""Entry points with the reference's names (reference main.py:13-76): ``train_GNN``
wires dataset -> loader -> GraphNet(3 blocks) -> CombinedModel -> ``train``; ``load_data`` / ``train_MLP`` are
the MLP baseline (main.py:13-29) on the same operators.

``train_GNN`` keeps the reference's one-graph-per-step semantics (DataLoader with
batch_size=1).  ``train_GNN_batched`` is the data-parallel form: block-diagonal
batches per rank, gradient all-reduce over NCCL when launched under torchrun.
"""
from __future__ import annotations

import torch
from torch.utils.data import DataLoader

from .utils.train_model import train


def num_nodes_for(method: str, resize_value: int) -> int:
    """The reference's rule (main.py:63-70), reproduced including the superpixel guess."""
    if method == "pixel":
        return resize_value * resize_value
    if method == "superpixel":
        return resize_value // 2
    if method == "patch":
        return (resize_value // 8) ** 2
    return resize_value * resize_value


def load_data(dataset_path, resize_value=128, batch_size=8):
    """ImageFolder -> Resize -> ToTensor batches, shuffled (reference main.py:13-18; host side, torchvision)."""
    import torchvision.datasets as datasets
    from torchvision import transforms
    transform = transforms.Compose([transforms.Resize((resize_value, resize_value)), transforms.ToTensor()])
    dataset = datasets.ImageFolder(root=dataset_path, transform=transform)
    return DataLoader(dataset, batch_size=batch_size, shuffle=True)


def train_MLP(epochs=30, channels=3, resize_value=128, batch_size=8, hidden_layers=2, output_path="weights/MLP",
              dataset_path="dataset"):
    """The MLP baseline (reference main.py:21-29): flattened pixels -> MLP(hidden 128, LayerNorm on the logits).
    The batches come from the host loader; ``train`` moves them to the model's device."""
    from .models.MLP import MLP
    input_dim = channels * resize_value * resize_value
    dataset = load_data(dataset_path, resize_value, batch_size)
    num_classes = len(dataset.dataset.classes)
    model = MLP(in_dim=input_dim, out_dim=num_classes, hidden_layers=hidden_layers).cuda()
    return train(model, dataset, epochs, patience=5, output_path=output_path)


def train_GNN(epochs=30, channels=3, resize_value=64, batch_size=8, hidden_layers=2, max_samples=None,
              method="pixel", use_cache=True, output_path="weights/GNN", dataset_path="dataset"):
    from .models.GNN import CombinedModel, GraphNet
    from .utils.dataloader import OptimizedDatasetLoader

    original_dataset = OptimizedDatasetLoader(dataset_path=dataset_path, resize_value=resize_value,
                                              method=method, use_cache=use_cache)
    num_classes = len(original_dataset.dataset.classes)
    if max_samples and max_samples < len(original_dataset):
        import random
        from torch.utils.data import Subset
        random.seed(42)
        indices = random.sample(range(len(original_dataset)), max_samples)
        dataset = Subset(original_dataset, indices)
        print(f"Using subset of {max_samples} samples for faster training")
    else:
        dataset = original_dataset
    dataloader = DataLoader(dataset, batch_size=1, shuffle=True, collate_fn=lambda batch: batch[0])
    num_nodes = num_nodes_for(method, resize_value)
    graph_net = GraphNet(num_local_features=3, space_dim=2, out_channels=1, n_blocks=3)
    model = CombinedModel(graph_net=graph_net, num_nodes=num_nodes, classes=num_classes).cuda()
    print(f"Training GNN with {method} method, {num_nodes} nodes")
    return train(model, dataloader, epochs, patience=5, output_path=output_path)


class BatchedShardLoader:
    """This rank's view of a dataset of ``(PIL image, label)`` items for data-parallel training: every epoch the indices
    are shuffled with a seed all ranks share, cut into global batches of ``batch_size`` graphs (truncated to a multiple
    of ``world_size`` so that the mean over equal shards is the global mean), and each rank takes its ``shard_range``
    of every batch.  Iterating yields ``((x, pos, edge_index), labels)`` with the shard as ONE block-diagonal graph
    built on the device (decode on the host, Pillow-exact resize + graph build on the GPU) - what ``train`` consumes."""

    def __init__(self, dataset, batch_size, resize_value, method="pixel", diagonals=False, patch_size=8, rank=0,
                 world_size=1, seed=0, shuffle=True):
        if method not in ("pixel", "patch"):
            raise ValueError(f"batched training needs a fixed node count per graph (pixel / patch), got method={method!r}")
        if batch_size < world_size:
            raise ValueError("batch_size is the GLOBAL batch: it must hold at least one graph per rank")
        self.dataset, self.batch_size, self.resize_value = dataset, int(batch_size), int(resize_value)
        self.method, self.diagonals, self.patch_size = method, bool(diagonals), int(patch_size)
        self.rank, self.world_size, self.seed, self.shuffle = int(rank), int(world_size), int(seed), bool(shuffle)
        self.epoch = 0
        self._pool = None               # utils.staging.DecodePool, created on first iteration (needs the CUDA device)

    def plan(self, epoch):
        """Index lists of this rank's shards for ``epoch`` (pure host logic)."""
        import random
        from .utils.distributed import shard_range
        order = list(range(len(self.dataset)))
        if self.shuffle:
            random.Random(self.seed * 1_000_003 + epoch).shuffle(order)
        shards = []
        for lo in range(0, len(order), self.batch_size):
            batch = order[lo:lo + self.batch_size]
            batch = batch[:len(batch) - len(batch) % self.world_size]
            if not batch:
                continue
            a, b = shard_range(len(batch), self.rank, self.world_size)
            shards.append(batch[a:b])
        return shards

    def __len__(self):
        return len(self.plan(0))

    def __iter__(self):
        from .utils.image_to_graph.batched import build_patch_graphs, build_pixel_graphs
        from .utils.staging import DecodePool
        shards = self.plan(self.epoch)
        self.epoch += 1
        if self._pool is None:
            self._pool = DecodePool()
        for idxs in shards:
            # dataset[i] opens and decodes the file (ImageFolder's loader): on the pool's threads, not one by one here
            items = list(self._pool.pool.map(self.dataset.__getitem__, idxs))
            pixels = self._pool.stage([img for img, _ in items], self.resize_value)
            if self.method == "pixel":
                gb = build_pixel_graphs(pixels, diagonals=self.diagonals)
            else:
                gb = build_patch_graphs(pixels, patch_size=self.patch_size)
            yield gb.as_tuple(), torch.tensor([int(lab) for _, lab in items], dtype=torch.long, device=pixels.device)


def train_GNN_batched(epochs=30, resize_value=64, batch_size=64, max_samples=None, method="pixel",
                      output_path="weights/GNN", dataset_path="dataset", dataset=None, seed=0, shuffle=True):
    """Data-parallel, batched form of ``train_GNN``: one optimizer step per GLOBAL batch of ``batch_size`` graphs.  Under
    ``torchrun`` (one process per GPU) every rank builds and processes its shard of each batch as a block-diagonal graph
    and the gradients meet in ONE all-reduce of the flat bucket (NCCL over NVLink); stand-alone it is the single-GPU
    batched loop.  Same model, loss, Adam(1e-3), checkpoints and log as the reference's ``train``."""
    import os
    import torch.distributed as dist
    from .models.GNN import CombinedModel, GraphNet
    from .utils.distributed import GradBucket, broadcast_parameters
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank() if dist.is_initialized() else 0
    if dataset is None:
        import torchvision.datasets as datasets
        dataset = datasets.ImageFolder(dataset_path)
    num_classes = len(dataset.classes) if hasattr(dataset, "classes") else 2
    if max_samples and max_samples < len(dataset):
        import random
        from torch.utils.data import Subset
        random.seed(42)
        dataset = Subset(dataset, random.sample(range(len(dataset)), max_samples))
    num_nodes = num_nodes_for(method, resize_value)
    graph_net = GraphNet(num_local_features=3, space_dim=2, out_channels=1, n_blocks=3)
    model = CombinedModel(graph_net=graph_net, num_nodes=num_nodes, classes=num_classes).cuda()
    broadcast_parameters(model)
    bucket = GradBucket(model.parameters())
    loader = BatchedShardLoader(dataset, batch_size, resize_value, method=method, rank=rank, world_size=world, seed=seed,
                                shuffle=shuffle)
    if rank == 0:
        print(f"Training GNN with {method} method, {num_nodes} nodes, global batch {batch_size} over {world} GPU(s)")
    return train(model, loader, epochs, patience=5, output_path=output_path, grad_sync=bucket.all_reduce), model


if __name__ == "__main__":
    print("start")
    train_GNN(epochs=100, resize_value=128, output_path="weights/GNN/dim128_3block")
