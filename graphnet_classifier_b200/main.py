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


if __name__ == "__main__":
    print("start")
    train_GNN(epochs=100, resize_value=128, output_path="weights/GNN/dim128_3block")
