"""``OptimizedDatasetLoader`` with the reference's constructor and item contract
(reference utils/dataloader.py:10-53): ``((x, pos, edge_index), label)`` per index,
float32 / float32 / int64 / int64 - built on the device, so the tensors are CUDA
tensors and ``edge_index`` carries its prebuilt CSR for the model's kernels.
"""
from __future__ import annotations

import numpy as np
import torch
from torch.utils.data import Dataset

from .image_to_graph.batched import build_patch_graphs, build_pixel_graphs
from .image_to_graph.image_to_graph_optimized import load_rgb_device
from .image_to_graph.image_to_graph_superpixel import image_to_graph_superpixel


class OptimizedDatasetLoader(Dataset):
    def __init__(self, dataset_path="dataset", resize_value=128, diagonals=False,
                 method="pixel", n_segments=100, patch_size=8, use_cache=True, dataset=None):
        self.dataset_path = dataset_path
        if dataset is None:
            import torchvision.datasets as datasets       # ImageFolder of PIL images, no transform
            dataset = datasets.ImageFolder(self.dataset_path)
        self.dataset = dataset
        self.resize_value = resize_value
        self.diagonals = diagonals
        self.method = method
        self.n_segments = n_segments
        self.patch_size = patch_size
        self.use_cache = use_cache
        print(f"Using {method} method with resize_value={resize_value}")
        if method == "pixel":
            print(f"Graph size: {resize_value*resize_value} nodes")
        elif method == "superpixel":
            print(f"Target superpixels: {n_segments}")
        elif method == "patch":
            print(f"Patch size: {patch_size}, patches: {(resize_value//patch_size)**2}")

    def __len__(self):
        return len(self.dataset)

    def _pixels(self, image) -> torch.Tensor:
        return load_rgb_device(image, self.resize_value)               # PIL decode on the host, Pillow-exact resize on the device

    def __getitem__(self, idx):
        image, label = self.dataset[idx]
        if self.method == "pixel":
            gb = build_pixel_graphs(self._pixels(image), diagonals=self.diagonals, use_cache=self.use_cache)
            sample = gb.as_tuple()
        elif self.method == "patch":
            gb = build_patch_graphs(self._pixels(image), patch_size=self.patch_size, use_cache=self.use_cache)
            sample = gb.as_tuple()
        elif self.method == "superpixel":
            x, pos, ei = image_to_graph_superpixel(image, self.resize_value, self.n_segments)
            dev = torch.device("cuda", torch.cuda.current_device())
            sample = (torch.tensor(x, dtype=torch.float32, device=dev),
                      torch.tensor(pos, dtype=torch.float32, device=dev),
                      torch.tensor(ei, dtype=torch.long, device=dev))
        else:
            raise ValueError(f"Unknown method: {self.method}")
        return sample, torch.tensor(label, dtype=torch.long, device=sample[0].device)
