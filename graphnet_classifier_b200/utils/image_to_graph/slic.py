"""Device SLIC (csrc/slic.cu): the segmentation stage of the superpixel builder.

The reference calls ``skimage.segmentation.slic(img, n_segments, compactness, start_label=0)``
(reference image_to_graph_superpixel.py:31).  scikit-image is un-vendored, unpinned and absent
here, so label parity is UNPINNED; this follows the published SLIC algorithm with
scikit-image's conventions (Lab colour space, colour / compactness, 10 iterations).
"""
from __future__ import annotations

import torch
from torch import Tensor

from ... import _lib
from ..._lib import check
from ...ops import _require_cuda, _stream, _workspace


def slic_labels(images: Tensor, n_segments: int = 100, compactness: float = 10.0, max_num_iter: int = 10) -> Tensor:
    """uint8 ``[B, H, W, 3]`` (or ``[H, W, 3]``) on the device -> int32 labels ``[B, H, W]``."""
    img = images if images.dim() == 4 else images.unsqueeze(0)
    _require_cuda(img)
    if img.dtype != torch.uint8 or img.shape[-1] != 3:
        raise TypeError("slic_labels expects uint8 [B, H, W, 3]")
    img = img.contiguous()
    B, H, W, _ = img.shape
    lib = _lib.load()
    labels = torch.empty(B, H, W, dtype=torch.int32, device=img.device)
    nbytes = int(lib.gnc_slic_workspace_bytes(B, H, W, int(n_segments)))
    work = _workspace(img.device, (nbytes + 7) // 8, torch.int64)
    check(lib.gnc_slic_labels_u8(img.data_ptr(), B, H, W, int(n_segments), float(compactness), int(max_num_iter),
                                 labels.data_ptr(), work.data_ptr(), _stream()), "slic_labels")
    return labels
