"""Baseline JPEG decode on the device (csrc/jpeg.cu): the file decode behind the reference's
``Image.open(path).convert('RGB')`` (utils/dataloader.py:34 through ImageFolder's loader,
utils/image_to_graph/image_to_graph_optimized.py:65-68, utils/inference.py:47), bit for bit Pillow's pixels.

    infos = [parse(data) for data in files]            # host: markers -> descriptor, None = not baseline JPEG
    images = decode_batch(files, device)               # list of uint8 [H, W, 3] device tensors (None where unsupported)

Only the compressed bytes cross PCIe (~14 x fewer than decoded pixels).  Files the device decoder does not cover
(progressive, CMYK, PNG, ...) come back as ``None``; ``utils/staging.DecodePool`` decodes those with Pillow on the host.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import numpy as np
import torch
from torch import Tensor

from .. import _lib, ops
from .._lib import GncJpegImage, check


def parse(data: bytes) -> Optional[GncJpegImage]:
    """Descriptor of a baseline JPEG file, or None when the device decoder does not cover the file."""
    info = GncJpegImage()
    rc = _lib.load().gnc_jpeg_parse(data, len(data), ctypes.byref(info))
    return info if rc == _lib.GNC_OK else None


def decode_batch(datas: Sequence[bytes], device=None, infos: Optional[Sequence[Optional[GncJpegImage]]] = None,
                 staging: Optional[dict] = None) -> List[Optional[Tensor]]:
    """JPEG file contents -> ``uint8 [H, W, 3]`` tensors on ``device``, in order; ``None`` for files outside the device
    decoder's scope.  ``infos`` = descriptors already parsed (e.g. on a thread pool); ``staging`` = a dict the caller
    keeps alive to reuse the pinned host buffers between calls."""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if dev.type != "cuda":
        raise RuntimeError("decode_batch decodes on a CUDA device (no CPU fallback)")
    if infos is None:
        infos = [parse(d) for d in datas]
    keep = [i for i, inf in enumerate(infos) if inf is not None]
    out: List[Optional[Tensor]] = [None] * len(datas)
    if not keep:
        return out
    n = len(keep)
    arr = (GncJpegImage * n)()
    stream_bytes = sum(len(datas[i]) for i in keep)
    staging = staging if staging is not None else {}
    if staging.get("event") is not None:
        staging["event"].synchronize()              # the previous call's copies out of the pinned buffers are done
    host_stream = staging.get("stream")
    if host_stream is None or host_stream.numel() < stream_bytes:
        host_stream = staging["stream"] = torch.empty(max(stream_bytes, 1 << 20), dtype=torch.uint8).pin_memory()
    host_view = host_stream.numpy()
    off = blocks = plane = pixels = 0
    for j, i in enumerate(keep):
        inf, data = infos[i], datas[i]
        ctypes.memmove(ctypes.byref(arr[j]), ctypes.byref(inf), ctypes.sizeof(GncJpegImage))
        host_view[off:off + len(data)] = np.frombuffer(data, dtype=np.uint8)
        arr[j].scan_offset = inf.scan_offset + off
        arr[j].block_offset, arr[j].coef_offset = blocks, 64 * blocks
        arr[j].plane_offset, arr[j].pixel_offset = plane, pixels
        off += len(data)
        blocks += inf.n_blocks
        plane += inf.plane_bytes
        pixels += inf.width * inf.height
    nbytes = ctypes.sizeof(arr)
    host_infos = staging.get("infos")
    if host_infos is None or host_infos.numel() < nbytes:
        host_infos = staging["infos"] = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8).pin_memory()
    ctypes.memmove(host_infos.data_ptr(), ctypes.addressof(arr), nbytes)
    d_stream = host_stream[:stream_bytes].to(dev, non_blocking=True)
    d_infos = host_infos[:nbytes].to(dev, non_blocking=True)
    staging["event"] = torch.cuda.Event()
    staging["event"].record()
    coef = torch.empty(64 * blocks, dtype=torch.int16, device=dev)
    planes = torch.empty(plane, dtype=torch.uint8, device=dev)
    rgb = torch.empty(3 * pixels, dtype=torch.uint8, device=dev)
    check(ops._call("jpeg_decode", 0.0, float(stream_bytes + 128 * blocks * 2 + plane * 2 + 3 * pixels),
                    _lib.load().gnc_jpeg_decode_rgb_u8, d_stream.data_ptr(), d_infos.data_ptr(), n, blocks, pixels,
                    coef.data_ptr(), planes.data_ptr(), rgb.data_ptr(), ops._stream()), "jpeg_decode")
    for j, i in enumerate(keep):
        h, w, p0 = arr[j].height, arr[j].width, arr[j].pixel_offset
        out[i] = rgb[3 * p0:3 * (p0 + h * w)].view(h, w, 3)
    return out
