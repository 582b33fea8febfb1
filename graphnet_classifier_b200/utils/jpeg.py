"""Baseline JPEG decode on the device (csrc/jpeg.cu): the file decode behind the reference's
``Image.open(path).convert('RGB')`` (utils/dataloader.py:34 through ImageFolder's loader,
utils/image_to_graph/image_to_graph_optimized.py:65-68, utils/inference.py:47), bit for bit Pillow's pixels.

    images = decode_batch(files, device)               # list of uint8 [H, W, 3] device tensors (None where unsupported)
    info = parse(data)                                 # host: markers -> descriptor, None = not a baseline JPEG

Only the compressed bytes cross PCIe (~14 x fewer than decoded pixels).  Files the device decoder does not cover
(progressive, CMYK, PNG, ...) come back as ``None``; ``utils/staging.DecodePool`` decodes those with Pillow on the host.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence

import torch
from torch import Tensor

from .. import _lib, ops
from .._lib import GncJpegImage, check

_HOST_THREADS = max(1, min(8, (os.cpu_count() or 2) - 1))


def parse(data: bytes) -> Optional[GncJpegImage]:
    """Descriptor of a baseline JPEG file, or None when the device decoder does not cover the file."""
    info = GncJpegImage()
    rc = _lib.load().gnc_jpeg_parse(data, len(data), ctypes.byref(info))
    return info if rc == _lib.GNC_OK else None


def decode_batch(datas: Sequence[bytes], device=None, staging: Optional[dict] = None) -> List[Optional[Tensor]]:
    """JPEG file contents -> ``uint8 [H, W, 3]`` tensors on ``device``, in order; ``None`` for files outside the device
    decoder's scope.  One host call parses the batch and packs descriptors and bytes into pinned memory
    (``gnc_jpeg_pack``, multi-threaded), one device call decodes it.  ``staging`` = a dict the caller keeps alive to reuse
    the pinned host buffers between calls."""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if dev.type != "cuda":
        raise RuntimeError("decode_batch decodes on a CUDA device (no CPU fallback)")
    lib = _lib.load()
    n = len(datas)
    out: List[Optional[Tensor]] = [None] * n
    if n == 0:
        return out
    staging = staging if staging is not None else {}
    if staging.get("event") is not None:
        staging["event"].synchronize()              # the previous call's copies out of the pinned buffers are done
    total_bytes = sum(len(d) for d in datas)
    host_stream = staging.get("stream")
    if host_stream is None or host_stream.numel() < total_bytes:
        host_stream = staging["stream"] = torch.empty(max(total_bytes, 1 << 20), dtype=torch.uint8).pin_memory()
    info_bytes = n * ctypes.sizeof(GncJpegImage)
    host_infos = staging.get("infos")
    if host_infos is None or host_infos.numel() < info_bytes:
        host_infos = staging["infos"] = torch.empty(max(info_bytes, 1 << 16), dtype=torch.uint8).pin_memory()
    ptrs = (ctypes.c_char_p * n)(*datas)
    sizes = (ctypes.c_int64 * n)(*[len(d) for d in datas])
    index = (ctypes.c_int32 * n)()
    totals = (ctypes.c_int64 * 5)()
    check(lib.gnc_jpeg_pack(ctypes.cast(ptrs, ctypes.c_void_p), ctypes.cast(sizes, ctypes.c_void_p), n, _HOST_THREADS,
                            host_stream.data_ptr(), host_stream.numel(), host_infos.data_ptr(),
                            ctypes.cast(index, ctypes.c_void_p), ctypes.cast(totals, ctypes.c_void_p)), "jpeg_pack")
    m, stream_bytes, blocks, plane, pixels = (int(v) for v in totals)
    if m == 0:
        return out
    d_stream = host_stream[:stream_bytes].to(dev, non_blocking=True)
    d_infos = host_infos[:m * ctypes.sizeof(GncJpegImage)].to(dev, non_blocking=True)
    staging["event"] = torch.cuda.Event()
    staging["event"].record()
    coef = torch.empty(64 * blocks, dtype=torch.int16, device=dev)
    planes = torch.empty(plane, dtype=torch.uint8, device=dev)
    rgb = torch.empty(3 * pixels, dtype=torch.uint8, device=dev)
    scratch = torch.empty(int(lib.gnc_jpeg_scratch_bytes(stream_bytes, m)), dtype=torch.uint8, device=dev)
    check(ops._call("jpeg_decode", 0.0, float(stream_bytes + 128 * blocks * 2 + plane * 2 + 3 * pixels),
                    lib.gnc_jpeg_decode_rgb_u8, d_stream.data_ptr(), d_infos.data_ptr(), m, blocks, pixels,
                    coef.data_ptr(), planes.data_ptr(), rgb.data_ptr(), scratch.data_ptr(), ops._stream()), "jpeg_decode")
    infos = ctypes.cast(host_infos.data_ptr(), ctypes.POINTER(GncJpegImage))
    for j in range(m):
        h, w, p0 = infos[j].height, infos[j].width, infos[j].pixel_offset
        out[index[j]] = rgb[3 * p0:3 * (p0 + h * w)].view(h, w, 3)
    return out
