#!/usr/bin/env python
"""Weight-gradient micro-benchmark: tcgen05 engine vs the fp32 CUDA-core kernel, plus error vs fp64."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphnet_classifier_b200 import ops, build, _lib
build.build()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 2_080_768
dev = "cuda"
dZ = torch.randn(M, 128, device=dev); X = torch.randn(M, 128, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        flush.zero_(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return min(ts)
def fp32_wgrad():
    lib = _lib.load(); dW = torch.empty(128, 128, device=dev)
    segs = ops._make_segs([X], [None]); n = int(lib.gnc_linear_wgrad_workspace(M, 128, 128)); ws = ops._workspace(dZ.device, n)
    ops.check(lib.gnc_linear_wgrad_f32(dZ.data_ptr(), 128, M, 128, segs, 1, dW.data_ptr(), 128, 0, ws.data_ptr(), n, ops._stream()))
    return dW
t_tc, t_fp = timeit(lambda: ops.tc_wgrad(dZ, X)), timeit(fp32_wgrad)
ref = (dZ[: min(M, 4_000_000)].double().t() @ X[: min(M, 4_000_000)].double()) if M <= 4_000_000 else None
gb = 8.0 * M * 128 / 1e9
print(f"M={M}: tc_wgrad {t_tc:.3f} ms ({gb / t_tc * 1e3:.0f} GB/s, {2.0 * M * 16384 / t_tc / 1e9:.1f} TFLOP/s fp32-eq)   fp32 wgrad {t_fp:.3f} ms")
if ref is not None:
    rel = lambda a: float((a.double() - ref).norm() / ref.norm())
    print(f"  rel-L2 error vs fp64: tc {rel(ops.tc_wgrad(dZ, X)):.2e}   fp32 {rel(fp32_wgrad()):.2e}")
