#!/usr/bin/env python
"""Is a kernel clock-limited by the board's power cap?  Loops one launch shape for a few seconds while nvidia-smi samples
SM clock and power draw every 50 ms; prints the medians under load next to the launch time.

    python scripts/power_probe.py [graphs] [seconds]
"""
import os, subprocess, sys, time, statistics, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphnet_classifier_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
r = 128; dev = "cuda"
N, E = B * r * r, B * 2 * r * (r - 1)
g = torch.Generator(device=dev).manual_seed(0)
mk = lambda *s: torch.randn(*s, device=dev, generator=g)
layers = [(mk(128, 128) / 11, mk(128) * 0.1) for _ in range(3)]
gamma, beta = torch.ones(128, device=dev), torch.zeros(128, device=dev)
v = torch.arange(r * r, device=dev).view(r, r)
src1 = torch.cat([v[:, :-1].reshape(-1), v[:-1, :].reshape(-1)]); dst1 = torch.cat([v[:, 1:].reshape(-1), v[1:, :].reshape(-1)])
off = (torch.arange(B, device=dev) * r * r).view(B, 1)
src = (src1.view(1, -1) + off).reshape(-1).int(); dst = (dst1.view(1, -1) + off).reshape(-1).int()
e = mk(E, 128); P = mk(N, 128); Q = mk(N, 128); out = torch.empty(E, 128, device=dev)
h = mk(N, 128)
graph = ops.build_csr(dst, N) if hasattr(ops, "build_csr") else None
big = torch.empty(1 << 30, dtype=torch.uint8, device=dev); big2 = torch.empty_like(big)
A = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16); Bm = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)

cases = {
    "edge MLP chained (3 layers, P/Q addends, LN, residual)": lambda: ops.tc_mlp_chain(e, layers, gather0=(P, src), gather1=(Q, dst), gamma=gamma, beta=beta, residual=e, out=out),
    "chain, 3 layers, no addends / LN / residual": lambda: ops.tc_mlp_chain(e, layers, out=out),
    "per-layer tc_linear (3xTF32), plain": lambda: ops.tc_linear(e, layers[0][0], bias=layers[0][1], relu=True, out=out),
    "P,Q products (multi)": lambda: ops.tc_linear_multi(h, [layers[0][0], layers[1][0]]),
    "device copy 1 GiB (torch)": lambda: big2.copy_(big),
    "cuBLAS bf16 8192^3 (torch.matmul)": lambda: torch.matmul(A, Bm),
}


def sample(stop, rows):
    while not stop.is_set():
        o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-i", "0"],
                           capture_output=True, text=True).stdout.strip().split(",")
        try:
            rows.append((float(o[0]), float(o[1]), o[2].strip()))
        except Exception:
            pass
        time.sleep(0.03)


for name, fn in cases.items():
    for _ in range(3): fn()
    torch.cuda.synchronize()
    time.sleep(1.0)                     # let the board cool to idle clocks between cases
    stop, rows = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, rows)); th.start()
    t0 = time.time(); n = 0
    s, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    while time.time() - t0 < secs:
        for _ in range(10): fn()
        n += 10
        torch.cuda.synchronize()
    en.record(); torch.cuda.synchronize()
    stop.set(); th.join()
    ms = s.elapsed_time(en) / n
    rows = rows[len(rows) // 3:]        # steady state
    clk = statistics.median(x[0] for x in rows); pw = statistics.median(x[1] for x in rows)
    cap = sum(1 for x in rows if x[2].startswith("Active")) / max(1, len(rows))
    print(f"{name:58s} {ms:8.3f} ms/launch  SM clock {clk:6.0f} MHz  power {pw:6.0f} W  sw_power_cap active in {100 * cap:3.0f}% of {len(rows)} samples")
