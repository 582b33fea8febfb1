#!/usr/bin/env python
"""Timeline of the chained-MLP kernel's roles in CTA 0 (gnc_debug_chain_trace): prints, for a few
steady-state tiles, each role's events in cycles relative to the tile's first MMA chunk."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from graphnet_classifier_b200 import ops, build, _lib
build.build()
lib = _lib.load()
mode = sys.argv[1] if len(sys.argv) > 1 else "plain"
dev = "cuda"
B, r = int(os.environ.get("TRACE_B", "64")), 128
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
kw = {}
if mode in ("ln", "full"):
    kw.update(gamma=gamma, beta=beta, residual=e)
if mode in ("full", "full_l2"):
    if mode == "full_l2":          # gather tables that stay L2-resident: isolates L2 latency from DRAM latency
        src, dst = src % 2048, dst % 2048
    kw.update(gamma=gamma, beta=beta, residual=e, gather0=(P, src), gather1=(Q, dst))
cap = 4096
buf = torch.zeros(4 * cap, dtype=torch.int64, device=dev)
ops.tc_mlp_chain(e, layers, out=out, **kw)
torch.cuda.synchronize()
lib.gnc_debug_chain_trace(buf.data_ptr(), cap)
ops.tc_mlp_chain(e, layers, out=out, **kw)
torch.cuda.synchronize()
lib.gnc_debug_chain_trace(None, 0)
tr = buf.cpu().view(4, cap).numpy()
names = {0: "MMA", 1: "EPI(q0,h0)", 2: "LOADER0", 3: "EPI(q0,h1)"}
ev = []
for role in range(4):
    for x in tr[role]:
        if x == 0:
            break
        ev.append((int(x) >> 8, role, int(x) & 0xff))
ev.sort()
# tile boundaries: MMA tag 0x10 (layer 0 chunk 0)
starts = [t for t, role, tag in ev if role == 0 and tag == 0x10]
print(f"mode={mode} tiles traced={len(starts)}; cycles per tile (steady): ", [starts[i + 1] - starts[i] for i in range(2, min(10, len(starts) - 1))])
t0, t1 = starts[4], starts[6]
tagname = lambda tag: (f"mma_go l{(tag-0x10)//4} c{(tag-0x10)%4}" if 0x10 <= tag < 0x20 else f"mma_commit l{tag-0x20}" if 0x20 <= tag < 0x30
                       else f"epi_dfull l{tag-0x30}" if 0x30 <= tag < 0x40 else f"epi_arrive l{(tag-0x40)//4} c{(tag-0x40)%4}" if 0x40 <= tag < 0x50
                       else ["e0_pre_cpwait", "e0_cp_landed", "e0_fetch_issued", "e0_ldtm_done", "sttm_issued", "sttm_done"][tag - 0x70] if 0x70 <= tag < 0x76
                       else "last_ldtm_done" if tag == 0x51 else "last_ln_done" if tag == 0x53 else f"last_step{tag-0x54}_stored" if 0x54 <= tag < 0x58
                       else "epi_last_dfull" if tag == 0x50 else "epi_tile_done" if tag == 0x52
                       else f"ld_landed c{tag-0x60}" if 0x60 <= tag < 0x64 else f"ld_aempty c{tag-0x64}" if 0x64 <= tag < 0x68 else f"ld_arrive c{tag-0x68}")
for t, role, tag in ev:
    if t0 - 2000 <= t < t1:
        print(f"{t - t0:8d}  {names[role]:12s} {tagname(tag)}")
