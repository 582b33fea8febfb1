#!/usr/bin/env python
"""A/B of library variants (graphnet_classifier_b200/variants/libgnc_<name>.so, built with different -D switches) on the
chained-kernel micro-benchmark: the variants run in separate processes, interleaved, several rounds (the GPU is under a
power cap, so neighbours in time are the fair comparison).

    python scripts/chain_ab.py base,elb,both [rounds] [case filter]
"""
import os, re, subprocess, sys, statistics, collections
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
names = sys.argv[1].split(",")
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
filt = sys.argv[3] if len(sys.argv) > 3 else "chain"
res = collections.defaultdict(lambda: collections.defaultdict(list))
for r in range(rounds):
    for n in names:
        env = dict(os.environ, GNC_LIB=os.path.join(root, "graphnet_classifier_b200", "variants", f"libgnc_{n}.so"))
        out = subprocess.run([sys.executable, os.path.join(root, "scripts", "chain_microbench.py"), "512", filt], env=env,
                             capture_output=True, text=True)
        if out.returncode != 0:
            print(n, "FAILED", out.stderr[-800:])
            continue
        for line in out.stdout.splitlines():
            m = re.match(r"(.+?)\s+rows=\d+\s+median\s+([\d.]+) ms", line)
            if m:
                res[m.group(1).strip()][n].append(float(m.group(2)))
for case, by in res.items():
    print(f"{case:36s} " + "  ".join(f"{n}: {statistics.median(v):7.3f} (min {min(v):7.3f})" for n, v in by.items()))
