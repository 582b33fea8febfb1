#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed
under profiles/.

    python scripts/summarize_ncu.py launches gpurun_out/launches.csv  > profiles/rNN_launches_summary.txt
    python scripts/summarize_ncu.py full     gpurun_out/prof_x.ncu-rep > profiles/rNN_x_full_summary.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
    "sm__cycles_elapsed.avg.per_second",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        name = row["Kernel Name"].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f"# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` ({path})")
    print("# cold-cache, serialised launches: compare SHARES, not absolutes")
    print(f"{'ms':>10} {'share':>7} {'launches':>9}  kernel")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v:10.3f} {100 * v / tot:6.2f}% {c:9d}  {k}")
    print(f"{tot:10.3f} 100.00%            total")


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full --clock-control none ({path}); one block per captured launch")
    for r in rows[2:]:
        print("---")
        for k in KEYS:
            if k in idx:
                print(f"{k} [{units[idx[k]]}] = {r[idx[k]][:110]}")
        if "dram__bytes_read.sum" in idx:
            def gb(k):
                u, v = units[idx[k]], float(r[idx[k]].replace(",", ""))
                return v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3}[u]
            print(f"traffic_GB (dram read + write) = {gb('dram__bytes_read.sum') + gb('dram__bytes_write.sum'):.4f}")


def traffic(path, key, *sources):
    """Record the DRAM bytes of the LAST captured launch in profiles/kernel_traffic.json under `key`, tagged with the
    hash of the kernel's sources (bench.py reports the figure only while the sources are unchanged).
        python scripts/summarize_ncu.py traffic gpurun_out/x.ncu-rep tc_chain2_edge_b512_r128 tc_chain.cu common.cuh"""
    import hashlib, json, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, r = rows[0], rows[1], rows[-1]
    idx = {h: i for i, h in enumerate(hdr)}

    def nbytes(k):
        return float(r[idx[k]].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[units[idx[k]]]

    h = hashlib.sha256()
    for n in sources:
        with open(os.path.join(root, "graphnet_classifier_b200", "csrc", n), "rb") as f:
            h.update(f.read())
    out = os.path.join(root, "profiles", "kernel_traffic.json")
    table = json.load(open(out)) if os.path.exists(out) else {}
    table[key] = {"kernel": r[idx["Kernel Name"]][:80], "source_sha": h.hexdigest()[:16], "sources": list(sources),
                  "dram_bytes_read": nbytes("dram__bytes_read.sum"), "dram_bytes_write": nbytes("dram__bytes_write.sum"),
                  "gpu_time_ms": r[idx["gpu__time_duration.sum"]] + " " + units[idx["gpu__time_duration.sum"]],
                  "from": os.path.relpath(path, root)}
    json.dump(table, open(out, "w"), indent=1)
    print(json.dumps(table[key]))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
