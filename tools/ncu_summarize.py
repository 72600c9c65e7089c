"""Turn gpurun_out/<tag>_launches.csv and <tag>_full_*.ncu-rep (tools/profile_round.sh) into the committed
summaries under profiles/:  <tag>_launch_shares.txt, <tag>_ncu_full.txt and ncu_traffic.json (DRAM bytes per
launch of each captured kernel -- bench.py reads it for roofline.traffic).

usage: python tools/ncu_summarize.py <tag> "<entry name of capture 0>" ["<entry name of capture 1>" ...]
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__cycles_active.avg",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]


def launch_shares(tag):
    path = os.path.join(OUT, f"{tag}_launches.csv")
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start + 1:]:
        if len(r) <= vi or not r[vi]:
            continue
        try:
            t = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        t *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")[:80]
        agg[name][0] += 1
        agg[name][1] += t
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    lines = [f"# {tag}: EADGAN_PROFILE_WINDOW=1 ncu --profile-from-start off --metrics gpu__time_duration.sum "
             "--clock-control none python bench.py --batch 1024 --steps 1 --warmup 3 --no-cpu-baseline",
             f"# ONE graph-replayed step at B=1024: {n} kernels, serialised kernel time {tot / 1e3:.2f} ms "
             "(cold-cache, serialised: compare SHARES with the bench, not absolutes)",
             f"{'kernel':82s} {'launches':>8s} {'total_ms':>9s} {'share':>6s}"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k:82s} {v[0]:8d} {v[1] / 1e3:9.3f} {100 * v[1] / tot:5.1f}%")
    with open(os.path.join(PROF, f"{tag}_launch_shares.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:30]))


def full(tag, names):
    traffic_path = os.path.join(PROF, "ncu_traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    out = [f"# {tag}: ncu --set full --clock-control none --import-source on, 2 launches per capture "
           "(tools/profile_round.sh; micro-benchmark tools/bench_layers.py at B=1024)"]
    for i, name in enumerate(names):
        rep = os.path.join(OUT, f"{tag}_full_{i}.ncu-rep")
        if not os.path.exists(rep):
            continue
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        out.append(f"\n== capture {i}: {name}")
        for r in rows[2:]:
            if len(r) < len(hdr):
                continue
            out.append("  kernel: " + r[hdr.index("Kernel Name")][:110])
            vals = {}
            for w in WANT:
                if w in hdr:
                    j = hdr.index(w)
                    vals[w] = (r[j], units[j])
                    out.append(f"    {w:70s} {r[j]:>14s} {units[j]}")
            try:
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                rd = float(vals["dram__bytes_read.sum"][0].replace(",", "")) * scale[vals["dram__bytes_read.sum"][1]]
                wr = float(vals["dram__bytes_write.sum"][0].replace(",", "")) * scale[vals["dram__bytes_write.sum"][1]]
                traffic[name] = {"dram_bytes_per_launch": rd + wr, "capture": f"{tag}_full_{i}",
                                 "tensor_pipe_pct": float(vals["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"][0])}
            except (KeyError, ValueError):
                pass
    with open(os.path.join(PROF, f"{tag}_ncu_full.txt"), "w") as f:
        f.write("\n".join(out) + "\n")
    with open(traffic_path, "w") as f:
        json.dump(traffic, f, indent=1, sort_keys=True)
    print("\n".join(out))


if __name__ == "__main__":
    tag = sys.argv[1]
    os.makedirs(PROF, exist_ok=True)
    if os.path.exists(os.path.join(OUT, f"{tag}_launches.csv")):
        launch_shares(tag)
    full(tag, sys.argv[2:])
