"""Summaries of ncu output for profiles/ (run here, no GPU needed).

    python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/rN_launches.md
    python tools/ncu_summary.py full     gpurun_out/prof.ncu-rep  > profiles/rN_full.md

`launches`: the per-launch gpu__time_duration.sum list (cold cache, serialised) aggregated per
kernel: launches, total, share of the profiled window, average.
`full`: one row per profiled launch of an `ncu --set full` report with the counters the roofline
uses (duration, DRAM bytes read+written, DRAM throughput %, tensor-pipe %, occupancy, registers).
"""
import collections
import csv
import io
import re
import subprocess
import sys


def short(name):
    name = re.sub(r"\(.*", "", name)
    for s in ("void ", "sig::", "<unnamed>::", "(anonymous namespace)::"):
        name = name.replace(s, "")
    return name[:72]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr, data = rows[0], rows[1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in data:
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print(f"{len(data)} launches, {tot / 1e3:.1f} us of kernel time (cold-cache, serialised: compare shares)\n")
    print("| kernel | launches | total us | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{n}` | {a[0]} | {a[1] / 1e3:.1f} | {a[1] / tot * 100:.1f}% | {a[1] / a[0] / 1e3:.2f} |")


METRICS = [
    ("gpu__time_duration.sum", "dur us", 1e-3),
    ("dram__bytes_read.sum", "dram rd MB", 1e-6),
    ("dram__bytes_write.sum", "dram wr MB", 1e-6),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %", 1),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %", 1),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %", 1),
    ("launch__registers_per_thread", "regs", 1),
]
UNIT_SCALE = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def full(path):
    """`path`: an .ncu-rep (read through `ncu --page raw --csv`) or that command's CSV output saved on the GPU box."""
    if path.endswith(".csv"):
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki, gi = hdr.index("Kernel Name"), hdr.index("Grid Size")
    cols = []
    for m, label, sc in METRICS:
        cand = [i for i, h in enumerate(hdr) if h == m]
        cols.append((cand[0] if cand else None, label, sc))
    print("| # | kernel | grid | " + " | ".join(l for _, l, _ in cols) + " |")
    print("|---|---|---|" + "---:|" * len(cols))
    for n, r in enumerate(data):
        vals = []
        for i, label, sc in cols:
            if i is None or r[i] == "":
                vals.append("-")
                continue
            v = float(r[i].replace(",", "")) * UNIT_SCALE.get(units[i], 1.0)
            vals.append(f"{v * sc:.1f}" if sc != 1 or "%" in label else f"{v:.0f}")
        print(f"| {n} | `{short(r[ki])}` | {r[gi]} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
