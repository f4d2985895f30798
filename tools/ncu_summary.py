#!/usr/bin/env python
"""Turn gpurun_out/ ncu artefacts into the small text summaries committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/x_launches.csv  > profiles/rNN_x_launches.md
    python tools/ncu_summary.py full     gpurun_out/x_gemm.ncu-rep  > profiles/rNN_x_gemm.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active cycles)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/CTA"),
    ("sm__cycles_active.avg", "SM active cycles"),
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        rows.append((row["Kernel Name"].split("(")[0], v))
    agg = collections.OrderedDict()
    for k, v in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in rows)
    print(f"# ncu launch list: {path}\n")
    print("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES)\n")
    print(f"{len(rows)} launches, {tot / 1e3:.1f} us total\n")
    print("| kernel | launches | total us | avg us | share |")
    print("|---|---:|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {t / 1e3:.1f} | {t / n / 1e3:.2f} | {100 * t / tot:.1f}% |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr, units = r[0], r[1]
    print(f"# ncu --set full: {path}\n")
    for row in r[2:]:
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        print(f"## `{d['Kernel Name'].split('(')[0]}` grid {d.get('Grid Size')} block {d.get('Block Size')}\n")
        print("| metric | value | unit |")
        print("|---|---:|---|")
        for key, label in KEYS:
            if key in d:
                print(f"| {label} (`{key}`) | {d[key]} | {u[key]} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
