#!/usr/bin/env python
"""Turn ncu outputs (brought back under gpurun_out/) into the small text summaries committed in profiles/.

  python profiles/summarize.py launches gpurun_out/launches_X.csv   > profiles/rNN_launches.md
  python profiles/summarize.py kernel   gpurun_out/prof_X.ncu-rep   > profiles/rNN_kernel_X.md
"""
import collections
import csv
import re
import subprocess
import sys


def _ms(row):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    return v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v


def launches(path):
    with open(path) as f:
        rows = list(csv.DictReader(l for l in f if not l.startswith("==")))
    def short(n):
        n = n.replace("void ", "").replace("<unnamed>::", "")
        n = re.sub(r"\(.*", "", n)
        return re.sub(r"^.*::", "", n) if "at::" in n or "cuda::" in n else n
    names = [short(r["Kernel Name"]) for r in rows]
    ps = [i for i, n in enumerate(names) if n.startswith("p_sample")]
    if len(ps) >= 2:
        a, b = ps[0] + 1, ps[1] + 1
        for i, j in zip(ps[:-1], ps[1:]):      # prefer a window without the chain's q_sample (+ its own noise launch)
            if not any(n.startswith("q_sample") for n in names[i + 1:j + 1]):
                a, b = i + 1, j + 1
                break
        title = f"one reverse step (launches {a}..{b} of the capture, p_sample to p_sample)"
    else:
        a, b = 0, len(rows)
        title = "whole capture"
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for i in range(a, b):
        agg[names[i]][0] += 1
        agg[names[i]][1] += _ms(rows[i])
        tot += _ms(rows[i])
    print(f"# ncu launch list -- {title}\n")
    print("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES).\n")
    print(f"{b - a} launches, sum of kernel durations {tot:.2f} ms\n")
    print("| kernel | launches | ms | share |\n|---|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if t / tot < 0.0005:
            continue
        print(f"| {k} | {n} | {t:.3f} | {100 * t / tot:.1f}% |")


KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "derived__lts__lts2xbar_bytes.sum.per_second", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg"]


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full capture: {path}\n")
    for r in rows[2:]:
        print(f"## {r[idx['Kernel Name']][:90]}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in idx:
                print(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |")
        print()


def aux(path):
    """one table row per captured kernel: duration, DRAM bytes / rate, SM and tensor-pipe activity, registers"""
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}

    def val(r, k, scale_unit=None):
        v = float(r[idx[k]].replace(",", ""))
        u = units[idx[k]]
        if scale_unit == "MB":
            v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u]
        if scale_unit == "us":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}[u]
        return v
    print("| kernel | grid | us | DRAM read MB | DRAM write MB | dram GB/s | dram % of peak | SM % | tensor pipe % | regs |")
    print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|")
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]].replace("void ", "").replace("<unnamed>::", ""))
        us = val(r, "gpu__time_duration.sum", "us")
        rd, wr = val(r, "dram__bytes_read.sum", "MB"), val(r, "dram__bytes_write.sum", "MB")
        print(f"| `{name}` | {r[idx['Grid Size']]} | {us:.1f} | {rd:.1f} | {wr:.1f} | {(rd + wr) / us * 1e3:.0f} | "
              f"{val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
              f"{val(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
              f"{val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
              f"{int(val(r, 'launch__registers_per_thread'))} |")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel, "aux": aux}[sys.argv[1]](sys.argv[2])
