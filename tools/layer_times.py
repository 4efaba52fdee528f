#!/usr/bin/env python
"""Per-launch timing of one UNet forward of the BASELINE model (CUDA events around every C-ABI call, eager mode).
Prints one row per launch: op, shape, mean ms, and the achieved TFLOP/s (convs) or GB/s (GroupNorm apply / stats).

  python tools/layer_times.py [--batch 32] [--size 256] [--reps 5] > gpurun_out/layer_times.md
"""
import argparse
import collections
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anoddpm_b200 import engine as eng          # noqa: E402
from anoddpm_b200.UNet import UNetModel          # noqa: E402


@torch.no_grad()
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--conv-impl", default="tc")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    import bench
    kw = dict(bench.UNET_KW)
    kw["img_size"] = a.size
    m = UNetModel(**kw).to(dev).eval()
    for p in m.parameters():                       # zero-initialised output convs would hide nothing, but keep it generic
        if float(p.abs().max()) == 0:
            p.data.normal_(0, 0.02)
    m.conv_impl = a.conv_impl
    x = torch.randn(a.batch, 1, a.size, a.size, device=dev)
    t = torch.full((a.batch,), 400, device=dev, dtype=torch.int64)
    for _ in range(3):
        m(x, t)
    real_call = eng.call
    records = []

    def timed_call(name, *args):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        info = ""
        work = None
        if name == "anoddpm_conv":
            d = args[0]._obj
            cin = d.Ca * d.ksize * d.ksize + d.Cr1 + d.Cr2
            info = f"{d.H}x{d.W} Ca{d.Ca} Cr{d.Cr1 + d.Cr2} Co{d.Cout} k{d.ksize} impl{d.impl}{'h' if d.operand_f16 else ''} res{d.res_mode if d.res else '-'}"
            work = ("TF", 2.0 * d.B * d.H * d.W * cin * d.Cout)
            eb = 2 if d.operand_f16 else 4
            byts = d.B * d.H * d.W * (d.Ca * eb + (d.Cr1 + d.Cr2) * eb + d.Cout * 4 + (d.Cout * 4 if d.res else 0))
            info += f" alg{byts / 1e6:.0f}MB"
        elif name == "anoddpm_gn_apply":
            C1, C2, B, H, W = args[3], args[5], args[10], args[11], args[12]
            mode, om = args[16], args[17]
            raw = args[1] is not None and getattr(args[1], "value", args[1]) is not None
            C = C1 + C2
            eo = 2 if om == 2 else 4
            Ho, Wo = (H // 2, W // 2) if mode == 1 else ((H * 2, W * 2) if mode == 2 else (H, W))
            byts = B * C * (H * W * 4 + Ho * Wo * eo + (H * W * eo if raw else 0))
            info = f"{H}x{W} C{C1}+{C2} mode{mode} out{om}{' +raw' if raw else ''}"
            work = ("GB", byts)
        elif name == "anoddpm_gn_stats":
            C, B, HW = args[2], args[3], args[4]
            info = f"HW{HW} C{C}"
            work = ("GB", B * C * HW * 4)
        elif name == "anoddpm_attention":
            B, T, C, heads = args[2], args[3], args[4], args[5]
            info = f"T{T} C{C} heads{heads}"
            work = ("TF", 4.0 * B * heads * T * T * (C // heads))
        e0.record()
        r = real_call(name, *args)
        e1.record()
        records.append((name, info, work, e0, e1))
        return r

    eng.call = timed_call
    per = collections.OrderedDict()
    for rep in range(a.reps):
        records.clear()
        m(x, t)
        torch.cuda.synchronize()
        for i, (name, info, work, e0, e1) in enumerate(records):
            per.setdefault(i, [name, info, work, 0.0])[3] += e0.elapsed_time(e1) / a.reps
    eng.call = real_call
    tot = sum(v[3] for v in per.values())
    print(f"# per-launch times, BASELINE UNet forward, batch {a.batch}, {a.size}x{a.size}, conv_impl={a.conv_impl}: "
          f"{len(per)} launches, {tot:.2f} ms (CUDA events around each call, eager; includes launch gaps)\n")
    print("| # | op | shape | ms | rate |\n|---:|---|---|---:|---:|")
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for i, (name, info, work, ms) in per.items():
        rate = ""
        if work:
            rate = f"{work[1] / ms / 1e9:.0f} {'TFLOP/s' if work[0] == 'TF' else 'GB/s'}" if work[0] == "TF" else \
                f"{work[1] / ms / 1e6:.0f} GB/s"
        print(f"| {i} | {name.replace('anoddpm_', '')} | {info} | {ms:.3f} | {rate} |")
        k = (name.replace("anoddpm_", ""), info)
        agg[k][0] += 1
        agg[k][1] += ms
        agg[k][2] += work[1] if work else 0
    print("\n## aggregated by op and shape\n\n| op | shape | n | ms total | share | rate |\n|---|---|---:|---:|---:|---:|")
    for (name, info), (n, ms, w) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        is_tf = name in ("conv", "attention")
        rate = (f"{w / ms / 1e9:.0f} TFLOP/s" if is_tf else f"{w / ms / 1e6:.0f} GB/s") if w else ""
        print(f"| {name} | {info} | {n} | {ms:.3f} | {100 * ms / tot:.1f}% | {rate} |")


if __name__ == "__main__":
    main()
