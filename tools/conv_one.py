#!/usr/bin/env python
"""Time one tensor-core conv configuration through the C ABI (CUDA events), for epilogue attribution and ncu captures.

  python tools/conv_one.py --ci 128 --co 128 --hw 256 --variants          # table over epilogue variants
  ncu --set full --import-source on -k regex:conv_tc3 -s 3 -c 1 -o gpurun_out/x python tools/conv_one.py --res 1 --stats
"""
import argparse
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anoddpm_b200 import _lib          # noqa: E402


def run(a, B, Ci, Co, HW, res, stats, emb, cr=0, n=10, gn=False, out16=False, out_f16=False, ci2=0):
    dev = torch.device("cuda:0")
    x = torch.randn(B, HW, HW, Ci - ci2, device=dev).half()
    w = (torch.randn(Co, 9 * Ci + cr, device=dev) * 0.03).half()
    r1 = torch.randn(B, HW, HW, cr, device=dev).half() if cr else None
    bias = torch.randn(Co, device=dev)
    out = torch.empty(B, HW, HW, Co, device=dev)
    d = _lib.ConvDesc()
    d.out, d.a, d.w, d.bias = out.data_ptr(), x.data_ptr(), w.data_ptr(), bias.data_ptr()
    d.B, d.H, d.W, d.Ca, d.Cout, d.ksize, d.impl = B, HW, HW, Ci - ci2, Co, 3, a.impl
    d.operand_f16 = 1
    keep0 = []
    if ci2:
        x2 = torch.randn(B, HW, HW, ci2, device=dev).half()
        keep0.append(x2)
        d.a2, d.Ca2 = x2.data_ptr(), ci2
    if gn:      # operand-side GroupNorm + SiLU: (scale, shift) table
        tab = torch.stack([torch.rand(B, Ci, device=dev) + 0.5, torch.randn(B, Ci, device=dev) * 0.3], dim=-1).contiguous()
        keep0.append(tab)
        d.gn_table = tab.data_ptr()
    if out16:
        o16 = torch.empty(B, HW, HW, Co, device=dev, dtype=torch.float16)
        keep0.append(o16)
        d.out16 = o16.data_ptr()
    if out_f16:
        out = torch.empty(B, HW, HW, Co, device=dev, dtype=torch.float16)
        d.out = out.data_ptr()
        d.out_f16 = 1
    if cr:
        d.r1, d.Cr1 = r1.data_ptr(), cr
    keep = []
    if res:
        shape = {1: (B, HW, HW, Co), 2: (B, 2 * HW, 2 * HW, Co), 3: (B, HW // 2, HW // 2, Co)}[res]
        rt = torch.randn(*shape, device=dev)
        keep.append(rt)
        d.res, d.res_mode = rt.data_ptr(), res
    if emb:
        et = torch.randn(B, Co, device=dev)
        keep.append(et)
        d.emb, d.emb_stride = et.data_ptr(), Co
    if stats:
        st = torch.zeros(8 * B * Co * 2, device=dev, dtype=torch.float64)
        keep.append(st)
        d.gn_stats = st.data_ptr()
    f = lambda: _lib.call("anoddpm_conv", ctypes.byref(d), _lib.stream())   # noqa: E731
    for _ in range(3):
        f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flop = 2.0 * B * HW * HW * Co * (9 * Ci + cr)
    return ms, flop / ms / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--ci", type=int, default=128)
    ap.add_argument("--co", type=int, default=128)
    ap.add_argument("--cr", type=int, default=0)
    ap.add_argument("--hw", type=int, default=256)
    ap.add_argument("--res", type=int, default=0)
    ap.add_argument("--impl", type=int, default=3)
    ap.add_argument("--stats", action="store_true")
    ap.add_argument("--emb", action="store_true")
    ap.add_argument("--variants", action="store_true")
    ap.add_argument("--gn", action="store_true", help="operand-side GroupNorm + SiLU (gn_table)")
    ap.add_argument("--out16", action="store_true")
    ap.add_argument("--outf16", action="store_true")
    ap.add_argument("--gn-attrib", action="store_true", help="attribute the operand-transform overhead (ANODDPM_TC3_DBG)")
    ap.add_argument("--conv2-attrib", action="store_true", help="attribute the conv2-form epilogue (residual / statistics / fp16 copy)")
    a = ap.parse_args()
    if a.conv2_attrib:
        print(f"| conv2 forms {a.ci}->{a.co} @{a.hw}^2 bs{a.batch} | RBUF | RESPF | ms | TFLOP/s |\n|---|---:|---:|---:|---:|")
        forms = (("plain", dict(res=0, stats=False, emb=False)),
                 ("+res1", dict(res=1, stats=False, emb=False)),
                 ("+res1+stats", dict(res=1, stats=True, emb=False)),
                 ("+res1+stats+fp16 copy", dict(res=1, stats=True, emb=False, out16=True)),
                 ("+stats+fp16 copy (no residual)", dict(res=0, stats=True, emb=False, out16=True)),
                 ("GN only", dict(res=0, stats=False, emb=False, gn=True)),
                 ("GN +res1", dict(res=1, stats=False, emb=False, gn=True)),
                 ("GN +res1+stats", dict(res=1, stats=True, emb=False, gn=True)),
                 ("GN +res1+stats+fp16 copy (production conv2)", dict(res=1, stats=True, emb=False, out16=True, gn=True)),
                 ("GN +res3+stats+fp16 copy (upsampled residual)", dict(res=3, stats=True, emb=False, out16=True, gn=True)))
        for name, kw in forms:
            combos = ((2, 0), (2, 1), (3, 0), (3, 1)) if kw["res"] else ((2, 0),)
            for rbuf, pf in combos:
                os.environ["ANODDPM_TC3_RBUF"] = str(rbuf)
                os.environ["ANODDPM_TC3_RESPF"] = str(pf)
                ms, tf = run(a, a.batch, a.ci, a.co, a.hw, **kw)
                print(f"| {name} | {rbuf} | {pf} | {ms:.3f} | {tf:.0f} |", flush=True)
        return
    if a.gn_attrib:
        print(f"| conv1 GN {a.ci}->{a.co} @{a.hw}^2 bs{a.batch}: transform variant | ms | TFLOP/s |\n|---|---:|---:|")
        for name, dbg in (("full transform", "0"), ("barrier forwarding only", "1"), ("read + compute, no write-back", "2")):
            os.environ["ANODDPM_TC3_DBG"] = dbg
            ms, tf = run(a, a.batch, a.ci, a.co, a.hw, res=0, stats=True, emb=True, out_f16=True, gn=True)
            print(f"| {name} | {ms:.3f} | {tf:.0f} |", flush=True)
        os.environ["ANODDPM_TC3_DBG"] = "0"
        os.environ["ANODDPM_TC3_NA"] = "3"
        for name, kw in (("conv1 GN, 3 A boxes + 2 weight stages", dict(res=0, stats=True, emb=True, out_f16=True, gn=True)),
                         ("conv2 GN + fp16 copy, 3 A boxes + 2 weight stages", dict(res=1, stats=True, emb=False, out16=True, gn=True))):
            ms, tf = run(a, a.batch, a.ci, a.co, a.hw, **kw)
            print(f"| {name} | {ms:.3f} | {tf:.0f} |", flush=True)
        os.environ["ANODDPM_TC3_NA"] = "2"
        ms, tf = run(a, a.batch, a.ci, a.co, a.hw, res=1, stats=True, emb=False, out16=True, gn=True)
        print(f"| conv2 GN + fp16 copy, 2 A boxes | {ms:.3f} | {tf:.0f} |", flush=True)
        ms, tf = run(a, a.batch, a.ci, a.co, a.hw, res=0, stats=True, emb=True, out_f16=True)
        print(f"| no operand transform (plain conv1) | {ms:.3f} | {tf:.0f} |")
        return
    if a.variants:
        print(f"| conv {a.ci}->{a.co} @{a.hw}^2 bs{a.batch} impl{a.impl} fp16 | ms | TFLOP/s |\n|---|---:|---:|")
        for name, res, stats, emb, cr in (("plain", 0, False, False, 0), ("+stats", 0, True, False, 0),
                                          ("+emb+stats", 0, True, True, 0), ("+res1", 1, False, False, 0),
                                          ("+res1+stats", 1, True, False, 0), ("+res3+stats", 3, True, False, 0),
                                          ("+skip1x1(256)+stats", 0, True, False, 256)):
            ms, tf = run(a, a.batch, a.ci, a.co, a.hw, res, stats, emb, cr)
            print(f"| {name} | {ms:.3f} | {tf:.0f} |")
        # production epilogues of a ResBlock, without / with the operand-side GroupNorm
        for name, kw in (("conv1: +emb+stats, fp16 out", dict(res=0, stats=True, emb=True, out_f16=True)),
                         ("conv1 GN: same, operand-side GroupNorm", dict(res=0, stats=True, emb=True, out_f16=True, gn=True)),
                         ("conv2: +res1+stats", dict(res=1, stats=True, emb=False)),
                         ("conv2 +fp16 copy", dict(res=1, stats=True, emb=False, out16=True)),
                         ("conv2 GN: +res1+stats+fp16 copy, operand-side GroupNorm", dict(res=1, stats=True, emb=False, out16=True, gn=True)),
                         ("conv2 GN, no fp16 copy", dict(res=1, stats=True, emb=False, gn=True))):
            ms, tf = run(a, a.batch, a.ci, a.co, a.hw, **kw)
            print(f"| {name} | {ms:.3f} | {tf:.0f} |", flush=True)
    else:
        ms, tf = run(a, a.batch, a.ci, a.co, a.hw, a.res, a.stats, a.emb, a.cr, gn=a.gn, out16=a.out16, out_f16=a.outf16)
        print(f"{ms:.3f} ms  {tf:.0f} TFLOP/s")


if __name__ == "__main__":
    main()
