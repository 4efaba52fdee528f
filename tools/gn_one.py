#!/usr/bin/env python
"""Time the GroupNorm apply passes at the full-resolution BASELINE shape (CUDA events, GB/s against algorithmic bytes)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from anoddpm_b200 import _lib          # noqa: E402


def t_ms(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    dev = torch.device("cuda:0")
    B, H, W = 32, 256, 256
    for C in (128, 256):
        xa = torch.randn(B, H, W, C, device=dev)
        st = torch.zeros(8, B, C, 2, device=dev, dtype=torch.float64)
        _lib.call("anoddpm_gn_stats", _lib.ptr(st), _lib.ptr(xa), C, B, H * W, _lib.stream())
        g1, b1 = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        oa = torch.empty(B, H, W, C, device=dev, dtype=torch.float16)
        ms = t_ms(lambda: _lib.call("anoddpm_gn_apply", _lib.ptr(oa), None, _lib.ptr(xa), C, None, 0, _lib.ptr(st), None,
                                    _lib.ptr(g1), _lib.ptr(b1), B, H, W, 32, 1e-5, 1, 0, 2, _lib.stream()))
        print(f"gn_apply      C{C} fp32->fp16: {ms:.3f} ms  {6.0 * xa.numel() / ms / 1e6:.0f} GB/s")
        if C == 256:
            x1, x2 = xa[..., :128].contiguous(), xa[..., 128:].contiguous()
            s1, s2 = st[:, :, :128].contiguous(), st[:, :, 128:].contiguous()
            raw = torch.empty_like(oa)
            ms = t_ms(lambda: _lib.call("anoddpm_gn_apply", _lib.ptr(oa), _lib.ptr(raw), _lib.ptr(x1), 128, _lib.ptr(x2), 128, _lib.ptr(s1),
                                        _lib.ptr(s2), _lib.ptr(g1), _lib.ptr(b1), B, H, W, 32, 1e-5, 1, 0, 2, _lib.stream()))
            print(f"gn_apply +raw C128+128 fp32->fp16: {ms:.3f} ms  {8.0 * xa.numel() / ms / 1e6:.0f} GB/s")
            del x1, x2, raw
        xh = xa.half()
        shc = torch.zeros(C, device=dev)
        ms = t_ms(lambda: _lib.call("anoddpm_gn_apply_f16in", _lib.ptr(oa), _lib.ptr(xh), C, _lib.ptr(st), _lib.ptr(g1), _lib.ptr(b1),
                                    _lib.ptr(shc), None, 0, B, H, W, 32, 1e-5, 1, 2, _lib.stream()))
        print(f"gn_apply_f16in C{C} fp16->fp16: {ms:.3f} ms  {4.0 * xa.numel() / ms / 1e6:.0f} GB/s")
        del xa, oa, xh


if __name__ == "__main__":
    main()
