#!/usr/bin/env python
"""Launch every non-conv kernel of the path once at the shape bench.py's aux_rooflines uses (for ncu captures).

  ncu --set full --clock-control none -k regex:'q_sample|p_sample|gn_apply|simplex3|attention_tc|attn_pack|adamw|sumsq|conv_head|conv_stem|head_proj|head_gather|gather_rows' \
      -f -o gpurun_out/aux python tools/aux_kernels.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench          # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    bench._time_ms = lambda fn, n=1, warm=0: (fn(), torch.cuda.synchronize(), 1.0)[2]     # one launch each, no timing
    out = bench.aux_rooflines(dev)
    # stem + fused head at the BASELINE shape
    from anoddpm_b200 import _lib
    x = torch.randn(32, 1, 256, 256, device=dev)
    w = torch.randn(128, 1, 3, 3, device=dev)
    b = torch.zeros(128, device=dev)
    h = torch.empty(32, 256, 256, 128, device=dev)
    _lib.call("anoddpm_conv_stem", _lib.ptr(h), _lib.ptr(x), _lib.ptr(w), _lib.ptr(b), 32, 1, 256, 256, 128, _lib.stream())
    st = torch.zeros(8, 32, 128, 2, device=dev, dtype=torch.float64)
    _lib.call("anoddpm_gn_stats", _lib.ptr(st), _lib.ptr(h), 128, 32, 65536, _lib.stream())
    wp = torch.randn(1, 9 * 128, device=dev)
    o = torch.empty(32, 1, 256, 256, device=dev)
    g1, b1 = torch.ones(128, device=dev), torch.zeros(128, device=dev)
    _lib.call("anoddpm_conv_head_gn", _lib.ptr(o), _lib.ptr(h), _lib.ptr(st), _lib.ptr(g1), _lib.ptr(b1), _lib.ptr(wp), None, 32, 128,
              256, 256, 1, 32, 1e-5, _lib.stream())
    # round 2: fused stem (features + statistics + fp16 copy), separable head (projection + gather), embedding-row gather
    h16 = torch.empty(32, 256, 256, 128, device=dev, dtype=torch.float16)
    st2 = torch.zeros(8, 32, 128, 2, device=dev, dtype=torch.float64)
    _lib.call("anoddpm_conv_stem_fused", _lib.ptr(h), _lib.ptr(h16), _lib.ptr(st2), _lib.ptr(x), _lib.ptr(w), _lib.ptr(b), 32, 1, 256, 256,
              128, _lib.stream())
    tab = torch.empty(32, 128, 2, device=dev)
    _lib.call("anoddpm_gn_table", _lib.ptr(tab), _lib.ptr(st2), 128, None, 0, _lib.ptr(g1), _lib.ptr(b1), None, None, 0, 32, 65536, 32, 1e-5,
              _lib.stream())
    ws = torch.empty(9 * 32 * 65536, device=dev)
    _lib.call("anoddpm_conv_head_sep", _lib.ptr(o), _lib.ptr(h), _lib.ptr(tab), _lib.ptr(wp), None, _lib.ptr(ws), 32, 128, 256, 256,
              _lib.stream())
    table = torch.randn(1000, 7168, device=dev)
    t = torch.randint(0, 1000, (32,), device=dev)
    rows = torch.empty(32, 7168, device=dev)
    _lib.call("anoddpm_gather_rows", _lib.ptr(rows), _lib.ptr(table), _lib.ptr(t), 32, 7168, 1000, _lib.stream())
    torch.cuda.synchronize()
    print(sorted(out))


if __name__ == "__main__":
    main()
