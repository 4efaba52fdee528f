#!/usr/bin/env python
"""Launch every non-conv kernel of the path once at the shape bench.py's aux_rooflines uses (for ncu captures).

  ncu --set full --clock-control none -k regex:'q_sample|p_sample|gn_apply|simplex3|attention_tc|attn_pack|adamw|sumsq|conv_head|conv_stem' \
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
    torch.cuda.synchronize()
    print(sorted(out))


if __name__ == "__main__":
    main()
