"""GPU parity of the native backward kernels (training step, SURVEY 8(a) row D11 / BASELINE configs[1]) against torch
autograd in fp32 on the same operands."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


WG_SHAPES = [  # B, H, W, Cin, Cout, ksize
    (2, 16, 16, 128, 128, 3), (1, 8, 8, 256, 128, 3), (2, 32, 32, 128, 256, 1), (1, 64, 64, 128, 128, 3),
    (3, 8, 8, 128, 384, 1), (2, 16, 32, 256, 256, 3), (8, 256, 256, 128, 128, 3),
]


@pytest.mark.parametrize("dx_boxes", [0, 1], ids=["halo-box", "box-per-tap"])
@pytest.mark.parametrize("shape", WG_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv_wgrad_vs_torch(shape, dx_boxes, monkeypatch):
    """dW of conv3x3(pad 1) / conv1x1 from NHWC fp16 operands (MN-major tcgen05 descriptors) vs torch's conv2d_weight on
    the same fp16 values in fp32."""
    from anoddpm_b200 import _lib
    B, H, W, Cin, Cout, k = shape
    if k == 1 and dx_boxes:
        pytest.skip("1x1 has a single tap")
    if dx_boxes and (W >= 16 or B * H * W > 100000):
        pytest.skip("box-per-tap cross-check: 8-pixel-wide tiles only (three 16-pixel boxes per stage exceed shared memory)")
    monkeypatch.setenv("ANODDPM_WGRAD_DX_BOXES", str(dx_boxes))
    g = torch.Generator().manual_seed(sum(shape))
    a = torch.randn(B, H, W, Cin, generator=g).half().to(DEV)
    dy = (torch.randn(B, H, W, Cout, generator=g) * 0.5).half().to(DEV)
    assert _lib.lib().anoddpm_conv_wgrad_supported(B, H, W, Cin, Cout, k)
    dw = torch.zeros(Cout, k * k, Cin, device=DEV)
    alpha = 0.25
    _lib.call("anoddpm_conv_wgrad", _lib.ptr(dw), _lib.ptr(dy), _lib.ptr(a), B, H, W, Cin, Cout, k, alpha, _lib.stream())
    torch.cuda.synchronize()
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref = torch.nn.grad.conv2d_weight(a.float().permute(0, 3, 1, 2), (Cout, Cin, k, k), dy.float().permute(0, 3, 1, 2),
                                          padding=k // 2)
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    ref = ref.permute(0, 2, 3, 1).reshape(Cout, k * k, Cin) * alpha
    r = rel(dw, ref)
    print(f"[parity] conv_wgrad {shape} dx_boxes={dx_boxes}: rel={r:.3e}")
    assert r < 2e-4
