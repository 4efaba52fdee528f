"""Library comparator (optional row): the reference's UNet forward expressed as stock PyTorch calls
(oracle.unet_oracle.forward = the reference's own GPU path: cuDNN / cuBLAS convolutions under PyTorch's default TF32 setting,
F.group_norm, F.silu, softmax attention) timed on the same B200 next to the native engine, at the benched configuration
(256^2, base 128, attn "16,8", 2 heads).  Not a parity test and not the reported baseline (that is the CPU run of bench.py):
it prints the practical bar a hand-written kernel path has to beat and checks that the two agree."""
import time

import pytest
import torch

pytestmark = pytest.mark.gpu


def _time(fn, n):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        y = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, y


def test_library_comparator_unet_forward_bs32():
    from anoddpm_b200.UNet import UNetModel
    from oracle import unet_oracle as UO
    kw = dict(img_size=256, base_channels=128, n_heads=2, attention_resolutions="16,8")
    m = UNetModel(**kw).eval()
    m.load_state_dict(UO.fill_state_dict(m.state_dict(), seed=1))
    m = m.cuda()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    B = 32
    x = torch.randn(B, 1, 256, 256, generator=torch.Generator().manual_seed(0)).cuda()
    t = torch.full((B,), 500, device="cuda", dtype=torch.int64)
    with torch.no_grad():
        ms_native, y_native = _time(lambda: m(x, t), 5)
        lib = lambda: UO.forward(sd, x, t, 256, 128, attention_resolutions="16,8", n_heads=2)      # noqa: E731
        ms_tf32, y_tf32 = _time(lib, 3)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            ms_fp32, y_fp32 = _time(lib, 2)
        finally:
            torch.backends.cudnn.allow_tf32 = True
    scale = float(y_fp32.abs().max())
    e_native = float((y_native - y_fp32).abs().max()) / scale
    e_tf32 = float((y_tf32 - y_fp32).abs().max()) / scale
    print(f"[comparator] UNet forward bs 32 @256^2: native engine {ms_native:.2f} ms | stock PyTorch (cuDNN/cuBLAS) TF32 "
          f"{ms_tf32:.2f} ms, fp32 {ms_fp32:.2f} ms | error vs the fp32 library run: native {e_native:.2e}, "
          f"library TF32 {e_tf32:.2e}")
    assert e_native < 1e-3
    assert ms_native < ms_tf32       # the hand-written path must beat the library path it replaces
