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


# ------------------------------------------------------------------------------------------- per-kernel backward checks
def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("case", [(2, 128, 0, 16, 16, 0, True), (2, 64, 64, 16, 16, 0, True), (1, 128, 0, 32, 32, 1, True),
                                  (2, 128, 0, 8, 8, 2, True), (2, 96, 0, 16, 16, 0, False)],
                         ids=lambda c: "x".join(map(str, c)))
def test_gn_silu_resample_backward_vs_autograd(case):
    """anoddpm_gn_silu_bwd (fp32 sources, optional concat, avg-pool / upsample modes, skip-path addend) vs torch autograd."""
    from anoddpm_b200 import _lib
    from anoddpm_b200.train import _gn_bwd, _stats
    B, C1, C2, H, W, mode, silu = case
    C = C1 + C2
    g = torch.Generator().manual_seed(sum(case[:5]))
    x1 = (torch.randn(B, C1, H, W, generator=g) * 1.5 + 0.2).to(DEV).requires_grad_()
    x2 = (torch.randn(B, C2, H, W, generator=g) * 0.7).to(DEV).requires_grad_() if C2 else None
    gn = torch.nn.GroupNorm(32, C).to(DEV)
    with torch.no_grad():
        gn.weight.copy_(torch.rand(C, generator=g) + 0.5)
        gn.bias.copy_(torch.randn(C, generator=g) * 0.2)
    xc = torch.cat([x1, x2], dim=1) if C2 else x1
    y = gn(xc)
    if silu:
        y = F.silu(y)
    if mode == 1:
        y = F.avg_pool2d(y, 2)
    elif mode == 2:
        y = F.interpolate(y, scale_factor=2, mode="nearest")
    Ho, Wo = y.shape[-2:]
    da = torch.randn(B, C, Ho, Wo, generator=g).to(DEV)
    # skip-path addend at the resolution the skip path sees
    add_mode = {0: 1, 1: 2, 2: 3}[mode]
    add = torch.randn(B, C, Ho, Wo, generator=g).to(DEV)
    skip = xc if mode == 0 else (F.avg_pool2d(xc, 2) if mode == 1 else F.interpolate(xc, scale_factor=2, mode="nearest"))
    loss = (y * da).sum() + (skip * add).sum()
    loss.backward()
    x1n, x2n = _nhwc(x1.detach()), (_nhwc(x2.detach()) if C2 else None)
    st1 = _stats(x1n, B, H * W)
    st2 = _stats(x2n, B, H * W) if C2 else None
    alpha = 0.5
    dx1, dx2, dg, db, _ = _gn_bwd(x1n, False, st1, x2n, st2, gn, None, None, 0, _nhwc(da).half(), True, _nhwc(add), add_mode, False,
                                  B, H, W, silu, mode, alpha)
    torch.cuda.synchronize()
    e = [rel(dx1.permute(0, 3, 1, 2), x1.grad), rel(dg, gn.weight.grad * alpha), rel(db, gn.bias.grad * alpha)]
    if C2:
        e.append(rel(dx2.permute(0, 3, 1, 2), x2.grad))
    print(f"[parity] gn_silu_bwd {case}: dx1 {e[0]:.2e} dgamma {e[1]:.2e} dbeta {e[2]:.2e}" + (f" dx2 {e[3]:.2e}" if C2 else ""))
    assert max(e) < 2e-3          # da enters as fp16 (2^-11 relative)


def test_batched_gemm_and_softmax_rows_vs_torch():
    from anoddpm_b200 import _lib
    from anoddpm_b200.train import _gemm
    g = torch.Generator().manual_seed(3)
    A = torch.randn(3, 2, 70, 45, generator=g).to(DEV)
    Bm = torch.randn(3, 2, 45, 130, generator=g).to(DEV)
    C = torch.empty(3, 2, 70, 130, device=DEV)
    _gemm(C, A, Bm, 70, 130, 45, 45, 1, 130, 1, 130, 0.5, n1=3, n2=2, sa=(2 * 70 * 45, 70 * 45), sb=(2 * 45 * 130, 45 * 130),
          sc=(2 * 70 * 130, 70 * 130))
    assert rel(C, 0.5 * A @ Bm) < 1e-5
    At = A.transpose(2, 3).contiguous()          # A^T stored: element (m, k) at k * 70 + m
    C2 = torch.zeros_like(C)
    _gemm(C2, At, Bm, 70, 130, 45, 1, 70, 130, 1, 130, 1.0, n1=3, n2=2, sa=(2 * 70 * 45, 70 * 45), sb=(2 * 45 * 130, 45 * 130),
          sc=(2 * 70 * 130, 70 * 130))
    assert rel(C2, A @ Bm) < 1e-5
    s = torch.randn(37, 200, generator=g).to(DEV)
    p = s.clone()
    _lib.call("anoddpm_softmax_rows", _lib.ptr(p), 37, 200, _lib.stream())
    assert rel(p, torch.softmax(s, dim=-1)) < 1e-6
    dp = torch.randn(37, 200, generator=g).to(DEV)
    ref = p * (dp - (dp * p).sum(-1, keepdim=True))
    _lib.call("anoddpm_softmax_bwd_rows", _lib.ptr(dp), _lib.ptr(p), 37, 200, _lib.stream())
    assert rel(dp, ref) < 1e-5


# ------------------------------------------------------------------------------------------- the whole training step
TRAIN_KW = dict(img_size=64, base_channels=64, n_heads=2, attention_resolutions="16,8")
TRAIN_KW_128 = dict(img_size=64, base_channels=128, n_heads=2, attention_resolutions="16,8")     # BASELINE channel widths


def _fresh_model(seed=1, kw=TRAIN_KW):
    from anoddpm_b200.UNet import UNetModel
    from oracle.unet_oracle import fill_state_dict
    m = UNetModel(**kw)
    m.load_state_dict(fill_state_dict(m.state_dict(), seed=seed))
    return m.to(DEV).train()


@pytest.mark.parametrize("TRAIN_KW", [TRAIN_KW, TRAIN_KW_128], ids=["base64", "base128"])
def test_train_step_matches_reference(TRAIN_KW):
    """Gradients of mean((eps_hat - noise)^2) wrt EVERY parameter: the native forward + hand-written backward against torch
    autograd through the fp32 oracle forward (same weights, same inputs; TF32 off)."""
    from oracle import unet_oracle as UO
    m = _fresh_model(kw=TRAIN_KW)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, 1, 64, 64, generator=g).to(DEV)
    t = torch.tensor([5, 500, 999], device=DEV)
    noise = torch.randn(3, 1, 64, 64, generator=g).to(DEV)
    y = m(x, t)
    assert y.requires_grad
    loss = ((y - noise) ** 2).mean()
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    assert all(v is not None for v in grads.values())
    # reference: functional oracle forward on leaf copies of the parameters
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point) for k, v in m.state_dict().items()}
        y_ref = UO.forward(sd, x, t, **TRAIN_KW)
        loss_ref = ((y_ref - noise) ** 2).mean()
        loss_ref.backward()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    print(f"[parity] train step: forward rel={rel(y.detach(), y_ref.detach()):.3e} loss {float(loss):.6f} vs {float(loss_ref):.6f}")
    errs = sorted(((rel(gq, sd[k].grad), float((gq - sd[k].grad).norm() / sd[k].grad.norm().clamp_min(1e-30)), k)
                   for k, gq in grads.items()), reverse=True)
    worst, worst_l2, worst_k = errs[0]
    l2s = sorted(e[1] for e in errs)
    print(f"[parity] train step: {len(errs)} parameter gradients, max-norm rel: median {errs[len(errs) // 2][0]:.3e} worst "
          f"{worst:.3e} ({worst_k}); L2 rel: median {l2s[len(l2s) // 2]:.3e} worst {l2s[-1]:.3e}")
    for e in errs[:12]:
        print(f"[parity]    {e[2]}: max-norm {e[0]:.3e}  L2 {e[1]:.3e}")
    # total gradient (what clip_grad_norm_ and AdamW see)
    flat = torch.cat([grads[k].reshape(-1) for k in grads])
    flat_ref = torch.cat([sd[k].grad.reshape(-1) for k in grads])
    tot = float((flat - flat_ref).norm() / flat_ref.norm())
    print(f"[parity] train step: whole gradient vector L2 rel={tot:.3e}, norm {float(flat.norm()):.6f} vs {float(flat_ref.norm()):.6f}")
    assert rel(y.detach(), y_ref.detach()) < 2e-3
    assert tot < 2e-3 and worst < 1e-2, (tot, worst, worst_k)


def test_p_loss_backward_and_optimizer_step_run_natively():
    """diffusion_training.py:99-107 on the native path: p_loss -> backward -> FusedAdamWEMA.step, loss decreases on a fixed batch."""
    import copy
    from anoddpm_b200.GaussianDiffusion import GaussianDiffusionModel, get_beta_schedule
    from anoddpm_b200.optim import FusedAdamWEMA
    m = _fresh_model(seed=2)
    ema = copy.deepcopy(m)
    opt = FusedAdamWEMA(m, ema, lr=2e-4, weight_decay=0.0, max_norm=1.0)
    D = GaussianDiffusionModel([64, 64], get_beta_schedule(1000, "linear"), noise="gauss")
    x = (torch.rand(4, 1, 64, 64, generator=torch.Generator().manual_seed(0)) * 2 - 1).to(DEV)
    noise = torch.randn(4, 1, 64, 64, generator=torch.Generator().manual_seed(1)).to(DEV)
    D.noise_fn = lambda a, b: noise
    t = torch.tensor([10, 300, 600, 900], device=DEV)
    losses = []
    for _ in range(6):
        opt.zero_grad()
        ld, x_t, eps = D.calc_loss(m, x, t)
        loss = ld["loss"].mean()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    print("[parity] native training losses:", " ".join(f"{v:.4f}" for v in losses))
    assert all(l == l for l in losses) and losses[-1] < losses[0]
