"""Round-2 parity tests (VERDICT r01 "next round" item 1 and the advisor's regressions), all through the C ABI:

  (a) the BENCHED configuration -- BASELINE model (256^2, base 128, attn "16,8", 2 heads), default "tc" path (fp16 operands,
      fp16 conv1 storage), batch > 1 -- over a multi-step reverse chain with replayed simplex seeds against the CPU oracle,
      per-step error growth printed; batch invariance (sample b of a batched forward == its bs-1 forward);
  (c) golden-backed p_loss / calc_loss / random_noise / simplex_randParam / calc_total_vlb (tests/golden/loss.npz, produced
      by the unmodified reference);
  (d) fp16 exponent range of the residual stream (x scaled by 1e3 and 1e-4) through ResBlocks with a fused 1x1 skip conv;
  advisor: stale packed weights after update_ema_params / FusedAdamWEMA.step, optimiser state_dict round trip, gradient
  re-binding after zero_grad(set_to_none=True), out-of-range timesteps.

Tolerance rule: measured on the B200 (printed by each test) x 1.5, never above the north star's 1e-3 where that bar applies.
"""
import copy
import random

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def build_model(kw, impl="tc", seed=1):
    from anoddpm_b200.UNet import UNetModel
    from oracle.unet_oracle import fill_state_dict
    m = UNetModel(**kw).eval()
    m.load_state_dict(fill_state_dict(m.state_dict(), seed=seed))
    m.conv_impl = impl
    return m.to(DEV)


BASE256 = dict(img_size=256, base_channels=128, n_heads=2, attention_resolutions="16,8")


# ---------------------------------------------------------------------------------------------- (a) benched configuration
def test_base256_batched_multistep_chain_vs_oracle():
    """bs 4, 256^2, tc/f16: q_sample to t=9 then 10 reverse steps with replayed per-sample simplex seeds; the CPU oracle
    (torch fp32 UNet + C simplex + diffusion restatement) runs the same chain.  North-star bar: 1e-3 relative."""
    from anoddpm_b200.GaussianDiffusion import GaussianDiffusionModel, get_beta_schedule, simplex_noise_from_seeds
    from oracle import diffusion_oracle as DO, unet_oracle as UO
    from oracle.simplex_oracle import OracleSimplex
    torch.set_num_threads(max(1, torch.get_num_threads()))
    B, td = 4, 10
    m = build_model(BASE256, "tc")
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    D = GaussianDiffusionModel([256, 256], get_beta_schedule(1000, "linear"), noise="simplex")
    x0 = torch.rand(B, 1, 256, 256, generator=torch.Generator().manual_seed(40)) * 2 - 1
    seeds = [[1000 * s + 17 * b + 3 for b in range(B)] for s in range(td + 1)]

    # GPU: step-by-step through the public API so that every intermediate x_t is visible
    it = iter(seeds)
    D.noise_fn = lambda x, t: simplex_noise_from_seeds(next(it), tuple(x.shape), t)
    seq = D.forward_backward(m, x0.to(DEV), "half", td, "noise_fn")
    assert len(seq) == td + 2
    assert m.engine().f16_convs >= 100 and m.engine().tc_convs >= 100

    tab = DO.Tables(DO.beta_schedule(1000, "linear"))
    it2 = iter(seeds)

    def oracle_noise(x, t):
        ss = next(it2)
        return torch.stack([torch.from_numpy(OracleSimplex(s).rand_3d_fixed_T_octaves((256, 256), np.array([int(tt)]), 6, 0.8, 64)
                                             .astype(np.float32)) for s, tt in zip(ss, t)])

    ref = DO.forward_backward(tab, lambda x, t: UO.forward(sd, x, t, **BASE256), x0, td, oracle_noise, oracle_noise, "half")
    errs = [rel(a, b) for a, b in zip(seq, ref)]
    print("[parity] base256 bs4 tc chain, per-step rel err:", " ".join(f"{e:.2e}" for e in errs))
    assert errs[1] == 0.0                       # q_sample + simplex are bit-exact
    assert max(errs) < 1e-3, errs


@pytest.mark.parametrize("impl", ["exact", "tc"])
def test_base256_batch_invariance(impl):
    """sample b of a bs-8 forward against its own bs-1 forward (tile runs that cross sample boundaries, register-held
    statistics flushed per sample, 8 statistic replicas).  The two runs add up the GroupNorm statistics in a different
    order (1e-7 relative); the exact path carries that to ~1.5e-6 at the output.  On the tensor-core paths the same
    perturbation moves a few operands across a 10-bit rounding boundary in every layer, so the two runs are two realisations
    of the path's own rounding noise: measured 4-6e-4 on tc / tc_tf32 / tc2 alike (tools/bisect_batch_inv.py), the level
    of the path's error against the fp32 reference -- a statistics or tiling bug would be orders of magnitude above it."""
    m = build_model(BASE256, impl)
    g = torch.Generator().manual_seed(8)
    x = torch.randn(8, 1, 256, 256, generator=g).to(DEV)
    t = torch.tensor([999, 0, 431, 7, 250, 800, 1, 500], device=DEV)
    with torch.no_grad():
        yb = m(x, t)
        worst = 0.0
        for b in (0, 3, 7):
            y1 = m(x[b:b + 1], t[b:b + 1])
            worst = max(worst, rel(yb[b:b + 1], y1))
    print(f"[parity] base256 batch invariance impl={impl} (bs 8 vs bs 1): {worst:.3e}")
    assert worst < (5e-6 if impl == "exact" else 1e-3)


# ---------------------------------------------------------------------------------------------- (c) loss path goldens
SMALL = dict(img_size=32, base_channels=32, n_heads=2)


@pytest.mark.parametrize("impl", ["exact", "tc"])
def test_calc_loss_and_p_loss_vs_reference_golden(golden, impl, monkeypatch):
    from anoddpm_b200.GaussianDiffusion import GaussianDiffusionModel, get_beta_schedule
    g = golden("loss")
    tol = 2e-5 if impl == "exact" else 3e-3
    m = build_model(SMALL, impl)
    x0, noise, t = (torch.from_numpy(g[k]).to(DEV) for k in ("x0", "noise", "t"))
    worst = 0.0
    for lt in ("l1", "l2", "hybrid", "other"):
        D = GaussianDiffusionModel([32, 32], get_beta_schedule(100, "linear"), loss_type=lt, noise="gauss")
        D.noise_fn = lambda x, tt: noise
        with torch.no_grad():
            loss, x_t, eps = D.calc_loss(m, x0, t)
        assert np.array_equal(x_t.cpu().numpy(), g[f"calc_loss/{lt}/x_t"])          # q_sample: bit-exact
        worst = max(worst, rel(eps.cpu(), torch.from_numpy(g[f"calc_loss/{lt}/eps"])))
        assert set(loss) == ({"loss", "vlb"} if lt == "hybrid" else {"loss"})
        for k in loss:
            worst = max(worst, rel(loss[k].cpu(), torch.from_numpy(g[f"calc_loss/{lt}/{k}"])))
    for lw in ("prop-t", "uniform"):
        D = GaussianDiffusionModel([32, 32], get_beta_schedule(100, "linear"), loss_type="l2", loss_weight=lw, noise="gauss")
        D.noise_fn = lambda x, tt: noise
        np.random.seed(11)
        tt, ww = D.sample_t_with_weights(3, torch.device(DEV))
        assert np.array_equal(tt.cpu().numpy(), g[f"p_loss/{lw}/t"]) and np.array_equal(ww.cpu().numpy(), g[f"p_loss/{lw}/w"])
        np.random.seed(11)
        with torch.no_grad():
            total, (ld, x_t, eps) = D.p_loss(m, x0, {"train_start": True, "sample_distance": 60})
        assert np.array_equal(x_t.cpu().numpy(), g[f"p_loss/{lw}/x_t"])
        worst = max(worst, rel(total.cpu(), torch.from_numpy(g[f"p_loss/{lw}/total"])),
                    rel(ld["loss"].cpu(), torch.from_numpy(g[f"p_loss/{lw}/loss"])))
    # loss_weight "none": t ~ torch.randint(0, min(sample_distance, T)) / (0, T) ON x_0.device (device RNG streams differ from
    # the CPU's, so the draw is replayed and its arguments are checked)
    D = GaussianDiffusionModel([32, 32], get_beta_schedule(100, "linear"), loss_type="l2", loss_weight="none", noise="gauss")
    D.noise_fn = lambda x, tt: noise
    real_randint = torch.randint
    for tag, args, hi in (("start", {"train_start": True, "sample_distance": 60}, 60),
                          ("full", {"train_start": False, "sample_distance": 60}, 100)):
        calls = []

        def fake(lo, high, size, device=None, **kw):
            calls.append((lo, high, tuple(size), str(device)))
            return torch.from_numpy(g[f"p_loss/none_{tag}/t"]).to(device)

        monkeypatch.setattr(torch, "randint", fake)
        with torch.no_grad():
            total, (ld, x_t, eps) = D.p_loss(m, x0, args)
        monkeypatch.setattr(torch, "randint", real_randint)
        assert calls == [(0, hi, (3,), str(x0.device))]
        assert np.array_equal(x_t.cpu().numpy(), g[f"p_loss/none_{tag}/x_t"])
        worst = max(worst, rel(total.cpu(), torch.from_numpy(g[f"p_loss/none_{tag}/total"])))
    print(f"[parity] calc_loss / p_loss impl={impl}: worst rel={worst:.3e}")
    assert worst < tol


def test_random_noise_and_randparam_rng_consumption_vs_reference_golden(golden):
    from anoddpm_b200.GaussianDiffusion import generate_simplex_noise, random_noise
    from anoddpm_b200.simplex import Simplex_CLASS
    g = golden("loss")
    random.seed(3), np.random.seed(4), torch.manual_seed(5)
    S = Simplex_CLASS()
    x1 = torch.zeros(1, 1, 32, 32, device=DEV)
    t1 = torch.tensor([7], device=DEV)
    for i, is_simplex in enumerate(g["random_noise/is_simplex"]):
        before = np.random.get_state()[1].copy(), np.random.get_state()[2]
        r = random_noise(S, x1, t1)
        after = np.random.get_state()[1], np.random.get_state()[2]
        advanced = not (np.array_equal(before[0], after[0]) and before[1] == after[1])
        assert advanced == bool(is_simplex), i                           # same branch as the reference, call by call
        assert r.shape == x1.shape and r.is_cuda
        if is_simplex:
            assert np.array_equal(r.cpu().numpy(), g["random_noise/out"][i])          # bit-exact simplex field
        else:
            assert 0.8 < float(r.std()) < 1.2                            # N(0,1) from the device RNG
    random.seed(9), np.random.seed(10)
    x2 = torch.zeros(1, 2, 16, 16, device=DEV)
    out = generate_simplex_noise(Simplex_CLASS(), x2, torch.tensor([33], device=DEV), True, in_channels=2)
    assert np.array_equal(out.cpu().numpy(), g["randparam/out"])
    assert random.random() == float(g["randparam/next_random"])          # one random.choice per channel was consumed
    assert int(np.random.randint(-10000000000, 10000000000)) == int(g["randparam/next_np"])


@pytest.mark.parametrize("impl", ["exact", "tc"])
def test_calc_total_vlb_vs_reference_golden(golden, impl, monkeypatch):
    from anoddpm_b200.GaussianDiffusion import GaussianDiffusionModel, get_beta_schedule
    g = golden("loss")
    m = build_model(SMALL, impl)
    x0 = torch.from_numpy(g["x0"]).to(DEV)
    T = 40
    D = GaussianDiffusionModel([32, 32], get_beta_schedule(T, "linear"), noise="gauss")
    bank = iter(torch.randn(T, 3, 1, 32, 32, generator=torch.Generator().manual_seed(31)).to(DEV))
    monkeypatch.setattr(torch, "randn_like", lambda x, *a, **k: next(bank))
    r = D.calc_total_vlb(x0, m, {"Batch_Size": 3})
    worst = {k: rel(v.cpu(), torch.from_numpy(g[f"total_vlb/{k}"])) for k, v in r.items() if k != "prior_vlb"}
    # prior_vlb = KL(q(x_T|x_0) || N(0,1)) ~ 1e-7 bits/dim here: a difference of nearly equal numbers, compared absolutely
    assert float((r["prior_vlb"].cpu() - torch.from_numpy(g["total_vlb/prior_vlb"])).abs().max()) < 1e-6
    print(f"[parity] calc_total_vlb impl={impl}:", {k: f"{v:.2e}" for k, v in worst.items()})
    assert set(r) == {"total_vlb", "prior_vlb", "vb", "x_0_mse", "mse"}
    assert max(worst.values()) < (5e-5 if impl == "exact" else 3e-3)


# ---------------------------------------------------------------------------------------------- (d) fp16 exponent range
@pytest.mark.parametrize("scale", [1e3, 1.0, 1e-4])
def test_resblock_with_fused_skip_conv_residual_stream_scale(scale):
    """The raw (un-normalised) residual stream is handed to the fused 1x1 skip conv as fp16: activations that are not
    O(1) -- x scaled by 1e3 and by 1e-4 -- must still match the fp32 torch module within the tensor-core tolerance."""
    from anoddpm_b200.UNet import ResBlock
    import torch.nn.functional as F
    from anoddpm_b200.engine import UNetEngine

    class Net(torch.nn.Module):        # two blocks: 64 -> 128 (1x1 skip conv) -> 128 (identity skip), 32x32 maps
        def __init__(self):
            super().__init__()
            self.b1 = ResBlock(64, 256, 0, out_channels=128)
            self.b2 = ResBlock(128, 256, 0, out_channels=128)

    torch.manual_seed(3)
    net = Net()
    for p in net.parameters():
        if float(p.abs().max()) == 0:
            p.data.uniform_(-1, 1).mul_(p[0].numel() ** -0.5)
    net = net.to(DEV).eval()
    x = (torch.randn(2, 64, 32, 32, generator=torch.Generator().manual_seed(5)) * scale).to(DEV)
    emb = torch.randn(2, 256, generator=torch.Generator().manual_seed(6)).to(DEV)
    with torch.no_grad():
        ref = net.b2(net.b1(x, emb), emb)
    # drive the engine's ResBlock routine directly (same code path as UNetModel.forward)
    eng = UNetEngine.__new__(UNetEngine)
    eng.launches = eng.fused_stats = eng.stats_passes = eng.tc_convs = eng.f16_convs = eng.f16_outputs = 0
    eng.use_tc, eng.tc_impl, eng.allow_f16, eng.h1_f16, eng.tc_min_channels = True, 3, True, True, 0
    eng.fuse_gn, eng.fuse_gn2, eng.gn_tables = True, True, 0
    eng.emb_w, eng.emb_b, eng.emb_off = [], [], 0
    y = eng.run_resblocks([net.b1, net.b2], x, emb)
    r = rel(y, ref)
    print(f"[parity] ResBlock chain with fused skip conv, residual stream x{scale:g}: rel={r:.3e}")
    assert eng.f16_convs >= 3 and eng.gn_tables >= 3       # GN2 of both blocks and GN1 of the second ran inside the convs
    assert r < 1.5e-3


# ---------------------------------------------------------------------------------------------- operand-side GroupNorm
def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("shape", [(2, 128, 64, 128, 32, 32), (1, 128, 0, 128, 64, 64), (3, 256, 128, 256, 16, 48),
                                   (1, 64, 0, 64, 256, 256)], ids=lambda s: "x".join(map(str, s)))
def test_conv_with_operand_side_groupnorm_vs_torch(shape):
    """anoddpm_conv(impl 3, gn_table): y = conv3x3(SiLU(GroupNorm(cat(x1, x2)))) + bias + residual from RAW fp16 copies of
    x1 / x2, with the fp16 copy of the output and its per-channel statistics from the epilogue."""
    import ctypes
    import torch.nn.functional as F
    from anoddpm_b200 import _lib
    from anoddpm_b200.engine import pack_conv_weight
    B, C1, C2, Cout, H, W = shape
    g = torch.Generator().manual_seed(sum(shape))
    x1 = (torch.randn(B, C1, H, W, generator=g) * 1.7 + 0.3).to(DEV)
    x2 = (torch.randn(B, C2, H, W, generator=g) * 0.6 - 0.2).to(DEV) if C2 else None
    C = C1 + C2
    gamma, beta = (torch.rand(C, generator=g) + 0.5).to(DEV), (torch.randn(C, generator=g) * 0.2).to(DEV)
    w = (torch.randn(Cout, C, 3, 3, generator=g) * (9 * C) ** -0.5).to(DEV)
    bias = torch.randn(Cout, generator=g).to(DEV)
    res = torch.randn(B, Cout, H, W, generator=g).to(DEV)
    REPL = 8
    x1n, x2n = _nhwc(x1), (_nhwc(x2) if C2 else None)
    st1 = torch.zeros(REPL * B * C1 * 2, device=DEV, dtype=torch.float64)
    _lib.call("anoddpm_gn_stats", _lib.ptr(st1), _lib.ptr(x1n), C1, B, H * W, _lib.stream())
    st2 = None
    if C2:
        st2 = torch.zeros(REPL * B * C2 * 2, device=DEV, dtype=torch.float64)
        _lib.call("anoddpm_gn_stats", _lib.ptr(st2), _lib.ptr(x2n), C2, B, H * W, _lib.stream())
    tab = torch.empty(B, C, 2, device=DEV)
    _lib.call("anoddpm_gn_table", _lib.ptr(tab), _lib.ptr(st1), C1, _lib.ptr(st2), C2, _lib.ptr(gamma), _lib.ptr(beta), None, None,
              0, B, H * W, 32, 1e-5, _lib.stream())
    a1, a2 = x1n.half(), (x2n.half() if C2 else None)
    wp = pack_conv_weight(w).contiguous().half()
    out = torch.empty(B, H, W, Cout, device=DEV)
    out16 = torch.empty(B, H, W, Cout, device=DEV, dtype=torch.float16)
    stats = torch.zeros(REPL * B * Cout * 2, device=DEV, dtype=torch.float64)
    resn = _nhwc(res)
    d = _lib.ConvDesc()
    d.out, d.a, d.w, d.bias = out.data_ptr(), a1.data_ptr(), wp.data_ptr(), bias.data_ptr()
    d.B, d.H, d.W, d.Ca, d.Cout, d.ksize = B, H, W, C1, Cout, 3
    if C2:
        d.a2, d.Ca2 = a2.data_ptr(), C2
    d.gn_table, d.out16, d.gn_stats = tab.data_ptr(), out16.data_ptr(), stats.data_ptr()
    d.res, d.res_mode = resn.data_ptr(), 1
    d.impl, d.operand_f16 = 3, 1
    assert _lib.lib().anoddpm_conv_tc3_supported(ctypes.byref(d))
    _lib.call("anoddpm_conv", ctypes.byref(d), _lib.stream())
    torch.cuda.synchronize()
    y = out.permute(0, 3, 1, 2)
    xc = torch.cat([x1] + ([x2] if C2 else []), dim=1)
    ref = F.conv2d(F.silu(F.group_norm(xc, 32, gamma, beta, 1e-5)), w, bias, padding=1) + res
    r = rel(y, ref)
    print(f"[parity] conv with operand-side GroupNorm {shape}: rel={r:.3e}")
    assert r < 1e-3
    assert torch.equal(out16, out.half())
    s = stats.view(REPL, B, Cout, 2).sum(0)
    assert torch.allclose(s[..., 0], out.double().sum(dim=(1, 2)), rtol=1e-6, atol=1e-3)
    assert torch.allclose(s[..., 1], (out.double() ** 2).sum(dim=(1, 2)), rtol=1e-6, atol=1e-3)


def test_base256_fused_groupnorm_path_matches_unfused_and_golden(golden):
    """BASELINE model: the default path (GroupNorm + SiLU inside the convs) against the reference golden and against the
    stand-alone-pass path (ANODDPM_FUSE_GN=0 equivalent)."""
    g = golden("unet")
    m = build_model(BASE256, "tc")
    x = torch.randn(1, 1, 256, 256, generator=torch.Generator().manual_seed(6)).to(DEV)
    t = torch.from_numpy(g["base256/t"]).to(DEV)
    crop = torch.from_numpy(g["base256/y_crop"]).to(DEV)
    scale = float(g["base256/y_stats"][2])
    errs = {}
    for fuse in (True, False):
        m.fuse_gn = fuse
        with torch.no_grad():
            y = m(x, t)
        errs[fuse] = float((y[0, 0, 100:132, 64:128] - crop).abs().max()) / scale
        eng = m.engine()
        if fuse:
            assert eng.gn_tables >= 50, eng.gn_tables
            launches_fused = eng.launches_per_forward
        else:
            assert eng.gn_tables <= 1       # only the separable head's table
    print(f"[parity] base256 tc: fused-GN rel={errs[True]:.3e}  separate-pass rel={errs[False]:.3e}  "
          f"launches {launches_fused} vs {eng.launches_per_forward}")
    assert errs[True] < 1e-3 and errs[False] < 1e-3


# ---------------------------------------------------------------------------------------------- advisor regressions
def test_engine_sees_ema_and_fused_optimiser_updates():
    """update_ema_params (.data-style foreach writes) and FusedAdamWEMA.step (raw-pointer kernels) do not bump
    p._version; the engine's packed weights / the CUDA-graph key must still follow them."""
    from anoddpm_b200.optim import FusedAdamWEMA
    from anoddpm_b200.UNet import update_ema_params
    from oracle import unet_oracle as UO
    m = build_model(SMALL, "exact")
    ema = copy.deepcopy(m)
    x = torch.randn(2, 1, 32, 32, generator=torch.Generator().manual_seed(1))
    t = torch.tensor([3, 70])

    def check(model, what):
        sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        with torch.no_grad():
            y = model(x.to(DEV), t.to(DEV)).cpu()
        r = rel(y, UO.forward(sd, x, t, **SMALL))
        assert r < 2e-5, (what, r)

    check(m, "fresh"), check(ema, "fresh ema")
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(1.05)
    update_ema_params(ema, m, 0.5)
    check(ema, "after update_ema_params")
    opt = FusedAdamWEMA(m, ema, lr=1e-2, weight_decay=0.0, max_norm=1.0, ema_decay=0.5)
    check(m, "after flattening")
    g = torch.Generator().manual_seed(2)
    for p in m.parameters():
        p.grad.copy_(torch.randn(p.shape, generator=g).to(DEV))
    opt.step()
    check(m, "after FusedAdamWEMA.step"), check(ema, "ema after FusedAdamWEMA.step")


def test_fused_optimiser_state_dict_and_grad_rebinding():
    from anoddpm_b200.optim import FusedAdamWEMA
    from anoddpm_b200.UNet import UNetModel
    torch.manual_seed(0)
    m1 = UNetModel(32, 32).to(DEV)
    for p in m1.parameters():
        if float(p.abs().max()) == 0:
            p.data.normal_(0, 0.02)
    m2, m3 = copy.deepcopy(m1), copy.deepcopy(m1)
    ref = torch.optim.AdamW(m1.parameters(), lr=1e-3, weight_decay=0.01)
    fused = FusedAdamWEMA(m2, None, lr=1e-3, weight_decay=0.01, max_norm=0.0)
    g = torch.Generator().manual_seed(1)

    def grads():
        return [torch.randn(p.shape, generator=g).to(DEV) * 0.01 for p in m1.parameters()]

    for it in range(2):
        gs = grads()
        m2.zero_grad(set_to_none=True)              # torch's default: detaches p.grad from the flat buffer
        for p1, p2, gr in zip(m1.parameters(), m2.parameters(), gs):
            p1.grad = gr.clone()
            p2.grad = gr.clone()                    # a fresh tensor, like autograd would assign
        ref.step()
        fused.step()
    for a, b in zip(m1.parameters(), m2.parameters()):
        assert torch.allclose(a, b, rtol=2e-5, atol=2e-7)
    # checkpoint: fused -> torch.optim.AdamW and fused -> fused, then one more identical step
    sd = fused.state_dict()
    ref2 = torch.optim.AdamW(m2.parameters(), lr=1e-3, weight_decay=0.01)
    ref2.load_state_dict(copy.deepcopy(sd))                        # torch accepts the layout
    m3.load_state_dict(m2.state_dict())
    fused3 = FusedAdamWEMA(m3, None, lr=5e-2, weight_decay=0.5, max_norm=0.0)
    fused3.load_state_dict(sd)
    assert fused3.steps == 2 and fused3.lr == 1e-3 and fused3.weight_decay == 0.01
    gs = grads()
    m4 = copy.deepcopy(m2)          # plain tensors for the torch optimiser
    ref4 = torch.optim.AdamW(m4.parameters(), lr=1e-3, weight_decay=0.01)
    ref4.load_state_dict(copy.deepcopy(sd))
    for p4, p3, gr in zip(m4.parameters(), m3.parameters(), gs):
        p4.grad = gr.clone()
        p3.grad.copy_(gr)
    ref4.step()
    fused3.step()
    for a, b in zip(m4.parameters(), m3.parameters()):
        assert torch.allclose(a, b, rtol=2e-5, atol=2e-7)


def test_out_of_range_timesteps():
    from anoddpm_b200.GaussianDiffusion import GaussianDiffusionModel, get_beta_schedule
    m = build_model(SMALL, "tc")
    D = GaussianDiffusionModel([32, 32], get_beta_schedule(100, "linear"), noise="gauss")
    x = torch.zeros(2, 1, 32, 32, device=DEV)
    with pytest.raises(IndexError):
        D.reverse_steps(m, x, 5, 7)                     # would run t down to -1
    with pytest.raises(IndexError):
        D.reverse_steps(m, x, 100, 3)                   # t >= T
    out = D.sample_p(lambda a, b: torch.zeros_like(a), x, torch.tensor([3, 100], device=DEV), denoise_fn=lambda a, b: torch.zeros_like(a))
    s = out["sample"]
    assert bool(torch.isfinite(s[0]).all()) and bool(torch.isnan(s[1]).all())     # poisoned, not an out-of-bounds read


def test_fused_stem_features_statistics_and_fp16_copy():
    """anoddpm_conv_stem_fused: Conv2d(1 -> C, 3x3) + the per-channel GroupNorm statistics of its output + the fp16 copy,
    one pass (UNet.py:279-281 followed by the first ResBlock's GroupNorm)."""
    import torch.nn.functional as F
    from anoddpm_b200 import _lib
    g = torch.Generator().manual_seed(12)
    for (B, H, W, C) in [(3, 32, 48, 128), (1, 256, 256, 128), (2, 16, 8, 64), (2, 16, 64, 128), (1, 8, 32, 64)]:
        x = torch.randn(B, 1, H, W, generator=g).to(DEV)
        w = (torch.randn(C, 1, 3, 3, generator=g) * 0.3).to(DEV)
        b = torch.randn(C, generator=g).to(DEV)
        out = torch.empty(B, H, W, C, device=DEV)
        out16 = torch.empty(B, H, W, C, device=DEV, dtype=torch.float16)
        st = torch.zeros(8, B, C, 2, device=DEV, dtype=torch.float64)
        _lib.call("anoddpm_conv_stem_fused", _lib.ptr(out), _lib.ptr(out16), _lib.ptr(st), _lib.ptr(x), _lib.ptr(w), _lib.ptr(b), B, 1, H, W,
                  C, _lib.stream())
        ref = F.conv2d(x.double(), w.double(), b.double(), padding=1).permute(0, 2, 3, 1)
        assert rel(out, ref.float()) < 1e-6
        assert torch.equal(out16, out.half())
        s = st.sum(0)
        assert torch.allclose(s[..., 0], ref.sum(dim=(1, 2)), rtol=1e-6, atol=1e-4)
        assert torch.allclose(s[..., 1], (ref ** 2).sum(dim=(1, 2)), rtol=1e-6, atol=1e-4)


@pytest.mark.parametrize("shape", [(2, 128, 32, 32), (3, 256, 24, 40), (1, 128, 256, 256)])
def test_separable_head_vs_torch(shape):
    """anoddpm_conv_head_sep (GroupNorm + SiLU + nine per-pixel channel reductions, then nine shifted adds) against the
    plain torch fp32 GroupNorm -> SiLU -> Conv2d(C -> 1) of reference UNet.py:384-388 and against anoddpm_conv_head_gn."""
    from anoddpm_b200 import _lib
    from anoddpm_b200._lib import call, ptr, stream
    from anoddpm_b200.engine import pack_conv_weight
    B, C, H, W = shape
    g = torch.Generator(device="cpu").manual_seed(5)
    x = (torch.randn(B, C, H, W, generator=g) * 1.7 + 0.3).cuda()
    gamma = (torch.rand(C, generator=g) + 0.5).cuda()
    beta = (torch.randn(C, generator=g) * 0.2).cuda()
    w = (torch.randn(1, C, 3, 3, generator=g) * 0.05).cuda()
    bias = torch.randn(1, generator=g).cuda()
    ref = F.conv2d(F.silu(F.group_norm(x.double(), 32, gamma.double(), beta.double(), 1e-5)), w.double(), bias.double(),
                   padding=1).float()       # fp64: torch's fp32 GPU convolution runs on TF32 by default
    xn = x.permute(0, 2, 3, 1).contiguous()
    st = torch.zeros(8 * B * C * 2, device="cuda", dtype=torch.float64)
    call("anoddpm_gn_stats", ptr(st), ptr(xn), C, B, H * W, stream())
    tab = torch.empty(B, C, 2, device="cuda")
    call("anoddpm_gn_table", ptr(tab), ptr(st), C, None, 0, ptr(gamma), ptr(beta), None, None, 0, B, H * W, 32, 1e-5, stream())
    wp = pack_conv_weight(w).contiguous().float()
    L = _lib.lib()
    assert L.anoddpm_conv_head_sep_supported(C, 1) == 1 and L.anoddpm_conv_head_sep_supported(192, 1) == 0
    ws = torch.empty(L.anoddpm_conv_head_sep_workspace_bytes(B, H, W) // 4, device="cuda")
    out = torch.empty(B, 1, H, W, device="cuda")
    call("anoddpm_conv_head_sep", ptr(out), ptr(xn), ptr(tab), ptr(wp), ptr(bias), ptr(ws), B, C, H, W, stream())
    old = torch.empty(B, 1, H, W, device="cuda")
    call("anoddpm_conv_head_gn", ptr(old), ptr(xn), ptr(st), ptr(gamma), ptr(beta), ptr(wp), ptr(bias), B, C, H, W, 1, 32, 1e-5,
         stream())
    torch.cuda.synchronize()
    e_ref = float((out - ref).abs().max() / ref.abs().max())
    e_old = float((out - old).abs().max() / ref.abs().max())
    print(f"separable head {shape}: vs torch {e_ref:.2e}, vs conv_head_gn {e_old:.2e}")
    assert e_ref < 5e-6 and e_old < 5e-6        # fp32 both ways: summation order + __expf only (measured ~1e-6)


def test_graph_loop_switches_leave_the_chain_unchanged(monkeypatch):
    """The captured reverse step with its optional pieces toggled (per-timestep embedding table, noise field on a forked
    stream, noise generated inside the posterior kernel) must produce the chain of the plain configuration: the table rows are bit-identical to the per-step
    computation and the noise kernel does not depend on where it runs."""
    from anoddpm_b200 import GaussianDiffusionModel, get_beta_schedule
    m = build_model(dict(img_size=64, base_channels=64, n_heads=2, attention_resolutions="16,8"), "tc")
    x0 = (torch.rand(2, 1, 64, 64, generator=torch.Generator().manual_seed(3)) * 2 - 1).to(DEV)
    outs = {}
    for name, env in (("plain", {"ANODDPM_TIME_TABLE": "0", "ANODDPM_NOISE_OVERLAP": "0", "ANODDPM_FUSED_NOISE": "0"}),
                      ("table", {"ANODDPM_TIME_TABLE": "1", "ANODDPM_NOISE_OVERLAP": "0", "ANODDPM_FUSED_NOISE": "0"}),
                      ("table+overlap", {"ANODDPM_TIME_TABLE": "1", "ANODDPM_NOISE_OVERLAP": "1", "ANODDPM_FUSED_NOISE": "0"}),
                      ("table+fused noise", {"ANODDPM_TIME_TABLE": "1", "ANODDPM_NOISE_OVERLAP": "0", "ANODDPM_FUSED_NOISE": "1"})):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        diff = GaussianDiffusionModel([64, 64], get_beta_schedule(1000, "linear"), noise="simplex")
        diff.simplex_seed_source = lambda step, n: [1000 + 17 * step + i for i in range(n)]
        with torch.no_grad():
            outs[name] = diff.reverse_steps(m, x0.clone(), 40, 6, "simplex")
    assert torch.isfinite(outs["plain"]).all()
    assert torch.equal(outs["plain"], outs["table"])
    assert torch.equal(outs["plain"], outs["table+overlap"])
    assert torch.equal(outs["plain"], outs["table+fused noise"])      # noise generated inside the posterior kernel: same bits
