"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, host logic of the
drop-in modules (no GPU compute calls)."""
import os
import re

import numpy as np
import pytest
import torch

from tests.conftest import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "anoddpm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(anoddpm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    from anoddpm_b200 import _lib
    L = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/anoddpm_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(syms)
    assert L.anoddpm_abi_version() == 1
    assert b"tcgen05" in L.anoddpm_build_info()


def test_error_channel_without_gpu():
    from anoddpm_b200 import _lib
    L = _lib.lib()
    rc = L.anoddpm_q_sample(None, None, None, None, None, None, 1, 16, None)
    assert rc < 0 and b"null pointer" in L.anoddpm_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(rc)


def test_host_permutation_tables_match_golden(golden):
    from anoddpm_b200.simplex import _init, init_tables, overflow
    g = golden("simplex")
    for i, s in enumerate(g["seeds"]):
        perm, pgi3 = _init(int(s))
        assert perm.dtype == np.int64 and np.array_equal(perm, g["perm"][i].astype(np.int64))
        assert np.array_equal(pgi3, g["pgi3"][i].astype(np.int64))
    assert init_tables([3, 12345]).shape == (2, 512)
    assert overflow(2 ** 63) == -2 ** 63


def test_simplex_class_host_state_and_rng_stream(golden):
    from anoddpm_b200.simplex import Simplex_CLASS
    g = golden("simplex")
    np.random.seed(0)
    S = Simplex_CLASS()                       # consumes one draw, like the reference constructor
    p0, _ = __import__("anoddpm_b200.simplex", fromlist=["_init"])._init(int(g["np_seed0_draws"][0]))
    assert np.array_equal(S._perm, p0)
    S.newSeed()                               # second draw
    assert int(np.random.randint(-10000000000, 10000000000)) == int(g["np_seed0_draws"][2])
    S.newSeed(12345)
    assert list(S._perm[:8]) == [33, 182, 149, 37, 26, 75, 1, 19]


def test_diffusion_tables_and_schedule_match_golden(golden):
    from anoddpm_b200.GaussianDiffusion import GaussianDiffusionModel, get_beta_schedule
    g = golden("diffusion")
    for T, name in [(100, "linear"), (1000, "linear"), (100, "cosine"), (1000, "cosine")]:
        D = GaussianDiffusionModel([64, 64], get_beta_schedule(T, name))
        for n in ("betas", "sqrt_alphas", "sqrt_betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
                  "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
                  "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
                  "posterior_mean_coef1", "posterior_mean_coef2"):
            assert np.array_equal(getattr(D, n), g[f"{name}{T}/{n}"]), n
        assert D.num_timesteps == T
    with pytest.raises(NotImplementedError):
        get_beta_schedule(10, "sqrt")
    D = GaussianDiffusionModel([64, 64], get_beta_schedule(10, "linear"), loss_weight="prop-t")
    assert list(D.weights) == list(range(10, 0, -1))


def test_unet_state_dict_layout_matches_reference(golden):
    import hashlib
    from anoddpm_b200.UNet import UNetModel
    g = golden("unet")
    cases = {"c1_64_b32_h2": dict(img_size=64, base_channels=32, n_heads=2),
             "s32_b32_attn168": dict(img_size=32, base_channels=32, attention_resolutions="16,8"),
             "s64_b64_hc32_rgb": dict(img_size=64, base_channels=64, n_head_channels=32, in_channels=3),
             "s128_b32_h4": dict(img_size=128, base_channels=32, n_heads=4)}
    for name, kw in cases.items():
        m = UNetModel(**kw)
        sd = m.state_dict()
        h = hashlib.sha1("\n".join(f"{k}:{tuple(v.shape)}" for k, v in sd.items()).encode()).hexdigest()
        assert h == str(g[f"{name}/keys_sha1"]), name
        assert sum(p.numel() for p in m.parameters()) == int(g[f"{name}/n_params"])
    with pytest.raises(ValueError):
        UNetModel(100, 32)
    m = UNetModel(64, 32)
    assert float(m.out[2].weight.abs().max()) == 0.0      # zero-initialised head like the reference
    assert m.attention_resolutions == "32,16,8" and m.num_heads == 1


def test_unet_fails_loudly_without_cuda():
    """no CPU path and no PyTorch-autograd fallback: inference and training forwards both refuse CPU tensors"""
    from anoddpm_b200.UNet import UNetModel
    m = UNetModel(32, 32)
    assert not hasattr(UNetModel, "allow_autograd_fallback") and not hasattr(m, "forward_torch")
    x, t = torch.zeros(1, 1, 32, 32), torch.zeros(1, dtype=torch.long)
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x, t)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x, t)


def test_ema_update():
    from anoddpm_b200.UNet import UNetModel, update_ema_params
    import copy
    a = UNetModel(32, 32)
    b = copy.deepcopy(a)
    for p in b.parameters():
        p.data.add_(1.0)
    update_ema_params(a, b, 0.9)
    w0 = UNetModel(32, 32).out[0].weight      # GroupNorm weight = 1 at init
    assert torch.allclose(a.out[0].weight, 0.9 * w0 + 0.1 * (w0 + 1))


def test_support_predicates_and_argument_checks_without_gpu():
    """Host-side contract of the tensor-core entry points: the predicates the engine plans with, workspace sizes, and the
    loud argument errors -- all decided before any CUDA call, so they run on the build container."""
    import ctypes
    from anoddpm_b200 import _lib
    L = _lib.lib()

    def desc(**kw):
        d = _lib.ConvDesc()
        d.B, d.H, d.W, d.Ca, d.Cr1, d.Cr2, d.Cout, d.ksize = 1, 256, 256, 128, 0, 0, 128, 3
        for k, v in kw.items():
            setattr(d, k, v)
        return d

    ok3 = lambda **kw: L.anoddpm_conv_tc3_supported(ctypes.byref(desc(**kw)))     # noqa: E731
    assert ok3() == 1 and ok3(operand_f16=1) == 1 and ok3(Ca=256, Cr1=128, Cr2=128) == 1
    assert ok3(H=8, W=8) == 0 and ok3(H=24, W=32) == 0            # 16x16 regions
    assert ok3(ksize=1) == 0                                      # 1x1 convs run on impl 2
    assert ok3(Ca=96, operand_f16=1) == 0 and ok3(Ca=96) == 1     # fp16 K-blocks are 64 channels, fp32 ones 32
    assert ok3(Cout=48) == 0 and ok3(Cout=160) == 0 and ok3(Cout=512) == 1
    res = ctypes.c_void_p(256)
    assert ok3(res=res, res_mode=1) == 1 and ok3(res=res, res_mode=3) == 1 and ok3(res=res, res_mode=2) == 0
    assert ok3(out_f16=1) == 1 and ok3(out_f16=1, res=res, res_mode=1) == 0
    ok2 = lambda **kw: L.anoddpm_conv_tc_supported(ctypes.byref(desc(**kw)))      # noqa: E731
    assert ok2(H=8, W=8, Ca=512, Cout=512) == 1 and ok2(ksize=1, H=16, W=16, Ca=512, Cout=1536) == 1
    assert ok2(Ca=20) == 0 and ok2(W=24, H=24) == 0

    assert L.anoddpm_attention_tc_supported(256, 512, 2) == 1 and L.anoddpm_attention_tc_supported(64, 512, 2) == 1
    assert L.anoddpm_attention_tc_supported(64, 512, 8) == 1            # head size 64 (num_head_channels=64)
    assert L.anoddpm_attention_tc_supported(1024, 256, 2) == 1          # T > 256: key-tiled tcgen05 kernel (32x32 maps)
    assert L.anoddpm_attention_tc_supported(1024, 256, 1) == 1 and L.anoddpm_attention_tc_supported(256, 512, 1) == 1
    assert L.anoddpm_attention_tc_supported(64, 512, 1) == 1            # single head of 512 channels (detection.py n_heads=1)
    assert L.anoddpm_attention_tc_supported(320, 256, 1) == 0           # T > 256 must be a multiple of 256
    assert L.anoddpm_attention_tc_supported(256, 500, 3) == 0
    assert L.anoddpm_attention_tc_workspace_bytes(32, 256, 512) == 3 * 32 * 256 * 512 * 2

    # argument errors carry a message and never reach the device
    d = desc(impl=3, res=res, res_mode=2)
    d.out, d.a, d.w = 256, 256, 256
    assert L.anoddpm_conv(ctypes.byref(d), None) < 0 and b"res_mode" in L.anoddpm_last_error()
    d = desc(impl=1, out_f16=1)
    d.out, d.a, d.w = 256, 256, 256
    assert L.anoddpm_conv(ctypes.byref(d), None) < 0 and b"fp16 output" in L.anoddpm_last_error()
    assert L.anoddpm_attention_tc(res, res, res, 1, 320, 256, 2, 0, None) < 0 and b"not supported" in L.anoddpm_last_error()
    assert L.anoddpm_gn_apply_f16in(res, res, 100, res, res, res, None, None, 0, 1, 8, 8, 32, 1e-5, 1, 2, None) < 0
    assert L.anoddpm_linear(res, 8, res, 8, res, None, 1, 8, 8, 7, None) < 0
    assert L.anoddpm_vlb_terms(None, None, None, None, None, None, None, None, 10, 1, 16, None) < 0
    assert L.anoddpm_conv_head_gn(res, res, res, res, res, res, None, 1, 1024, 8, 8, 1, 32, 1e-5, None) < 0
    assert b"at most" in L.anoddpm_last_error()


def test_round2_planning_entry_points_host_only():
    """Host-side planners added in round 2 (no GPU): separable-head support / workspace, split-K workspace of the small-map
    conv kernel (follows the pair-tile planner: bs 32 fills the machine with narrow tiles, tiny batches split K), argument
    errors of the new entry points."""
    import ctypes
    from anoddpm_b200 import _lib
    from anoddpm_b200._lib import ConvDesc
    L = _lib.lib()
    assert L.anoddpm_conv_head_sep_supported(128, 1) == 1 and L.anoddpm_conv_head_sep_supported(256, 1) == 1
    assert L.anoddpm_conv_head_sep_supported(128, 2) == 0 and L.anoddpm_conv_head_sep_supported(192, 1) == 0
    assert L.anoddpm_conv_head_sep_workspace_bytes(32, 256, 256) == 9 * 32 * 256 * 256 * 4

    def desc(B, H, W, Ca, Cout, ks=3, impl=2):
        d = ConvDesc()
        d.B, d.H, d.W, d.Ca, d.Cout, d.ksize, d.impl, d.operand_f16 = B, H, W, Ca, Cout, ks, impl, 1
        return d
    assert L.anoddpm_conv_workspace_bytes(ctypes.byref(desc(32, 8, 8, 512, 512))) == 0          # 64 narrow tiles: no split
    assert L.anoddpm_conv_workspace_bytes(ctypes.byref(desc(32, 16, 16, 512, 1536, ks=1))) == 0  # qkv 1x1: plenty of tiles
    need = L.anoddpm_conv_workspace_bytes(ctypes.byref(desc(4, 8, 8, 256, 256)))
    assert need > 0 and need % 4 == 0
    assert L.anoddpm_conv_workspace_bytes(ctypes.byref(desc(4, 8, 8, 256, 256, impl=3))) == 0   # split-K is impl 2 only
    assert L.anoddpm_conv_workspace_bytes(None) == 0
    # argument errors (checked before any CUDA call)
    assert L.anoddpm_gather_rows(None, None, None, 1, 8, 10, None) < 0
    assert L.anoddpm_conv_head_sep(None, None, None, None, None, None, 1, 128, 8, 8, None) < 0
    assert b"anoddpm_conv_head_sep" in L.anoddpm_last_error()
