"""numpy/torch oracles of the diffusion maths and the UNet vs goldens from the unmodified reference."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import diffusion_oracle as DO
from oracle import unet_oracle as UO

CASES = [(100, "linear"), (1000, "linear"), (100, "cosine"), (1000, "cosine")]


@pytest.mark.parametrize("T,name", CASES)
def test_tables_and_steps_bit_exact(golden, T, name):
    g = golden("diffusion")
    tab = DO.Tables(DO.beta_schedule(T, name))
    k = f"{name}{T}/"
    for n in DO.Tables.NAMES:
        assert np.array_equal(getattr(tab, n), g[k + n]), n
    x, noise, eps, t = (torch.from_numpy(g[k + n]) for n in ("x", "noise", "eps", "t"))
    assert np.array_equal(DO.sample_q(tab, x, t, noise).numpy(), g[k + "sample_q"])
    assert np.array_equal(DO.sample_q_gradual(tab, x, t, noise).numpy(), g[k + "sample_q_gradual"])
    pmv = DO.p_mean_variance(tab, x, t, eps)
    for kk, v in pmv.items():
        assert np.array_equal(v.contiguous().numpy(), g[k + "pmv_" + kk]), kk
    sp = DO.sample_p(tab, x, t, eps, noise)
    assert np.array_equal(sp["sample"].numpy(), g[k + "sample_p"])
    assert np.array_equal(sp["pred_x_0"].numpy(), g[k + "sample_p_pred_x_0"])


@pytest.mark.parametrize("T,name", [(100, "linear"), (1000, "linear"), (1000, "cosine")])
def test_vlb_helpers_bit_exact(golden, T, name):
    """normal_kl / discretised log-likelihood / calc_vlb_xt / prior_vlb of the oracle == the unmodified reference."""
    g = golden("vlb")
    tab = DO.Tables(DO.beta_schedule(T, name))
    k = f"{name}{T}/"
    x0, x_t, eps, t = (torch.from_numpy(g[k + n]) for n in ("x0", "x_t", "eps", "t"))
    r = DO.vlb_xt(tab, x0, x_t, t, eps)
    assert np.array_equal(r["output"].numpy(), g[k + "vlb"])
    assert np.array_equal(r["pred_x_0"].numpy(), g[k + "pred_x_0"])
    pmv = DO.p_mean_variance(tab, x_t, t, eps)
    tm = DO.coef(tab.posterior_mean_coef1, t) * x0 + DO.coef(tab.posterior_mean_coef2, t) * x_t
    tlv = DO.coef(tab.posterior_log_variance_clipped, t).expand(x_t.shape)
    assert np.array_equal(DO.normal_kl(tm, tlv, pmv["mean"], pmv["log_variance"]).numpy(), g[k + "normal_kl"])
    assert np.array_equal(DO.disc_ll(x0, pmv["mean"], 0.5 * pmv["log_variance"]).numpy(), g[k + "disc_ll"])
    assert np.array_equal(DO.prior_vlb(tab, x0).numpy(), g[k + "prior_vlb"])
    assert float((x0 > 0.999).float().mean()) > 0 and float((x0 < -0.999).float().mean()) > 0   # both open-ended bins hit


def test_survey_table_sanity():
    tab = DO.Tables(DO.beta_schedule(100, "linear"))
    assert np.allclose(tab.sqrt_alphas_cumprod[:2], [0.99949987, 0.99799444], atol=1e-8)
    assert np.allclose(tab.posterior_mean_coef1[:2], [1.0, 0.7508179], atol=1e-7)
    assert np.allclose(tab.posterior_log_variance_clipped[:2], [-7.19384716] * 2, atol=1e-8)
    with pytest.raises(NotImplementedError):
        DO.beta_schedule(10, "quadratic")


UNET_CASES = {
    "c1_64_b32_h2": (dict(img_size=64, base_channels=32, n_heads=2), 1),
    "s32_b32_attn168": (dict(img_size=32, base_channels=32, attention_resolutions="16,8"), 1),
    "s64_b64_hc32_rgb": (dict(img_size=64, base_channels=64, n_head_channels=32), 3),
    "s128_b32_h4": (dict(img_size=128, base_channels=32, n_heads=4), 1),
}


def make_state_dict(kw, in_channels, seed=1):
    """Reference-layout state_dict with deterministic key-addressed values, built from the oracle's plan only."""
    from tests.util_unet import empty_state_dict
    return UO.fill_state_dict(empty_state_dict(in_channels=in_channels, **{k: v for k, v in kw.items()
                                                                         if k in ("img_size", "base_channels",
                                                                                  "attention_resolutions")}), seed)


@pytest.mark.parametrize("name", list(UNET_CASES))
def test_unet_oracle_matches_reference_golden(golden, name):
    g = golden("unet")
    kw, cin = UNET_CASES[name]
    sd = make_state_dict(kw, cin)
    keys = hashlib.sha1("\n".join(f"{k}:{tuple(v.shape)}" for k, v in sd.items()).encode()).hexdigest()
    assert keys == str(g[f"{name}/keys_sha1"])          # state_dict layout == reference layout
    assert sum(v.numel() for v in sd.values()) == int(g[f"{name}/n_params"])
    x, t = torch.from_numpy(g[f"{name}/x"]), torch.from_numpy(g[f"{name}/t"])
    taps = {}
    with torch.no_grad():
        y = UO.forward(sd, x, t, taps=taps, **kw)
    ref = g[f"{name}/y"]
    assert np.abs(y.numpy() - ref).max() <= 1e-5 * np.abs(ref).max()
    for k in ("down.0", "down.1", "down.3", "middle", "up.0"):
        v = taps[k]
        s = v.flatten()[:: max(1, v.numel() // 512)][:512].numpy()
        assert np.allclose(s, g[f"{name}/tap/{k}/sample"], rtol=1e-4, atol=1e-5), k


def test_config1_loop_oracle(golden):
    """BASELINE config 1 (64x64, T=100, base 32, replayed Gaussian noise), 12-step variant."""
    g = golden("loop")
    kw = dict(img_size=64, base_channels=32, n_heads=2)
    sd = make_state_dict(kw, 1)
    tab = DO.Tables(DO.beta_schedule(100, "linear"))
    x0 = torch.from_numpy(g["td12/x0"])
    bank = iter(torch.randn(13, 1, 1, 64, 64, generator=torch.Generator().manual_seed(99)))
    nf = lambda x, t: next(bank)
    seq = DO.forward_backward(tab, lambda x, t: UO.forward(sd, x, t, **kw), x0, 12, nf, nf, "half")
    assert len(seq) == 14
    assert np.allclose(seq[1].numpy(), g["td12/x_T"], atol=1e-6)
    ref = g["td12/final"]
    assert np.abs(seq[-1].numpy() - ref).max() <= 1e-4 * np.abs(ref).max()
