"""GPU parity (through the C ABI): simplex noise and diffusion elementwise kernels vs goldens from the
unmodified reference and vs the CPU oracle.  Integer/table path and fp64 fields: bit-exact."""
import hashlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_point_values_bit_exact(golden):
    from anoddpm_b200.simplex import Simplex_CLASS
    g = golden("simplex")
    S = Simplex_CLASS()
    S.newSeed(12345)
    P = g["points"]
    for i in list(range(40)) + list(range(3005, 3045)):        # named + random + lattice-aligned points
        assert S.noise3(*P[i]) == g["noise3"][i], i
        assert S.noise2(P[i][0], P[i][1]) == g["noise2"][i], i
    # all 6005 points through the grid kernel: per-point 1x1x1 grids are slow, so use the diagonal of small grids
    idx = np.arange(0, 6005, 7)
    for i in idx[:200]:
        v = S.noise3array([P[i][0]], [P[i][1]], [P[i][2]])
        assert v[0, 0, 0] == g["noise3"][i]


def test_fields_bit_exact(golden):
    from anoddpm_b200.simplex import Simplex_CLASS
    g = golden("simplex")
    S = Simplex_CLASS()
    S.newSeed(12345)
    assert np.array_equal(S.rand_3d_fixed_T_octaves((64, 64), np.array([17]), 6, 0.8, 64), g["fixedT_64_t17"])
    assert np.array_equal(S.rand_3d_fixed_T_octaves((48, 80), np.array([999]), 3, 0.6, 16), g["fixedT_48x80_t999_o3"])
    big = S.rand_3d_fixed_T_octaves((256, 256), np.array([500]), 6, 0.8, 64)
    assert big.shape == (1, 256, 256) and big.dtype == np.float64
    assert sha(big) == str(g["fixedT_256_t500_sha1"])
    assert np.array_equal(big[0, 96:160, 100:132], g["fixedT_256_t500_crop"])
    assert np.array_equal(S.rand_3d_octaves((4, 32, 32), 8, 0.8, 64), g["vol_4x32x32_o8"])
    assert np.array_equal(S.noise3array(np.arange(9) / 3.0, np.arange(7) / 5.0, np.arange(5) / 7.0),
                          g["noise3array_5x7x9"])
    with pytest.raises(ValueError):      # like the reference, T with more than one entry does not broadcast
        S.rand_3d_fixed_T_octaves((16, 16), np.array([1, 2]), 2, 0.5, 8)


def test_config5_volume_bit_exact(golden):
    """BASELINE config 5: (32,256,256) x 8 octaves."""
    from anoddpm_b200.simplex import Simplex_CLASS
    g = golden("simplex")
    S = Simplex_CLASS()
    S.newSeed(12345)
    vol = S.rand_3d_octaves((32, 256, 256), 8, 0.8, 64)
    assert sha(vol) == str(g["vol_32x256x256_o8_sha1"])
    assert sha(vol.astype(np.float32)) == str(g["vol_32x256x256_o8_f32_sha1"])


def test_2d_octaves_vs_oracle():
    from anoddpm_b200.simplex import Simplex_CLASS
    from oracle.simplex_oracle import OracleSimplex
    S = Simplex_CLASS()
    S.newSeed(77)
    O = OracleSimplex(77)
    a = S.rand_2d_octaves((32, 32), 4, 0.7, 16)
    y, x = np.arange(32), np.arange(32)
    ref = np.zeros((32, 32))
    f, amp = 16, 1
    for _ in range(4):
        ref += amp * np.array([[O.noise2(xx / f, yy / f) for xx in x] for yy in y])
        f /= 2
        amp *= 0.7
    assert np.array_equal(a, ref)


def test_generate_simplex_noise_b1_matches_reference(golden):
    from anoddpm_b200.GaussianDiffusion import generate_simplex_noise
    from anoddpm_b200.simplex import Simplex_CLASS
    g = golden("simplex")
    np.random.seed(42)
    S = Simplex_CLASS()
    x = torch.zeros(1, 2, 32, 32, device=DEV)
    n = generate_simplex_noise(S, x, torch.tensor([123], device=DEV), in_channels=2)
    assert n.dtype == torch.float32 and n.shape == (1, 2, 32, 32)
    assert np.array_equal(n.cpu().numpy(), g["gsn_B1_C2_t123"])


def test_batched_simplex_is_per_sample_reference():
    """Batched semantic: sample b == the reference at B=1 after newSeed(seed[b]) (checked against the oracle)."""
    from anoddpm_b200.GaussianDiffusion import generate_simplex_noise, simplex_noise_from_seeds
    from anoddpm_b200.simplex import Simplex_CLASS
    from oracle.simplex_oracle import OracleSimplex
    np.random.seed(7)
    seeds = [int(np.random.randint(-10000000000, 10000000000)) for _ in range(3)]
    np.random.seed(7)
    S = Simplex_CLASS.__new__(Simplex_CLASS)
    t = torch.tensor([5, 250, 999], device=DEV)
    x = torch.zeros(3, 1, 40, 24, device=DEV)
    n = generate_simplex_noise(S, x, t).cpu().numpy()
    for b in range(3):
        ref = OracleSimplex(seeds[b]).rand_3d_fixed_T_octaves((40, 24), np.array([int(t[b])]), 6, 0.8, 64)
        assert np.array_equal(n[b, 0], ref[0].astype(np.float32)), b
    assert np.array_equal(S._perm, OracleSimplex(seeds[2])._perm)
    sh = simplex_noise_from_seeds([11], (3, 1, 40, 24), t, shared_seed=True).cpu().numpy()
    for b in range(3):
        ref = OracleSimplex(11).rand_3d_fixed_T_octaves((40, 24), np.array([int(t[b])]), 6, 0.8, 64)
        assert np.array_equal(sh[b, 0], ref[0].astype(np.float32))


@pytest.mark.parametrize("freq", [48.0, 10.0, 64.0, 0.5, 2.0 ** -3])
def test_octave_frequencies_division_and_power_of_two_paths(freq):
    """Power-of-two frequencies take the multiply-by-reciprocal path of the octave kernel, the others the IEEE division of the
    reference (simplex.py:75-93): both bit-exact against the oracle."""
    from anoddpm_b200.simplex import Simplex_CLASS
    from oracle.simplex_oracle import OracleSimplex
    S = Simplex_CLASS()
    S.newSeed(4242)
    O = OracleSimplex(4242)
    for T in (7, 901):
        a = S.rand_3d_fixed_T_octaves((40, 24), np.array([T]), 5, 0.7, freq)
        b = O.rand_3d_fixed_T_octaves((40, 24), np.array([T]), 5, 0.7, freq)
        assert np.array_equal(a[0], b[0])
    v = S.rand_3d_octaves((3, 16, 24), 4, 0.6, freq)
    assert np.array_equal(v, O.rand_3d_octaves((3, 16, 24), 4, 0.6, freq))


@pytest.mark.parametrize("T,name", [(100, "linear"), (1000, "linear"), (100, "cosine"), (1000, "cosine")])
def test_diffusion_steps_bit_exact_vs_reference_golden(golden, T, name):
    from anoddpm_b200.GaussianDiffusion import GaussianDiffusionModel, get_beta_schedule
    g = golden("diffusion")
    k = f"{name}{T}/"
    D = GaussianDiffusionModel([24, 40], get_beta_schedule(T, name))
    x, noise, eps, t = (torch.from_numpy(g[k + n]).to(DEV) for n in ("x", "noise", "eps", "t"))
    assert np.array_equal(D.sample_q(x, t, noise).cpu().numpy(), g[k + "sample_q"])
    assert np.array_equal(D.sample_q_gradual(x, t, noise).cpu().numpy(), g[k + "sample_q_gradual"])
    out = D.sample_p(lambda a, b: eps, x, t, denoise_fn=lambda a, b: noise)
    assert set(out) == {"sample", "pred_x_0"}
    assert np.array_equal(out["sample"].cpu().numpy(), g[k + "sample_p"])
    assert np.array_equal(out["pred_x_0"].cpu().numpy(), g[k + "sample_p_pred_x_0"])
    pmv = D.p_mean_variance(None, x, t, eps)
    for kk in ("mean", "variance", "log_variance", "pred_x_0"):
        assert np.allclose(pmv[kk].cpu().numpy(), g[k + "pmv_" + kk], rtol=1e-6, atol=1e-7), kk


def test_elementwise_odd_sizes_and_full_size_property():
    """ragged (non multiple-of-4, unaligned) sizes take the scalar path; BASELINE-size linearity property."""
    from anoddpm_b200.GaussianDiffusion import GaussianDiffusionModel, get_beta_schedule
    from oracle import diffusion_oracle as DO
    D = GaussianDiffusionModel([7, 9], get_beta_schedule(1000, "linear"))
    tab = DO.Tables(DO.beta_schedule(1000, "linear"))
    g = torch.Generator().manual_seed(3)
    x, n, e = (torch.randn(5, 3, 7, 9, generator=g) for _ in range(3))
    t = torch.tensor([0, 999, 1, 500, 250])
    assert torch.equal(D.sample_q(x.to(DEV), t.to(DEV), n.to(DEV)).cpu(), DO.sample_q(tab, x, t, n))
    o = D.sample_p(lambda a, b: e.to(DEV), x.to(DEV), t.to(DEV), denoise_fn=lambda a, b: n.to(DEV))
    r = DO.sample_p(tab, x, t, e, n)
    assert torch.equal(o["sample"].cpu(), r["sample"]) and torch.equal(o["pred_x_0"].cpu(), r["pred_x_0"])
    # full BASELINE size (32,1,256,256): q_sample(x, t, 0) == coef*x exactly and is linear in the noise term
    D = GaussianDiffusionModel([256, 256], get_beta_schedule(1000, "linear"))
    x = torch.randn(32, 1, 256, 256, device=DEV)
    t = torch.randint(0, 1000, (32,), device=DEV)
    z = torch.zeros_like(x)
    a = torch.from_numpy(D.sqrt_alphas_cumprod).to(DEV)[t].float().view(-1, 1, 1, 1)
    assert torch.equal(D.sample_q(x, t, z), a * x + 0 * z)
    assert torch.equal(D.sample_q(z, t, x) + D.sample_q(x, t, z), D.sample_q(x, t, x)) or \
        torch.allclose(D.sample_q(z, t, x) + D.sample_q(x, t, z), D.sample_q(x, t, x), atol=1e-6)


def test_fused_detection_metrics_match_reference_formulas():
    """one-pass device metrics == the reference's evaluation.py formulas evaluated per image."""
    from anoddpm_b200 import evaluation as E
    g = torch.Generator().manual_seed(8)
    B = 5
    image = (torch.rand(B, 1, 64, 48, generator=g) * 2 - 1).to(DEV)
    recon = (image + torch.randn(B, 1, 64, 48, generator=g).to(DEV) * 0.6)
    mask = (torch.rand(B, 1, 64, 48, generator=g) > 0.7).float().to(DEV)
    m = E.detection_metrics(image, recon, mask)
    for b in range(B):
        im, rc, mk = image[b:b + 1], recon[b:b + 1], mask[b:b + 1]
        thr = ((im - rc).square() > 0.5).float()
        assert abs(float(m["dice"][b]) - float(E.dice_coeff(im, rc, mk, mse=thr))) < 1e-6
        assert abs(float(m["precision"][b]) - float(E.precision(mk, thr))) < 1e-6
        assert abs(float(m["recall"][b]) - float(E.recall(mk, thr))) < 1e-6
        assert abs(float(m["fpr"][b]) - float(E.FPR(mk, thr))) < 1e-6
        assert abs(float(m["iou"][b]) - E.IoU(mk, thr)) < 1e-6
        assert abs(float(m["psnr"][b]) - float(E.PSNR(rc, im))) < 1e-3
    # property at BASELINE size: counts partition the image
    c = E.detection_counts(torch.zeros(32, 1, 256, 256, device=DEV), torch.ones(32, 1, 256, 256, device=DEV),
                           torch.ones(32, 1, 256, 256, device=DEV))
    assert torch.equal(c[:, 0], torch.full((32,), 65536.0, device=DEV, dtype=torch.float64)) and float(c[:, 1:4].abs().sum()) == 0


@pytest.mark.parametrize("T,name", [(100, "linear"), (1000, "linear"), (1000, "cosine")])
def test_vlb_term_fused_kernel_vs_reference_golden(golden, T, name):
    """calc_vlb_xt (GaussianDiffusion.py:384-397): fused KL / decoder-NLL kernel under no_grad, torch formulas with autograd."""
    from anoddpm_b200 import GaussianDiffusionModel, get_beta_schedule
    g = golden("vlb")
    k = f"{name}{T}/"
    D = GaussianDiffusionModel([24, 40], get_beta_schedule(T, name))
    x0, x_t, eps, t = (torch.from_numpy(g[k + n]).to(DEV) for n in ("x0", "x_t", "eps", "t"))
    ref = torch.from_numpy(g[k + "vlb"])
    with torch.no_grad():
        r = D.calc_vlb_xt(None, x0, x_t, t, eps)
    assert r["output"].dtype == torch.float32 and r["output"].shape == (5,)
    assert torch.allclose(r["output"].cpu(), ref, rtol=2e-5, atol=1e-6), (r["output"].cpu() - ref).abs().max()
    assert np.array_equal(r["pred_x_0"].cpu().numpy(), g[k + "pred_x_0"])            # same mul / sub / clamp, no contraction
    e2 = eps.clone().requires_grad_(True)                                             # hybrid-loss path: gradient w.r.t. the estimate
    r2 = D.calc_vlb_xt(None, x0, x_t, t, e2)
    assert torch.allclose(r2["output"].detach().cpu(), ref, rtol=2e-5, atol=1e-6)
    r2["output"].sum().backward()
    assert bool(torch.isfinite(e2.grad).all()) and float(e2.grad.abs().max()) > 0
    assert torch.allclose(D.prior_vlb(x0, {"Batch_Size": 5}).cpu(), torch.from_numpy(g[k + "prior_vlb"]), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("shape", [(3, 40, 24), (2, 64, 64)])
def test_fused_noise_posterior_kernel_is_bit_identical_to_the_two_kernels(shape):
    """anoddpm_p_sample_simplex (noise3 octaves evaluated inside the posterior kernel: SURVEY 8(d)'s 12 B/px form of sample_p)
    against anoddpm_simplex_noise_steps followed by anoddpm_p_sample, including an out-of-range timestep (poisoned)."""
    from anoddpm_b200 import GaussianDiffusionModel, get_beta_schedule
    from anoddpm_b200._lib import call, ptr, stream
    from anoddpm_b200.simplex import init_tables
    B, H, W = shape
    T = 1000
    D = GaussianDiffusionModel([H, W], get_beta_schedule(T, "linear"), noise="simplex")
    tab = D._tables(torch.device(DEV))
    g = torch.Generator().manual_seed(21)
    x_t = torch.randn(B, 1, H, W, generator=g).to(DEV)
    eps = torch.randn(B, 1, H, W, generator=g).to(DEV)
    steps = 3
    tables = torch.from_numpy(init_tables([7 + 13 * i for i in range(steps * B)])).view(steps, B * 512).to(DEV)
    step = torch.tensor([2], device=DEV, dtype=torch.int64)
    for tvals in ([5, 999, 0][:B] + [17] * max(0, B - 3), [1000] + [3] * (B - 1)):
        t = torch.tensor(tvals[:B], device=DEV, dtype=torch.int64)
        noise = torch.empty(B, 1, H, W, device=DEV)
        call("anoddpm_simplex_noise_steps", ptr(noise), B, 1, H, W, ptr(tables), ptr(step), 0, ptr(t), 6, 0.8, 64.0, stream())
        ref_s, ref_p = torch.empty_like(x_t), torch.empty_like(x_t)
        call("anoddpm_p_sample", ptr(ref_s), ptr(ref_p), ptr(x_t), ptr(eps), ptr(noise), ptr(t), ptr(tab[4:]), T, B, H * W, stream())
        out_s, out_p = torch.empty_like(x_t), torch.empty_like(x_t)
        call("anoddpm_p_sample_simplex", ptr(out_s), ptr(out_p), ptr(x_t), ptr(eps), ptr(tables), ptr(step), ptr(t), ptr(tab[4:]), T, B,
             H, W, 6, 0.8, 64.0, stream())
        torch.cuda.synchronize()
        for a, b in ((out_s, ref_s), (out_p, ref_p)):      # bit-identical; a poisoned (out-of-range t) sample is NaN in both
            assert torch.equal(torch.isnan(a), torch.isnan(b))
            assert torch.equal(torch.nan_to_num(a).view(torch.int32), torch.nan_to_num(b).view(torch.int32))
        assert bool(torch.isnan(ref_s[0]).all()) == (tvals[0] >= T)
