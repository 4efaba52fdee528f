"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

The GPU box has no /root/reference, so the outputs are committed; tests compare the oracle
(oracle/) and the CUDA path (anoddpm_b200/) against these files.  Everything here CALLS the
reference; there is no arithmetic of our own except input generation.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refload  # noqa: E402

simplex, UNet, GD = refload.load()
from oracle.unet_oracle import fill_state_dict  # noqa: E402  (weight filler only: deterministic inputs)


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def gen_simplex():
    out = {}
    seeds = [3, 12345, -9999999999, 9999999999, 1, -1, 2 ** 62 + 12345, -2 ** 63 + 5, 6569750267]
    out["seeds"] = np.array([str(s) for s in seeds])
    perms, pgis = [], []
    for s in seeds:
        p, g = simplex._init(s)
        perms.append(p), pgis.append(g)
    out["perm"] = np.stack(perms).astype(np.int16)
    out["pgi3"] = np.stack(pgis).astype(np.int16)
    S = simplex.Simplex_CLASS()
    S.newSeed(12345)
    rng = np.random.default_rng(1234)
    pts = rng.uniform(-40, 40, size=(3000, 3))
    f = 2.0 ** rng.integers(1, 7, size=(3000, 1))
    lat = rng.integers(0, 1000, size=(3000, 3)) / f          # exact dyadic: lattice-boundary hits are common
    named = np.array([[0, 0, 0], [0.5, 0.25, 0.125], [1, 1, 1], [3.984375, 2.0, 7.796875], [100.3, -4.2, 55.5]])
    P = np.concatenate([named, pts, lat])
    out["points"] = P
    out["noise3"] = np.array([S.noise3(*p) for p in P])
    out["noise2"] = np.array([S.noise2(p[0], p[1]) for p in P])
    # octave fields (seed 12345): small ones stored whole, the 256x256 hot-path field by crop + hash + moments
    out["fixedT_64_t17"] = S.rand_3d_fixed_T_octaves((64, 64), np.array([17]), 6, 0.8, 64)
    out["fixedT_48x80_t999_o3"] = S.rand_3d_fixed_T_octaves((48, 80), np.array([999]), 3, 0.6, 16)
    big = S.rand_3d_fixed_T_octaves((256, 256), np.array([500]), 6, 0.8, 64)
    out["fixedT_256_t500_sha1"] = np.array(sha(big))
    out["fixedT_256_t500_f32_sha1"] = np.array(sha(big.astype(np.float32)))
    out["fixedT_256_t500_crop"] = big[0, 96:160:1, 100:132]
    out["fixedT_256_t500_stats"] = np.array([big.sum(), big.mean(), big.std(), big.min(), big.max()])
    vol = S.rand_3d_octaves((32, 256, 256), 8, 0.8, 64)      # BASELINE config 5
    out["vol_32x256x256_o8_sha1"] = np.array(sha(vol))
    out["vol_32x256x256_o8_f32_sha1"] = np.array(sha(vol.astype(np.float32)))
    out["vol_32x256x256_o8_stats"] = np.array([vol.sum(), vol.mean(), vol.std(), vol.min(), vol.max()])
    out["vol_32x256x256_o8_crop"] = vol[::8, 120:136, 64:96]
    out["vol_4x32x32_o8"] = S.rand_3d_octaves((4, 32, 32), 8, 0.8, 64)
    out["noise3array_5x7x9"] = S.noise3array(np.arange(9) / 3.0, np.arange(7) / 5.0, np.arange(5) / 7.0)
    # host RNG stream consumed by newSeed() (simplex.py:19-22)
    np.random.seed(0)
    out["np_seed0_draws"] = np.array([np.random.randint(-10000000000, 10000000000) for _ in range(8)], dtype=np.int64)
    # generate_simplex_noise at B=1 (the only batch the reference supports)
    np.random.seed(42)
    x = torch.zeros(1, 2, 32, 32)
    out["gsn_B1_C2_t123"] = GD.generate_simplex_noise(simplex.Simplex_CLASS(), x, torch.tensor([123]),
                                                       in_channels=2).numpy()
    np.savez_compressed(os.path.join(HERE, "simplex.npz"), **out)
    print("simplex.npz", {k: getattr(v, "shape", None) for k, v in out.items()})


def gen_diffusion():
    out = {}
    for T, name in [(100, "linear"), (1000, "linear"), (100, "cosine"), (1000, "cosine")]:
        D = GD.GaussianDiffusionModel([64, 64], GD.get_beta_schedule(T, name))
        for n in ("betas", "sqrt_alphas", "sqrt_betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
                  "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
                  "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
                  "posterior_mean_coef1", "posterior_mean_coef2"):
            out[f"{name}{T}/{n}"] = getattr(D, n)
        g = torch.Generator().manual_seed(T)
        x = torch.rand(4, 1, 24, 40, generator=g) * 2 - 1
        n1 = torch.randn(4, 1, 24, 40, generator=g)
        e1 = torch.randn(4, 1, 24, 40, generator=g) * 1.5
        t = torch.tensor([0, 1, T // 2, T - 1])
        out[f"{name}{T}/x"], out[f"{name}{T}/noise"], out[f"{name}{T}/eps"], out[f"{name}{T}/t"] = \
            x.numpy(), n1.numpy(), e1.numpy(), t.numpy()
        out[f"{name}{T}/sample_q"] = D.sample_q(x, t, n1).numpy()
        out[f"{name}{T}/sample_q_gradual"] = D.sample_q_gradual(x, t, n1).numpy()
        pmv = D.p_mean_variance(None, x, t, e1)
        for k, v in pmv.items():
            out[f"{name}{T}/pmv_{k}"] = v.contiguous().numpy()
        sp = D.sample_p(lambda a, b: e1, x, t, denoise_fn=lambda a, b: n1)
        out[f"{name}{T}/sample_p"] = sp["sample"].numpy()
        out[f"{name}{T}/sample_p_pred_x_0"] = sp["pred_x_0"].numpy()
        out[f"{name}{T}/predict_eps_from_x_0"] = D.predict_eps_from_x_0(x, t[:1].repeat(4) + 5, pmv["pred_x_0"]).numpy()
    np.savez_compressed(os.path.join(HERE, "diffusion.npz"), **out)
    print("diffusion.npz", len(out), "arrays")


UNET_CASES = {
    # name: (ctor kwargs, batch, timesteps)
    "c1_64_b32_h2": (dict(img_size=64, base_channels=32, n_heads=2), 2, [3, 77]),                 # BASELINE config 1 model
    "s32_b32_attn168": (dict(img_size=32, base_channels=32, attention_resolutions="16,8"), 3, [0, 999, 500]),
    "s64_b64_hc32_rgb": (dict(img_size=64, base_channels=64, n_head_channels=32, in_channels=3), 1, [12]),
    "s128_b32_h4": (dict(img_size=128, base_channels=32, n_heads=4), 1, [250]),
}


def gen_unet():
    out = {}
    for name, (kw, B, ts) in UNET_CASES.items():
        torch.manual_seed(0)
        m = UNet.UNetModel(**kw).eval()
        m.load_state_dict(fill_state_dict(m.state_dict(), seed=1))
        g = torch.Generator().manual_seed(5)
        x = torch.randn(B, kw.get("in_channels", 1), kw["img_size"], kw["img_size"], generator=g)
        t = torch.tensor(ts)
        acts = {}
        hooks = []
        for mod_name in ("down.0", "down.1", "down.3", "middle", "up.0", f"up.{len(m.up) - 1}"):
            mod = m.get_submodule(mod_name)
            hooks.append(mod.register_forward_hook(lambda _m, _i, o, n=mod_name: acts.__setitem__(n, o.detach())))
        with torch.no_grad():
            y = m(x, t)
        for h in hooks:
            h.remove()
        out[f"{name}/x"], out[f"{name}/t"], out[f"{name}/y"] = x.numpy(), t.numpy(), y.numpy()
        for k, v in acts.items():   # intermediate taps: moments + a strided sample (small)
            out[f"{name}/tap/{k}/stats"] = np.array([v.mean().item(), v.std().item(), v.abs().max().item()])
            out[f"{name}/tap/{k}/sample"] = v.flatten()[:: max(1, v.numel() // 512)][:512].numpy()
        out[f"{name}/n_params"] = np.array(sum(p.numel() for p in m.parameters()))
        out[f"{name}/keys_sha1"] = np.array(hashlib.sha1(
            "\n".join(f"{k}:{tuple(v.shape)}" for k, v in m.state_dict().items()).encode()).hexdigest())
    # the BASELINE model (256^2, base 128, attn "16,8", 2 heads): one bs-1 forward
    torch.manual_seed(0)
    kw = dict(img_size=256, base_channels=128, n_heads=2, attention_resolutions="16,8")
    m = UNet.UNetModel(**kw).eval()
    m.load_state_dict(fill_state_dict(m.state_dict(), seed=1))
    g = torch.Generator().manual_seed(6)
    x = torch.randn(1, 1, 256, 256, generator=g)
    t = torch.tensor([431])
    with torch.no_grad():
        y = m(x, t)
    out["base256/x_seed"] = np.array(6)
    out["base256/t"] = t.numpy()
    out["base256/y_f16"] = y.numpy().astype(np.float16)          # compact copy (parity bar is 1e-3 relative)
    out["base256/y_crop"] = y.numpy()[0, 0, 100:132, 64:128]
    out["base256/y_stats"] = np.array([y.mean().item(), y.std().item(), y.abs().max().item()])
    out["base256/n_params"] = np.array(sum(p.numel() for p in m.parameters()))
    out["base256/keys_sha1"] = np.array(hashlib.sha1(
        "\n".join(f"{k}:{tuple(v.shape)}" for k, v in m.state_dict().items()).encode()).hexdigest())
    np.savez_compressed(os.path.join(HERE, "unet.npz"), **out)
    print("unet.npz", len(out), "arrays")


def gen_loop():
    """BASELINE config 1: 64x64, T=100, base 32, Gaussian noise replayed through a callable."""
    out = {}
    torch.manual_seed(0)
    m = UNet.UNetModel(64, 32, n_heads=2).eval()
    m.load_state_dict(fill_state_dict(m.state_dict(), seed=1))
    D = GD.GaussianDiffusionModel([64, 64], GD.get_beta_schedule(100, "linear"), noise="gauss")
    g = torch.Generator().manual_seed(0)
    x0 = torch.rand(1, 1, 64, 64, generator=g) * 2 - 1
    for tag, td in (("td100", 100), ("td12", 12)):
        gz = torch.Generator().manual_seed(99)
        bank = torch.randn(td + 1, 1, 1, 64, 64, generator=gz)
        it = iter(bank)
        D.noise_fn = lambda x, t: next(it)
        seq = D.forward_backward(m, x0, see_whole_sequence="half", t_distance=td, denoise_fn=lambda x, t: next(it))
        assert len(seq) == td + 2
        out[f"{tag}/x0"] = x0.numpy()
        out[f"{tag}/final"] = seq[-1].numpy()
        out[f"{tag}/x_T"] = seq[1].numpy()
        out[f"{tag}/step_absmax"] = np.array([s.abs().max().item() for s in seq])
        out[f"{tag}/step_mean"] = np.array([s.mean().item() for s in seq])
    # "whole" sequence variant, short
    gz = torch.Generator().manual_seed(7)
    bank = torch.randn(2 * 6, 1, 1, 64, 64, generator=gz)
    it = iter(bank)
    D.noise_fn = lambda x, t: next(it)
    seq = D.forward_backward(m, x0, see_whole_sequence="whole", t_distance=6, denoise_fn="noise_fn")
    assert len(seq) == 13
    out["whole6/final"] = seq[-1].numpy()
    out["whole6/mid"] = seq[6].numpy()
    # simplex-noise loop at B=1 with the host RNG pinned (np.random.seed) -- the reference's own noise path
    np.random.seed(2024)
    D2 = GD.GaussianDiffusionModel([64, 64], GD.get_beta_schedule(100, "linear"), noise="simplex")
    y = D2.forward_backward(m, x0, see_whole_sequence=None, t_distance=8, denoise_fn="noise_fn")
    out["simplex_td8/final"] = y.numpy()
    np.savez_compressed(os.path.join(HERE, "loop.npz"), **out)
    print("loop.npz", len(out), "arrays")


def gen_vlb():
    """VLB helpers (GaussianDiffusion.py:43-93, 384-397, 436-447): calc_vlb_xt with a given noise estimate, its two
    elementwise ingredients, prior_vlb."""
    out = {}
    for T, name in [(100, "linear"), (1000, "linear"), (1000, "cosine")]:
        D = GD.GaussianDiffusionModel([24, 40], GD.get_beta_schedule(T, name))
        g = torch.Generator().manual_seed(7 * T + len(name))
        x0 = torch.rand(5, 1, 24, 40, generator=g) * 2 - 1
        x0[0, 0, :4] = 1.0          # saturated pixels exercise the x > 0.999 / x < -0.999 branches
        x0[1, 0, :4] = -1.0
        noise = torch.randn(5, 1, 24, 40, generator=g)
        eps = noise + 0.3 * torch.randn(5, 1, 24, 40, generator=g)
        t = torch.tensor([0, 0, 1, T // 2, T - 1])
        x_t = D.sample_q(x0, t, noise)
        k = f"{name}{T}"
        out[f"{k}/x0"], out[f"{k}/x_t"], out[f"{k}/eps"], out[f"{k}/t"] = x0.numpy(), x_t.numpy(), eps.numpy(), t.numpy()
        r = D.calc_vlb_xt(None, x0, x_t, t, eps)
        out[f"{k}/vlb"], out[f"{k}/pred_x_0"] = r["output"].numpy(), r["pred_x_0"].numpy()
        pmv = D.p_mean_variance(None, x_t, t, eps)
        tm, _, tlv = D.q_posterior_mean_variance(x0, x_t, t)
        out[f"{k}/normal_kl"] = GD.normal_kl(tm, tlv, pmv["mean"], pmv["log_variance"]).numpy()
        out[f"{k}/disc_ll"] = GD.discretised_gaussian_log_likelihood(x0, pmv["mean"], log_scales=0.5 * pmv["log_variance"]).numpy()
        out[f"{k}/prior_vlb"] = D.prior_vlb(x0, {"Batch_Size": 5}).numpy()
    np.savez_compressed(os.path.join(HERE, "vlb.npz"), **out)
    print("vlb.npz", len(out), "arrays")


def gen_loss():
    """Training-loss path (GaussianDiffusion.py:140-147, 220-226, 399-478): calc_loss / p_loss for every loss type and
    weighting, random_noise's branch sequence, simplex_randParam's RNG consumption, calc_total_vlb.  Every random draw
    is pinned (explicit t, replayed noise, seeded host RNGs) so the GPU run can consume identical numbers."""
    import random
    out = {}
    kw = dict(img_size=32, base_channels=32, n_heads=2)
    torch.manual_seed(0)
    m = UNet.UNetModel(**kw).eval()
    m.load_state_dict(fill_state_dict(m.state_dict(), seed=1))
    g = torch.Generator().manual_seed(21)
    x0 = torch.rand(3, 1, 32, 32, generator=g) * 2 - 1
    noise = torch.randn(3, 1, 32, 32, generator=g)
    t = torch.tensor([0, 5, 99])
    out["x0"], out["noise"], out["t"] = x0.numpy(), noise.numpy(), t.numpy()
    for lt in ("l1", "l2", "hybrid", "other"):
        D = GD.GaussianDiffusionModel([32, 32], GD.get_beta_schedule(100, "linear"), loss_type=lt, noise="gauss")
        D.noise_fn = lambda x, tt: noise
        with torch.no_grad():
            loss, x_t, eps = D.calc_loss(m, x0, t)
        for k, v in loss.items():
            out[f"calc_loss/{lt}/{k}"] = v.numpy()
        out[f"calc_loss/{lt}/x_t"], out[f"calc_loss/{lt}/eps"] = x_t.numpy(), eps.numpy()
    # p_loss: weighted timestep sampling (np.random) and uniform sampling (torch.randint on x_0.device)
    for lw in ("prop-t", "uniform"):
        D = GD.GaussianDiffusionModel([32, 32], GD.get_beta_schedule(100, "linear"), loss_type="l2", loss_weight=lw, noise="gauss")
        D.noise_fn = lambda x, tt: noise
        np.random.seed(11)
        tt, ww = D.sample_t_with_weights(3, torch.device("cpu"))
        out[f"p_loss/{lw}/t"], out[f"p_loss/{lw}/w"] = tt.numpy(), ww.numpy()
        np.random.seed(11)
        with torch.no_grad():
            total, (ld, x_t, eps) = D.p_loss(m, x0, {"train_start": True, "sample_distance": 60})
        out[f"p_loss/{lw}/total"], out[f"p_loss/{lw}/loss"] = total.numpy(), ld["loss"].numpy()
        out[f"p_loss/{lw}/x_t"] = x_t.numpy()
    D = GD.GaussianDiffusionModel([32, 32], GD.get_beta_schedule(100, "linear"), loss_type="l2", loss_weight="none", noise="gauss")
    D.noise_fn = lambda x, tt: noise
    for tag, args in (("start", {"train_start": True, "sample_distance": 60}), ("full", {"train_start": False, "sample_distance": 60})):
        torch.manual_seed(5)
        t_draw = torch.randint(0, 60 if tag == "start" else 100, (3,))
        torch.manual_seed(5)
        with torch.no_grad():
            total, (ld, x_t, eps) = D.p_loss(m, x0, args)
        out[f"p_loss/none_{tag}/t"], out[f"p_loss/none_{tag}/total"] = t_draw.numpy(), total.numpy()
        out[f"p_loss/none_{tag}/loss"], out[f"p_loss/none_{tag}/x_t"] = ld["loss"].numpy(), x_t.numpy()
    # random_noise: branch sequence under pinned RNGs (B = 1)
    random.seed(3), np.random.seed(4), torch.manual_seed(5)
    S = simplex.Simplex_CLASS()
    x1 = torch.zeros(1, 1, 32, 32)
    t1 = torch.tensor([7])
    outs, flags = [], []
    for _ in range(8):
        before = np.random.get_state()[1].copy(), np.random.get_state()[2]
        r = GD.random_noise(S, x1, t1)
        after = np.random.get_state()[1], np.random.get_state()[2]
        flags.append(not (np.array_equal(before[0], after[0]) and before[1] == after[1]))
        outs.append(r.numpy())
    out["random_noise/out"], out["random_noise/is_simplex"] = np.stack(outs), np.array(flags)
    # simplex_randParam: output and RNG consumption (one random.choice per channel; the drawn parameters are discarded)
    random.seed(9), np.random.seed(10)
    x2 = torch.zeros(1, 2, 16, 16)
    out["randparam/out"] = GD.generate_simplex_noise(simplex.Simplex_CLASS(), x2, torch.tensor([33]), True, in_channels=2).numpy()
    out["randparam/next_random"] = np.array(random.random())
    out["randparam/next_np"] = np.array(np.random.randint(-10000000000, 10000000000))
    # calc_total_vlb with torch.randn_like replayed from a bank (the reference draws it on x_0.device)
    T = 40          # linear schedule scaled by 1000/T: betas stay below 1 for T >= 21
    D = GD.GaussianDiffusionModel([32, 32], GD.get_beta_schedule(T, "linear"), noise="gauss")
    bank = torch.randn(T, 3, 1, 32, 32, generator=torch.Generator().manual_seed(31))
    it = iter(bank)
    orig = torch.randn_like
    torch.randn_like = lambda x, *a, **k: next(it)
    try:
        r = D.calc_total_vlb(x0, m, {"Batch_Size": 3})
    finally:
        torch.randn_like = orig
    for k, v in r.items():
        out[f"total_vlb/{k}"] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "loss.npz"), **out)
    print("loss.npz", len(out), "arrays")


def gen_eval():
    """Detection metrics of the reference's evaluation.py:26-87 on synthetic (image, reconstruction, mask) triples: the
    values detection.py:229-250 computes per image."""
    import importlib
    ev = importlib.import_module("evaluation")
    assert ev.__file__.startswith("/root/reference")
    out = {}
    g = torch.Generator().manual_seed(77)
    for i, (H, W) in enumerate([(64, 64), (32, 48), (256, 256)]):
        image = torch.rand(1, 1, H, W, generator=g) * 2 - 1
        recon = (image + torch.randn(1, 1, H, W, generator=g) * 0.6).clamp(-1, 1)
        mask = (torch.rand(1, 1, H, W, generator=g) > 0.7).float()
        k = f"case{i}"
        out[f"{k}/image"], out[f"{k}/recon"], out[f"{k}/mask"] = image.numpy(), recon.numpy(), mask.numpy()
        mse = (image - recon).square()
        fpr, tpr, _ = ev.ROC_AUC(mask.to(torch.uint8), mse)
        out[f"{k}/auc"] = np.array(ev.AUC_score(fpr, tpr))
        thr = (mse > 0.5).float()
        out[f"{k}/dice"] = ev.dice_coeff(image, recon, mask, mse=thr).numpy()
        out[f"{k}/dice_default"] = ev.dice_coeff(image, recon, mask).numpy()
        out[f"{k}/precision"] = ev.precision(mask, thr).numpy()
        out[f"{k}/recall"] = ev.recall(mask, thr).numpy()
        out[f"{k}/fpr"] = ev.FPR(mask, thr).numpy()
        out[f"{k}/iou"] = np.array(ev.IoU(mask, thr))
        out[f"{k}/psnr"] = np.array(ev.PSNR(recon, image))
    np.savez_compressed(os.path.join(HERE, "eval.npz"), **out)
    print("eval.npz", len(out), "arrays")


def gen_pipeline():
    """Input pipeline: the deterministic part of MRIDataset's transform (dataset.py:585-593 -- the
    torchvision calls the reference makes, without the RandomAffine jitter) on slices of a synthetic cached volume."""
    from torchvision import transforms
    g = np.random.default_rng(5)
    out = {}
    vol = g.random((256, 6, 192), dtype=np.float32)
    out["vol"] = vol
    for name, size in (("s256", (256, 256)), ("s64", (64, 64)), ("s128x96", (128, 96))):
        tf = transforms.Compose([transforms.ToPILImage(), transforms.CenterCrop(235),
                                 transforms.Resize(size, transforms.InterpolationMode.BILINEAR), transforms.ToTensor(),
                                 transforms.Normalize((0.5), (0.5))])
        for s in (0, 3):
            image = vol[:, s:s + 1, :].reshape(256, 192).astype(np.float32)          # dataset.py:631
            out[f"{name}/slice{s}"] = tf(image).numpy()
    np.savez_compressed(os.path.join(HERE, "pipeline.npz"), **out)
    print("pipeline.npz", len(out), "arrays")


if __name__ == "__main__":
    which = sys.argv[1:] or ["simplex", "diffusion", "unet", "loop", "vlb", "loss", "eval", "pipeline"]
    for w in which:
        globals()["gen_" + w]()
