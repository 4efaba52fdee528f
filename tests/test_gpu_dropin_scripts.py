"""Script-level drop-in proof: with ``dropin/`` first on sys.path the reference scripts'
own import lines resolve to this package, and their call sequences -- the training loop of diffusion_training.py:40-107 +
save() :156-189, then anomalous_metric_calculation of detection.py:155-250 (load_state_dict(ckpt["ema"]) from the saved
.pt -> forward_backward(..., denoise_fn=args["noise_fn"]) -> squared error, ROC/AUC, Dice, precision / recall / IoU / FPR)
-- run UNCHANGED on the GPU, with results matching the CPU oracle driven by the same seeds."""
import copy
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from tests.conftest import ROOT

pytestmark = pytest.mark.gpu
DEV = "cuda"
ARGS = {"img_size": [64, 64], "base_channels": 64, "channel_mults": "", "dropout": 0, "num_heads": 2, "num_head_channels": -1,
        "T": 1000, "beta_schedule": "linear", "loss_weight": "none", "loss-type": "l2", "noise_fn": "simplex", "lr": 1e-4,
        "weight_decay": 0.0, "Batch_Size": 2, "train_start": True, "sample_distance": 800, "EPOCHS": 2, "arg_num": 0,
        "dataset": "mri", "channels": ""}


def _dropin_modules():
    path = os.path.join(ROOT, "dropin")
    sys.path.insert(0, path)
    try:
        for n in ("GaussianDiffusion", "UNet", "simplex", "evaluation"):
            sys.modules.pop(n, None)
        mods = [importlib.import_module(n) for n in ("GaussianDiffusion", "UNet", "simplex", "evaluation")]
    finally:
        sys.path.remove(path)
    assert all(m.__file__.startswith(path) for m in mods)
    return mods


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_training_script_then_detection_script_through_dropin(tmp_path):
    GD, UN, SX, EV = _dropin_modules()
    from GaussianDiffusion import GaussianDiffusionModel, get_beta_schedule          # the scripts' own import lines
    from UNet import UNetModel, update_ema_params
    from oracle import diffusion_oracle as DO, unet_oracle as UO
    from oracle.simplex_oracle import OracleSimplex
    args = dict(ARGS)
    device = torch.device(DEV)
    kw = dict(img_size=64, base_channels=64, n_heads=2)

    # ---------------- diffusion_training.py:40-107 (2 optimiser steps on a synthetic batch)
    torch.manual_seed(0)
    model = UNetModel(args["img_size"][0], args["base_channels"], channel_mults=args["channel_mults"], dropout=args["dropout"],
                      n_heads=args["num_heads"], n_head_channels=args["num_head_channels"], in_channels=1)
    model.load_state_dict(UO.fill_state_dict(model.state_dict(), seed=3))            # (the zero-initialised tensors carry signal)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    betas = get_beta_schedule(args["T"], args["beta_schedule"])
    diffusion = GaussianDiffusionModel(args["img_size"], betas, loss_weight=args["loss_weight"], loss_type=args["loss-type"],
                                       noise=args["noise_fn"], img_channels=1)
    ema = copy.deepcopy(model)
    model.to(device)
    ema.to(device)
    optimiser = torch.optim.AdamW(model.parameters(), lr=args["lr"], weight_decay=args["weight_decay"], betas=(0.9, 0.999))
    g = torch.Generator().manual_seed(5)
    batches = [(torch.rand(2, 1, 64, 64, generator=g) * 2 - 1) for _ in range(2)]
    losses, ts = [], []
    real_randint = torch.randint
    for i, xb in enumerate(batches):
        np.random.seed(100 + i)
        t_fix = real_randint(0, 800, (2,), generator=torch.Generator().manual_seed(50 + i))
        torch.randint = lambda lo, hi, size, device=None, **k: t_fix.to(device)     # p_loss draws t on x.device: pinned
        try:
            x = xb.to(device)
            loss, estimates = diffusion.p_loss(model, x, args)
        finally:
            torch.randint = real_randint
        noisy, est = estimates[1], estimates[2]
        optimiser.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1)
        optimiser.step()
        update_ema_params(ema, model)
        losses.append(float(loss.data.cpu()))
        ts.append(t_fix)
    # ---------------- save() :156-189, reference checkpoint format
    os.makedirs(tmp_path / "model" / "diff-params-ARGS=0", exist_ok=True)
    ckpt_path = tmp_path / "model" / "diff-params-ARGS=0" / "params-final.pt"
    torch.save({"n_epoch": args["EPOCHS"], "model_state_dict": model.state_dict(), "optimizer_state_dict": optimiser.state_dict(),
                "ema": ema.state_dict(), "args": args}, ckpt_path)

    # ---------------- the same two steps on the oracle (fp32 torch on the GPU, C simplex, same seeds / t / batches)
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        psd = {k: v.clone().to(device).requires_grad_(True) for k, v in sd0.items()}
        esd = {k: v.clone().to(device) for k, v in sd0.items()}
        opt_ref = torch.optim.AdamW(list(psd.values()), lr=args["lr"], weight_decay=args["weight_decay"], betas=(0.9, 0.999))
        tab = DO.Tables(DO.beta_schedule(1000, "linear"))
        ref_losses = []
        for i, xb in enumerate(batches):
            np.random.seed(100 + i)
            seeds = [int(np.random.randint(-10000000000, 10000000000)) for _ in range(2)]     # b-major, one channel
            t = ts[i]
            noise = torch.stack([torch.from_numpy(OracleSimplex(s).rand_3d_fixed_T_octaves((64, 64), np.array([int(tt)]), 6, 0.8, 64)
                                                  .astype(np.float32)) for s, tt in zip(seeds, t)])
            x_t = DO.sample_q(tab, xb, t, noise)
            eps = UO.forward(psd, x_t.to(device), t.to(device), **kw)
            loss = ((eps - noise.to(device)) ** 2).mean(dim=(1, 2, 3)).mean()
            opt_ref.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(list(psd.values()), 1)
            opt_ref.step()
            with torch.no_grad():
                for k in esd:
                    esd[k].mul_(0.9999).add_(psd[k].detach(), alpha=1 - 0.9999)
            ref_losses.append(float(loss))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    print(f"[parity] drop-in training replay: losses {losses} vs oracle {ref_losses}")
    assert all(abs(a - b) / abs(b) < 2e-3 for a, b in zip(losses, ref_losses))
    # AdamW's first steps move every weight by ~lr * sign(g): compare the UPDATE, in units of lr
    upd = max(float(((model.state_dict()[k] - sd0[k].to(device)) - (psd[k].detach() - sd0[k].to(device))).abs().max()) for k in psd) / args["lr"]
    ema_err = max(rel(ema.state_dict()[k], esd[k]) for k in esd)
    print(f"[parity] drop-in training replay: worst parameter-update difference {upd:.3f} lr, ema rel {ema_err:.2e}")
    assert ema_err < 1e-5

    # ---------------- detection.py:155-250 on the saved checkpoint
    output = torch.load(ckpt_path, map_location="cpu", weights_only=False)
    a2 = output["args"]
    unet = UNetModel(a2["img_size"][0], a2["base_channels"], channel_mults=a2["channel_mults"], in_channels=1)
    # (detection.py builds the UNet with the constructor-default single head -- SURVEY F6; the checkpoint has the same tensor shapes)
    diff = GaussianDiffusionModel(a2["img_size"], get_beta_schedule(a2["T"], a2["beta_schedule"]), loss_weight=a2["loss_weight"],
                                  loss_type=a2["loss-type"], noise=a2["noise_fn"], img_channels=1)
    unet.load_state_dict(output["ema"])
    unet.to(device)
    unet.eval()
    gi = torch.Generator().manual_seed(9)
    image = (torch.rand(1, 1, 64, 64, generator=gi) * 2 - 1)
    mask = (torch.rand(1, 1, 64, 64, generator=gi) > 0.8).float()
    np.random.seed(777)
    out = diff.forward_backward(unet, image.to(device), see_whole_sequence=None, t_distance=20, denoise_fn=a2["noise_fn"])
    mse = (image.to(device) - out).square()
    fpr_s, tpr_s, _ = EV.ROC_AUC(mask.to(torch.uint8), mse)
    auc = EV.AUC_score(fpr_s, tpr_s)
    mse_t = (mse > 0.5).float()
    dice = EV.dice_coeff(image.to(device), out.to(device), mask.to(device), mse=mse_t).cpu().item()
    prec, rec = EV.precision(mask.to(device), mse_t).cpu().numpy(), EV.recall(mask.to(device), mse_t).cpu().numpy()
    iou, fpr = EV.IoU(mask.to(device), mse_t), EV.FPR(mask.to(device), mse_t).cpu().numpy()
    # oracle chain with the same host RNG stream (B = 1: one np.random seed per noise call, like the reference)
    np.random.seed(777)
    esd_cpu = {k: v.cpu() for k, v in output["ema"].items()}

    def oracle_noise(x, t):
        s = int(np.random.randint(-10000000000, 10000000000))
        return torch.from_numpy(OracleSimplex(s).rand_3d_fixed_T_octaves((64, 64), np.array([int(t[0])]), 6, 0.8, 64).astype(np.float32))[None]

    tab = DO.Tables(DO.beta_schedule(1000, "linear"))
    ref = DO.forward_backward(tab, lambda x, t: UO.forward(esd_cpu, x, t, img_size=64, base_channels=64), image, 20, oracle_noise,
                              oracle_noise)
    r = rel(out.cpu(), ref)
    auc_dev = float(EV.auc_histogram(mask.to(device), mse)[0])
    print(f"[parity] drop-in detection replay: forward_backward(t_distance=20) rel={r:.3e}; AUC {auc:.4f} (device histogram "
          f"{auc_dev:.4f}) dice {dice:.4f} precision {float(prec):.4f} recall {float(rec):.4f} IoU {iou:.4f} FPR {float(fpr):.4f}")
    assert r < 1e-3
    assert abs(auc_dev - auc) < 2e-3 and 0.0 <= auc <= 1.0
