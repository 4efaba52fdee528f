"""Detection metrics of the reference's ``evaluation`` module (reference evaluation.py:26-76), plus a fused
device-side path: ``detection_metrics`` computes the squared error, the 0.5 threshold and
the confusion counts of a whole batch in ONE kernel pass and derives Dice / IoU / precision / recall / FPR / PSNR from
six numbers per image -- no per-image device-to-host copies of full-size tensors.

The individually named functions keep the reference's signatures and (peculiar) definitions: with ``real_mask`` as
"real" and the thresholded error as "recon", the reference calls ``(real == 1) & (recon == 0)`` a false positive
(evaluation.py:59-76); that naming is preserved, not corrected.
"""
import torch

from . import _lib
from ._lib import call, ptr, stream


def dice_coeff(real, recon, real_mask, smooth=0.000001, mse=None):
    if mse is None:
        mse = ((real - recon).square() > 0.5).float()
    intersection = torch.sum(mse * real_mask, dim=[1, 2, 3])
    union = torch.sum(mse, dim=[1, 2, 3]) + torch.sum(real_mask, dim=[1, 2, 3])
    return torch.mean((2. * intersection + smooth) / (union + smooth), dim=0)


def PSNR(recon, real):
    mse = torch.mean((real - recon).square())
    return (20 * torch.log10(torch.max(real) / torch.sqrt(mse))).detach().cpu().numpy()


def IoU(real, recon):
    inter = torch.logical_and(real.bool(), recon.bool()).sum().item()
    union = torch.logical_or(real.bool(), recon.bool()).sum().item()
    return inter / (union + 1e-8)


def precision(real_mask, recon_mask):
    TP = ((real_mask == 1) & (recon_mask == 1))
    FP = ((real_mask == 1) & (recon_mask == 0))
    return torch.sum(TP).float() / ((torch.sum(TP) + torch.sum(FP)).float() + 1e-6)


def recall(real_mask, recon_mask):
    TP = ((real_mask == 1) & (recon_mask == 1))
    FN = ((real_mask == 0) & (recon_mask == 1))
    return torch.sum(TP).float() / ((torch.sum(TP) + torch.sum(FN)).float() + 1e-6)


def FPR(real_mask, recon_mask):
    FP = ((real_mask == 1) & (recon_mask == 0))
    TN = ((real_mask == 0) & (recon_mask == 0))
    return torch.sum(FP).float() / ((torch.sum(FP) + torch.sum(TN)).float() + 1e-6)


def ROC_AUC(real_mask, square_error):
    """(fpr, tpr, thresholds) of the pixel-wise squared error as an anomaly score (evaluation.py:79-83: sklearn's
    roc_curve on the flattened arrays, on the host -- the reference's definition; ``auc_histogram`` is the device path)"""
    from sklearn.metrics import roc_curve
    if isinstance(real_mask, torch.Tensor):
        return roc_curve(real_mask.detach().cpu().numpy().flatten(), square_error.detach().cpu().numpy().flatten())
    return roc_curve(real_mask.flatten(), square_error.flatten())


def AUC_score(fpr, tpr):
    from sklearn.metrics import auc
    return auc(fpr, tpr)


def auc_histogram(real_mask, square_error, bins=65536):
    """Area under the ROC curve on the DEVICE, no host copy of the images: the scores are binned (uniform bins over
    [0, max]), positives / negatives counted per bin, and the trapezoid rule over the bin edges gives
    AUC = sum_b neg_b * (pos_above_b + pos_b / 2) / (P * N) -- exactly sklearn's roc_curve + auc for scores quantised to
    the bins (ties inside a bin get half credit).  Per image when the inputs are batched (dim 0)."""
    m = real_mask.reshape(real_mask.shape[0], -1) > 0
    s = square_error.reshape(square_error.shape[0], -1).float()
    out = []
    for b in range(s.shape[0]):
        hi = s[b].max().clamp_min(1e-30)
        idx = (s[b] / hi * (bins - 1)).round().long().clamp_(0, bins - 1)
        pos = torch.bincount(idx[m[b]], minlength=bins).double()
        neg = torch.bincount(idx[~m[b]], minlength=bins).double()
        P, N = pos.sum(), neg.sum()
        above = P - torch.cumsum(pos, 0)                       # positives in strictly higher bins
        out.append((neg * (above + 0.5 * pos)).sum() / (P * N).clamp_min(1.0))
    return torch.stack(out)


def detection_counts(image, recon, mask, threshold=0.5):
    """(B, 6) float64 on the device: n11, n10, n01, n00, sum of squared error, max(image) per image."""
    _lib.require_cuda(image, "image")
    image, recon, mask = (t.contiguous().float() for t in (image, recon, mask))
    B = image.shape[0]
    counts = torch.zeros((B, 6), device=image.device, dtype=torch.float64)
    counts[:, 5] = float("-inf")
    call("anoddpm_detection_counts", ptr(counts), ptr(image), ptr(recon), ptr(mask), float(threshold), B,
         image.numel() // B, stream())
    return counts


def detection_metrics(image, recon, mask, threshold=0.5, smooth=0.000001):
    """Per-image metrics of reference detection.py:229-250 as (B,) float64 tensors on the device:
    dice, iou, precision, recall, fpr, psnr, mse (same formulas as the functions above, evaluated per image)."""
    c = detection_counts(image, recon, mask, threshold)
    n11, n10, n01, n00, se, mx = (c[:, i] for i in range(6))
    n = image.numel() // image.shape[0]
    thr_sum, mask_sum = n11 + n01, n11 + n10
    return {"dice": (2.0 * n11 + smooth) / (thr_sum + mask_sum + smooth),
            "iou": n11 / (n11 + n10 + n01 + 1e-8),
            "precision": n11 / (n11 + n10 + 1e-6),
            "recall": n11 / (n11 + n01 + 1e-6),
            "fpr": n10 / (n10 + n00 + 1e-6),
            "mse": se / n,
            "psnr": 20 * torch.log10(mx / torch.sqrt(se / n))}
