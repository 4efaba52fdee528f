"""Detection metrics pinned to goldens produced by the reference's own evaluation.py
(tests/golden/make_golden.py gen_eval): the individually named functions on CPU tensors, and -- on the GPU -- the fused
counts kernel + derived metrics and the device ROC area."""
import numpy as np
import pytest
import torch


def _case(g, k):
    return tuple(torch.from_numpy(g[f"{k}/{n}"]) for n in ("image", "recon", "mask"))


@pytest.mark.parametrize("k", ["case0", "case1", "case2"])
def test_named_metric_functions_vs_reference_golden(golden, k):
    from anoddpm_b200 import evaluation as EV
    g = golden("eval")
    image, recon, mask = _case(g, k)
    mse = (image - recon).square()
    thr = (mse > 0.5).float()
    assert np.allclose(EV.dice_coeff(image, recon, mask, mse=thr).numpy(), g[f"{k}/dice"], rtol=1e-6)
    assert np.allclose(EV.dice_coeff(image, recon, mask).numpy(), g[f"{k}/dice_default"], rtol=1e-6)
    assert np.allclose(EV.precision(mask, thr).numpy(), g[f"{k}/precision"], rtol=1e-6)
    assert np.allclose(EV.recall(mask, thr).numpy(), g[f"{k}/recall"], rtol=1e-6)
    assert np.allclose(EV.FPR(mask, thr).numpy(), g[f"{k}/fpr"], rtol=1e-6)
    assert np.isclose(EV.IoU(mask, thr), float(g[f"{k}/iou"]), rtol=1e-6)
    assert np.isclose(float(EV.PSNR(recon, image)), float(g[f"{k}/psnr"]), rtol=1e-5)
    fpr, tpr, _ = EV.ROC_AUC(mask.to(torch.uint8), mse)
    assert np.isclose(EV.AUC_score(fpr, tpr), float(g[f"{k}/auc"]), rtol=1e-9)
    assert abs(float(EV.auc_histogram(mask, mse)[0]) - float(g[f"{k}/auc"])) < 2e-4        # binned scores (65536 bins)


@pytest.mark.gpu
def test_fused_metrics_kernel_vs_reference_golden(golden):
    """anoddpm_detection_counts + detection_metrics (one pass over a batch) against the reference's per-image values"""
    from anoddpm_b200 import evaluation as EV
    g = golden("eval")
    for k in ("case0", "case1", "case2"):
        image, recon, mask = (t.cuda() for t in _case(g, k))
        m = EV.detection_metrics(image, recon, mask)
        assert np.isclose(float(m["dice"][0]), float(g[f"{k}/dice"]), rtol=1e-6)
        assert np.isclose(float(m["precision"][0]), float(g[f"{k}/precision"]), rtol=1e-6)
        assert np.isclose(float(m["recall"][0]), float(g[f"{k}/recall"]), rtol=1e-6)
        assert np.isclose(float(m["fpr"][0]), float(g[f"{k}/fpr"]), rtol=1e-6)
        assert np.isclose(float(m["iou"][0]), float(g[f"{k}/iou"]), rtol=1e-6)
        assert np.isclose(float(m["psnr"][0]), float(g[f"{k}/psnr"]), rtol=1e-5)
        mse = (image - recon).square()
        assert abs(float(EV.auc_histogram(mask, mse)[0]) - float(g[f"{k}/auc"])) < 2e-4
    # batched: three images of one size in one launch == the per-image values
    image, recon, mask = (t.cuda() for t in _case(g, "case0"))
    im3 = torch.cat([image, image.flip(-1), -image])
    re3 = torch.cat([recon, recon.flip(-1), -recon])
    ma3 = torch.cat([mask, mask.flip(-1), mask])
    m3 = EV.detection_metrics(im3, re3, ma3)
    for i in range(3):
        m1 = EV.detection_metrics(im3[i:i + 1], re3[i:i + 1], ma3[i:i + 1])
        assert all(np.isclose(float(m3[key][i]), float(m1[key][0]), rtol=1e-9, equal_nan=True) for key in m3)
    assert np.isclose(float(m3["dice"][0]), float(g["case0/dice"]), rtol=1e-6)


@pytest.mark.gpu
def test_gpu_input_pipeline_vs_reference_transform(golden):
    """anoddpm_slice_resize_norm: slice extraction + CenterCrop + PIL 8-bit bilinear Resize + ToTensor +
    Normalize on the device, bit-exact against the torchvision calls of the reference's MRIDataset (tests/golden/pipeline.npz)."""
    from anoddpm_b200.pipeline import MRISlicePipeline
    g = golden("pipeline")
    vol = g["vol"]
    for name, size in (("s256", (256, 256)), ("s64", (64, 64)), ("s128x96", (128, 96))):
        P = MRISlicePipeline(img_size=size, crop=235).add_volume("a", vol).add_volume("b", vol[::-1].copy())
        x = P.batch(["a", "a", "b"], [0, 3, 3])
        assert x.shape == (3, 1, *size) and x.is_cuda
        assert np.array_equal(x[0].cpu().numpy(), g[f"{name}/slice0"])
        assert np.array_equal(x[1].cpu().numpy(), g[f"{name}/slice3"])
        assert not np.array_equal(x[2].cpu().numpy(), g[f"{name}/slice3"])           # the flipped volume is a different image
