"""GPU-side input pipeline for cached MRI volumes (reference dataset.py:575-643 ``MRIDataset``).

The reference loads a cached ``<id>.npy`` volume ([256][D][192] fp32), takes ``image[:, s:s+1, :]`` and runs a
torchvision transform on the CPU per sample (ToPILImage, RandomAffine, CenterCrop(235), Resize(bilinear), ToTensor,
Normalize(0.5, 0.5)).  Here the volumes go to the device once through pinned staging buffers (asynchronous copies on a side
stream) and one kernel produces the whole normalised batch (``anoddpm_slice_resize_norm``): slice extraction by strides,
CenterCrop, PIL-compatible bilinear resampling and normalisation -- the deterministic part of the transform; the random
affine jitter is not applied.  Checkpoint I/O keeps the reference's ``.pt`` format (plain ``torch.save`` / ``torch.load`` of
``state_dict``s; exercised by tests/test_gpu_dropin_scripts.py).
"""
import numpy as np
import torch

from . import _lib
from ._lib import call, ptr


class MRISlicePipeline:
    def __init__(self, img_size=(256, 256), crop=235, device="cuda"):
        self.img_size = tuple(img_size)
        self.crop = int(crop)
        self.device = torch.device(device)
        self._vols = {}          # key -> device tensor [Hs][D][Ws]
        self._stream = torch.cuda.Stream(device=self.device)

    def add_volume(self, key, volume):
        """volume: numpy [Hs][D][Ws] fp32 (the cached .npy) or a path to it; uploaded once through a pinned buffer"""
        if isinstance(volume, str):
            volume = np.load(volume)
        host = torch.from_numpy(np.ascontiguousarray(volume, dtype=np.float32)).pin_memory()
        with torch.cuda.stream(self._stream):
            self._vols[key] = host.to(self.device, non_blocking=True)
        self._keep = host
        return self

    def batch(self, keys, slice_idx):
        """(B, 1, H, W) fp32 on the device: slice ``slice_idx[i]`` of volume ``keys[i]``, cropped, resized, normalised"""
        torch.cuda.current_stream(self.device).wait_stream(self._stream)
        vols = [self._vols[k] for k in keys]
        Hs, D, Ws = vols[0].shape
        if any(tuple(v.shape) != (Hs, D, Ws) for v in vols):
            raise ValueError("volumes of one batch must share their shape")
        base = min(v.data_ptr() for v in vols)
        offs = torch.tensor([(v.data_ptr() - base) // 4 + int(s) * Ws for v, s in zip(vols, slice_idx)], dtype=torch.int64,
                            device=self.device)
        B = len(keys)
        out = torch.empty((B, 1, *self.img_size), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            call("anoddpm_slice_resize_norm", ptr(out), _lib.c_vp(base), ptr(offs), B, Hs, Ws, D * Ws, self.crop, self.img_size[0],
                 self.img_size[1], _lib.stream())
        return out
