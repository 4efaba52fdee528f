"""Fused optimiser step of the reference training loop (reference diffusion_training.py:75,102-107):

    optimiser.zero_grad(); loss.backward(); clip_grad_norm_(model.parameters(), 1); optimiser.step(); update_ema_params(ema, model)

``FusedAdamWEMA`` re-points the parameters (and gradients, and the EMA copy) of the model into flat fp32 buffers and
runs clip -> AdamW -> EMA as two kernels (csrc/optim.cu); with ``torch.distributed`` initialised the flat gradient is
averaged with ONE all-reduce before the step (data-parallel training).  The backward pass itself is still
PyTorch autograd (UNetModel.allow_autograd_fallback); only the optimiser / EMA / gradient exchange are native here.
"""
import torch
import torch.distributed as dist

from . import _lib
from ._lib import call, ptr, stream


def _flatten_into(params, flat):
    off = 0
    for p in params:
        n = p.numel()
        flat[off:off + n].copy_(p.data.reshape(-1))
        p.data = flat[off:off + n].view_as(p)
        off += n


class FusedAdamWEMA:
    def __init__(self, model, ema_model=None, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_norm=1.0,
                 ema_decay=0.9999):
        self.params = [p for p in model.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        _lib.require_cuda(self.params[0], "model parameters")
        self.n = sum(p.numel() for p in self.params)
        # offsets keep every tensor 16-byte aligned only if all sizes are multiples of 4; pad the flat buffers instead
        self.flat_p = torch.zeros(self.n, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(self.n, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(self.n, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(self.n, device=dev, dtype=torch.float32)
        _flatten_into(self.params, self.flat_p)
        off = 0
        for p in self.params:                      # autograd accumulates straight into the flat gradient buffer
            p.grad = self.flat_g[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.flat_ema = None
        if ema_model is not None:
            named = dict(ema_model.named_parameters())
            ema_params = [named[k] for k, p in model.named_parameters() if p.requires_grad]
            self.flat_ema = torch.zeros(self.n, device=dev, dtype=torch.float32)
            _flatten_into(ema_params, self.flat_ema)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.max_norm, self.ema_decay = max_norm, ema_decay
        self.steps = 0
        self._sumsq = torch.zeros(1, device=dev, dtype=torch.float64)

    def zero_grad(self):
        self.flat_g.zero_()

    def grad_norm(self):
        """global L2 norm of the (already reduced) gradient, as a device tensor"""
        return self._sumsq.sqrt()

    def step(self):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat_g)                      # the single gradient all-reduce of the training step
            self.flat_g.div_(dist.get_world_size())
        self.steps += 1
        self._sumsq.zero_()
        if self.max_norm and self.max_norm > 0:
            call("anoddpm_sumsq", ptr(self._sumsq), ptr(self.flat_g), self.n, stream())
        b1, b2 = self.betas
        call("anoddpm_adamw_ema_step", ptr(self.flat_p), ptr(self.flat_g), ptr(self.exp_avg), ptr(self.exp_avg_sq),
             ptr(self.flat_ema), ptr(self._sumsq), self.n, float(self.lr), float(b1), float(b2), float(self.eps),
             float(self.weight_decay), int(self.steps), float(self.max_norm or 0.0), float(self.ema_decay), stream())
