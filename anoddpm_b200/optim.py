"""Fused optimiser step of the reference training loop (reference diffusion_training.py:75,102-107):

    optimiser.zero_grad(); loss.backward(); clip_grad_norm_(model.parameters(), 1); optimiser.step(); update_ema_params(ema, model)

``FusedAdamWEMA`` re-points the parameters (and gradients, and the EMA copy) of the model into flat fp32 buffers and
runs clip -> AdamW -> EMA as two kernels (csrc/optim.cu); with ``torch.distributed`` initialised the flat gradient is
averaged before the step (data-parallel training): in buckets laid out in REVERSE parameter order, so the
native backward (which produces the gradients of the last layers first) can start reducing a bucket as soon as it is
complete while earlier layers are still being differentiated (``GradReducer.mark_ready`` / ``finish``); a plain
``step()`` without those calls reduces all buckets at once.
"""
import torch
import torch.distributed as dist

from . import _lib
from ._lib import call, ptr, stream


class GradReducer:
    """Bucketed data-parallel averaging of one flat gradient buffer (training row).

    Buckets are contiguous ranges of the flat buffer cut at parameter boundaries, numbered from the END of the buffer:
    backward produces the gradients of the last layers first, so bucket 0 (the tail) is complete first and its
    all-reduce overlaps the differentiation of the earlier layers.  ``mark_ready(offset)`` says "every gradient at
    flat offset >= offset is final"; complete buckets are launched asynchronously in order.  ``finish()`` launches what
    is left, waits, and divides by the world size.  Works on any backend (NCCL on the GPUs, gloo in the CPU tests)."""

    def __init__(self, flat_g, offsets, sizes, bucket_elems=16 << 20, group=None):
        self.flat_g, self.group = flat_g, group
        self.buckets = []                      # (start, end), tail first
        end = int(flat_g.numel())
        cur_end = end
        for off, n in sorted(zip(offsets, sizes), reverse=True):
            if cur_end - off >= bucket_elems:
                self.buckets.append((off, cur_end))
                cur_end = off
        if cur_end > 0:
            self.buckets.append((0, cur_end))
        self.reset()

    def reset(self):
        self.next_bucket = 0
        self.handles = []
        self._done = None

    def track(self, offsets, sizes):
        """per-tensor completion tracking for ``tensor_done``: tensors sorted by offset"""
        order = sorted(range(len(offsets)), key=lambda i: offsets[i])
        self._offsets = [offsets[i] for i in order]
        self._rank = {i: r for r, i in enumerate(order)}
        self._n = len(order)

    def tensor_done(self, i):
        """the gradient of tensor i is final (autograd's post-accumulate hook): buckets whose whole range is final are
        reduced at once, even when the tensors of a layer complete in arbitrary order"""
        if not self.active():
            return
        if self._done is None:
            self._done = [False] * self._n
            self._tail = self._n               # tensors [tail, n) are final
        self._done[self._rank[i]] = True
        while self._tail > 0 and self._done[self._tail - 1]:
            self._tail -= 1
        self.mark_ready(self._offsets[self._tail] if self._tail < self._n else self.flat_g.numel())

    def active(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _launch(self, k):
        s, e = self.buckets[k]
        self.handles.append(dist.all_reduce(self.flat_g[s:e], group=self.group, async_op=True))

    def mark_ready(self, offset):
        if not self.active():
            return
        while self.next_bucket < len(self.buckets) and self.buckets[self.next_bucket][0] >= offset:
            self._launch(self.next_bucket)
            self.next_bucket += 1

    def finish(self):
        if not self.active():
            return
        self.mark_ready(0)
        for h in self.handles:
            h.wait()
        self.flat_g.div_(dist.get_world_size(self.group))
        self.reset()


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
        self.model, self.ema_model = model, ema_model
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
        self.offsets = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += p.numel()
        self._bind_grads()
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
        self.reducer = GradReducer(self.flat_g, self.offsets, [p.numel() for p in self.params])
        self.reducer.track(self.offsets, [p.numel() for p in self.params])
        # overlap: as soon as autograd has accumulated a parameter's gradient it reports in; complete tail buckets go out on
        # NCCL while the backward of the earlier layers is still running
        for i, p in enumerate(self.params):
            p.register_post_accumulate_grad_hook(lambda _p, i=i: self.reducer.tensor_done(i))

    def _bind_grads(self):
        """p.grad = view into the flat gradient buffer (gradients accumulate straight into it)"""
        for p, off in zip(self.params, self.offsets):
            p.grad = self.flat_g[off:off + p.numel()].view_as(p)

    def _collect_grads(self):
        """model.zero_grad(set_to_none=True) (torch's default) or an assignment to p.grad detaches a parameter from the
        flat buffer: copy such gradients in and re-bind, so that step() never silently applies a zero gradient."""
        for p, off in zip(self.params, self.offsets):
            g = p.grad
            want = self.flat_g.data_ptr() + 4 * off
            if g is None:
                self.flat_g[off:off + p.numel()].zero_()
            elif g.data_ptr() != want:
                self.flat_g[off:off + p.numel()].copy_(g.detach().reshape(-1))
            else:
                continue
            p.grad = self.flat_g[off:off + p.numel()].view_as(p)

    def zero_grad(self, set_to_none=False):
        self.flat_g.zero_()
        self._bind_grads()

    # ---- checkpointing (reference diffusion_training.py:173,184 save optimiser.state_dict(); :77 loads it) ----------
    def state_dict(self):
        """torch.optim.AdamW-compatible layout: {"state": {i: {"step", "exp_avg", "exp_avg_sq"}}, "param_groups": [...]}
        (tensors are copies shaped like their parameters), so a checkpoint written here resumes under
        torch.optim.AdamW and vice versa."""
        state = {}
        for i, (p, off) in enumerate(zip(self.params, self.offsets)):
            n = p.numel()
            state[i] = {"step": torch.tensor(float(self.steps)),
                        "exp_avg": self.exp_avg[off:off + n].view_as(p).clone(),
                        "exp_avg_sq": self.exp_avg_sq[off:off + n].view_as(p).clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay,
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                 "fused": None, "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        groups = sd["param_groups"]
        ids = [i for g in groups for i in g["params"]]
        if len(ids) != len(self.params):
            raise ValueError(f"optimiser state has {len(ids)} parameters, the model has {len(self.params)}")
        g0 = groups[0]
        self.lr, self.betas, self.eps = g0["lr"], tuple(g0["betas"]), g0["eps"]
        self.weight_decay = g0["weight_decay"]
        steps = 0
        for pos, i in enumerate(ids):
            st = sd["state"].get(i)
            p, off = self.params[pos], self.offsets[pos]
            n = p.numel()
            if st is None:
                self.exp_avg[off:off + n].zero_()
                self.exp_avg_sq[off:off + n].zero_()
                continue
            if tuple(st["exp_avg"].shape) != tuple(p.shape):
                raise ValueError(f"optimiser state {i}: shape {tuple(st['exp_avg'].shape)} != parameter {tuple(p.shape)}")
            self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps = max(steps, int(float(st["step"])))
        self.steps = steps

    def grad_norm(self):
        """global L2 norm of the (already reduced) gradient, as a device tensor"""
        return self._sumsq.sqrt()

    def step(self):
        self._collect_grads()
        self.reducer.finish()                 # gradient exchange: buckets not already launched by the backward + wait
        self.steps += 1
        self._sumsq.zero_()
        if self.max_norm and self.max_norm > 0:
            call("anoddpm_sumsq", ptr(self._sumsq), ptr(self.flat_g), self.n, stream())
        b1, b2 = self.betas
        call("anoddpm_adamw_ema_step", ptr(self.flat_p), ptr(self.flat_g), ptr(self.exp_avg), ptr(self.exp_avg_sq),
             ptr(self.flat_ema), ptr(self._sumsq), self.n, float(self.lr), float(b1), float(b2), float(self.eps),
             float(self.weight_decay), int(self.steps), float(self.max_norm or 0.0), float(self.ema_decay), stream())
        # the kernels wrote parameters (and the EMA copy) through raw pointers: p._version is unchanged, so tell the
        # inference engines that their packed-weight caches (and CUDA graphs keyed on them) are stale
        for m in (self.model, self.ema_model):
            eng = getattr(m, "_engine", None) if m is not None else None
            if eng is not None:
                eng.invalidate()
