"""Drop-in for the reference module ``GaussianDiffusion`` (reference GaussianDiffusion.py).

Same functions, class, method names, return conventions and public numpy tables.  The hot path --
``sample_q`` / ``sample_q_gradual`` / ``sample_p`` / ``forward_backward`` and simplex noise generation --
runs on the sm_100a kernels (csrc/diffusion.cu, csrc/simplex.cu) through the C ABI:

  * coefficient tables are uploaded once per device as fp32 (the reference re-uploads a T-entry fp64
    table on every ``extract``, GaussianDiffusion.py:32-36);
  * ``sample_p`` is one fused kernel after the UNet (x0 prediction, clamp, posterior mean, noise mix);
  * simplex noise is written straight into the (B,C,H,W) fp32 noise tensor on the device, with no
    device->host sync of ``t`` (the reference does D2H t -> numba CPU -> H2D per call).

Batched simplex semantic (the reference only supports B == 1): sample b / channel c gets its
own seed, drawn from ``np.random.randint(-1e10, 1e10)`` in (b-major, c-minor) order, i.e. the result
for sample b equals the reference run at B=1 after ``newSeed(seed[b,c])``; B=1 consumes the host RNG
exactly like the reference.

The VLB / hybrid-loss helpers (normal_kl, discretised_gaussian_log_likelihood, calc_vlb_xt,
calc_total_vlb, prior_vlb) and the detection_A/B experiment drivers are kept as plain torch / python
(outside the accelerated path).
"""
import os
import random

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream
from .simplex import Simplex_CLASS, init_tables


def get_beta_schedule(num_diffusion_steps, name="cosine"):
    """fp64 beta schedule (GaussianDiffusion.py:12-29)."""
    if name == "cosine":
        f = lambda t: np.cos((t + 0.008) / 1.008 * np.pi / 2) ** 2
        betas = [min(1 - f((i + 1) / num_diffusion_steps) / f(i / num_diffusion_steps), 0.999)
                 for i in range(num_diffusion_steps)]
        return np.array(betas)
    if name == "linear":
        scale = 1000 / num_diffusion_steps
        return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_steps, dtype=np.float64)
    raise NotImplementedError(f"unknown beta schedule: {name}")


def extract(arr, timesteps, broadcast_shape, device):
    """fp64 table -> gather by t -> fp32 -> broadcast (GaussianDiffusion.py:32-36).  Kept for API parity;
    the kernels read device-resident fp32 tables instead."""
    res = torch.from_numpy(arr).to(device=timesteps.device)[timesteps].float()
    while len(res.shape) < len(broadcast_shape):
        res = res[..., None]
    return res.expand(broadcast_shape).to(device)


def mean_flat(tensor):
    return torch.mean(tensor, dim=list(range(1, len(tensor.shape))))


def normal_kl(mean1, logvar1, mean2, logvar2):
    return 0.5 * (-1 + logvar2 - logvar1 + torch.exp(logvar1 - logvar2) + ((mean1 - mean2) ** 2) * torch.exp(-logvar2))


def approx_standard_normal_cdf(x):
    return 0.5 * (1.0 + torch.tanh(np.sqrt(2.0 / np.pi) * (x + 0.044715 * torch.pow(x, 3))))


def discretised_gaussian_log_likelihood(x, means, log_scales):
    assert x.shape == means.shape == log_scales.shape
    centered_x = x - means
    inv_stdv = torch.exp(-log_scales)
    cdf_plus = approx_standard_normal_cdf(inv_stdv * (centered_x + 1.0 / 255.0))
    cdf_min = approx_standard_normal_cdf(inv_stdv * (centered_x - 1.0 / 255.0))
    log_cdf_plus = torch.log(cdf_plus.clamp(min=1e-12))
    log_one_minus_cdf_min = torch.log((1.0 - cdf_min).clamp(min=1e-12))
    cdf_delta = cdf_plus - cdf_min
    log_probs = torch.where(x < -0.999, log_cdf_plus,
                            torch.where(x > 0.999, log_one_minus_cdf_min, torch.log(cdf_delta.clamp(min=1e-12))))
    assert log_probs.shape == x.shape
    return log_probs


_RANDOM_PARAMS = [(2, 0.6, 16), (6, 0.6, 32), (7, 0.7, 32), (10, 0.8, 64), (5, 0.8, 16), (4, 0.6, 16), (1, 0.6, 64),
                  (7, 0.8, 128), (6, 0.9, 64), (2, 0.85, 128), (2, 0.85, 64), (2, 0.85, 32), (2, 0.85, 16),
                  (2, 0.85, 8), (2, 0.85, 4), (2, 0.85, 2), (1, 0.85, 128), (1, 0.85, 64), (1, 0.85, 32),
                  (1, 0.85, 16), (1, 0.85, 8), (1, 0.85, 4), (1, 0.85, 2)]


def draw_simplex_seeds(n):
    """n host seeds, consumed exactly like n calls of ``Simplex_CLASS.newSeed()`` (simplex.py:19-22)."""
    return [int(np.random.randint(-10000000000, 10000000000)) for _ in range(n)]


def simplex_noise_from_seeds(seeds, shape, t, octave=6, persistence=0.8, frequency=64, shared_seed=False):
    """fp32 noise tensor ``shape`` = (B,C,H,W) on t.device from explicit seeds: (B*C) seeds in b-major /
    c-minor order, or C seeds when ``shared_seed``.  One kernel launch; t stays on the device."""
    B, C, H, W = shape
    _lib.require_cuda(t, "t")
    tables = torch.from_numpy(init_tables(seeds))
    if tables.shape[0] != (C if shared_seed else B * C):
        raise ValueError("wrong number of simplex seeds")
    tables_dev = tables.pin_memory().to(t.device, non_blocking=True)
    noise = torch.empty((B, C, H, W), device=t.device, dtype=torch.float32)
    t64 = t.to(torch.int64).contiguous()
    with torch.cuda.device(t.device):
        call("anoddpm_simplex_noise", ptr(noise), B, C, H, W, ptr(tables_dev), int(shared_seed), ptr(t64), int(octave),
             float(persistence), float(frequency), stream())
    return noise


def generate_simplex_noise(Simplex_instance, x, t, random_param=False, octave=6, persistence=0.8, frequency=64,
                           in_channels=1):
    """GaussianDiffusion.py:96-137, batched per sample (see module docstring)."""
    _lib.require_cuda(x, "x")
    B = x.shape[0]
    H, W = x.shape[-2:]
    seeds = []
    for _ in range(B):
        for _c in range(in_channels):
            seeds.append(draw_simplex_seeds(1)[0])
            if random_param:
                random.choice(_RANDOM_PARAMS)  # drawn and discarded: the reference overwrites that field
    # leave the instance in the state the reference leaves it in: tables of the last seed drawn
    Simplex_instance._load_seed(seeds[-1])
    noise = simplex_noise_from_seeds(seeds, (B, in_channels, H, W), t.to(x.device), octave, persistence, frequency)
    if tuple(x.shape) != tuple(noise.shape):
        out = torch.empty(x.shape, device=x.device)
        out[:, :in_channels] = noise  # same broadcasting as the reference's noise[:, i] assignment
        return out
    return noise


def random_noise(Simplex_instance, x, t):
    param = random.choice(["gauss", "simplex"])
    if param == "gauss":
        return torch.randn_like(x)
    return generate_simplex_noise(Simplex_instance, x, t)


class GaussianDiffusionModel:
    """Diffusion process of reference GaussianDiffusion.py:150-623 (not an nn.Module; numpy tables are public)."""

    TABLE_ROWS = ("sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "sqrt_alphas", "sqrt_betas",
                  "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1",
                  "posterior_mean_coef2")

    def __init__(self, img_size, betas, img_channels=1, loss_type="l2", loss_weight="none", noise="gauss"):
        super().__init__()
        if noise == "gauss":
            self.noise_fn = lambda x, t: torch.randn_like(x)
            self._noise_kind = "gauss"
        else:
            self.simplex = Simplex_CLASS()
            if noise == "simplex_randParam":
                self.noise_fn = lambda x, t: generate_simplex_noise(self.simplex, x, t, True, in_channels=img_channels)
                self._noise_kind = "simplex_randParam"
            elif noise == "random":
                self.noise_fn = lambda x, t: random_noise(self.simplex, x, t)
                self._noise_kind = "random"
            else:
                self.noise_fn = lambda x, t: generate_simplex_noise(self.simplex, x, t, False, in_channels=img_channels)
                self._noise_kind = "simplex"
        self._default_noise_fn = self.noise_fn
        # CUDA-graph fast path of the reverse loop (see _reverse_loop); seeds of step `i` come from simplex_seed_source
        self.use_cuda_graphs = os.environ.get("ANODDPM_CUDA_GRAPHS", "1") == "1"
        self.simplex_seed_source = None      # optional callable(step_index, n) -> list of n seeds (default: np.random)
        self._graphs = {}

        self.img_size = img_size
        self.img_channels = img_channels
        self.loss_type = loss_type
        self.num_timesteps = len(betas)
        if loss_weight == "prop-t":
            self.weights = np.arange(self.num_timesteps, 0, -1)
        elif loss_weight == "uniform":
            self.weights = np.ones(self.num_timesteps)
        self.loss_weight = loss_weight

        alphas = 1 - betas
        self.betas = betas
        self.sqrt_alphas = np.sqrt(alphas)
        self.sqrt_betas = np.sqrt(betas)
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)
        self._dev_tables = {}

    # ---- device-resident fp32 tables -------------------------------------------------------------
    def _model_var(self):
        model_var = np.append(self.posterior_variance[1], self.betas[1:])  # GaussianDiffusion.py:282-283
        return model_var, np.log(model_var)

    def _tables(self, device):
        """fp32(round(fp64)) rows [8 coefficient tables | sigma0], built from the CURRENT public numpy tables."""
        key = str(device)
        if key not in self._dev_tables:
            T = self.num_timesteps
            rows = [torch.from_numpy(np.asarray(getattr(self, n), dtype=np.float64)).float() for n in self.TABLE_ROWS]
            _, logvar = self._model_var()
            # sigma0[t] = [t != 0] * exp(0.5 * fp32(logvar[t])), evaluated in fp32 like the reference (:314-317)
            sigma = torch.exp(0.5 * torch.from_numpy(logvar).float())
            sigma[0] = 0.0
            tab = torch.stack(rows + [sigma]).contiguous()
            assert tab.shape == (9, T)
            self._dev_tables[key] = tab.to(device)
        return self._dev_tables[key]

    @staticmethod
    def _prep(x, what):
        _lib.require_cuda(x, what)
        return x.contiguous().float()

    def _t(self, t, x):
        return t.to(device=x.device, dtype=torch.int64).contiguous()

    # ---- reference API ---------------------------------------------------------------------------
    def sample_t_with_weights(self, b_size, device):
        p = self.weights / np.sum(self.weights)
        indices_np = np.random.choice(len(p), size=b_size, p=p)
        indices = torch.from_numpy(indices_np).long().to(device)
        weights_np = 1 / len(p) * p[indices_np]
        weights = torch.from_numpy(weights_np).float().to(device)
        return indices, weights

    def predict_x_0_from_eps(self, x_t, t, eps):
        return (extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape, x_t.device) * x_t
                - extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape, x_t.device) * eps)

    def predict_eps_from_x_0(self, x_t, t, pred_x_0):
        return (extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape, x_t.device) * x_t - pred_x_0) \
            / extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape, x_t.device)

    def q_mean_variance(self, x_0, t):
        mean = extract(self.sqrt_alphas_cumprod, t, x_0.shape, x_0.device) * x_0
        variance = extract(1.0 - self.alphas_cumprod, t, x_0.shape, x_0.device)
        log_variance = extract(self.log_one_minus_alphas_cumprod, t, x_0.shape, x_0.device)
        return mean, variance, log_variance

    def q_posterior_mean_variance(self, x_0, x_t, t):
        posterior_mean = (extract(self.posterior_mean_coef1, t, x_t.shape, x_t.device) * x_0
                          + extract(self.posterior_mean_coef2, t, x_t.shape, x_t.device) * x_t)
        posterior_var = extract(self.posterior_variance, t, x_t.shape, x_t.device)
        posterior_log_var_clipped = extract(self.posterior_log_variance_clipped, t, x_t.shape, x_t.device)
        return posterior_mean, posterior_var, posterior_log_var_clipped

    def p_mean_variance(self, model, x_t, t, estimate_noise=None):
        """dict mean / variance / log_variance / pred_x_0 (GaussianDiffusion.py:269-296).  Full dict kept for
        the VLB helpers; the sampling path uses the fused kernel in sample_p instead."""
        if estimate_noise is None:
            estimate_noise = model(x_t, t)
        model_var, model_logvar = self._model_var()
        model_var = extract(model_var, t, x_t.shape, x_t.device)
        model_logvar = extract(model_logvar, t, x_t.shape, x_t.device)
        pred_x_0 = self.predict_x_0_from_eps(x_t, t, estimate_noise).clamp(-1, 1)
        model_mean, _, _ = self.q_posterior_mean_variance(pred_x_0, x_t, t)
        return {"mean": model_mean, "variance": model_var, "log_variance": model_logvar, "pred_x_0": pred_x_0}

    def _draw_denoise_noise(self, x_t, t, denoise_fn):
        if type(denoise_fn) == str:
            if denoise_fn == "gauss" or denoise_fn == "random":
                return torch.randn_like(x_t)
            if denoise_fn == "noise_fn":
                return self.noise_fn(x_t, t).float()
            return generate_simplex_noise(self.simplex, x_t, t, False, in_channels=self.img_channels).float()
        return denoise_fn(x_t, t)

    def sample_p(self, model, x_t, t, denoise_fn="gauss"):
        """One reverse step (GaussianDiffusion.py:298-318): UNet forward, noise draw, one fused kernel."""
        x_t = self._prep(x_t, "x_t")
        tt = self._t(t, x_t)
        eps = self._prep(model(x_t, t), "model output")
        noise = self._prep(self._draw_denoise_noise(x_t, t, denoise_fn), "denoise noise")
        if noise.shape != x_t.shape:
            noise = noise.expand_as(x_t).contiguous()
        tab = self._tables(x_t.device)
        sample, pred = torch.empty_like(x_t), torch.empty_like(x_t)
        B = x_t.shape[0]
        with torch.cuda.device(x_t.device):     # launch on the tensors' device, whichever device is current
            call("anoddpm_p_sample", ptr(sample), ptr(pred), ptr(x_t), ptr(eps), ptr(noise), ptr(tt), ptr(tab[4:]),
                 self.num_timesteps, B, x_t.numel() // B, stream())
        return {"sample": sample, "pred_x_0": pred}

    def _q(self, x, t, noise, row):
        x = self._prep(x, "x")
        noise = self._prep(noise, "noise")
        if noise.shape != x.shape:
            noise = noise.expand_as(x).contiguous()
        tt = self._t(t, x)
        tab = self._tables(x.device)
        out = torch.empty_like(x)
        B = x.shape[0]
        with torch.cuda.device(x.device):
            call("anoddpm_q_sample", ptr(out), ptr(x), ptr(noise), ptr(tt), ptr(tab[row]), ptr(tab[row + 1]), B,
                 x.numel() // B, stream())
        return out

    def sample_q(self, x_0, t, noise):
        """q(x_t | x_0) (GaussianDiffusion.py:361-371)."""
        return self._q(x_0, t, noise, 0)

    def sample_q_gradual(self, x_t, t, noise):
        """q(x_t | x_{t-1}) (GaussianDiffusion.py:373-382)."""
        return self._q(x_t, t, noise, 2)

    def forward_backward(self, model, x, see_whole_sequence="half", t_distance=None, denoise_fn="gauss"):
        """Noise to t_distance then denoise back (GaussianDiffusion.py:320-359) -- the detection loop."""
        assert see_whole_sequence == "whole" or see_whole_sequence == "half" or see_whole_sequence == None
        if t_distance == 0:
            return x.detach()
        if t_distance is None:
            t_distance = self.num_timesteps
        seq = [x.cpu().detach()] if see_whole_sequence else []  # (the reference copies x to the host even when unused)
        B = x.shape[0]
        if see_whole_sequence == "whole":
            for t in range(int(t_distance)):
                t_batch = torch.full((B,), t, device=x.device, dtype=torch.int64)
                noise = self.noise_fn(x, t_batch).float()
                with torch.no_grad():
                    x = self.sample_q_gradual(x, t_batch, noise)
                seq.append(x.cpu().detach())
        else:
            t_tensor = torch.full((B,), t_distance - 1, device=x.device, dtype=torch.int64)
            x = self.sample_q(x, t_tensor, self.noise_fn(x, t_tensor).float())
            if see_whole_sequence == "half":
                seq.append(x.cpu().detach())
        if not see_whole_sequence:
            return self.reverse_steps(model, x, int(t_distance) - 1, int(t_distance), denoise_fn).detach()
        for t in range(int(t_distance) - 1, -1, -1):
            t_batch = torch.full((B,), t, device=x.device, dtype=torch.int64)
            with torch.no_grad():
                x = self.sample_p(model, x, t_batch, denoise_fn)["sample"]
            seq.append(x.cpu().detach())
        return seq

    # ---- the reverse loop (GaussianDiffusion.py:351-357), optionally as a replayed CUDA graph ------------------
    def _graph_mode(self, model, x, denoise_fn):
        """Which noise source a captured step may use, or None if this call must run eagerly."""
        from .UNet import UNetModel
        if not (self.use_cuda_graphs and isinstance(model, UNetModel) and isinstance(x, torch.Tensor) and x.is_cuda
                and x.dim() == 4 and type(denoise_fn) == str):
            return None
        if denoise_fn in ("gauss", "random"):
            return "gauss"
        if denoise_fn == "noise_fn":
            if self.noise_fn is not self._default_noise_fn:
                return None                      # user-supplied noise function: arbitrary python, cannot be captured
            return {"gauss": "gauss", "simplex": "simplex", "simplex_randParam": "simplex_randParam"}.get(self._noise_kind)
        return "simplex"                         # any other string -> default-parameter simplex (reference :309-310)

    def _draw_step_seeds(self, step, n):
        """n = B*C seeds of one noise call (b-major, c-minor)."""
        if self.simplex_seed_source is not None:
            seeds = list(self.simplex_seed_source(step, n))
            if len(seeds) != n:
                raise ValueError("simplex_seed_source returned the wrong number of seeds")
            return seeds
        return draw_simplex_seeds(n)

    @torch.no_grad()
    def reverse_steps(self, model, x, t_start, n_steps, denoise_fn="gauss"):
        """n_steps reverse steps t_start, t_start-1, ... (the loop body of forward_backward).  On a CUDA UNetModel with
        a capturable noise source the step (UNet forward + noise + fused posterior + t -= 1) is captured ONCE as a CUDA
        graph and replayed: no per-step host work (simplex seed tables of all steps are built and uploaded up front,
        consuming the host RNG in the same order as the step-by-step loop)."""
        if not (0 <= int(t_start) < self.num_timesteps) or int(n_steps) > int(t_start) + 1:
            raise IndexError(f"reverse_steps: t runs from {int(t_start)} down to {int(t_start) - int(n_steps) + 1}, outside "
                             f"[0, {self.num_timesteps}) (the reference raises IndexError in extract())")
        mode = self._graph_mode(model, x, denoise_fn)
        if mode is None or n_steps <= 2:
            for t in range(int(t_start), int(t_start) - int(n_steps), -1):
                t_batch = torch.full((x.shape[0],), t, device=x.device, dtype=torch.int64)
                x = self.sample_p(model, x, t_batch, denoise_fn)["sample"]
            return x
        return self._graph_loop(model, x, int(t_start), int(n_steps), mode)

    def _graph_loop(self, model, x, t_start, n_steps, mode):
        with torch.cuda.device(x.device):
            return self._graph_loop_on_device(model, x, t_start, n_steps, mode)

    def _graph_loop_on_device(self, model, x, t_start, n_steps, mode):
        x = self._prep(x, "x")
        B, C, H, W = x.shape
        dev = x.device
        eng = model.engine()
        eng._refresh()
        key = (id(model), eng._sig, model.conv_impl, (B, C, H, W), str(dev), mode)
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) >= 4:
                self._graphs.clear()
            g = self._capture_step(model, (B, C, H, W), dev, mode)
            self._graphs[key] = g
        if mode != "gauss":
            chans = self.img_channels
            seeds = []
            for i in range(n_steps):
                seeds.extend(self._draw_step_seeds(g["step_base"] + i, B * chans))
                if mode == "simplex_randParam":
                    for _ in range(B * chans):
                        random.choice(_RANDOM_PARAMS)     # drawn and discarded like the reference
            g["step_base"] += n_steps
            self.simplex._load_seed(seeds[-1])
            tabs = torch.from_numpy(init_tables(seeds)).view(n_steps, B * chans * 512)
            if n_steps > g["tables"].shape[0]:
                raise RuntimeError("more reverse steps than the captured table buffer holds")
            g["tables"][:n_steps].copy_(tabs.pin_memory(), non_blocking=True)
        g["x"].copy_(x)
        g["t"].fill_(t_start)
        g["step"].zero_()
        for _ in range(n_steps):
            g["graph"].replay()
        return g["x"].clone()

    def _capture_step(self, model, shape, dev, mode):
        B, C, H, W = shape
        chans = self.img_channels
        tab = self._tables(dev)
        st = {"x": torch.zeros(shape, device=dev), "t": torch.zeros(B, device=dev, dtype=torch.int64),
              "step": torch.zeros(1, device=dev, dtype=torch.int64), "step_base": 0,
              "tables": torch.zeros((self.num_timesteps, B * chans * 512), device=dev, dtype=torch.int16)}
        x_out = torch.empty(shape, device=dev)

        # the step's noise field does not depend on the UNet output: it is generated on a second, LOW-priority stream forked
        # at the top of the step and joined before the posterior kernel, so its FP64 work fills the SMs the narrow levels of
        # the UNet (8x8 / 16x16 maps: 64 busy CTAs) leave idle instead of extending the step (ANODDPM_NOISE_OVERLAP=0: serial)
        # everything the UNet derives from t alone (timestep MLP + per-block embedding projections), for all T timesteps, once
        use_ttab = os.environ.get("ANODDPM_TIME_TABLE", "1") != "0"
        if use_ttab:
            st["time_table"] = model.engine().time_table(self.num_timesteps, dev)    # kept alive with the graph that reads it
        overlap = mode != "gauss" and os.environ.get("ANODDPM_NOISE_OVERLAP", "0") == "1"
        noise_stream = torch.cuda.Stream(device=dev, priority=0) if overlap else None
        # noise generated inside the posterior kernel (anoddpm_p_sample_simplex, bit-identical to the two kernels): measured
        # 1 % SLOWER per step than the separate noise kernel (23.17 vs 22.93 ms on one box), so it is opt-in
        fused_noise = (mode != "gauss" and not overlap and chans == 1 and C == 1
                       and os.environ.get("ANODDPM_FUSED_NOISE", "0") == "1")

        def make_noise():
            noise = torch.empty((B, chans, H, W), device=dev, dtype=torch.float32)
            call("anoddpm_simplex_noise_steps", ptr(noise), B, chans, H, W, ptr(st["tables"]), ptr(st["step"]), 0,
                 ptr(st["t"]), 6, 0.8, 64.0, stream())
            if chans != C:
                noise = noise.expand(shape).contiguous()
            return noise

        def step():
            if mode == "gauss":
                eps = model.engine().forward(st["x"], st["t"], time_table=use_ttab)
                noise = torch.randn_like(st["x"])
            elif overlap:
                main = torch.cuda.current_stream(dev)
                noise_stream.wait_stream(main)
                with torch.cuda.stream(noise_stream):
                    noise = make_noise()
                eps = model.engine().forward(st["x"], st["t"], time_table=use_ttab)
                main.wait_stream(noise_stream)
            elif fused_noise:
                # single-channel images: the noise value of a pixel is generated inside the posterior kernel (no noise tensor)
                eps = model.engine().forward(st["x"], st["t"], time_table=use_ttab)
                call("anoddpm_p_sample_simplex", ptr(x_out), None, ptr(st["x"]), ptr(eps), ptr(st["tables"]), ptr(st["step"]),
                     ptr(st["t"]), ptr(tab[4:]), self.num_timesteps, B, H, W, 6, 0.8, 64.0, stream())
                st["x"].copy_(x_out)
                st["t"].sub_(1)
                st["step"].add_(1)
                return
            else:
                eps = model.engine().forward(st["x"], st["t"], time_table=use_ttab)
                noise = make_noise()
            call("anoddpm_p_sample", ptr(x_out), None, ptr(st["x"]), ptr(eps), ptr(noise), ptr(st["t"]), ptr(tab[4:]),
                 self.num_timesteps, B, st["x"].numel() // B, stream())
            st["x"].copy_(x_out)
            st["t"].sub_(1)
            st["step"].add_(1)

        # warm-up on a side stream (sets kernel attributes, fills the engine's weight caches), then capture
        st["t"].fill_(self.num_timesteps - 1)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        st["t"].fill_(self.num_timesteps - 1)
        st["step"].zero_()
        # captured on a HIGH-priority stream: kernel nodes inherit the priority of the stream they were captured from, so the
        # UNet's CTAs are placed ahead of the noise kernel's wherever both are pending
        with torch.cuda.graph(graph, stream=torch.cuda.Stream(device=dev, priority=-1) if overlap else None):
            step()
        torch.cuda.synchronize(dev)
        st["graph"] = graph
        return st

    # aliases named in BASELINE.json's north star
    q_sample = sample_q
    p_sample = sample_p
    p_sample_loop = forward_backward

    # ---- VLB / loss (plain torch) --------------------------------------------------
    def _vlb_logvars(self, device):
        """[2][T] fp32 = fp32(posterior_log_variance_clipped), fp32(log model variance): what extract() hands calc_vlb_xt"""
        key = "vlb" + str(device)
        if key not in self._dev_tables:
            _, logvar = self._model_var()
            rows = [torch.from_numpy(np.asarray(self.posterior_log_variance_clipped, dtype=np.float64)).float(),
                    torch.from_numpy(logvar).float()]
            self._dev_tables[key] = torch.stack(rows).contiguous().to(device)
        return self._dev_tables[key]

    def calc_vlb_xt(self, model, x_0, x_t, t, estimate_noise=None):
        """bits/dim of one step (GaussianDiffusion.py:384-397).  Without autograd (calc_total_vlb, evaluation) the KL and
        the decoder NLL come out of one fused kernel; with a gradient flowing into the noise estimate (hybrid loss) the
        torch formulas below are used."""
        if estimate_noise is None:
            estimate_noise = model(x_t, t)
        if x_t.is_cuda and not (torch.is_grad_enabled() and estimate_noise.requires_grad):
            x0c, xtc = self._prep(x_0, "x_0"), self._prep(x_t, "x_t")
            epc = self._prep(estimate_noise.detach(), "noise estimate")
            tt = self._t(t, xtc)
            B = xtc.shape[0]
            per = xtc.numel() // B
            acc = torch.zeros(B, 2, device=xtc.device, dtype=torch.float64)
            pred = torch.empty_like(xtc)
            _lib.call("anoddpm_vlb_terms", _lib.ptr(acc), _lib.ptr(pred), _lib.ptr(x0c), _lib.ptr(xtc), _lib.ptr(epc), _lib.ptr(tt),
                      _lib.ptr(self._tables(xtc.device)[4:]), _lib.ptr(self._vlb_logvars(xtc.device)), self.num_timesteps, B, per,
                      _lib.stream())
            terms = (acc / (per * np.log(2.0))).float()
            return {"output": torch.where(tt == 0, terms[:, 1], terms[:, 0]), "pred_x_0": pred}
        true_mean, _, true_log_var = self.q_posterior_mean_variance(x_0, x_t, t)
        output = self.p_mean_variance(model, x_t, t, estimate_noise)
        kl = mean_flat(normal_kl(true_mean, true_log_var, output["mean"], output["log_variance"])) / np.log(2.0)
        decoder_nll = -discretised_gaussian_log_likelihood(x_0, output["mean"], log_scales=0.5 * output["log_variance"])
        decoder_nll = mean_flat(decoder_nll) / np.log(2.0)
        nll = torch.where((t == 0), decoder_nll, kl)
        return {"output": nll, "pred_x_0": output["pred_x_0"]}

    def calc_loss(self, model, x_0, t):
        noise = self.noise_fn(x_0, t).float()
        x_t = self.sample_q(x_0, t, noise)
        estimate_noise = model(x_t, t)
        loss = {}
        if self.loss_type == "l1":
            loss["loss"] = mean_flat((estimate_noise - noise).abs())
        elif self.loss_type == "hybrid":
            loss["vlb"] = self.calc_vlb_xt(model, x_0, x_t, t, estimate_noise)["output"]
            loss["loss"] = loss["vlb"] + mean_flat((estimate_noise - noise).square())
        else:
            loss["loss"] = mean_flat((estimate_noise - noise).square())
        return loss, x_t, estimate_noise

    def p_loss(self, model, x_0, args):
        if self.loss_weight == "none":
            if args["train_start"]:
                t = torch.randint(0, min(args["sample_distance"], self.num_timesteps), (x_0.shape[0],), device=x_0.device)
            else:
                t = torch.randint(0, self.num_timesteps, (x_0.shape[0],), device=x_0.device)
            weights = 1
        else:
            t, weights = self.sample_t_with_weights(x_0.shape[0], x_0.device)
        loss, x_t, eps_t = self.calc_loss(model, x_0, t)
        return ((loss["loss"] * weights).mean(), (loss, x_t, eps_t))

    def prior_vlb(self, x_0, args):
        t = torch.tensor([self.num_timesteps - 1] * args["Batch_Size"], device=x_0.device)
        qt_mean, _, qt_log_variance = self.q_mean_variance(x_0, t)
        kl_prior = normal_kl(mean1=qt_mean, logvar1=qt_log_variance, mean2=torch.tensor(0.0, device=x_0.device),
                             logvar2=torch.tensor(0.0, device=x_0.device))
        return mean_flat(kl_prior) / np.log(2.0)

    def calc_total_vlb(self, x_0, model, args):
        vb, x_0_mse, mse = [], [], []
        for t in reversed(list(range(self.num_timesteps))):
            t_batch = torch.tensor([t] * args["Batch_Size"], device=x_0.device)
            noise = torch.randn_like(x_0)
            x_t = self.sample_q(x_0=x_0, t=t_batch, noise=noise)
            with torch.no_grad():
                out = self.calc_vlb_xt(model, x_0=x_0, x_t=x_t, t=t_batch)
            vb.append(out["output"])
            x_0_mse.append(mean_flat((out["pred_x_0"] - x_0) ** 2))
            eps = self.predict_eps_from_x_0(x_t, t_batch, out["pred_x_0"])
            mse.append(mean_flat((eps - noise) ** 2))
        vb, x_0_mse, mse = torch.stack(vb, dim=1), torch.stack(x_0_mse, dim=1), torch.stack(mse, dim=1)
        prior_vlb = self.prior_vlb(x_0, args)
        total_vlb = vb.sum(dim=1) + prior_vlb
        return {"total_vlb": total_vlb, "prior_vlb": prior_vlb, "vb": vb, "x_0_mse": x_0_mse, "mse": mse}

    # ---- detection experiment drivers (callers of the hot path; reference :480-623) -----------------
    def _denoise_from(self, model, x, t_distance, denoise_fn="gauss"):
        for t in range(int(t_distance) - 1, -1, -1):
            t_batch = torch.full((x.shape[0],), t, device=x.device, dtype=torch.int64)
            with torch.no_grad():
                x = self.sample_p(model, x, t_batch, denoise_fn)["sample"]
        return x

    def detection_A_fixedT(self, model, x_0, args, mask, end_freq=6):
        """Frequency sweep at fixed t_distance = 250 (GaussianDiffusion.py:596-623); returns the image strip."""
        t_distance = 250
        output = torch.empty((6 * end_freq, 1, *args["img_size"]), device=x_0.device)
        for i in range(1, end_freq + 1):
            freq = 2 ** i
            noise_fn = lambda x, t: generate_simplex_noise(self.simplex, x, t, False, frequency=freq).float()
            t_tensor = torch.full((x_0.shape[0],), t_distance - 1, device=x_0.device, dtype=torch.int64)
            x = self.sample_q(x_0, t_tensor, noise_fn(x_0, t_tensor).float())
            x_noised = x.clone().detach()
            x = self._denoise_from(model, x, t_distance, denoise_fn=noise_fn)
            mse = ((x_0 - x).square() * 2) - 1
            mse_threshold = ((mse > 0).float() * 2) - 1
            output[(i - 1) * 6:i * 6, ...] = torch.cat((x_0, x_noised, x, mse, mse_threshold, mask))
        return output

    @staticmethod
    def _save_grid(out, path, row_size=4):
        """Figure output of the reference drivers (plt.imshow(gridify_output(out, 4)); savefig).  Plotting is outside
        the accelerated path: done only if matplotlib / torchvision are importable, silently skipped otherwise."""
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
            import torchvision.utils
        except Exception:
            return False
        img = ((out + 1) * 127.5).clamp(0, 255).to(torch.uint8)
        grid = torchvision.utils.make_grid(img, nrow=row_size, pad_value=-1).cpu().permute(1, 2, 0)
        plt.imshow(grid, cmap="gray")
        plt.axis("off")
        plt.savefig(path)
        plt.clf()
        return True

    def _noise_then_denoise(self, model, x_0, t_distance, noise_fn, denoise_fn="gauss"):
        """x_0 -> q-sample at index t_distance -> t_distance reverse steps (the loop the reference copy-pastes in
        detection_A / detection_B, GaussianDiffusion.py:500-510, 555-565)."""
        t_tensor = torch.full((x_0.shape[0],), t_distance, device=x_0.device, dtype=torch.int64)
        x = self.sample_q(x_0, t_tensor, noise_fn(x_0, t_tensor).float())
        return self.reverse_steps(model, x, int(t_distance) - 1, int(t_distance), denoise_fn)

    batch_repeats = True      # detection_A / detection_B: run the total_avg repeats of one t_distance as ONE batched chain set

    def _repeat_chains(self, model, x_0, t_distance, total_avg):
        """the `total_avg` independent noise -> denoise chains of one image (reference :500-510, 555-565 loops them one by
        one).  Batched: the image is repeated along the batch axis and every repeat draws its own noise (per-sample simplex
        seeds / independent Gaussian draws), so the repeats share every kernel launch of the UNet instead of running
        total_avg x t_distance single-image forwards.  The host RNG is consumed step-major instead of chain-major: the
        repeats are equally independent draws, but not the same numbers as the sequential order (batch_repeats = False)."""
        if self.batch_repeats and x_0.shape[0] == 1 and total_avg > 1:
            x_rep = x_0.expand(total_avg, *x_0.shape[1:]).contiguous()
            return self._noise_then_denoise(model, x_rep, t_distance, self.noise_fn)
        output = torch.empty((total_avg, *x_0.shape[1:]), device=x_0.device)
        for avg in range(total_avg):
            output[avg, ...] = self._noise_then_denoise(model, x_0, t_distance, self.noise_fn)
        return output

    def detection_A(self, model, x_0, args, file, mask, total_avg=2):
        """Frequency / t_distance sweep with fixed-frequency simplex noise (GaussianDiffusion.py:480-527): every
        sampling chain runs on the native kernels; the PNG output is written when matplotlib is available."""
        out_dir = f"./diffusion-videos/ARGS={args['arg_num']}/Anomalous/{file[0]}/{file[1]}/A"
        os.makedirs(out_dir, exist_ok=True)
        for i in range(7, 0, -1):
            freq = 2 ** i
            self.noise_fn = lambda x, t: generate_simplex_noise(self.simplex, x, t, False, frequency=freq,
                                                                in_channels=self.img_channels)
            for t_distance in range(50, int(args["T"] * 0.6), 50):
                output = self._repeat_chains(model, x_0, t_distance, total_avg)
                output_mean = torch.mean(output, dim=0).reshape(1, 1, *args["img_size"])
                mse = ((output_mean - x_0).square() * 2) - 1
                mse_threshold = ((mse > 0).float() * 2) - 1
                out = torch.cat([x_0, output[:3], output_mean, mse, mse_threshold, mask])
                n_prev = len(os.listdir(out_dir))
                self._save_grid(out, f"{out_dir}/freq={i}-t={t_distance}-{n_prev + 1}.png")

    def detection_B(self, model, x_0, args, file, mask, denoise_fn="gauss", total_avg=5):
        """t_distance sweep with octave simplex ("octave") or Gaussian noise (GaussianDiffusion.py:529-594).  Returns
        the per-t_distance list like the reference (whose evaluation.heatmap returns None)."""
        assert type(file) == tuple
        out_dir = f"./diffusion-videos/ARGS={args['arg_num']}/Anomalous/{file[0]}/{file[1]}/{denoise_fn}"
        os.makedirs(out_dir, exist_ok=True)
        if denoise_fn == "octave":
            end = int(args["T"] * 0.6)
            self.noise_fn = lambda x, t: generate_simplex_noise(self.simplex, x, t, False, frequency=64, octave=6,
                                                                persistence=0.8).float()
        else:
            end = int(args["T"] * 0.8)
            self.noise_fn = lambda x, t: torch.randn_like(x)
        dice_coeff = []
        for t_distance in range(50, end, 50):
            output = self._repeat_chains(model, x_0, t_distance, total_avg)
            output_mean = torch.mean(output, dim=[0]).reshape(1, 1, *args["img_size"])
            n_prev = len(os.listdir(out_dir))
            dice = None   # reference: evaluation.heatmap(...) draws a figure and returns None
            mse = ((output_mean - x_0).square() * 2) - 1
            mse_threshold = ((mse > 0).float() * 2) - 1
            out = torch.cat([x_0, output[:3], output_mean, mse, mse_threshold, mask])
            self._save_grid(out, f"{out_dir}/t={t_distance}-{n_prev + 1}.png")
            dice_coeff.append(dice)
        return dice_coeff
