"""Native training step of the UNet denoiser (reference diffusion_training.py:99-107: ``loss.backward()`` through
``UNetModel.forward``, UNet.py:390-406).

The forward runs the same sm_100a kernels as inference (GroupNorm passes that materialise the fp16 operands, tcgen05
convolutions with fused bias / embedding / skip / residual, fused attention) and keeps what the backward needs; the backward
is hand-written per block on our own kernels:

  * data gradients of every 3x3 / 1x1 convolution: ``anoddpm_conv`` with the spatially flipped, transposed weight pack
    (tcgen05 implicit GEMM, fp16 operands);
  * weight gradients: ``anoddpm_conv_wgrad`` (tcgen05, MN-major operands straight from the NHWC tensors, split-K);
  * GroupNorm + SiLU (+ avg-pool / upsample) backward: ``anoddpm_gn_silu_bwd`` (two passes; also adds the gradient that
    reaches the block input on its skip path and emits the bias / timestep-embedding gradients of the first conv);
  * attention core, timestep MLPs: ``anoddpm_gemm_batched`` + softmax rows; stem / head: dedicated kernels.

torch.autograd is used only as the TAPE between blocks (it calls each block's backward and sums the gradients of tensors
with two consumers -- the skip tensors of the up path); no cuDNN / cuBLAS kernel takes part.  Activation gradients are
fp16 tensor-core operands, so they carry a power-of-two loss scale chosen from the incoming gradient; parameter gradients
are un-scaled where they are produced (fp32).
"""
import ctypes
import math
import os

import torch

from . import _lib
from ._lib import ConvDesc, call, ptr, stream
from .engine import STAT_REPLICAS, UNetEngine, _T, pack_conv_weight


class _Ctx:
    """Per-forward shared state: the loss scale is fixed by the first backward call (the output gate)."""

    def __init__(self):
        self.scale = 1.0


def _f16_ok(*channels):
    return all(c % 64 == 0 for c in channels)


# ---------------------------------------------------------------------------------------------------------- helpers
def _stats(x, B, HW):
    """[REPLICAS][B][C][2] fp64 per-channel (sum, sum of squares) of an NHWC fp32 tensor"""
    C = x.shape[-1]
    st = torch.zeros(STAT_REPLICAS * B * C * 2, device=x.device, dtype=torch.float64)
    call("anoddpm_gn_stats", ptr(st), ptr(x), C, B, HW, stream())
    return st


def _gn_apply(x1, st1, x2, st2, gn, B, H, W, silu, mode, want_raw=False):
    """fp16 resample(act(GN(cat(x1, x2)))) (+ the fp16 raw concatenated copy, or the fp32 pooled input in mode 1)"""
    C1 = x1.shape[-1]
    C2 = x2.shape[-1] if x2 is not None else 0
    C = C1 + C2
    Ho, Wo = (H // 2, W // 2) if mode == 1 else ((H * 2, W * 2) if mode == 2 else (H, W))
    out = torch.empty((B, Ho, Wo, C), device=x1.device, dtype=torch.float16)
    raw = None
    if want_raw:
        raw = torch.empty((B, Ho, Wo, C), device=x1.device, dtype=torch.float16 if mode == 0 else torch.float32)
    call("anoddpm_gn_apply", ptr(out), ptr(raw), ptr(x1), C1, ptr(x2), C2, ptr(st1), ptr(st2), ptr(gn.weight), ptr(gn.bias),
         B, H, W, 32, float(gn.eps), int(silu), mode, 2, stream())
    return out, raw


def _conv(a16, w_packed16, bias, B, H, W, cout, ksize, r=None, emb=None, emb_stride=0, res=None, res_mode=0, out_f16=False,
          want_stats=False, impl=None):
    """anoddpm_conv with fp16 operands; returns (out, stats or None)"""
    L = _lib.lib()
    d = ConvDesc()
    d.B, d.H, d.W, d.Ca, d.Cout, d.ksize = B, H, W, a16.shape[-1], cout, ksize
    d.operand_f16 = 1
    if r is not None:
        d.r1, d.Cr1 = r.data_ptr(), r.shape[-1]
    if impl is None:
        impl = 3 if L.anoddpm_conv_tc3_supported(ctypes.byref(d)) else 2
        if impl == 2 and not L.anoddpm_conv_tc_supported(ctypes.byref(d)):
            raise NotImplementedError(f"no tensor-core conv for Ca={d.Ca} Cout={cout} H={H} W={W} ksize={ksize}")
    if impl != 3:
        out_f16 = False
    out = torch.empty((B, H, W, cout), device=a16.device, dtype=torch.float16 if out_f16 else torch.float32)
    d.out, d.a, d.w = out.data_ptr(), a16.data_ptr(), w_packed16.data_ptr()
    d.bias = bias.data_ptr() if bias is not None else None
    if emb is not None:
        d.emb, d.emb_stride = emb.data_ptr(), emb_stride
    if res is not None:
        d.res, d.res_mode = res.data_ptr(), res_mode
    d.impl, d.out_f16 = impl, 1 if out_f16 else 0
    st = None
    if want_stats and (impl == 3 or (H * W) % 32 == 0):
        st = torch.zeros(STAT_REPLICAS * B * cout * 2, device=a16.device, dtype=torch.float64)
        d.gn_stats = st.data_ptr()
    call("anoddpm_conv", ctypes.byref(d), stream())
    return out, st, out_f16


def _ksize(w):
    return 1 if (w.dim() == 3 or w.shape[-1] == 1) else 3


def _pack_fwd(weights):
    """forward pack [Cout][taps*Cin (+ Cskip)] fp16 of one or two OIHW / OI1 parameters (one native pass each)"""
    cout = weights[0].shape[0]
    widths = [w.shape[1] * _ksize(w) ** 2 for w in weights]
    ktot = sum(widths)
    out = torch.empty((cout, ktot), device=weights[0].device, dtype=torch.float16)
    off = 0
    for w, n in zip(weights, widths):
        wd = w.detach()
        assert wd.is_contiguous() and wd.dtype == torch.float32
        call("anoddpm_pack_conv_w", ptr(out), ptr(wd), cout, w.shape[1], _ksize(w), ktot, off, 0, stream())
        off += n
    return out


def _pack_dgrad(w):
    """data-gradient pack: d_a = conv(dy, w') with w'[ci][tap'][co] = w[co][ci][2-ky][2-kx] (flipped, transposed)"""
    wd = w.detach()
    assert wd.is_contiguous() and wd.dtype == torch.float32
    co, ci, k = w.shape[0], w.shape[1], _ksize(w)
    out = torch.empty((ci, k * k * co), device=w.device, dtype=torch.float16)
    call("anoddpm_pack_conv_w", ptr(out), ptr(wd), co, ci, k, k * k * co, 0, 1, stream())
    return out


# Data gradients that feed an ELEMENTWISE consumer (the GroupNorm backward) stay fp32 by default: an fp16 store there is
# a 2^-11 rounding per element per layer that nothing averages (measured: per-tensor gradient error 2.2e-3 with fp16
# stores vs the fp32 reference); operands of the next tensor-core GEMM are fp16 either way.
DA_F16 = os.environ.get("ANODDPM_TRAIN_DA_F16", "0") == "1"


def _dgrad(w, dy16, B, H, W, want_f16=None):
    """gradient wrt the conv input: [B][H][W][Cin] fp32 (fp16 on request, halo-reuse kernel only); returns (tensor, is_f16)"""
    if want_f16 is None:
        want_f16 = DA_F16
    ksize = _ksize(w)
    cin = w.shape[1]
    out, _, f16 = _conv(dy16, _pack_dgrad(w), None, B, H, W, cin, ksize, out_f16=want_f16)
    return out, f16


def _pad_c(t, mult=128):
    C = t.shape[-1]
    Cp = (C + mult - 1) // mult * mult
    if Cp == C:
        return t
    out = torch.zeros(t.shape[:-1] + (Cp,), device=t.device, dtype=t.dtype)
    out[..., :C] = t
    return out


def _wgrad(dy16, a16, B, H, W, ksize, alpha, like):
    """weight gradient shaped like the parameter `like` (OIHW or OI1), fp32"""
    cout, cin = dy16.shape[-1], a16.shape[-1]
    dyp, ap = _pad_c(dy16), _pad_c(a16)          # channel counts that are not multiples of 128 run zero-padded
    cop, cip = dyp.shape[-1], ap.shape[-1]
    if not _lib.lib().anoddpm_conv_wgrad_supported(B, H, W, cip, cop, ksize):
        raise NotImplementedError(f"no tensor-core weight gradient for H={H} W={W} (needs H % 8 == 0, W == 8 or W % 16 == 0)")
    dw = torch.zeros((cop, ksize * ksize, cip), device=dy16.device, dtype=torch.float32)
    call("anoddpm_conv_wgrad", ptr(dw), ptr(dyp), ptr(ap), B, H, W, cip, cop, ksize, float(alpha), stream())
    dw = dw[:cout, :, :cin]
    g = dw.reshape(cout, ksize, ksize, cin).permute(0, 3, 1, 2)
    return g.reshape(like.shape).contiguous()


def _cast_colsum(d_out, alpha, want_colsum=True):
    """fp16 copy of an fp32 gradient [.., C] and alpha * column sums (the bias gradient of the producing conv)"""
    C = d_out.shape[-1]
    N = d_out.numel() // C
    y = torch.empty(d_out.shape, device=d_out.device, dtype=torch.float16)
    cs = torch.zeros(C, device=d_out.device, dtype=torch.float32) if want_colsum else None
    call("anoddpm_cast_colsum", ptr(y), ptr(d_out), ptr(cs), N, C, 1.0, float(alpha), stream())
    return y, cs


def _gn_bwd(x1, x1_f16, st1, x2, st2, gn, shift_c, shift_bc, shift_stride, da, da_f16, add, add_mode, add_f16, B, H, W, silu,
            mode, alpha, want_dsum=False):
    """returns dx1, dx2, dgamma, dbeta, dsum"""
    C1 = x1.shape[-1]
    C2 = x2.shape[-1] if x2 is not None else 0
    C = C1 + C2
    dev = x1.device
    dx1 = torch.empty(x1.shape, device=dev, dtype=torch.float16 if x1_f16 else torch.float32)
    dx2 = torch.empty(x2.shape, device=dev, dtype=torch.float32) if x2 is not None else None
    dg = torch.zeros(C, device=dev, dtype=torch.float32)
    db = torch.zeros(C, device=dev, dtype=torch.float32)
    dsum = torch.zeros((B, C), device=dev, dtype=torch.float32) if want_dsum else None
    ws = torch.empty(B * C * 2, device=dev, dtype=torch.float32)
    call("anoddpm_gn_silu_bwd", ptr(dx1), ptr(dx2), ptr(dg), ptr(db), ptr(dsum), ptr(ws), ptr(x1), int(x1_f16), C1, ptr(x2), C2,
         ptr(st1), ptr(st2), ptr(gn.weight), ptr(gn.bias), ptr(shift_c), ptr(shift_bc), int(shift_stride), ptr(da), int(da_f16),
         ptr(add), int(add_mode), int(add_f16), B, H, W, 32, float(gn.eps), int(silu), int(mode), float(alpha), stream())
    return dx1, dx2, dg, db, dsum


def _gemm(c, a, b, M, N, K, a_ms, a_ks, b_ks, b_ns, ldc, alpha=1.0, accumulate=False, n1=1, n2=1, sa=(0, 0), sb=(0, 0), sc=(0, 0)):
    call("anoddpm_gemm_batched", ptr(c), ptr(a), ptr(b), M, N, K, a_ms, a_ks, b_ks, b_ns, ldc, n1, n2, sa[0], sa[1], sb[0], sb[1],
         sc[0], sc[1], float(alpha), int(accumulate), stream())


class _View:
    """a raw device address inside a tensor (kept alive by `base`): lets the batched GEMM address q / k / v slices"""

    def __init__(self, base, offset_elems):
        self.base, self.off = base, offset_elems

    def data_ptr(self):
        return self.base.data_ptr() + 4 * self.off


# ---------------------------------------------------------------------------------------------------------- Functions
class _OutputGate(torch.autograd.Function):
    """identity; its backward fixes the loss scale for this backward pass from the incoming gradient"""

    @staticmethod
    def forward(ctx, cx, y):
        ctx.cx = cx
        return y.view_as(y)

    @staticmethod
    def backward(ctx, g):
        amax = float(g.abs().max())
        s = 1.0
        if amax > 0 and math.isfinite(amax):
            s = 2.0 ** max(-24, min(24, math.floor(math.log2(1024.0 / amax))))
        ctx.cx.scale = s
        return None, g * s


class _StemFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cx, x, w, b):
        B, Cin, H, W = x.shape
        C0 = w.shape[0]
        xc = x.contiguous().float()
        h = torch.empty((B, H, W, C0), device=x.device, dtype=torch.float32)
        call("anoddpm_conv_stem", ptr(h), ptr(xc), ptr(w), ptr(b), B, Cin, H, W, C0, stream())
        ctx.cx, ctx.shape = cx, (B, Cin, H, W, C0)
        ctx.save_for_backward(xc)
        return h

    @staticmethod
    def backward(ctx, dh):
        (xc,) = ctx.saved_tensors
        B, Cin, H, W, C0 = ctx.shape
        if Cin != 1:
            raise NotImplementedError("native stem weight gradient: single-channel images (the reference's datasets)")
        alpha = 1.0 / ctx.cx.scale
        dy16, db = _cast_colsum(dh.contiguous(), alpha)
        dw = torch.zeros((C0, 9), device=dh.device, dtype=torch.float32)
        call("anoddpm_stem_wgrad", ptr(dw), ptr(dy16), ptr(xc), B, H, W, C0, float(alpha), stream())
        return None, None, dw.view(C0, 1, 3, 3), db


class _HeadFn(torch.autograd.Function):
    """GroupNorm -> SiLU -> conv3x3(C -> 1) (UNet.py:384-388), output NCHW"""

    @staticmethod
    def forward(ctx, cx, h, st, gn_w, gn_b, w, b, gn):
        B, H, W, C = h.shape
        if st is None:
            st = _stats(h, B, H * W)
        out = torch.empty((B, w.shape[0], H, W), device=h.device, dtype=torch.float32)
        wp = pack_conv_weight(w.detach()).contiguous().float()
        call("anoddpm_conv_head_gn", ptr(out), ptr(h), ptr(st), ptr(gn_w), ptr(gn_b), ptr(wp), ptr(b), B, C, H, W, w.shape[0],
             32, float(gn.eps), stream())
        ctx.cx, ctx.gn = cx, gn
        ctx.save_for_backward(h, st, wp, w)
        return out

    @staticmethod
    def backward(ctx, dy):
        h, st, wp, w = ctx.saved_tensors
        B, H, W, C = h.shape
        if w.shape[0] != 1:
            raise NotImplementedError("native head gradient: single-channel output")
        alpha = 1.0 / ctx.cx.scale
        gn = ctx.gn
        a16, _ = _gn_apply(h, st, None, None, gn, B, H, W, True, 0)         # recomputed, not stored
        da16 = torch.empty((B, H, W, C), device=h.device, dtype=torch.float16)
        dwp = torch.zeros((9, C), device=h.device, dtype=torch.float32)
        dbias = torch.zeros(1, device=h.device, dtype=torch.float32)
        dyc = dy.contiguous().float()
        call("anoddpm_head_bwd", ptr(da16), ptr(dwp), ptr(dbias), ptr(dyc), ptr(a16), ptr(wp), B, H, W, C, float(alpha), stream())
        dh, _, dg, db, _ = _gn_bwd(h, False, st, None, None, gn, None, None, 0, da16, True, None, 0, 0, B, H, W, True, 0, alpha)
        dw = dwp.view(3, 3, C).permute(2, 0, 1).reshape(1, C, 3, 3).contiguous()
        return None, dh, None, dg, db, dw, dbias, None


class _TimeMLPFn(torch.autograd.Function):
    """emb = W2 SiLU(W1 temb + b1) + b2 (UNet.py:271-276); temb carries no gradient"""

    @staticmethod
    def forward(ctx, cx, temb, w1, b1, w2, b2):
        B, base = temb.shape
        E = w1.shape[0]
        e1 = torch.empty((B, E), device=temb.device, dtype=torch.float32)
        call("anoddpm_linear", ptr(e1), E, ptr(temb), base, ptr(w1), ptr(b1), B, E, base, 0, stream())
        emb = torch.empty((B, E), device=temb.device, dtype=torch.float32)
        call("anoddpm_linear", ptr(emb), E, ptr(e1), E, ptr(w2), ptr(b2), B, E, E, 1, stream())
        ctx.cx = cx
        ctx.save_for_backward(temb, e1, w2)
        return emb

    @staticmethod
    def backward(ctx, d_emb):
        temb, e1, w2 = ctx.saved_tensors
        B, base = temb.shape
        E = e1.shape[1]
        alpha = 1.0 / ctx.cx.scale
        d_emb = d_emb.contiguous()
        dev = temb.device
        se1 = torch.nn.functional.silu(e1)                                  # [B, E] (a handful of elements)
        dw2 = torch.empty((E, E), device=dev, dtype=torch.float32)
        _gemm(dw2, d_emb, se1, E, E, B, 1, E, E, 1, E, alpha)               # dw2[n][k] = sum_b d_emb[b][n] se1[b][k]
        ones = torch.ones(B, device=dev, dtype=torch.float32)
        db2 = torch.empty(E, device=dev, dtype=torch.float32)
        _gemm(db2, ones, d_emb, 1, E, B, 0, 1, E, 1, E, alpha)
        d_se1 = torch.empty((B, E), device=dev, dtype=torch.float32)
        _gemm(d_se1, d_emb, w2, B, E, E, E, 1, E, 1, E)                     # d_se1[b][k] = sum_n d_emb[b][n] w2[n][k]
        d_e1 = torch.empty((B, E), device=dev, dtype=torch.float32)
        call("anoddpm_silu_bwd", ptr(d_e1), ptr(d_se1), ptr(e1), B * E, stream())
        dw1 = torch.empty((E, base), device=dev, dtype=torch.float32)
        _gemm(dw1, d_e1, temb, E, base, B, 1, E, base, 1, base, alpha)
        db1 = torch.empty(E, device=dev, dtype=torch.float32)
        _gemm(db1, ones, d_e1, 1, E, B, 0, 1, E, 1, E, alpha)
        return None, None, dw1, db1, dw2, db2


class _EmbProjFn(torch.autograd.Function):
    """rows[b, off_k : off_k + C_k] = W_k SiLU(emb[b]) + b_k for every ResBlock k (UNet.py:185-188), one launch"""

    @staticmethod
    def forward(ctx, cx, emb, *wb):
        ws, bs = wb[0::2], wb[1::2]
        wcat = torch.cat([w.detach() for w in ws], dim=0).contiguous().float()
        bcat = torch.cat([b.detach() for b in bs], dim=0).contiguous().float()
        B, E = emb.shape
        S = wcat.shape[0]
        rows = torch.empty((B, S), device=emb.device, dtype=torch.float32)
        call("anoddpm_linear", ptr(rows), S, ptr(emb), E, ptr(wcat), ptr(bcat), B, S, E, 1, stream())
        ctx.cx, ctx.sizes = cx, [w.shape[0] for w in ws]
        ctx.save_for_backward(emb, wcat)
        return rows

    @staticmethod
    def backward(ctx, d_rows):
        emb, wcat = ctx.saved_tensors
        B, E = emb.shape
        S = wcat.shape[0]
        alpha = 1.0 / ctx.cx.scale
        dev = emb.device
        d_rows = d_rows.contiguous()
        se = torch.nn.functional.silu(emb)
        dw = torch.empty((S, E), device=dev, dtype=torch.float32)
        _gemm(dw, d_rows, se, S, E, B, 1, S, E, 1, E, alpha)
        ones = torch.ones(B, device=dev, dtype=torch.float32)
        dbias = torch.empty(S, device=dev, dtype=torch.float32)
        _gemm(dbias, ones, d_rows, 1, S, B, 0, 1, S, 1, S, alpha)
        # d_se[b][k] = sum_n d_rows[b][n] wcat[n][k]: the reduction (S ~ 12 K) is split over the batch dimension of the GEMM
        KS = 256
        nz = (S + KS - 1) // KS
        if S % KS == 0:
            part = torch.empty((nz, B, E), device=dev, dtype=torch.float32)
            _gemm(part, d_rows, wcat, B, E, KS, S, 1, E, 1, E, n1=nz, n2=1, sa=(KS, 0), sb=(KS * E, 0), sc=(B * E, 0))
            d_se = part.sum(dim=0)
        else:
            d_se = torch.empty((B, E), device=dev, dtype=torch.float32)
            _gemm(d_se, d_rows, wcat, B, E, S, S, 1, E, 1, E)
        d_emb = torch.empty((B, E), device=dev, dtype=torch.float32)
        call("anoddpm_silu_bwd", ptr(d_emb), ptr(d_se), ptr(emb), B * E, stream())
        grads = []
        off = 0
        for n in ctx.sizes:
            grads += [dw[off:off + n], dbias[off:off + n]]
            off += n
        return (None, d_emb) + tuple(grads)


class _ResFn(torch.autograd.Function):
    """One ResBlock (UNet.py:156-217) on NHWC tensors.  inputs: x1, x2 (or None), rows [B, S] (all blocks' embedding rows),
    then gn1.w, gn1.b, conv1.w, conv1.b, gn2.w, gn2.b, conv2.w, conv2.b [, skip.w, skip.b]"""

    @staticmethod
    def forward(ctx, cx, rb, emb_off, x1, st1, x2, st2, rows, *p):
        gn1, gn2 = rb.in_layers[0], rb.out_layers[0]
        conv1_w, conv1_b, conv2_w, conv2_b = p[2], p[3], p[6], p[7]
        skip_conv = len(p) > 8
        mode = {"same": 0, "down": 1, "up": 2}[rb.resample_mode]
        B, H, W, C1 = x1.shape
        C2 = x2.shape[-1] if x2 is not None else 0
        cin, cout = C1 + C2, conv1_w.shape[0]
        if not _f16_ok(cin, cout):
            raise NotImplementedError(f"native training needs channel counts that are multiples of 64 (got {cin} -> {cout})")
        if rb.out_layers[2].p > 0:
            raise NotImplementedError("native training: dropout > 0 is not implemented")
        Ho, Wo = (H // 2, W // 2) if mode == 1 else ((H * 2, W * 2) if mode == 2 else (H, W))
        if st1 is None:        # statistics come out of the producing conv's epilogue where there is one
            st1 = _stats(x1, B, H * W)
        if x2 is not None and st2 is None:
            st2 = _stats(x2, B, H * W)
        a1, raw = _gn_apply(x1, st1, x2, st2, gn1, B, H, W, True, mode, want_raw=skip_conv or mode == 1)
        S = rows.shape[1]
        emb = _View(rows, emb_off)
        h, st_h, h16 = _conv(a1, _pack_fwd([conv1_w]), conv1_b, B, Ho, Wo, cout, 3, emb=emb, emb_stride=S, out_f16=True,
                             want_stats=True)
        if st_h is None:
            st_h = _stats(h, B, Ho * Wo)
        if h16:
            a2 = torch.empty((B, Ho, Wo, cout), device=x1.device, dtype=torch.float16)
            call("anoddpm_gn_apply_f16in", ptr(a2), ptr(h), cout, ptr(st_h), ptr(gn2.weight), ptr(gn2.bias), ptr(conv1_b), ptr(emb),
                 S, B, Ho, Wo, 32, float(gn2.eps), 1, 2, stream())
        else:
            a2, _ = _gn_apply(h, st_h, None, None, gn2, B, Ho, Wo, True, 0)
        if skip_conv:
            assert mode == 0
            out, st_out, _ = _conv(a2, _pack_fwd([conv2_w, p[8]]), (conv2_b + p[9]).detach().float().contiguous(), B, Ho, Wo, cout, 3,
                                   r=raw, want_stats=True)
        elif mode == 1:
            out, st_out, _ = _conv(a2, _pack_fwd([conv2_w]), conv2_b, B, Ho, Wo, cout, 3, res=raw, res_mode=1, want_stats=True)   # raw = pooled x
        else:
            out, st_out, _ = _conv(a2, _pack_fwd([conv2_w]), conv2_b, B, Ho, Wo, cout, 3, res=x1, res_mode=1 if mode == 0 else 3,
                                   want_stats=True)
        if st_out is None:
            st_out = _stats(out, B, Ho * Wo)
        ctx.cx, ctx.rb, ctx.emb_off, ctx.meta = cx, rb, emb_off, (mode, skip_conv, B, H, W, Ho, Wo, C1, C2, cout, h16, S)
        ctx.x2_none = x2 is None
        ctx.save_for_backward(x1, x2 if x2 is not None else x1, st1, st2 if st2 is not None else st1, a1,
                              raw if (skip_conv and raw is not None) else a1, h, st_h, a2, rows, *p)
        ctx.mark_non_differentiable(st_out)
        return out, st_out

    @staticmethod
    def backward(ctx, d_out, _d_st=None):
        x1, x2, st1, st2, a1, raw, h, st_h, a2, rows = ctx.saved_tensors[:10]
        p = ctx.saved_tensors[10:]
        mode, skip_conv, B, H, W, Ho, Wo, C1, C2, cout, h16, S = ctx.meta
        if ctx.x2_none:
            x2 = st2 = None
        rb = ctx.rb
        gn1, gn2 = rb.in_layers[0], rb.out_layers[0]
        conv1_w, conv1_b, conv2_w = p[2], p[3], p[6]
        alpha = 1.0 / ctx.cx.scale
        d_out = d_out.contiguous()
        emb = _View(rows, ctx.emb_off)
        # conv2: bias + weight + data gradients
        dy16, db2 = _cast_colsum(d_out, alpha)
        dw2 = _wgrad(dy16, a2, B, Ho, Wo, 3, alpha, conv2_w)
        d_a2, d_a2_f16 = _dgrad(conv2_w, dy16, B, Ho, Wo)
        # GroupNorm 2 + SiLU backward -> dh (fp16 when h is the fp16 conv sum), bias / embedding-row gradients of conv1
        if h16:
            dh, _, dg2, dbt2, dsum = _gn_bwd(h, True, st_h, None, None, gn2, conv1_b, emb, S, d_a2, d_a2_f16, None, 0, 0, B, Ho, Wo,
                                             True, 0, alpha, want_dsum=True)
            dh16 = dh
        else:
            dh, _, dg2, dbt2, dsum = _gn_bwd(h, False, st_h, None, None, gn2, None, None, 0, d_a2, d_a2_f16, None, 0, 0, B, Ho, Wo,
                                             True, 0, alpha, want_dsum=True)
            dh16, _ = _cast_colsum(dh, alpha, want_colsum=False)
        db1 = dsum.sum(dim=0) * alpha                                           # [cout]  (B x cout numbers)
        d_rows = torch.zeros((B, S), device=d_out.device, dtype=torch.float32)
        d_rows[:, ctx.emb_off:ctx.emb_off + cout] = dsum                        # still carries the loss scale
        # conv1
        dw1 = _wgrad(dh16, a1, B, Ho, Wo, 3, alpha, conv1_w)
        d_a1, d_a1_f16 = _dgrad(conv1_w, dh16, B, Ho, Wo)
        # skip path gradient reaching the block input
        grads_skip = ()
        if skip_conv:
            dws = _wgrad(dy16, raw, B, Ho, Wo, 1, alpha, p[8])
            d_raw, d_raw_f16 = _dgrad(p[8], dy16, B, Ho, Wo, want_f16=False)
            add, add_mode, add_f16 = d_raw, 1, d_raw_f16
            grads_skip = (dws, db2)
        else:
            add, add_mode, add_f16 = d_out, {0: 1, 1: 2, 2: 3}[mode], False
        dx1, dx2, dg1, dbt1, _ = _gn_bwd(x1, False, st1, x2, st2, gn1, None, None, 0, d_a1, d_a1_f16, add, add_mode, add_f16, B, H, W,
                                         True, mode, alpha)
        return (None, None, None, dx1, None, dx2, None, d_rows, dg1, dbt1, dw1, db1, dg2, dbt2, dw2, db2) + grads_skip


class _AttnFn(torch.autograd.Function):
    """AttentionBlock (UNet.py:95-153) on an NHWC tensor: GN -> qkv 1x1 -> attention -> proj 1x1 -> + x.
    params: norm.w, norm.b, to_qkv.w, to_qkv.b, proj_out.w, proj_out.b"""

    @staticmethod
    def forward(ctx, cx, ab, x, st, *p):
        B, H, W, C = x.shape
        T, heads = H * W, ab.num_heads
        gn = ab.norm
        if st is None:
            st = _stats(x, B, T)
        xn, _ = _gn_apply(x, st, None, None, gn, B, H, W, False, 0)
        qkv, _, _ = _conv(xn, _pack_fwd([p[2]]), p[3], B, H, W, 3 * C, 1)
        a = torch.empty((B, H, W, C), device=x.device, dtype=torch.float16)
        L = _lib.lib()
        if L.anoddpm_attention_tc_supported(T, C, heads):
            ws = torch.empty(L.anoddpm_attention_tc_workspace_bytes(B, T, C), device=x.device, dtype=torch.uint8)
            call("anoddpm_attention_tc", ptr(a), ptr(qkv), ptr(ws), B, T, C, heads, 2, stream())
        else:
            call("anoddpm_attention", ptr(a), ptr(qkv), B, T, C, heads, 2, stream())
        out, st_out, _ = _conv(a, _pack_fwd([p[4]]), p[5], B, H, W, C, 1, res=x, res_mode=1, want_stats=True)
        if st_out is None:
            st_out = _stats(out, B, T)
        ctx.cx, ctx.ab = cx, ab
        ctx.save_for_backward(x, st, xn, qkv, a, *p)
        ctx.mark_non_differentiable(st_out)
        return out, st_out

    @staticmethod
    def backward(ctx, d_out, _d_st=None):
        x, st, xn, qkv, a = ctx.saved_tensors[:5]
        p = ctx.saved_tensors[5:]
        ab = ctx.ab
        B, H, W, C = x.shape
        T, heads = H * W, ab.num_heads
        ch = C // heads
        alpha = 1.0 / ctx.cx.scale
        dev = x.device
        d_out = d_out.contiguous()
        dy16, dbp = _cast_colsum(d_out, alpha)
        dwp = _wgrad(dy16, a, B, H, W, 1, alpha, p[4])
        d_a, _ = _dgrad(p[4], dy16, B, H, W, want_f16=False)                    # fp32 [B][T][C]
        # attention core backward per (sample, head): P = softmax(s2 Q K^T); dV = P^T dO; dP = dO V^T;
        # dL = P o (dP - rowsum(dP o P)); dQ = s2 dL K; dK = s2 dL^T Q            (UNet.py:137-153)
        s2 = 1.0 / math.sqrt(ch)
        C3 = 3 * C
        P = torch.empty((B, heads, T, T), device=dev, dtype=torch.float32)
        dP = torch.empty((B, heads, T, T), device=dev, dtype=torch.float32)
        dqkv = torch.empty((B, T, C3), device=dev, dtype=torch.float32)
        q, k, v = _View(qkv, 0), _View(qkv, ch), _View(qkv, 2 * ch)
        dq, dk, dv = _View(dqkv, 0), _View(dqkv, ch), _View(dqkv, 2 * ch)
        sq, sp, so = (T * C3, 3 * ch), (heads * T * T, T * T), (T * C, ch)       # batch strides (sample, head)
        _gemm(P, q, k, T, T, ch, C3, 1, 1, C3, T, s2, n1=B, n2=heads, sa=sq, sb=sq, sc=sp)           # S = s2 Q K^T
        call("anoddpm_softmax_rows", ptr(P), B * heads * T, T, stream())
        _gemm(dv, P, d_a, T, ch, T, 1, T, C, 1, C3, 1.0, n1=B, n2=heads, sa=sp, sb=so, sc=sq)         # dV = P^T dO
        _gemm(dP, d_a, v, T, T, ch, C, 1, 1, C3, T, 1.0, n1=B, n2=heads, sa=so, sb=sq, sc=sp)         # dP = dO V^T
        call("anoddpm_softmax_bwd_rows", ptr(dP), ptr(P), B * heads * T, T, stream())
        _gemm(dq, dP, k, T, ch, T, T, 1, C3, 1, C3, s2, n1=B, n2=heads, sa=sp, sb=sq, sc=sq)          # dQ = s2 dL K
        _gemm(dk, dP, q, T, ch, T, 1, T, C3, 1, C3, s2, n1=B, n2=heads, sa=sp, sb=sq, sc=sq)          # dK = s2 dL^T Q
        dq16, dbq = _cast_colsum(dqkv, alpha) if C3 <= 1024 else (None, None)
        if dq16 is None:       # 3C > 1024 channels: column sums in three slices of C
            dq16 = torch.empty(dqkv.shape, device=dev, dtype=torch.float16)
            dbq = torch.zeros(C3, device=dev, dtype=torch.float32)
            flat = dqkv.view(B * T, C3)
            for i in range(3):
                sl = flat[:, i * C:(i + 1) * C].contiguous()
                y, cs = _cast_colsum(sl, alpha)
                dq16.view(B * T, C3)[:, i * C:(i + 1) * C] = y
                dbq[i * C:(i + 1) * C] = cs
        dwq = _wgrad(dq16.view(B, H, W, C3), xn, B, H, W, 1, alpha, p[2])
        d_xn, d_xn_f16 = _dgrad(p[2], dq16.view(B, H, W, C3), B, H, W, want_f16=False)
        dx, _, dg, db, _ = _gn_bwd(x, False, st, None, None, ab.norm, None, None, 0, d_xn, d_xn_f16, d_out, 1, False, B, H, W, False, 0,
                                   alpha)
        return None, None, dx, None, dg, db, dwq, dbq, dwp, dbp


# ---------------------------------------------------------------------------------------------------------- driver
def forward_train(model, x, time):
    """UNetModel.forward with gradients: every kernel native, torch.autograd as the tape between blocks."""
    from .UNet import AttentionBlock, ResBlock
    _lib.require_cuda(x, "UNetModel input")
    if model.dropout > 0:
        raise NotImplementedError("native training: dropout > 0 is not implemented")
    with torch.cuda.device(x.device):
        cx = _Ctx()
        B = x.shape[0]
        dev = x.device
        base = model.model_channels
        t = time.to(device=dev, dtype=torch.int64).contiguous()
        eng = model.engine()
        temb = torch.empty((B, base), device=dev, dtype=torch.float32)
        call("anoddpm_timestep_embedding", ptr(temb), ptr(t), ptr(eng._freq(base, dev)), B, base, stream())
        l1, l2 = model.time_embedding[1], model.time_embedding[3]
        emb = _TimeMLPFn.apply(cx, temb, l1.weight, l1.bias, l2.weight, l2.bias)
        res_blocks = [m for m in model.modules() if isinstance(m, ResBlock)]
        offs, off, wb = {}, 0, []
        for rb in res_blocks:
            lin = rb.embed_layers[1]
            offs[id(rb)] = off
            off += lin.weight.shape[0]
            wb += [lin.weight, lin.bias]
        rows = _EmbProjFn.apply(cx, emb, *wb)

        def run_seq(seq, h, skip=None):
            """h, skip: (tensor, statistics-or-None) pairs"""
            for layer in seq:
                if isinstance(layer, ResBlock):
                    ps = [layer.in_layers[0].weight, layer.in_layers[0].bias, layer.in_layers[2].weight, layer.in_layers[2].bias,
                          layer.out_layers[0].weight, layer.out_layers[0].bias, layer.out_layers[3].weight, layer.out_layers[3].bias]
                    if isinstance(layer.skip_connection, torch.nn.Conv2d):
                        if layer.skip_connection.kernel_size != (1, 1):
                            raise NotImplementedError("ResBlock(use_conv=True) skip has no native kernel")
                        ps += [layer.skip_connection.weight, layer.skip_connection.bias]
                    x2, st2 = skip if skip is not None else (None, None)
                    h = _ResFn.apply(cx, layer, offs[id(layer)], h[0], h[1], x2, st2, rows, *ps)
                    skip = None
                elif isinstance(layer, AttentionBlock):
                    assert skip is None
                    h = _AttnFn.apply(cx, layer, h[0], h[1], layer.norm.weight, layer.norm.bias, layer.to_qkv.weight,
                                      layer.to_qkv.bias, layer.proj_out.weight, layer.proj_out.bias)
                else:
                    raise NotImplementedError(f"{type(layer).__name__} has no native training kernels")
            return h

        stem = model.down[0][0]
        h = (_StemFn.apply(cx, x, stem.weight, stem.bias), None)
        skips = [h]
        for seq in list(model.down)[1:]:
            h = run_seq(seq, h)
            skips.append(h)
        h = run_seq(model.middle, h)
        for seq in model.up:
            h = run_seq(seq, h, skip=skips.pop())
        gn, conv = model.out[0], model.out[2]
        out = _HeadFn.apply(cx, h[0], h[1], gn.weight, gn.bias, conv.weight, conv.bias, gn)
        return _OutputGate.apply(cx, out)
