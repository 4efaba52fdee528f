"""UNet inference engine: walks a ``UNetModel`` module tree once, packs its weights for the kernels and
executes the forward pass as a sequence of C-ABI calls on the current CUDA stream.

Data layout in HBM: activations NHWC (pixel-major, channels contiguous = the K axis of the implicit GEMM, so a K-block of one
pixel is one 128-byte TMA row): the residual stream fp32 (+ a raw fp16 copy where a later conv reads it as an operand),
conv1 outputs fp16 (convolution sum without bias / embedding); conv weights packed [Cout][Ktot] K-major with Ktot ordered
tap-major / channel-minor followed by the 1x1 skip-connection channels; GroupNorm statistics fp64 (sum, sum of squares) per
(sample, CHANNEL) in 8 replicas, emitted by the epilogue of the kernel that produced the tensor.

Fusion map of one same-resolution ResBlock (reference UNet.py:202-217), 4 launches instead of ~20 library calls:
    gn_table(GN1 of x [, skip])  -> conv3x3 whose loader applies GroupNorm + SiLU to the raw fp16 boxes of x [and of the
       skip tensor: torch.cat folded into the K loop] (+bias, +timestep-embedding row, +statistics of h, fp16 output)
    gn_table(GN2 of h)           -> conv3x3 with the same operand transform (+bias, +1x1 skip conv on the raw fp16 copies as
       extra K-blocks | +identity residual, +statistics of the output, fp32 output + fp16 copy for the next block)
Resampling blocks, attention inputs and tensors without an fp16 copy still take a stand-alone GroupNorm pass
(gn_apply: +SiLU, +avg-pool / nearest-upsample, +fp16 rounding, +raw copy / pooled residual).
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import ConvDesc, call, ptr, stream

STAT_REPLICAS = 8     # ANODDPM_STAT_REPLICAS of include/anoddpm_b200.h
_MODE = {"same": 0, "down": 1, "up": 2}
_RES_FOR_MODE = {"same": 1, "down": 2, "up": 3}


def round_tf32_(w):
    """In-place round-to-nearest (ties away) of fp32 values to TF32 precision, like cvt.rna.tf32.f32."""
    i = w.view(torch.int32)
    i.add_(0x1000).bitwise_and_(-8192)
    return w


def pack_conv_weight(w):
    """OIHW (or OI1 for Conv1d) -> [Cout][taps*Cin], tap-major / channel-minor."""
    if w.dim() == 3:
        w = w.unsqueeze(-1)
    co, ci, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(co, kh * kw * ci)


class _T:
    """An NHWC activation and (optionally) its per-channel statistics [B][C][2] fp64 (sum, sum of squares).
    shift = (bias, emb_rows): an fp16 conv output that stores the convolution sum only; the statistics describe
    t + bias[c] + emb_rows[b, c], and the GroupNorm pass adds that offset back in fp32."""
    __slots__ = ("t", "st", "shift", "t16")

    def __init__(self, t, st=None, shift=None, t16=None):
        self.t, self.st, self.shift = t, st, shift
        self.t16 = t16      # optional fp16 copy of an fp32 residual-stream tensor: the raw operand of a conv that applies
        #                     GroupNorm + SiLU while it loads (gn_table), and of a fused 1x1 skip conv


class _Conv:
    """One fused convolution site (weights packed lazily per implementation)."""

    def __init__(self, weights, biases, ksize):
        self.weights = weights      # list of parameter tensors: main conv first, then optional 1x1 skip conv
        self.biases = biases
        self.ksize = ksize
        self.cache = {}

    def packed(self, impl, f16=False):
        """impl 0: exact fp32; impl >= 1: TF32-rounded fp32, or IEEE fp16 (round to nearest) when f16."""
        key = "f16" if f16 else ("tf32" if impl >= 1 else "f32")
        if key not in self.cache:
            parts = [pack_conv_weight(w.detach()) for w in self.weights]
            w = torch.cat(parts, dim=1).contiguous().float()
            if key == "tf32":
                round_tf32_(w)
            elif key == "f16":
                w = w.clamp(-65504.0, 65504.0).half().contiguous()
            b = self.biases[0].detach().float()
            for extra in self.biases[1:]:
                b = b + extra.detach().float()
            self.cache[key] = (w, b.contiguous())
        return self.cache[key]


class UNetEngine:
    def __init__(self, model):
        self.model = model
        self._sig = None
        self._time_table = None
        self._ws, self._ws_keep = {}, []
        self.split_k, self.split_convs, self.sep_head = False, 0, True      # per-forward switches (set again by _forward)
        self._gen = 0               # bumped by invalidate(): raw-pointer / .data writes do not bump p._version
        self.launches = 0           # native kernel launches issued (for bench.py's gpu_launches)

    # ------------------------------------------------------------------ packing
    def invalidate(self):
        """Forget every derived weight copy (packed / rounded conv weights, concatenated embedding projections, head
        weights) and every CUDA graph keyed on them.  Called by the update paths that write parameters behind autograd's
        back (update_ema_params' .data writes, FusedAdamWEMA.step's raw-pointer kernels), which leave p._version
        unchanged."""
        self._gen += 1

    def _signature(self):
        return (self._gen,) + tuple((p.data_ptr(), p._version, p.device) for p in self.model.parameters())

    def _build_res(self, rb):
        if not isinstance(rb.skip_connection, (torch.nn.Identity, torch.nn.Conv2d)) or (
                isinstance(rb.skip_connection, torch.nn.Conv2d) and rb.skip_connection.kernel_size != (1, 1)):
            raise NotImplementedError("ResBlock(use_conv=True) skip has no native kernel")
        conv1, conv2 = rb.in_layers[2], rb.out_layers[3]
        cout = conv1.out_channels
        e = {"kind": "res", "mode": rb.resample_mode, "cin": conv1.in_channels, "cout": cout,
             "gn1": rb.in_layers[0], "gn2": rb.out_layers[0],
             "conv1": _Conv([conv1.weight], [conv1.bias], 3), "emb_off": self.emb_off}
        self.emb_w.append(rb.embed_layers[1].weight)
        self.emb_b.append(rb.embed_layers[1].bias)
        self.emb_off += cout
        if isinstance(rb.skip_connection, torch.nn.Conv2d):
            e["conv2"] = _Conv([conv2.weight, rb.skip_connection.weight], [conv2.bias, rb.skip_connection.bias], 3)
            e["skip_conv"] = True
        else:
            e["conv2"] = _Conv([conv2.weight], [conv2.bias], 3)
            e["skip_conv"] = False
        return e

    def _build(self):
        from .UNet import AttentionBlock, ResBlock
        m = self.model
        self.emb_w, self.emb_b, self.emb_off = [], [], 0
        self._time_table = None     # per-timestep embedding rows of the previous weight version

        build_res = self._build_res

        def build_attn(ab):
            return {"kind": "attn", "c": ab.in_channels, "heads": ab.num_heads, "gn": ab.norm,
                    "qkv": _Conv([ab.to_qkv.weight], [ab.to_qkv.bias], 1),
                    "proj": _Conv([ab.proj_out.weight], [ab.proj_out.bias], 1)}

        def build_entry(seq):
            out = []
            for layer in seq:
                if isinstance(layer, ResBlock):
                    out.append(build_res(layer))
                elif isinstance(layer, AttentionBlock):
                    out.append(build_attn(layer))
                elif isinstance(layer, torch.nn.Conv2d):
                    out.append({"kind": "stem", "conv": layer})
                else:
                    raise NotImplementedError(f"{type(layer).__name__} has no native kernel "
                                              "(biggan_updown=False resampling is not on the reference's path)")
            return out

        self.down = [build_entry(seq) for seq in m.down]
        self.middle = build_entry(m.middle)
        self.up = [build_entry(seq) for seq in m.up]
        self.head = {"gn": m.out[0], "conv": m.out[2]}
        self.emb_cat_w = torch.cat([w.detach() for w in self.emb_w], dim=0).contiguous().float()
        self.emb_cat_b = torch.cat([b.detach() for b in self.emb_b], dim=0).contiguous().float()
        self.head_w = pack_conv_weight(self.head["conv"].weight.detach()).contiguous().float()

    def _stats_pool_size(self, B):
        """doubles needed for the per-channel statistics of all conv / stem / attention outputs of one forward"""
        total = 0
        for entry in self.down + [self.middle] + self.up:
            for e in entry:
                if e["kind"] == "res":
                    total += 2 * e["cout"]
                elif e["kind"] == "attn":
                    total += e["c"]
                else:
                    total += e["conv"].out_channels
        return (total + 64) * B * 2 * STAT_REPLICAS

    def _refresh(self):
        sig = self._signature()
        if sig != self._sig:
            self._build()
            self._sig = sig

    def _freq(self, dim, dev):
        key = (dim, str(dev))
        if getattr(self, "_freq_cache", None) is None:
            self._freq_cache = {}
        if key not in self._freq_cache:
            import numpy as np
            half = dim // 2     # same expression as the reference (UNet.py:52-54), evaluated on the host once
            self._freq_cache[key] = torch.exp(torch.arange(half) * -(np.log(10000) / half)).float().to(dev)
        return self._freq_cache[key]

    # ------------------------------------------------------------------ op wrappers
    def _stats_slot(self, B, C):
        """[B][C][2] fp64 slice of the per-forward statistics pool (zeroed once per forward)."""
        n = B * C * 2 * STAT_REPLICAS
        if self.stats_used + n > self.stats_pool.numel():
            raise RuntimeError("statistics pool exhausted")
        sl = self.stats_pool[self.stats_used:self.stats_used + n]
        self.stats_used += n
        return sl

    def _stats_of(self, x, B, HW):
        """per-channel statistics of activation x (_T): taken from its producer's epilogue, else one reduction pass"""
        if x.st is None:
            C = x.t.shape[-1]
            x.st = self._stats_slot(B, C)
            call("anoddpm_gn_stats", ptr(x.st), ptr(x.t), C, B, HW, stream())
            self.launches += 1
            self.stats_passes += 1
        return x.st

    def _gn_apply(self, srcs, gn, B, H, W, silu, mode, out_mode, want_raw=False):
        """srcs: one or two _T (channel-concatenated).  out_mode: 0 fp32, 1 fp32 rounded to TF32, 2 fp16."""
        x1 = srcs[0]
        x2 = srcs[1] if len(srcs) > 1 else None
        if x1.shift is not None:
            assert x2 is None and mode == 0 and not want_raw and x1.st is not None and x1.t.dtype == torch.float16
            C = x1.t.shape[-1]
            bias, emb = x1.shift
            out = torch.empty((B, H, W, C), device=x1.t.device, dtype=torch.float16 if out_mode == 2 else torch.float32)
            call("anoddpm_gn_apply_f16in", ptr(out), ptr(x1.t), C, ptr(x1.st), ptr(gn.weight), ptr(gn.bias), ptr(bias),
                 ptr(emb), self.emb_all.shape[1] if emb is not None else 0, B, H, W, 32, float(gn.eps), int(silu),
                 int(out_mode), stream())
            self.launches += 1
            return out
        st1 = self._stats_of(x1, B, H * W)
        st2 = self._stats_of(x2, B, H * W) if x2 is not None else None
        C1 = x1.t.shape[-1]
        C2 = x2.t.shape[-1] if x2 is not None else 0
        C = C1 + C2
        Ho, Wo = (H // 2, W // 2) if mode == 1 else ((H * 2, W * 2) if mode == 2 else (H, W))
        odt = torch.float16 if out_mode == 2 else torch.float32
        out = torch.empty((B, Ho, Wo, C), device=x1.t.device, dtype=odt)
        raw = None
        if want_raw:      # mode 0: rounded concatenated copy of the input; mode 1: fp32 avg-pooled input
            raw = torch.empty((B, Ho, Wo, C), device=x1.t.device, dtype=odt if mode == 0 else torch.float32)
        call("anoddpm_gn_apply", ptr(out), ptr(raw), ptr(x1.t), C1, ptr(x2.t) if x2 is not None else None, C2,
             ptr(st1), ptr(st2), ptr(gn.weight), ptr(gn.bias), B, H, W, 32, float(gn.eps), int(silu), mode,
             int(out_mode), stream())
        self.launches += 1
        return (out, raw) if want_raw else out

    def _gn_table(self, srcs, gn, B, HW):
        """[B][C][2] (scale, shift) of GroupNorm over one tensor or the concatenation of two (anoddpm_gn_table): the conv
        that consumes the tensor(s) normalises + activates its operand boxes with it (no normalised copy in HBM)."""
        x1 = srcs[0]
        x2 = srcs[1] if len(srcs) > 1 else None
        st1 = self._stats_of(x1, B, HW)
        st2 = self._stats_of(x2, B, HW) if x2 is not None else None
        C1 = x1.t.shape[-1]
        C2 = x2.t.shape[-1] if x2 is not None else 0
        bias, emb = x1.shift if x1.shift is not None else (None, None)
        assert x1.shift is None or x2 is None
        tab = torch.empty((B, C1 + C2, 2), device=x1.t.device, dtype=torch.float32)
        call("anoddpm_gn_table", ptr(tab), ptr(st1), C1, ptr(st2), C2, ptr(gn.weight), ptr(gn.bias), ptr(bias), ptr(emb),
             self.emb_all.shape[1] if emb is not None else 0, B, HW, 32, float(gn.eps), stream())
        self.launches += 1
        self.gn_tables += 1
        return tab

    def _conv(self, site, a, B, H, W, cout, plan, r=(), emb=None, res=None, res_mode=0, want_stats=True, out_f16=False,
              a2=None, gn_table=None, want16=False):
        impl, f16 = plan
        d = ConvDesc()
        d.B, d.H, d.W = B, H, W
        d.Ca, d.Cout, d.ksize = a.shape[-1], cout, site.ksize
        if a2 is not None:
            d.a2, d.Ca2 = a2.data_ptr(), a2.shape[-1]
        if gn_table is not None:
            assert impl == 3 and f16
            d.gn_table = gn_table.data_ptr()
        out16 = None
        if want16 and impl == 3 and not out_f16:
            out16 = torch.empty((B, H, W, cout), device=a.device, dtype=torch.float16)
            d.out16 = out16.data_ptr()
        d.Cr1 = r[0].shape[-1] if len(r) > 0 else 0
        d.Cr2 = r[1].shape[-1] if len(r) > 1 else 0
        d.operand_f16 = 1 if f16 else 0
        assert a.dtype == (torch.float16 if f16 else torch.float32)
        w, b = site.packed(impl, f16)
        out = torch.empty((B, H, W, cout), device=a.device, dtype=torch.float16 if out_f16 else torch.float32)
        d.out_f16 = 1 if out_f16 else 0
        d.out, d.a, d.w, d.bias = out.data_ptr(), a.data_ptr(), w.data_ptr(), b.data_ptr()
        d.r1 = r[0].data_ptr() if len(r) > 0 else None
        d.r2 = r[1].data_ptr() if len(r) > 1 else None
        if emb is not None:
            d.emb, d.emb_stride = emb.data_ptr(), self.emb_all.shape[1]
        else:
            d.emb, d.emb_stride = None, 0
        d.res = res.data_ptr() if res is not None else None
        d.res_mode = res_mode
        d.impl = impl
        d.gn_stats = None
        st = None
        if want_stats and (impl == 3 or (impl == 2 and (H * W) % 32 == 0)):
            # the epilogue also accumulates the per-channel statistics of its output: no statistics pass for any
            # GroupNorm that later consumes this tensor (alone or concatenated with a skip tensor)
            st = self._stats_slot(B, cout)
            d.gn_stats = st.data_ptr()
            self.fused_stats += 1
        if impl == 2 and self.split_k:
            # small maps: split-K over a persistent, zero-initialised workspace (arrival counters reset themselves)
            need = _lib.lib().anoddpm_conv_workspace_bytes(ctypes.byref(d))
            if need > 0:
                ws = self._workspace(need, a.device)
                d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
                self.split_convs += 1
        call("anoddpm_conv", ctypes.byref(d), stream())
        self.launches += 1
        self.tc_convs += 1 if impl else 0
        self.f16_convs += 1 if f16 else 0
        if out_f16:
            assert st is not None and res is None
            self.f16_outputs += 1
            return _T(out, st, shift=(b, emb))
        return _T(out, st, t16=out16)

    def _workspace(self, nbytes, dev):
        """Persistent zeroed scratch for the split-K convolutions of this engine (grow-only; earlier buffers stay alive because
        captured CUDA graphs may still point at them)."""
        ws = self._ws.get(dev)
        if ws is None or ws.numel() < nbytes:
            ws = torch.zeros(int(nbytes), device=dev, dtype=torch.uint8)
            self._ws_keep.append(ws)
            self._ws[dev] = ws
        return ws

    def _plan(self, H, W, Ca, Cr, cout, ksize):
        """(impl, f16) for a conv with main source Ca channels and (single, concatenated) skip source Cr channels."""
        L = _lib.lib()
        if not self.use_tc or Ca < self.tc_min_channels:
            # short-K convs (fewer than tc_min_channels input channels) stay on the fp32 FFMA kernel: 10-bit-mantissa operand
            # rounding is averaged over only 9*Ca products there, which is what kept the base-32 models above 1e-3
            return 0, False
        d = ConvDesc()
        d.B, d.H, d.W, d.Ca, d.Cr1, d.Cr2, d.Cout, d.ksize = 1, H, W, Ca, Cr, 0, cout, ksize
        f16 = self.allow_f16 and Ca % 64 == 0 and Cr % 64 == 0
        d.operand_f16 = 1 if f16 else 0
        if self.tc_impl == 3 and L.anoddpm_conv_tc3_supported(ctypes.byref(d)):
            return 3, f16
        if L.anoddpm_conv_tc_supported(ctypes.byref(d)):
            impl = min(self.tc_impl, 2)
            return impl, (f16 and impl == 2)
        return 0, False

    @staticmethod
    def _out_mode(plan):
        impl, f16 = plan
        return 2 if f16 else (1 if impl else 0)

    # ------------------------------------------------------------------ blocks
    def _res(self, e, srcs, B, H, W, need16=False):
        """srcs: one or two _T (the up path concatenates h with a skip tensor).  need16: a later convolution will read the
        raw fp16 copy of this block's output (operand-side GroupNorm or fused 1x1 skip conv)."""
        mode = _MODE[e["mode"]]
        Ho, Wo = (H // 2, W // 2) if mode == 1 else ((H * 2, W * 2) if mode == 2 else (H, W))
        c_src = sum(x.t.shape[-1] for x in srcs)
        plan1 = self._plan(Ho, Wo, e["cin"], 0, e["cout"], 3)
        plan2 = self._plan(Ho, Wo, e["cout"], c_src if e["skip_conv"] else 0, e["cout"], 3)
        emb = self.emb_all[:, e["emb_off"]:]
        have16 = all(x.t16 is not None for x in srcs)
        fuse1 = self.fuse_gn and plan1 == (3, True) and mode == 0 and have16
        # tensor-core path: the fused 1x1 skip conv reads a rounded (TF32 / fp16), channel-concatenated copy of x: the fp16
        # copies written by the producers' epilogues, else a copy written by the pass that normalises x
        raw = None
        pooled = None
        if fuse1:
            # GroupNorm 1 + SiLU happen inside conv1: its loader rewrites the raw fp16 boxes of x (and of the skip tensor)
            tab1 = self._gn_table(srcs, e["gn1"], B, H * W)
            h = self._conv(e["conv1"], srcs[0].t16, B, Ho, Wo, e["cout"], plan1, emb=emb, gn_table=tab1,
                           a2=srcs[1].t16 if len(srcs) > 1 else None, out_f16=self.h1_f16)
            if e["skip_conv"] and plan2 == (3, True):
                raw = [x.t16 for x in srcs]
        else:
            if e["skip_conv"] and plan2[0] and self._out_mode(plan1) == self._out_mode(plan2):
                a1, raw = self._gn_apply(srcs, e["gn1"], B, H, W, True, mode, self._out_mode(plan1), want_raw=True)
                raw = [raw]
            elif mode == 1 and not e["skip_conv"] and plan2[0] == 3:
                # the halo-reuse kernel takes a same-resolution residual: the pass that normalises x also avg-pools it
                a1, pooled = self._gn_apply(srcs, e["gn1"], B, H, W, True, mode, self._out_mode(plan1), want_raw=True)
            else:
                a1 = self._gn_apply(srcs, e["gn1"], B, H, W, True, mode, self._out_mode(plan1))
            # conv1's output feeds GroupNorm 2 only: stored as fp16 without its per-channel offset where the fast kernel runs
            h = self._conv(e["conv1"], a1, B, Ho, Wo, e["cout"], plan1, emb=emb,
                           out_f16=self.h1_f16 and plan1[0] == 3 and plan1[1])
            del a1
        if e["skip_conv"] and raw is None and plan2[1]:
            plan2 = (plan2[0], False)      # no fp16 raw copy available: run conv2 with TF32 operands
        # GroupNorm 2 inside conv2: in kernel-microseconds about equal to the stand-alone fp16 -> fp16 pass + plain conv2 (conv2's
        # residual / statistics epilogue already loads the SM next to the MMAs), and the step time under the board's power cap
        # is the same within noise (26.4-26.5 ms both ways, A/B on one box); kept on: 4 B/element less HBM traffic
        fuse2 = self.fuse_gn and self.fuse_gn2 and plan2 == (3, True) and h.shift is not None
        if fuse2:
            # GroupNorm 2 + SiLU inside conv2 (the table carries the bias + embedding offset the fp16 h leaves out)
            tab2 = self._gn_table([h], e["gn2"], B, Ho * Wo)
            a2, kw = h.t, {"gn_table": tab2}
        else:
            a2, kw = self._gn_apply([h], e["gn2"], B, Ho, Wo, True, 0, self._out_mode(plan2)), {}
            del h
        kw["want16"] = need16 and self.fuse_gn
        if e["skip_conv"]:
            rs = raw if raw is not None else [x.t for x in srcs]
            return self._conv(e["conv2"], a2, B, Ho, Wo, e["cout"], plan2, r=rs, **kw), Ho, Wo
        assert len(srcs) == 1
        if pooled is not None:
            return self._conv(e["conv2"], a2, B, Ho, Wo, e["cout"], plan2, res=pooled, res_mode=1, **kw), Ho, Wo
        return self._conv(e["conv2"], a2, B, Ho, Wo, e["cout"], plan2, res=srcs[0].t,
                          res_mode=_RES_FOR_MODE[e["mode"]], **kw), Ho, Wo

    def _attn(self, e, x, B, H, W):
        C, T = e["c"], H * W
        plan_qkv = self._plan(H, W, C, 0, 3 * C, 1)
        plan_proj = self._plan(H, W, C, 0, C, 1)
        xn = self._gn_apply([x], e["gn"], B, H, W, False, 0, self._out_mode(plan_qkv))
        qkv = self._conv(e["qkv"], xn, B, H, W, 3 * C, plan_qkv, want_stats=False).t
        del xn
        om = self._out_mode(plan_proj)
        a = torch.empty((B, H, W, C), device=x.t.device, dtype=torch.float16 if om == 2 else torch.float32)
        L = _lib.lib()
        if self.use_tc and L.anoddpm_attention_tc_supported(T, C, e["heads"]):
            # tcgen05 attention: fp16 pack pass + one CTA per (128 queries, sample x head)
            ws = torch.empty(L.anoddpm_attention_tc_workspace_bytes(B, T, C), device=x.t.device, dtype=torch.uint8)
            call("anoddpm_attention_tc", ptr(a), ptr(qkv), ptr(ws), B, T, C, e["heads"], om, stream())
            self.launches += 2
            self.tc_attn += 1
        else:
            call("anoddpm_attention", ptr(a), ptr(qkv), B, T, C, e["heads"], om, stream())
            self.launches += 1
        return self._conv(e["proj"], a, B, H, W, C, plan_proj, res=x.t, res_mode=1)

    def _entry(self, entry, srcs, B, H, W, need16=False):
        """need16: the entry's OUTPUT will be read as a raw fp16 operand (skip tensor of the up path / next entry's conv1)."""
        h = None
        for i, e in enumerate(entry):
            cur = srcs if h is None else [h]
            if e["kind"] == "stem":
                raise RuntimeError("stem handled by forward()")
            if e["kind"] == "res":
                nxt = entry[i + 1] if i + 1 < len(entry) else None
                want = need16 if nxt is None else (nxt["kind"] == "res" and nxt["mode"] == "same")
                h, H, W = self._res(e, cur, B, H, W, need16=want)
            else:
                assert len(cur) == 1
                h = self._attn(e, cur[0], B, H, W)
        return h, H, W

    # ------------------------------------------------------------------ stand-alone ResBlock chain (tests, tools)
    @torch.no_grad()
    def run_resblocks(self, blocks, x, time_embed):
        """ResBlock modules applied in sequence to an NCHW tensor with the (B, E) time embedding (before the blocks'
        own SiLU -> Linear), on exactly the routines UNetModel.forward uses.  The caller sets use_tc / tc_impl /
        allow_f16 / h1_f16."""
        self.emb_w, self.emb_b, self.emb_off = [], [], 0
        entries = [self._build_res(rb) for rb in blocks]
        w = torch.cat([t.detach() for t in self.emb_w], dim=0).contiguous().float()
        b = torch.cat([t.detach() for t in self.emb_b], dim=0).contiguous().float()
        B, _, H, W = x.shape
        E, S = time_embed.shape[1], w.shape[0]
        te = time_embed.contiguous().float()
        with torch.cuda.device(x.device):
            self.emb_all = torch.empty((B, S), device=x.device, dtype=torch.float32)
            call("anoddpm_linear", ptr(self.emb_all), S, ptr(te), E, ptr(w), ptr(b), B, S, E, 1, stream())
            self.stats_pool = torch.zeros((sum(2 * e["cout"] for e in entries) + max(e["cin"] for e in entries) + 64)
                                          * B * 2 * STAT_REPLICAS, device=x.device, dtype=torch.float64)
            self.stats_used = 0
            h = _T(x.permute(0, 2, 3, 1).contiguous().float())
            for i, e in enumerate(entries):
                h, H, W = self._res(e, [h], B, H, W, need16=i + 1 < len(entries))
        return h.t.permute(0, 3, 1, 2).contiguous()

    # ------------------------------------------------------------------ timestep embedding
    def _embedding_rows(self, t, dev):
        """(len(t), S) rows: timestep embedding MLP (UNet.py:396-397) + every ResBlock's SiLU -> Linear projection
        (UNet.py:206-207), four small launches."""
        m = self.model
        n = t.numel()
        base, E = m.model_channels, m.model_channels * 4
        temb = torch.empty((n, base), device=dev, dtype=torch.float32)
        call("anoddpm_timestep_embedding", ptr(temb), ptr(t), ptr(self._freq(base, dev)), n, base, stream())
        l1, l2 = m.time_embedding[1], m.time_embedding[3]
        e1 = torch.empty((n, E), device=dev, dtype=torch.float32)
        call("anoddpm_linear", ptr(e1), E, ptr(temb), base, ptr(l1.weight), ptr(l1.bias), n, E, base, 0, stream())
        emb = torch.empty((n, E), device=dev, dtype=torch.float32)      # SiLU(time embedding): every ResBlock applies it
        call("anoddpm_linear", ptr(emb), E, ptr(e1), E, ptr(l2.weight), ptr(l2.bias), n, E, E, 3, stream())
        S = self.emb_cat_w.shape[0]
        rows = torch.empty((n, S), device=dev, dtype=torch.float32)
        call("anoddpm_linear", ptr(rows), S, ptr(emb), E, ptr(self.emb_cat_w), ptr(self.emb_cat_b), n, S, E, 0, stream())
        return rows

    @torch.no_grad()
    def time_table(self, T, dev):
        """Build (once per weight version) the [T][S] table of embedding rows for timesteps 0..T-1: everything the UNet
        computes from t alone.  The sampling loop calls this before it captures its step; forward() then gathers one row per
        sample (bit-identical: each row's summation order does not depend on the batch it is computed in).  Dropped whenever
        the weights change (_build)."""
        self._refresh()
        if self._time_table is None or self._time_table.shape[0] != T or self._time_table.device != dev:
            with torch.cuda.device(dev):
                t = torch.arange(T, device=dev, dtype=torch.int64)
                self._time_table = self._embedding_rows(t, dev)
        return self._time_table

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, x, t, time_table=False):
        """time_table: take the timestep-embedding rows from the table built by time_table() (the sampling loop, whose t is
        always inside [0, T)) instead of computing them for this batch."""
        _lib.require_cuda(x, "UNetModel input")
        with torch.cuda.device(x.device):       # kernels launch on the tensors' device, whichever device is current
            return self._forward(x, t, time_table)

    def _forward(self, x, t, time_table=False):
        m = self.model
        p0 = next(m.parameters())
        if p0.device != x.device:
            raise RuntimeError(f"UNetModel parameters are on {p0.device} but the input is on {x.device}")
        self._refresh()
        l_start = self.launches
        if m.conv_impl not in ("tc", "tc_tf32", "tc2", "tc1", "exact"):
            raise ValueError(f"conv_impl must be 'tc', 'tc_tf32', 'tc2', 'tc1' or 'exact', got {m.conv_impl!r}")
        self.use_tc = m.conv_impl != "exact"
        # tc / tc_tf32: halo-reuse CTA-pair kernel (impl 3) where it applies, else impl 2; tc2: persistent CTA pairs;
        # tc1: single-CTA tiles.  "tc" additionally feeds fp16 operands (kind::f16: TF32's 10-bit mantissa at twice
        # the rate and half the bytes) wherever every channel count is a multiple of 64; the others use kind::tf32.
        self.tc_impl = {"tc": 3, "tc_tf32": 3, "tc2": 2, "tc1": 1, "exact": 0}[m.conv_impl]
        self.allow_f16 = m.conv_impl == "tc"
        self.tc_min_channels = int(getattr(m, "tc_min_channels", 0)) if m.conv_impl == "tc" else 0
        self.h1_f16 = self.allow_f16 and getattr(m, "h1_fp16", True)
        # GroupNorm + SiLU applied to the operand boxes inside the 3x3 convs (no normalised copies in HBM)
        self.fuse_gn = self.h1_f16 and getattr(m, "fuse_gn", True)
        self.fuse_gn2 = getattr(m, "fuse_gn2", True)
        self.sep_head = getattr(m, "sep_head", True) and os.environ.get("ANODDPM_SEP_HEAD", "1") != "0"
        # split-K of small-map convs: measured SLOWER than narrow tiles at bs 32 (the last arriver's reduction is a serial tail);
        # conv_tc2 picks a narrower pair tile instead (tc2_pick_bn).  ANODDPM_SPLIT_K=1 enables it (tiny batches)
        self.split_k = os.environ.get("ANODDPM_SPLIT_K", "0") == "1"
        self.split_convs = 0
        self.gn_tables = 0
        self.f16_outputs = 0
        self.f16_convs = 0
        self.fused_stats = 0
        self.stats_passes = 0
        self.tc_convs = 0
        self.tc_attn = 0
        in_dtype = x.dtype
        x = x.contiguous().float()
        B, Cin, H, W = x.shape
        if Cin != m.in_channels:
            raise RuntimeError(f"expected {m.in_channels} input channels, got {Cin}")
        t = t.to(device=x.device, dtype=torch.int64).contiguous()
        if t.numel() != B:
            raise RuntimeError("time must have one entry per sample")
        dev = x.device
        S = self.emb_cat_w.shape[0]
        ttab = self._time_table if time_table else None
        if ttab is not None and ttab.device == dev and S % 4 == 0:
            # sampling loop: the rows were computed once for every timestep (time_table()); one gather per step
            self.emb_all = torch.empty((B, S), device=dev, dtype=torch.float32)
            call("anoddpm_gather_rows", ptr(self.emb_all), ptr(ttab), ptr(t), B, S, ttab.shape[0], stream())
            self.launches += 1
        else:
            self.emb_all = self._embedding_rows(t, dev)
            self.launches += 4
        # per-channel statistics of every activation that feeds a GroupNorm, bump-allocated, one memset per forward
        self.stats_pool = torch.zeros(self._stats_pool_size(B), device=dev, dtype=torch.float64)
        self.stats_used = 0

        stem = self.down[0][0]["conv"]
        C0 = stem.out_channels
        h = torch.empty((B, H, W, C0), device=dev, dtype=torch.float32)
        if Cin == 1 and W % 8 == 0 and 256 % max(C0 // 4, 1) == 0 and C0 % 4 == 0 and self.use_tc:
            # one pass: features + their GroupNorm statistics (+ the fp16 copy the first conv's operand transform reads)
            st0 = self._stats_slot(B, C0)
            h16 = torch.empty((B, H, W, C0), device=dev, dtype=torch.float16) if self.fuse_gn else None
            call("anoddpm_conv_stem_fused", ptr(h), ptr(h16), ptr(st0), ptr(x), ptr(stem.weight), ptr(stem.bias), B, Cin, H, W, C0,
                 stream())
            h = _T(h, st0, t16=h16)
            self.fused_stats += 1
        else:
            call("anoddpm_conv_stem", ptr(h), ptr(x), ptr(stem.weight), ptr(stem.bias), B, Cin, H, W, C0, stream())
            h = _T(h)
        self.launches += 1
        skips = [(h, H, W)]
        for entry in self.down[1:]:
            h, H, W = self._entry(entry, [h], B, H, W, need16=True)      # every down output is a skip tensor of the up path
            skips.append((h, H, W))
        h, H, W = self._entry(self.middle, [h], B, H, W, need16=True)
        for i, entry in enumerate(self.up):
            sk, sh, sw = skips.pop()
            assert (sh, sw) == (H, W)
            h, H, W = self._entry(entry, [h, sk], B, H, W, need16=i + 1 < len(self.up))
        hd = self.head
        out = torch.empty((B, m.out_channels, H, W), device=dev, dtype=torch.float32)
        C = h.t.shape[-1]
        if (self.sep_head and _lib.lib().anoddpm_conv_head_sep_supported(C, m.out_channels) and (H * W) % 64 == 0
                and hd["gn"].num_groups == 32):
            # single-output head, separable by tap: one halo-free streaming pass (GroupNorm + SiLU + nine channel reductions per
            # pixel) into nine L2-resident maps, then the nine shifted adds
            tab = self._gn_table([h], hd["gn"], B, H * W)
            ws = torch.empty(9 * B * H * W, device=dev, dtype=torch.float32)
            call("anoddpm_conv_head_sep", ptr(out), ptr(h.t), ptr(tab), ptr(self.head_w), ptr(hd["conv"].bias), ptr(ws), B, C, H, W,
                 stream())
            self.launches += 1      # two kernels (projection + gather); the second one is counted by the common += 1 below
        elif C <= 512 and m.out_channels <= 4:
            # GroupNorm + SiLU applied while the head conv stages its tiles: the last feature map is read once
            gn = hd["gn"]
            call("anoddpm_conv_head_gn", ptr(out), ptr(h.t), ptr(self._stats_of(h, B, H * W)), ptr(gn.weight), ptr(gn.bias),
                 ptr(self.head_w), ptr(hd["conv"].bias), B, C, H, W, m.out_channels, 32, float(gn.eps), stream())
        else:
            a = self._gn_apply([h], hd["gn"], B, H, W, True, 0, 0)
            call("anoddpm_conv_head", ptr(out), ptr(a), ptr(self.head_w), ptr(hd["conv"].bias), B, C, H, W,
                 m.out_channels, stream())
        self.launches += 1
        self.emb_all = None
        self.stats_pool = None
        self.launches_per_forward = self.launches - l_start
        return out.to(in_dtype)
