#!/usr/bin/env python
"""bench.py -- detection throughput of the AnoDDPM hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2], per GPU): GaussianDiffusionModel.forward_backward on a synthetic
(32,1,256,256) batch, T=1000 linear schedule, t_distance=800, simplex noise (6 octaves, 0.8, 64), UNet
256/base 128/attn "16,8"/2 heads with random-init weights (reference zero-inits re-drawn).
A STEP is one reverse-diffusion step of that chain -- `sample_p` = UNet forward + batched simplex noise +
fused posterior kernel -- over the whole batch; the K timed steps are K consecutive iterations of a real
chain (t = 799, 798, ...).  One image costs t_distance = 800 steps, so
    images/s = n_gpus * B / (800 * seconds_per_step).
N > 1: 32 images per GPU (weak scaling, configs[3] = 256 images over 8 GPUs), no data-path collective.
`--full-chain` times whole 800-step chains instead (one chain = one step) to confirm the same number.

Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
Working set per step (activations, GBs) is far larger than the 126 MB L2, so no explicit L2 flush.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

T_DISTANCE = 800
UNET_KW = dict(img_size=256, base_channels=128, n_heads=2, attention_resolutions="16,8")
GFLOP_PER_IMAGE_FWD = 554.5          # BASELINE.md section 3 (277.26 GMAC)
METRIC = "detection images/sec (256x256, T=1000 partial reverse t_distance=800, simplex)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.rows = []
        self.mark_at = 0

    def mark(self):
        """the timed region starts here (start() is called before the warm-up so that nvidia-smi is already streaming)"""
        self.mark_at = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        rows, note = self.rows[self.mark_at:], None
        if not rows and self.rows:      # timed region shorter than the sampling interval: last warm-up samples (same load)
            rows, note = self.rows[-2:], "timed region shorter than the 100 ms sampling interval: last warm-up samples (same load)"
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm)}
        if note:
            out["note"] = note
        return out


def build_unet(device):
    from anoddpm_b200 import UNetModel
    from anoddpm_b200.synth import fill_state_dict   # deterministic key-addressed synthetic weights
    m = UNetModel(**UNET_KW).eval()
    m.load_state_dict(fill_state_dict(m.state_dict(), seed=1))
    return m.to(device)


# ------------------------------------------------------------------------------------------------ native arm
def run_native(args):
    from anoddpm_b200 import GaussianDiffusionModel, dist as D, get_beta_schedule
    from anoddpm_b200.GaussianDiffusion import simplex_noise_from_seeds
    rank, world, local = D.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (sm_100a); there is no CPU fallback for the native arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B = args.batch
    model = build_unet(dev)
    if args.conv_impl:
        model.conv_impl = args.conv_impl
    diff = GaussianDiffusionModel([256, 256], get_beta_schedule(1000, "linear"), noise="simplex")
    lo = rank * B
    gidx = range(lo, lo + B)
    launches = {"n": 0}
    step_counter = {"s": 0}

    def noise_fn(x, t):   # batched simplex noise, per-image seeds by GLOBAL index (world-size invariant)
        seeds = D.seeds_for_step(args.seed, step_counter["s"], gidx)
        step_counter["s"] += 1
        launches["n"] += 1
        return simplex_noise_from_seeds(seeds, tuple(x.shape), t)

    # the reverse loop draws its simplex seeds through this hook (same per-image seeds whether the step runs eagerly or
    # as a replayed CUDA graph); the forward jump (sample_q) uses noise_fn above
    diff.simplex_seed_source = lambda step, n: D.seeds_for_step(args.seed, 100000 + step, gidx)
    if args.no_graphs:
        diff.use_cuda_graphs = False
    g = torch.Generator().manual_seed(1000 + rank)
    x_host = (torch.rand(B, 1, 256, 256, generator=g) * 2 - 1).pin_memory()
    eng = model.engine()

    @torch.no_grad()
    def reverse_steps(x, t_hi, n):   # n consecutive iterations of forward_backward's reverse loop (public API)
        x = diff.reverse_steps(model, x, t_hi, n, "simplex")
        launches["n"] += 2 * n      # simplex noise + fused posterior per step (UNet launches are counted by the engine)
        return x

    def chain(x_h, td):   # the public API call, host buffers in, host buffer out
        x_d = x_h.to(dev, non_blocking=True)
        t_start = torch.full((B,), td - 1, device=dev, dtype=torch.int64)
        x_d = diff.sample_q(x_d, t_start, noise_fn(x_d, t_start))
        y = diff.reverse_steps(model, x_d, td - 1, td, "simplex")     # = forward_backward(model, x, None, td, "simplex")
        return y.cpu()

    clocks = ClockSampler(local)
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host_enqueue_ms = None
    if args.full_chain:
        for _ in range(args.warmup):
            chain(x_host, T_DISTANCE)
        torch.cuda.synchronize(); D.barrier()
        clocks.mark(); l0, e0 = launches["n"], eng.launches
        ev0.record()
        for _ in range(args.steps):
            chain(x_host, T_DISTANCE)
        ev1.record(); torch.cuda.synchronize(); D.barrier()
        ms = ev0.elapsed_time(ev1)
        steps_per_image = 1.0
    else:
        # device-timed steps and the end-to-end chain are measured ALTERNATELY (two rounds each, same power / thermal state);
        # `value` and `e2e` are the means of their rounds
        t_start = torch.full((B,), T_DISTANCE - 1, device=dev, dtype=torch.int64)
        x = diff.sample_q(x_host.to(dev), t_start, noise_fn(x_host.to(dev), t_start))
        launches["n"] += 1
        x = reverse_steps(x, T_DISTANCE - 1, args.warmup)
        chain(x_host, 3)
        torch.cuda.synchronize(); D.barrier()
        clocks.mark()
        k_e2e = max(3, min(args.steps, 25))
        ms_rounds, e2e_rounds, t_hi = [], [], T_DISTANCE - 1 - args.warmup
        l0 = e0 = None
        for rnd in range(2):
            torch.cuda.synchronize(); D.barrier()
            if rnd == 0:
                l0, e0 = launches["n"], eng.launches
            ev0.record()
            h0 = time.perf_counter()
            x = reverse_steps(x, t_hi, args.steps)
            if rnd == 0:
                host_enqueue_ms = (time.perf_counter() - h0) * 1e3 / args.steps   # host time to ENQUEUE one step (no sync)
            ev1.record(); torch.cuda.synchronize(); D.barrier()
            if rnd == 0:
                l1, e1n = launches["n"], eng.launches
            ms_rounds.append(ev0.elapsed_time(ev1))
            t_hi -= args.steps
            t0 = time.perf_counter()
            ev0.record()
            y = chain(x_host, k_e2e)
            ev1.record(); torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) * 1e3
            e2e_rounds.append(max(ev0.elapsed_time(ev1), wall) / k_e2e)   # .cpu() completes after the last event: slower clock
        ms = float(np.mean(ms_rounds))
        steps_per_image = float(T_DISTANCE)
    clk = clocks.stop()
    per_fwd = getattr(eng, "launches_per_forward", 0)
    fwd_timed = args.steps * (T_DISTANCE if args.full_chain else 1)
    # eager: the engine counted every launch; graph replay: launches per captured step x replays
    if args.full_chain:
        n_launch = (launches["n"] - l0) + ((eng.launches - e0) if args.no_graphs else per_fwd * fwd_timed)
    else:
        n_launch = (l1 - l0) + ((e1n - e0) if args.no_graphs else per_fwd * fwd_timed)
    ms = D.max_over_ranks(ms, dev)
    ms_per_step = ms / args.steps
    value = world * B / (steps_per_image * ms_per_step / 1e3)

    # ---- end-to-end through the public API with host buffers (H2D of the batch + D2H of the result inside the timed region)
    if args.full_chain:
        k_e2e = 3
        chain(x_host, 2)
        torch.cuda.synchronize(); D.barrier()
        t0 = time.perf_counter()
        ev0.record()
        y = chain(x_host, k_e2e)
        ev1.record(); torch.cuda.synchronize()
        e2e_rounds = [max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3) / k_e2e]
    e2e_ms = D.max_over_ranks(float(np.mean(e2e_rounds)), dev)
    img_bytes = B * 256 * 256 * 4
    e2e = {"value": world * B / (T_DISTANCE * e2e_ms / 1e3), "unit": "images/s",
           "h2d_bytes_per_step": int(img_bytes / k_e2e + B * 1024), "d2h_bytes_per_step": int(img_bytes / k_e2e),
           "api": f"x_pinned.to(cuda) -> GaussianDiffusionModel.sample_q + reverse_steps (= forward_backward(unet, x, None, "
                  f"t_distance={k_e2e}, 'simplex')) -> .cpu(); per-step time x 800 steps/image; measured alternately with the "
                  f"device-timed steps (rounds: {[round(v, 3) for v in e2e_rounds]} ms/step e2e, "
                  f"{[round(v / args.steps, 3) for v in (ms_rounds if not args.full_chain else [ms])]} ms/step device)",
           "finite": bool(torch.isfinite(y).all())}

    out = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": {"tc": "f16 operands (10-bit mantissa, = TF32 precision) / f32 accumulate", "exact": "f32"}.get(model.conv_impl, "tf32"), "data": "synthetic",
           "config": {"workload": "BASELINE configs[2] per GPU: forward_backward 256x256x1 bs=32 T=1000 t_distance=800 "
                                  "simplex(6,0.8,64), UNet 256/base128/attn16,8/2 heads (554.5 GFLOP/image/step); "
                                  + ("step = one full 800-step chain" if args.full_chain else
                                     "step = one reverse step (sample_p: UNet fwd + simplex noise + fused posterior) of the chain; "
                                     "images/s = n_gpus*B/(800*s_per_step)"),
                      "global_batch": world * B, "per_gpu_batch": B, "t_distance": T_DISTANCE,
                      "parallelism": f"dp{world} (images sharded by index, no collective in the loop)",
                      "l2": "working set per step >> 126 MB L2 (no explicit flush)", "conv_impl": model.conv_impl,
                      "cuda_graph": bool(diff.use_cuda_graphs)},
           "e2e": e2e, "secondary": {"t_distance": 250, "value": value * T_DISTANCE / 250.0, "unit": "images/s",
                                     "note": "the reference's own ROC sweep uses t_distance=250 (detection.py:503): same per-step time, 250 steps/image"},
           "gpu_launches": int(n_launch), "clocks": clk, "tc_convs_per_forward": eng.tc_convs,
           "host_enqueue_ms_per_step": host_enqueue_ms, "fused_gn_stats_per_forward": getattr(eng, "fused_stats", 0),
           "gn_stats_passes_per_forward": getattr(eng, "stats_passes", 0)}

    if rank == 0 and world == 1 and not args.no_default_model:
        # secondary: the UNet the reference's scripts actually construct when they do not override the attention settings
        # (UNet.py:232 attention_resolutions="32,16,8"; detection.py:168 n_heads=1): attention at 32x32 (T = 1024, one head of
        # 256 channels), 16x16 and 8x8 (one head of 512) -- all on the key-tiled tcgen05 attention kernel; 562.6 GFLOP/image
        from anoddpm_b200 import UNetModel
        from anoddpm_b200.synth import fill_state_dict
        x = None
        torch.cuda.empty_cache()
        m2 = UNetModel(img_size=256, base_channels=128).eval()
        m2.load_state_dict(fill_state_dict(m2.state_dict(), seed=1))
        m2 = m2.to(dev)
        xs = diff.sample_q(x_host.to(dev), torch.full((B,), T_DISTANCE - 1, device=dev, dtype=torch.int64),
                           noise_fn(x_host.to(dev), torch.full((B,), T_DISTANCE - 1, device=dev, dtype=torch.int64)))
        xs = diff.reverse_steps(m2, xs, T_DISTANCE - 1, 4, "simplex")
        torch.cuda.synchronize()
        ev0.record()
        xs = diff.reverse_steps(m2, xs, T_DISTANCE - 5, 8, "simplex")
        ev1.record(); torch.cuda.synchronize()
        ms2 = ev0.elapsed_time(ev1) / 8
        out["secondary_default_model"] = {"model": "UNetModel(256, 128): attention_resolutions='32,16,8', n_heads=1 (562.6 GFLOP/image)",
                                          "ms_per_step": ms2, "value": B / (T_DISTANCE * ms2 / 1e3), "unit": "images/s",
                                          "tc_attention_sites": m2.engine().tc_attn, "finite": bool(torch.isfinite(xs).all())}
        del m2, xs
    if rank == 0:
        x = None
        torch.cuda.empty_cache()
        out["roofline"] = roofline(model, dev, B)
        out["aux_rooflines"] = aux_rooflines(dev)
        out["cpu_baseline"] = cpu_baseline() if world == 1 and not args.no_cpu_baseline else None
        print(json.dumps(out))
    D.barrier()


def _time_ms(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel/shape in its PRODUCTION form (conv1 of a
# ResBlock, 128->128 @256^2 bs 32: operand-side GroupNorm + SiLU, +bias/embedding statistics, fp16 output) from the committed
# `ncu --set full` capture profiles/r02_ncu_conv_tc3_gn.md (algorithmic: 0.537 GB read + 0.537 GB written)
NCU_TRAFFIC_DOMINANT = {"bytes": 0.600738e9 + 0.498947e9, "source": "profiles/r02_ncu_conv_tc3_gn.md"}


def roofline(model, dev, B):
    """Dominant kernel = the tensor-core conv (conv_tc3_kernel).  Live CUDA-event timing, through the same C-ABI entry
    point the engine uses, of the three conv shapes that carry 61% of the UNet MACs, each in the
    THREE forms it takes: `plain` (no epilogue extras -- the favourable variant round 1 quoted), `conv1` = the production
    first conv of a ResBlock (operand-side GroupNorm + SiLU on raw fp16 boxes, + embedding row, statistics, fp16 output) and
    `conv2` = the production second conv (operand-side GroupNorm, + fp32 residual, statistics, fp32 output + fp16 copy).
    The headline `achieved` is the PRODUCTION conv1 form of the most frequent shape (128->128 @256^2)."""
    import ctypes
    from anoddpm_b200 import _lib
    pk = peaks()
    impl = {"tc": 3, "tc_tf32": 3, "tc2": 2, "tc1": 1}.get(model.conv_impl, 0)
    f16 = model.conv_impl == "tc"
    fused = f16 and impl == 3 and getattr(model, "fuse_gn", True)

    def run(Ci, Co, HW, form):
        a = torch.randn(B, HW, HW, Ci, device=dev)
        w = torch.randn(Co, 9 * Ci, device=dev) * 0.03
        if f16:
            a, w = a.half(), w.half()
        bias = torch.zeros(Co, device=dev)
        keep = [a, w, bias]
        d = _lib.ConvDesc()
        d.B, d.H, d.W, d.Ca, d.Cout, d.ksize, d.impl = B, HW, HW, Ci, Co, 3, impl
        d.operand_f16 = 1 if f16 else 0
        out_f16 = form == "conv1" and f16 and impl == 3
        out = torch.empty(B, HW, HW, Co, device=dev, dtype=torch.float16 if out_f16 else torch.float32)
        d.out, d.a, d.w, d.bias, d.out_f16 = out.data_ptr(), a.data_ptr(), w.data_ptr(), bias.data_ptr(), 1 if out_f16 else 0
        abytes = a.numel() * a.element_size() + out.numel() * out.element_size() + w.numel() * w.element_size()
        if form != "plain":
            st = torch.zeros(8 * B * Co * 2, device=dev, dtype=torch.float64)
            keep.append(st)
            d.gn_stats = st.data_ptr()
            if fused:
                tab = torch.stack([torch.rand(B, Ci, device=dev) + 0.5, torch.randn(B, Ci, device=dev) * 0.3], dim=-1).contiguous()
                keep.append(tab)
                d.gn_table = tab.data_ptr()
            if form == "conv1":
                emb = torch.randn(B, Co, device=dev)
                keep.append(emb)
                d.emb, d.emb_stride = emb.data_ptr(), Co
            else:
                res = torch.randn(B, HW, HW, Co, device=dev)
                keep.append(res)
                d.res, d.res_mode = res.data_ptr(), 1
                abytes += res.numel() * 4
                if fused:
                    o16 = torch.empty(B, HW, HW, Co, device=dev, dtype=torch.float16)
                    keep.append(o16)
                    d.out16 = o16.data_ptr()
                    abytes += o16.numel() * 2
        ms = _time_ms(lambda: _lib.call("anoddpm_conv", ctypes.byref(d), _lib.stream()))
        flop = 2.0 * B * HW * HW * Co * 9 * Ci
        return {"shape": f"conv3x3 {Ci}->{Co} @{HW}^2 bs{B}", "form": form, "ms": ms, "tflops": flop / ms / 1e9, "flop": flop,
                "algorithmic_bytes": abytes}

    res = []
    tot_flop, tot_ms = 0.0, 0.0
    for (Ci, Co, HW, count) in ((128, 128, 256, 9), (256, 128, 256, 3), (128, 128, 128, 9)):
        forms = ("plain", "conv1", "conv2") if impl == 3 else ("plain",)
        for form in forms:
            r = run(Ci, Co, HW, form)
            r["count_per_forward"] = count
            res.append(r)
            if form == forms[-2 if len(forms) > 1 else 0]:       # weight the production conv1 form
                tot_flop += r["flop"] * count
                tot_ms += r["ms"] * count
    dom = next(r for r in res if r["form"] == ("conv1" if impl == 3 else "plain"))
    kname = {3: "conv_tc3_kernel (tcgen05.mma.cta_group::2 kind::%s, single-box halo-reuse main loop, persistent CTA pairs, TMA-fed "
                "implicit GEMM%s, TMA-store epilogue)" % ("f16" if f16 else "tf32", ", transform warps apply GroupNorm + SiLU to the landed "
                                                          "operand boxes" if fused else ""),
             2: "conv_tc2_kernel (tcgen05.mma.cta_group::2, persistent CTA pairs)", 1: "conv_tc_kernel (tcgen05.mma kind::tf32)",
             0: "conv_simt_kernel"}[impl]
    peak = pk["bf16_burst"]      # the kernel is timed alone here -> burst figure (B200_PROFILING.md)
    plain = res[0]
    return {"bound": "tensor", "kernel": kname, "shape": dom["shape"], "form": dom["form"] + " (production epilogue"
            + (", operand-side GroupNorm)" if fused else ")"), "achieved": dom["tflops"], "peak": peak,
            "unit": "TFLOP/s", "frac": dom["tflops"] / peak,
            "peak_note": f"dense bf16 cuBLAS burst, {pk['src']} (MEASURED_PEAKS.json); sustained figure {pk['bf16_sustained']}: frac "
                         f"{dom['tflops'] / pk['bf16_sustained']:.3f}" + ("" if f16 else "; kind::tf32 issues at half the bf16 rate"),
            "plain_variant": {"achieved": plain["tflops"], "frac": plain["tflops"] / peak,
                              "note": "same shape with no epilogue extras and no operand transform (the variant BENCH_r01 quoted)"},
            "algorithmic": "2*M*N*K flop per launch, M=B*H*W, N=Cout, K=9*Cin", "flop_per_launch": dom["flop"],
            "algorithmic_bytes": dom["algorithmic_bytes"],
            "traffic": NCU_TRAFFIC_DOMINANT["bytes"] if (fused and B == 32) else None,
            "traffic_source": NCU_TRAFFIC_DOMINANT["source"],
            "weighted_3_shapes_tflops": tot_flop / tot_ms / 1e9, "per_shape": res}


def aux_rooflines(dev):
    """HBM-bound fused elementwise kernels at a size that exercises HBM (bs 256: 64 MiB per tensor, > 126 MB L2 in total),
    GroupNorm apply at full resolution, and the simplex generator (BASELINE config 5, FP64-ALU bound)."""
    from anoddpm_b200 import GaussianDiffusionModel, _lib, get_beta_schedule
    from anoddpm_b200.simplex import init_tables
    pk = peaks()
    out = {}
    D = GaussianDiffusionModel([256, 256], get_beta_schedule(1000, "linear"))
    Bq = 256
    x, n, e = (torch.randn(Bq, 1, 256, 256, device=dev) for _ in range(3))
    t = torch.randint(0, 1000, (Bq,), device=dev)
    tab = D._tables(dev)
    o = torch.empty_like(x)
    per = x.numel() // Bq
    ms = _time_ms(lambda: _lib.call("anoddpm_q_sample", _lib.ptr(o), _lib.ptr(x), _lib.ptr(n), _lib.ptr(t), _lib.ptr(tab[0]),
                                    _lib.ptr(tab[1]), Bq, per, _lib.stream()))
    gb = 12.0 * x.numel() / 1e9
    out["q_sample"] = {"bound": "hbm", "shape": "(256,1,256,256)", "bytes": gb * 1e9, "ms": ms, "achieved": gb / ms * 1e3,
                       "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gb / ms * 1e3 / pk["hbm_gbs"]}
    ms = _time_ms(lambda: _lib.call("anoddpm_p_sample", _lib.ptr(o), None, _lib.ptr(x), _lib.ptr(e), _lib.ptr(n), _lib.ptr(t),
                                    _lib.ptr(tab[4:]), 1000, Bq, per, _lib.stream()))
    gb = 16.0 * x.numel() / 1e9
    out["p_sample"] = {"bound": "hbm", "shape": "(256,1,256,256)", "bytes": gb * 1e9, "ms": ms, "achieved": gb / ms * 1e3,
                       "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gb / ms * 1e3 / pk["hbm_gbs"]}
    del x, n, e, o
    # GroupNorm apply, (32,256,256,128) fp32 -> fp16
    xa = torch.randn(32, 256, 256, 128, device=dev)
    st = torch.zeros(8, 32, 128, 2, device=dev, dtype=torch.float64)      # [ANODDPM_STAT_REPLICAS][B][C][2]
    _lib.call("anoddpm_gn_stats", _lib.ptr(st), _lib.ptr(xa), 128, 32, 65536, _lib.stream())
    g1, b1 = torch.ones(128, device=dev), torch.zeros(128, device=dev)
    oa = torch.empty(32, 256, 256, 128, device=dev, dtype=torch.float16)
    ms = _time_ms(lambda: _lib.call("anoddpm_gn_apply", _lib.ptr(oa), None, _lib.ptr(xa), 128, None, 0, _lib.ptr(st), None,
                                    _lib.ptr(g1), _lib.ptr(b1), 32, 256, 256, 32, 1e-5, 1, 0, 2, _lib.stream()))
    gb = 6.0 * xa.numel() / 1e9
    out["gn_apply"] = {"bound": "hbm", "shape": "(32,256,256,128) fp32 -> fp16", "bytes": gb * 1e9, "ms": ms,
                       "achieved": gb / ms * 1e3, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gb / ms * 1e3 / pk["hbm_gbs"]}
    # the same pass on an fp16 conv1 output (4 B/elem)
    xh = xa.half()
    shc = torch.zeros(128, device=dev)
    ms = _time_ms(lambda: _lib.call("anoddpm_gn_apply_f16in", _lib.ptr(oa), _lib.ptr(xh), 128, _lib.ptr(st), _lib.ptr(g1), _lib.ptr(b1),
                                    _lib.ptr(shc), None, 0, 32, 256, 256, 32, 1e-5, 1, 2, _lib.stream()))
    gb = 4.0 * xa.numel() / 1e9
    out["gn_apply_f16in"] = {"bound": "hbm", "shape": "(32,256,256,128) fp16 -> fp16", "bytes": gb * 1e9, "ms": ms,
                             "achieved": gb / ms * 1e3, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gb / ms * 1e3 / pk["hbm_gbs"]}
    del xa, oa, xh
    # tcgen05 attention at the BASELINE shape (T=256, C=512, 2 heads): tiny (4.3 GFLOP), latency- not pipe-bound
    L = _lib.lib()
    qkv = torch.randn(32, 256, 3 * 512, device=dev)
    ao = torch.empty(32, 256, 512, device=dev, dtype=torch.float16)
    ws = torch.empty(L.anoddpm_attention_tc_workspace_bytes(32, 256, 512), device=dev, dtype=torch.uint8)
    ms = _time_ms(lambda: _lib.call("anoddpm_attention_tc", _lib.ptr(ao), _lib.ptr(qkv), _lib.ptr(ws), 32, 256, 512, 2, 2, _lib.stream()))
    fl = 4.0 * 32 * 2 * 256 * 256 * 256
    out["attention_tc"] = {"bound": "tensor (latency-bound at this size)", "shape": "B32 T256 C512 heads2 (pack + attention kernel)",
                           "ms": ms, "flop": fl, "achieved": fl / ms / 1e9, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                           "frac": fl / ms / 1e9 / pk["bf16_sustained"],
                           "bytes": 32 * 256 * 512 * (12 + 2) , "gbs": 32 * 256 * 512 * 14 / ms / 1e6}
    del qkv, ao, ws
    # fused clip + AdamW + EMA step (training row T1): 4 B/param (norm pass) + 20 B read + 16 B written per parameter
    npar = 131_650_000
    pp, gg, mm, vv, ee = (torch.randn(npar, device=dev) * 0.01 for _ in range(5))
    vv.abs_()
    ssq = torch.zeros(1, device=dev, dtype=torch.float64)

    def opt_step():
        ssq.zero_()
        _lib.call("anoddpm_sumsq", _lib.ptr(ssq), _lib.ptr(gg), npar, _lib.stream())
        _lib.call("anoddpm_adamw_ema_step", _lib.ptr(pp), _lib.ptr(gg), _lib.ptr(mm), _lib.ptr(vv), _lib.ptr(ee), _lib.ptr(ssq), npar,
                  1e-4, 0.9, 0.999, 1e-8, 0.0, 10, 1.0, 0.9999, _lib.stream())
    ms = _time_ms(opt_step)
    gb = 40.0 * npar / 1e9
    out["adamw_ema_step"] = {"bound": "hbm", "shape": "131.65 M parameters (clip + AdamW + EMA, 2 kernels)", "bytes": gb * 1e9, "ms": ms,
                             "achieved": gb / ms * 1e3, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gb / ms * 1e3 / pk["hbm_gbs"]}
    del pp, gg, mm, vv, ee
    # simplex: BASELINE config 5 = 3-D field (32,256,256) x 8 octaves (there is no 4-D generator in the reference)
    tabs = torch.from_numpy(init_tables([12345])).to(dev)
    vol = torch.empty(32, 256, 256, device=dev)
    ms = _time_ms(lambda: _lib.call("anoddpm_simplex_octaves", _lib.ptr(vol), 0, 32, 256, 256, _lib.ptr(tabs), None, 8, 0.8, 64.0,
                                    _lib.stream()))
    # FP64 roofline denominator, measured like MEASURED_PEAKS.json measures bf16: dependent-FMA probe, best of 10 (burst) and
    # a back-to-back loop of ~1 s (sustained)
    probe_out = torch.zeros(8, device=dev, dtype=torch.float64)
    blocks, iters = 148 * 8, 20000
    probe = lambda: _lib.call("anoddpm_fp64_fma_probe", _lib.ptr(probe_out), blocks, iters, _lib.stream())   # noqa: E731
    flop_probe = 2.0 * 8 * iters * 256 * blocks
    best = min(_time_ms(probe, n=1, warm=1) for _ in range(10))
    n_sus = max(3, int(1000.0 / max(best, 1e-3)))
    sus = _time_ms(probe, n=n_sus, warm=0)
    fp64_burst, fp64_sus = flop_probe / best / 1e9, flop_probe / sus / 1e9
    gflops = 146.0 * 8 * 32 * 65536 / ms / 1e6
    out["simplex_config5"] = {"bound": "fp64 alu", "shape": "(32,256,256) x 8 octaves, fp32 out", "ms": ms,
                              "mvox_per_s": 32 * 65536 / ms / 1e3, "gflops_fp64": gflops, "achieved": gflops / 1e3,
                              "peak": fp64_burst, "unit": "TFLOP/s", "frac": gflops / 1e3 / fp64_burst,
                              "peak_note": f"measured FP64 FMA probe (anoddpm_fp64_fma_probe): burst {fp64_burst:.1f}, sustained {fp64_sus:.1f} "
                                           "TFLOP/s counting an fma as 2 flop; the generator is compiled -fmad=false (bit-exact vs the CPU: "
                                           "separate mul and add, 1 flop per issue slot), so its ceiling is HALF this figure",
                              "frac_of_unfused_ceiling": gflops / 1e3 / (fp64_burst / 2),
                              "algorithmic": "~146 fp64 flop per voxel-octave (SURVEY App. C.4)"}
    return out


# ------------------------------------------------------------------------------------------------ CPU arms
CPU_STEPS = 10          # timed reverse steps of the CPU arm (median), after CPU_WARM warm-up steps
CPU_WARM = 2


def cpu_reverse_step_seconds(steps=CPU_STEPS, threads=None, warm=CPU_WARM):
    """The oracle port of ONE reverse step at bs 1 (UNet fwd fp32 on torch CPU + C simplex + posterior): `steps` timed
    iterations after `warm` warm-up ones, fixed thread count; returns (median, all samples).  The SAME function feeds
    `cpu_baseline` of the native arm and `--impl reference`."""
    from oracle import diffusion_oracle as DO, unet_oracle as UO
    from oracle.simplex_oracle import OracleSimplex
    from tests.util_unet import empty_state_dict
    threads = threads or cpu_threads()
    torch.set_num_threads(threads)
    kw = dict(UNET_KW)
    sd = UO.fill_state_dict(empty_state_dict(img_size=256, base_channels=128, attention_resolutions="16,8"), seed=1)
    tab = DO.Tables(DO.beta_schedule(1000, "linear"))
    x = torch.rand(1, 1, 256, 256, generator=torch.Generator().manual_seed(0)) * 2 - 1
    times = []
    with torch.no_grad():
        for i in range(steps + warm):
            t = torch.tensor([T_DISTANCE - 1 - i])
            t0 = time.perf_counter()
            eps = UO.forward(sd, x, t, **kw)
            z = torch.from_numpy(OracleSimplex(1000 + i, threads=threads).rand_3d_fixed_T_octaves(
                (256, 256), np.array([int(t)]), 6, 0.8, 64).astype(np.float32))[None]
            x = DO.sample_p(tab, x, t, eps, z)["sample"]
            if i >= warm:
                times.append(time.perf_counter() - t0)
    return float(np.median(times)), times


def cpu_threads():
    """torch's CPU conv kernels stop scaling (and oversubscribe badly) beyond a few dozen threads at batch size 1"""
    return min(os.cpu_count() or 1, 32)


def cpu_baseline():
    threads = cpu_threads()
    s, samples = cpu_reverse_step_seconds()
    # simplex microbench (config 5) on the host cores with the C oracle
    from oracle.simplex_oracle import OracleSimplex
    O = OracleSimplex(12345, threads=threads)
    O.rand_3d_octaves((4, 64, 64), 2, 0.8, 64)
    t0 = time.perf_counter()
    O.rand_3d_octaves((32, 256, 256), 8, 0.8, 64)
    sx = time.perf_counter() - t0
    return {"value": 1.0 / (T_DISTANCE * s), "unit": "images/s", "cores": threads, "kind": "port",
            "sample": f"median of {len(samples)} reverse steps at bs 1 (256x256, same UNet) after {CPU_WARM} warm-up steps, "
                      f"{s:.3f} s/step (min {min(samples):.3f}, max {max(samples):.3f}), extrapolated x800 steps/image",
            "s_per_step": s,
            "simplex_config5_mvox_per_s": 32 * 65536 / sx / 1e6}


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path.  The reference is Python and cannot
    travel to the GPU box, so this times the oracle port (pinned bit-identical to it on CPU) on the host cores -- the same
    function, thread count and sample as the native arm's `cpu_baseline`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = cpu_threads()
    s, samples = cpu_reverse_step_seconds()
    v = 1.0 / (T_DISTANCE * s)
    cb = {"value": v, "unit": "images/s", "cores": threads, "kind": "port",
          "sample": f"median of {len(samples)} reverse steps at bs 1 after {CPU_WARM} warm-up steps, {s:.3f} s/step "
                    f"(min {min(samples):.3f}, max {max(samples):.3f}), x800 steps/image"}
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus,
                      "steps": len(samples), "warmup": CPU_WARM, "ms_per_step": s * 1e3, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "same chain as the native arm, oracle port on host CPU, bs 1 sample", "t_distance": T_DISTANCE},
                      "cpu_baseline": cb,
                      "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------------ training step (config 2)
def run_train(args):
    """BASELINE configs[1] / SURVEY 8(d) config 2: one training step at bs 8 per GPU, 256x256, simplex noise.
    NOT the headline metric (never run by default).  Everything in the step is native: simplex noise, q_sample, the UNet
    forward and the hand-written backward (anoddpm_b200/train.py: tcgen05 data / weight gradients, GroupNorm / attention
    backward kernels), the bucketed gradient all-reduce and clip + AdamW + EMA (csrc/optim.cu)."""
    import copy
    from anoddpm_b200 import GaussianDiffusionModel, UNetModel, dist as D, get_beta_schedule
    from anoddpm_b200.optim import FusedAdamWEMA
    rank, world, local = D.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B = 8
    model = build_unet(dev).train()
    ema = copy.deepcopy(model)
    opt = FusedAdamWEMA(model, ema, lr=1e-4, weight_decay=0.0, betas=(0.9, 0.999), max_norm=1.0, ema_decay=0.9999)
    diff = GaussianDiffusionModel([256, 256], get_beta_schedule(1000, "linear"), loss_weight="none", noise="simplex")
    targs = {"train_start": True, "sample_distance": 800, "Batch_Size": B}
    np.random.seed(rank)
    g = torch.Generator().manual_seed(2000 + rank)
    x_host = (torch.rand(B, 1, 256, 256, generator=g) * 2 - 1).pin_memory()
    losses = []

    def step():     # diffusion_training.py:91-107: batch to the device, p_loss, backward, optimiser + EMA, loss to the host
        x = x_host.to(dev, non_blocking=True)
        opt.zero_grad()
        loss, _ = diff.p_loss(model, x, targs)
        loss.backward()
        opt.step()
        losses.append(loss.data.cpu())

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    D.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    clk = sampler.stop()
    ms = D.max_over_ranks(e0.elapsed_time(e1) / args.steps, dev)
    D.barrier()
    if rank == 0:
        # roofline of the dominant backward kernel: the tcgen05 weight gradient on the most frequent shape (128 -> 128 @256^2)
        from anoddpm_b200 import _lib
        pk = peaks()
        a16 = torch.randn(B, 256, 256, 128, device=dev).half()
        dy16 = torch.randn(B, 256, 256, 128, device=dev).half()
        dw = torch.zeros(128, 9, 128, device=dev)
        wms = _time_ms(lambda: _lib.call("anoddpm_conv_wgrad", _lib.ptr(dw), _lib.ptr(dy16), _lib.ptr(a16), B, 256, 256, 128, 128, 3, 1.0,
                                         _lib.stream()))
        wflop = 2.0 * B * 65536 * 128 * 128 * 9
        roof = {"bound": "tensor", "kernel": "conv_wgrad_kernel (tcgen05.mma kind::f16, MN-major operands from NHWC, halo box shared by 3 taps, "
                                              "one-wave split-K)", "shape": f"dW of conv3x3 128->128 @256^2 bs{B}", "achieved": wflop / wms / 1e9,
                "peak": pk["bf16_burst"], "unit": "TFLOP/s", "frac": wflop / wms / 1e9 / pk["bf16_burst"], "ms": wms,
                "algorithmic": "2*B*H*W*Cout*Cin*9 flop per launch", "traffic": None}
        v = world * B * 65536 / (ms / 1e3)
        print(json.dumps({"metric": "training px/s (256x256 bs 8 per GPU, simplex, AdamW + EMA)", "value": v,
                          "e2e": {"value": v, "unit": "px/s", "h2d_bytes_per_step": B * 65536 * 4, "d2h_bytes_per_step": 4,
                                  "api": "x_pinned.to(cuda) -> GaussianDiffusionModel.p_loss -> loss.backward() -> FusedAdamWEMA.step() -> "
                                         "loss.data.cpu(): the timed step IS the end-to-end step"},
                          "roofline": roof, "gpu_launches": "about 2800 per step (tools/train_profile.py)",
                          "unit": "px/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f16 tensor-core operands (forward and backward), f32 accumulation / gradients / optimiser",
                          "data": "synthetic", "config": {"workload": "BASELINE configs[1]: train step bs 8 per GPU", "global_batch": world * B},
                          "native_parts": "simplex noise, q_sample, UNet forward + backward (tcgen05 dgrad / wgrad, GroupNorm + attention "
                                          "backward kernels), bucketed gradient all-reduce + clip + AdamW + EMA",
                          "library_parts": "none (torch.autograd only orders the per-block backward calls and sums skip-tensor gradients)",
                          "train_flop_per_step": 3 * 2 * 277.26e9 * B * world,
                          "loss_first_last": [float(losses[0]), float(losses[-1])], "clocks": clk}))
    D.barrier()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--train", action="store_true", help="secondary metric: BASELINE configs[1] training step (see run_train)")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU (BASELINE: 32)")
    ap.add_argument("--seed", type=int, default=2024)
    ap.add_argument("--full-chain", action="store_true", help="a step = one whole 800-step chain")
    ap.add_argument("--conv-impl", default=None, choices=[None, "tc", "tc_tf32", "tc2", "tc1", "exact"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-default-model", action="store_true", help="skip the secondary line for the constructor-default UNet")
    ap.add_argument("--no-graphs", action="store_true", help="run the reverse loop eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.train:
        run_train(args)
    else:
        run_native(args)
    from anoddpm_b200 import dist as D
    D.shutdown()


if __name__ == "__main__":
    main()
