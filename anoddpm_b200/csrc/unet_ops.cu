// anoddpm_b200/csrc/unet_ops.cu -- CUDA-core (fp32-exact) kernels of the UNet denoiser:
// timestep embedding, linear, GroupNorm statistics / apply(+SiLU,+resample), the exact-fp32
// implicit-GEMM convolution, stem / head convolutions and QKV attention.
//
// Activations are NHWC fp32.  The tensor-core convolution (tcgen05, kind::tf32) lives in
// conv_tc.cu and shares anoddpm_conv_desc with the exact kernel here; the exact kernel is the
// strict-parity mode (fp32 multiplicands) and the path for shapes the tensor-core tiles do not take.
#include "common.cuh"
#include "anoddpm_b200.h"
#include <stdlib.h>
#include <cuda_fp16.h>
#include <math.h>

namespace {

struct alignas(16) F4 { float x, y, z, w; };
__device__ __forceinline__ F4 ld4(const float* p) { return *reinterpret_cast<const F4*>(p); }
__device__ __forceinline__ void st4(float* p, F4 v) { *reinterpret_cast<F4*>(p) = v; }
__device__ __forceinline__ float silu(float y) { return y / (1.0f + expf(-y)); }
// ~2 ulp variant for outputs that are about to be rounded to a 10-bit mantissa (TF32 / fp16 operands)
__device__ __forceinline__ float silu_fast(float y) { return __fdividef(y, 1.0f + __expf(-y)); }
__device__ __forceinline__ float round_tf32(float v) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return __uint_as_float(u);
}

// ---------------------------------------------------------------- timestep embedding (UNet.py:50-57)
__global__ void tembed_kernel(float* __restrict__ emb, const long long* __restrict__ t,
                              const float* __restrict__ freq, int B, int dim) {
    const int half = dim / 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * half) return;
    const int b = i / half, j = i - b * half;
    const float a = (float)t[b] * freq[j];
    emb[(size_t)b * dim + j] = sinf(a);
    emb[(size_t)b * dim + half + j] = cosf(a);
}

// out[b][:] = table[idx[b]][:] (rows of S floats, S % 4 == 0); an index outside [0, T) poisons the row with NaN
__global__ void __launch_bounds__(256) gather_rows_kernel(float* __restrict__ out, const float* __restrict__ table,
                                                          const long long* __restrict__ idx, int B, int S4, int T) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= B * S4) return;
    const int b = i / S4, c = i - b * S4;
    const long long r = idx[b];
    F4 v;
    if (r >= 0 && r < T) v = ld4(table + ((size_t)r * S4 + c) * 4);
    else v.x = v.y = v.z = v.w = __int_as_float(0x7fc00000);
    st4(out + ((size_t)b * S4 + c) * 4, v);
}


// ---------------------------------------------------------------- linear: one warp per (b, n)
// warp = (output feature n, group of LB samples): the weight row is read once per group instead of once per sample; per
// (b, n) the summation order is unchanged (lane-strided partial sums, then a butterfly).  act: bit 0 = SiLU on the
// input, bit 1 = SiLU on the output.
constexpr int LB = 8;
__global__ void __launch_bounds__(256) linear_kernel(float* __restrict__ y, int ldy, const float* __restrict__ x,
                                                     int ldx, const float* __restrict__ w,
                                                     const float* __restrict__ bias, int B, int N, int K, int act) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int ngroups = (B + LB - 1) / LB;
    if (warp >= ngroups * N) return;
    const int bg = warp / N, n = warp - bg * N;
    const int b0 = bg * LB;
    const float* wr = w + (size_t)n * K;
    float acc[LB];
#pragma unroll
    for (int i = 0; i < LB; ++i) acc[i] = 0.f;
    for (int k = lane; k < K; k += 32) {
        const float wv = wr[k];
#pragma unroll
        for (int i = 0; i < LB; ++i) {
            if (b0 + i < B) {
                float xv = x[(size_t)(b0 + i) * ldx + k];
                if (act & 1) xv = silu(xv);
                acc[i] = fmaf(xv, wv, acc[i]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    }
    if (lane < LB && b0 + lane < B) {
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < LB; ++i) if (lane == i) v = acc[i];
        v += bias ? bias[n] : 0.f;
        if (act & 2) v = silu(v);
        y[(size_t)(b0 + lane) * ldy + n] = v;
    }
}

// ---------------------------------------------------------------- GroupNorm statistics (per channel)
// stats[b][c] = (sum, sum of squares) over the H*W pixels of channel c, fp64.  Per-CHANNEL (not per-group) sums
// are what every producer of a tensor can emit without knowing how a consumer will group or concatenate it: the
// consumer (gn_apply) adds up the channels of each group, possibly across two concatenated tensors.
// grid (chunks, B); a CTA reduces a pixel range of one sample; each thread owns a fixed channel vector,
// accumulates in fp32, then fp64 shared + global atomics.
template <int V>
__global__ void __launch_bounds__(256) gn_stats_kernel(double* __restrict__ stats, const float* __restrict__ x, int C,
                                                       int HW, int pix_per_cta) {
    extern __shared__ double sh[];   // [C][2]
    const int b = blockIdx.y, ncv = C / V;
    for (int i = threadIdx.x; i < C * 2; i += blockDim.x) sh[i] = 0.0;
    __syncthreads();
    const int p0 = blockIdx.x * pix_per_cta;
    const int p1 = min(HW, p0 + pix_per_cta);
    int tpr, rows;
    if (ncv >= 256) { tpr = 256; rows = 1; }
    else { tpr = ncv; rows = 256 / ncv; }
    const int r = threadIdx.x / tpr, lane = threadIdx.x - r * tpr;
    if (r < rows) {
        for (int cv = lane; cv < ncv; cv += tpr) {
            const int c = cv * V;
            const float* src = x + (size_t)b * HW * C + c;
            float s[V], q[V];
#pragma unroll
            for (int j = 0; j < V; ++j) { s[j] = 0.f; q[j] = 0.f; }
            for (int p = p0 + r; p < p1; p += rows) {
                if (V == 4) {
                    const F4 v = ld4(src + (size_t)p * C);
                    s[0] += v.x; s[1 % V] += v.y; s[2 % V] += v.z; s[3 % V] += v.w;
                    q[0] += v.x * v.x; q[1 % V] += v.y * v.y; q[2 % V] += v.z * v.z; q[3 % V] += v.w * v.w;
                } else {
                    const float v = src[(size_t)p * C];
                    s[0] += v;
                    q[0] += v * v;
                }
            }
#pragma unroll
            for (int j = 0; j < V; ++j) {
                atomicAdd(&sh[(c + j) * 2], (double)s[j]);
                atomicAdd(&sh[(c + j) * 2 + 1], (double)q[j]);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * 2; i += blockDim.x) atomicAdd(&stats[(size_t)b * C * 2 + i], sh[i]);
}

// ---------------------------------------------------------------- GroupNorm apply (+SiLU, +resample)
// grid (pixel chunks, B).  Output NHWC (B, Ho, Wo, C).  mode 0 same, 1 avg-pool 2x2 (of activated values),
// 2 nearest x2.  Like gn_stats, a thread owns a fixed channel vector (scale/shift computed once, no per-element
// integer division) and walks output pixels; 4 pixels are in flight per thread for memory-level parallelism.
template <int V>
__device__ __forceinline__ void gn_load(float (&v)[V], const float* p) {
    if (V == 4) {
        const F4 t = ld4(p);
        v[0] = t.x; v[1 % V] = t.y; v[2 % V] = t.z; v[3 % V] = t.w;
    } else {
        v[0] = p[0];
    }
}
template <int V>
__device__ __forceinline__ void gn_store(float* p, const float (&v)[V]) {
    if (V == 4) {
        F4 o; o.x = v[0]; o.y = v[1 % V]; o.z = v[2 % V]; o.w = v[3 % V];
        st4(p, o);
    } else {
        p[0] = v[0];
    }
}
// fp16 store (round to nearest, saturating to the finite range): `p` indexes HALF elements
template <int V>
__device__ __forceinline__ void gn_store_h(__half* p, const float (&v)[V]) {
    if (V == 4) {
        const __half2 lo = __floats2half2_rn(fminf(fmaxf(v[0], -65504.f), 65504.f), fminf(fmaxf(v[1 % V], -65504.f), 65504.f));
        const __half2 hi = __floats2half2_rn(fminf(fmaxf(v[2 % V], -65504.f), 65504.f), fminf(fmaxf(v[3 % V], -65504.f), 65504.f));
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t*>(&lo);
        o.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(p) = o;
    } else {
        p[0] = __float2half_rn(fminf(fmaxf(v[0], -65504.f), 65504.f));
    }
}

template <int V, int NT>   // NT = taps per output pixel: 4 for the avg-pool mode, else 1
__global__ void __launch_bounds__(256) gn_apply_kernel(float* __restrict__ out, float* __restrict__ raw_out,
                                                       const float* __restrict__ x1, int C1,
                                                       const float* __restrict__ x2, int C2,
                                                       const double* __restrict__ stats1,
                                                       const double* __restrict__ stats2,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       int H, int W, int groups, float eps, int do_silu, int mode,
                                                       int rtf32, int pix_per_cta) {
    __shared__ float s_mean[64], s_rstd[64];
    const int b = blockIdx.y, C = C1 + C2, gs = C / groups, ncv = C / V;
    {
        // group statistics = sum over the group's gs channels (which may straddle the two sources) and over the statistic
        // replicas; thread (g, r) adds up replica r of group g, 8 consecutive lanes are then combined by shuffles
        const int g = threadIdx.x / ANODDPM_STAT_REPLICAS, r = threadIdx.x % ANODDPM_STAT_REPLICAS;
        const int nb = gridDim.y;   // batch size
        double s = 0.0, q = 0.0;
        if (g < groups) {
            for (int c = g * gs; c < (g + 1) * gs; ++c) {
                const bool first = c < C1;
                const int Cs = first ? C1 : C2, cc = first ? c : c - C1;
                const double* st = (first ? stats1 : stats2) + (size_t)r * nb * Cs * 2 + ((size_t)b * Cs + cc) * 2;
                s += st[0];
                q += st[1];
            }
        }
#pragma unroll
        for (int o = 1; o < ANODDPM_STAT_REPLICAS; o <<= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (g < groups && r == 0) {
            const double n = (double)H * W * gs;
            const double m = s / n;
            double var = q / n - m * m;
            if (var < 0) var = 0;
            s_mean[g] = (float)m;
            s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
        }
    }
    __syncthreads();
    const int Ho = mode == 1 ? H / 2 : (mode == 2 ? H * 2 : H);
    const int Wo = mode == 1 ? W / 2 : (mode == 2 ? W * 2 : W);
    const int npo = Ho * Wo;
    // the pixel loop runs over OUTPUT pixels, except for the nearest-upsample mode: there it runs over INPUT pixels and every
    // normalised + activated value is computed once and stored to its four children (the output-pixel form evaluated each
    // SiLU four times: 339 us for the 128^2 -> 256^2 level)
    const int npl = mode == 2 ? H * W : npo;
    const int p0 = blockIdx.x * pix_per_cta;
    const int p1 = min(npl, p0 + pix_per_cta);
    int tpr, rows;
    if (ncv >= 256) { tpr = 256; rows = 1; }
    else { tpr = ncv; rows = 256 / ncv; }
    const int r = threadIdx.x / tpr, lane = threadIdx.x - r * tpr;
    if (r >= rows) return;
    for (int cv = lane; cv < ncv; cv += tpr) {
        const int c = cv * V;
        const float* src;
        int cs, cc;
        if (c < C1) { src = x1; cs = C1; cc = c; }
        else { src = x2; cs = C2; cc = c - C1; }
        src += (size_t)b * H * W * cs + cc;
        float sc[V], bi[V];
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const int g = (c + j) / gs;
            sc[j] = s_rstd[g] * gamma[c + j];
            bi[j] = beta[c + j] - sc[j] * s_mean[g];
        }
        const bool f16 = rtf32 == 2;     // out_mode: 0 fp32, 1 fp32 rounded to TF32, 2 fp16
        float* dst = out + (f16 ? 0 : (size_t)b * npo * C + c);
        float* rdst = raw_out ? raw_out + (f16 ? 0 : (size_t)b * npo * C + c) : nullptr;
        __half* dsth = reinterpret_cast<__half*>(out) + (size_t)b * npo * C + c;
        __half* rdsth = reinterpret_cast<__half*>(raw_out) + (size_t)b * npo * C + c;
        constexpr int U = NT == 4 ? 1 : 4;
        for (int pb = p0 + r; pb < p1; pb += rows * U) {
            float v[U][NT][V];   // [pixel in flight][tap][channel]
            constexpr int ntap = NT;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int po = pb + u * rows;
                if (po < p1) {
                    if (mode == 0 || mode == 2) {
                        gn_load<V>(v[u][0], src + (size_t)po * cs);
                    } else {
                        const int ho = po / Wo, wo = po - ho * Wo;
                        if (NT == 1) {
                            gn_load<V>(v[u][0], src + ((size_t)(ho >> 1) * W + (wo >> 1)) * cs);
                        } else {
                            const float* q = src + ((size_t)(2 * ho) * W + 2 * wo) * cs;
                            gn_load<V>(v[u][0], q);
                            gn_load<V>(v[u][1 % NT], q + cs);
                            gn_load<V>(v[u][2 % NT], q + (size_t)W * cs);
                            gn_load<V>(v[u][3 % NT], q + (size_t)W * cs + cs);
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int po = pb + u * rows;
                if (po < p1) {
                    float acc[V];
#pragma unroll
                    for (int j = 0; j < V; ++j) acc[j] = 0.f;
#pragma unroll
                    for (int tp = 0; tp < ntap; ++tp) {
#pragma unroll
                        for (int j = 0; j < V; ++j) {
                            float y = v[u][tp][j] * sc[j] + bi[j];
                            if (do_silu) y = rtf32 ? silu_fast(y) : silu(y);
                            acc[j] += y;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < V; ++j) {
                        if (mode == 1) acc[j] *= 0.25f;
                        if (rtf32 == 1) acc[j] = round_tf32(acc[j]);
                    }
                    if (mode == 2) {
                        const int hi = po / W, wi = po - hi * W;
                        const size_t o0 = ((size_t)(2 * hi) * Wo + 2 * wi) * C, o1 = o0 + (size_t)Wo * C;
                        if (f16) {
                            gn_store_h<V>(dsth + o0, acc); gn_store_h<V>(dsth + o0 + C, acc);
                            gn_store_h<V>(dsth + o1, acc); gn_store_h<V>(dsth + o1 + C, acc);
                        } else {
                            gn_store<V>(dst + o0, acc); gn_store<V>(dst + o0 + C, acc);
                            gn_store<V>(dst + o1, acc); gn_store<V>(dst + o1 + C, acc);
                        }
                    } else if (f16) gn_store_h<V>(dsth + (size_t)po * C, acc);
                    else gn_store<V>(dst + (size_t)po * C, acc);
                    if (mode == 1) {
                        if (raw_out) {   // fp32 avg-pooled copy of the INPUT: the residual of a downsampling ResBlock
                            float pr[V];
#pragma unroll
                            for (int j = 0; j < V; ++j)
                                pr[j] = (v[u][0][j] + v[u][1 % NT][j] + v[u][2 % NT][j] + v[u][3 % NT][j]) * 0.25f;
                            gn_store<V>(raw_out + (size_t)b * npo * C + c + (size_t)po * C, pr);
                        }
                    } else if (rdst) {   // mode 0: rounded concatenated copy of the INPUT for the fused skip conv
                        if (f16) {
                            gn_store_h<V>(rdsth + (size_t)po * C, v[u][0]);
                        } else {
                            float rv[V];
#pragma unroll
                            for (int j = 0; j < V; ++j) rv[j] = round_tf32(v[u][0][j]);
                            gn_store<V>(rdst + (size_t)po * C, rv);
                        }
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------- GroupNorm apply, fp16 input (+ per-channel shift)
// x holds fp16 values of (h - shift), shift[b][c] = shift_c[c] + shift_bc[b][c] (conv bias + timestep embedding, which
// the producing conv left out so that the stored half carries only the convolution sum); stats are those of h.
// y = act((x + shift - mean) * rstd * gamma + beta), same resolution, one source.  Thread = 8 channels (16 B in).
template <int OM>   // output mode: 0 fp32, 1 fp32 rounded to TF32, 2 fp16
__global__ void __launch_bounds__(256) gn_apply_h_kernel(void* __restrict__ out_v, const __half* __restrict__ x, int C,
                                                         const double* __restrict__ stats, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, const float* __restrict__ shift_c,
                                                         const float* __restrict__ shift_bc, int shift_stride, int HW,
                                                         int groups, float eps, int do_silu, int pix_per_cta) {
    __shared__ float s_mean[64], s_rstd[64];
    const int b = blockIdx.y, gs = C / groups, ncv = C / 8;
    {
        const int g = threadIdx.x / ANODDPM_STAT_REPLICAS, r = threadIdx.x % ANODDPM_STAT_REPLICAS;
        const int nb = gridDim.y;
        double s = 0.0, q = 0.0;
        if (g < groups) {
            for (int c = g * gs; c < (g + 1) * gs; ++c) {
                const double* st = stats + (size_t)r * nb * C * 2 + ((size_t)b * C + c) * 2;
                s += st[0];
                q += st[1];
            }
        }
#pragma unroll
        for (int o = 1; o < ANODDPM_STAT_REPLICAS; o <<= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (g < groups && r == 0) {
            const double n = (double)HW * gs;
            const double m = s / n;
            double var = q / n - m * m;
            if (var < 0) var = 0;
            s_mean[g] = (float)m;
            s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
        }
    }
    __syncthreads();
    const int p0 = blockIdx.x * pix_per_cta;
    const int p1 = min(HW, p0 + pix_per_cta);
    int tpr, rows;
    if (ncv >= 256) { tpr = 256; rows = 1; }
    else { tpr = ncv; rows = 256 / ncv; }
    const int r = threadIdx.x / tpr, lane = threadIdx.x - r * tpr;
    if (r >= rows) return;
    for (int cv = lane; cv < ncv; cv += tpr) {
        const int c = cv * 8;
        float sc[8], bi[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int g = (c + j) / gs;
            float sh = shift_c ? shift_c[c + j] : 0.f;
            if (shift_bc) sh += shift_bc[(size_t)b * shift_stride + c + j];
            sc[j] = s_rstd[g] * gamma[c + j];
            bi[j] = beta[c + j] + sc[j] * (sh - s_mean[g]);
        }
        const __half* src = x + (size_t)b * HW * C + c;
        constexpr int U = 4;
        for (int pb = p0 + r; pb < p1; pb += rows * U) {
            uint4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int po = pb + u * rows;
                if (po < p1) v[u] = *reinterpret_cast<const uint4*>(src + (size_t)po * C);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int po = pb + u * rows;
                if (po >= p1) continue;
                const __half2* h2 = reinterpret_cast<const __half2*>(&v[u]);
                float y[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(h2[j]);
                    y[2 * j] = f.x * sc[2 * j] + bi[2 * j];
                    y[2 * j + 1] = f.y * sc[2 * j + 1] + bi[2 * j + 1];
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (do_silu) y[j] = OM ? silu_fast(y[j]) : silu(y[j]);
                    if (OM == 1) y[j] = round_tf32(y[j]);
                }
                const size_t o = ((size_t)b * HW + po) * C + c;
                if (OM == 2) {
                    uint4 w;
                    uint32_t* wp = reinterpret_cast<uint32_t*>(&w);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const __half2 hh = __floats2half2_rn(fminf(fmaxf(y[2 * j], -65504.f), 65504.f),
                                                             fminf(fmaxf(y[2 * j + 1], -65504.f), 65504.f));
                        wp[j] = *reinterpret_cast<const uint32_t*>(&hh);
                    }
                    *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(out_v) + o) = w;
                } else {
                    float* op = reinterpret_cast<float*>(out_v) + o;
                    F4 a, bq;
                    a.x = y[0]; a.y = y[1]; a.z = y[2]; a.w = y[3];
                    bq.x = y[4]; bq.y = y[5]; bq.z = y[6]; bq.w = y[7];
                    st4(op, a);
                    st4(op + 4, bq);
                }
            }
        }
    }
}

// ---------------------------------------------------------------- GroupNorm as a per-(sample, channel) affine table
// table[b][c] = (scale, shift): GroupNorm(x + off)[b, c] = x * scale + shift.  Feeds the operand transform inside the
// tensor-core convolution (conv_tc3.cu, gn_table): the consumer normalises while it loads, no normalised copy exists.
// One CTA per sample; thread (g, r) adds up replica r of group g exactly like gn_apply_kernel, over one or two tensors.
__global__ void __launch_bounds__(256) gn_table_kernel(float2* __restrict__ table, const double* __restrict__ st1, int C1,
                                                       const double* __restrict__ st2, int C2,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const float* __restrict__ shift_c, const float* __restrict__ shift_bc,
                                                       int shift_stride, int HW, int groups, float eps) {
    __shared__ float s_mean[64], s_rstd[64];
    const int b = blockIdx.x, nb = gridDim.x, C = C1 + C2, gs = C / groups;
    {
        const int g = threadIdx.x / ANODDPM_STAT_REPLICAS, r = threadIdx.x % ANODDPM_STAT_REPLICAS;
        double s = 0.0, q = 0.0;
        if (g < groups) {
            for (int c = g * gs; c < (g + 1) * gs; ++c) {
                const double* st = c < C1 ? st1 + (size_t)r * nb * C1 * 2 + ((size_t)b * C1 + c) * 2
                                          : st2 + (size_t)r * nb * C2 * 2 + ((size_t)b * C2 + (c - C1)) * 2;
                s += st[0];
                q += st[1];
            }
        }
#pragma unroll
        for (int o = 1; o < ANODDPM_STAT_REPLICAS; o <<= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (g < groups && r == 0) {
            const double n = (double)HW * gs;
            const double m = s / n;
            double var = q / n - m * m;
            if (var < 0) var = 0;
            s_mean[g] = (float)m;
            s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / gs;
        float sh = shift_c ? shift_c[c] : 0.f;
        if (shift_bc) sh += shift_bc[(size_t)b * shift_stride + c];
        const float sc = s_rstd[g] * gamma[c];
        table[(size_t)b * C + c] = make_float2(sc, beta[c] + sc * (sh - s_mean[g]));
    }
}

// ---------------------------------------------------------------- exact fp32 implicit-GEMM conv
// CTA tile 64 pixels x 64 output channels, K chunks of 16 input channels of one tap.
constexpr int CBM = 64, CBN = 64, CBK = 16;

struct ConvArgs {
    anoddpm_conv_desc d;
    int M;     // B*H*W
    int Ktot;  // packed weight row length
};

__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvArgs A) {
    __shared__ float As[CBK][CBM + 4];
    __shared__ float Bs[CBK][CBN + 4];
    const anoddpm_conv_desc& d = A.d;
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * CBM, n0 = blockIdx.y * CBN;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    // loader coordinates
    const int lp = tid >> 2, lc4 = (tid & 3) * 4;  // A: pixel row lp, channel offset lc4 ; B: cout row lp, k offset lc4
    const int p = m0 + lp;
    int pb = 0, ph = 0, pw = 0;
    const bool pvalid = p < A.M;
    if (pvalid) { pw = p % d.W; ph = (p / d.W) % d.H; pb = p / (d.W * d.H); }
    const int nrow = n0 + lp;
    const bool nvalid = nrow < d.Cout;

    int kbase = 0;  // running offset into the packed K axis
    for (int seg = 0; seg < 3; ++seg) {
        const float* src = static_cast<const float*>(seg == 0 ? d.a : (seg == 1 ? d.r1 : d.r2));
        const int Cs = seg == 0 ? d.Ca : (seg == 1 ? d.Cr1 : d.Cr2);
        if (src == nullptr || Cs == 0) continue;
        const int ks = seg == 0 ? d.ksize : 1;
        const int ntaps = ks * ks, padr = ks / 2;
        for (int tap = 0; tap < ntaps; ++tap) {
            const int dy = tap / ks - padr, dx = tap % ks - padr;
            const int hh = ph + dy, ww = pw + dx;
            const bool inb = pvalid && hh >= 0 && hh < d.H && ww >= 0 && ww < d.W;
            const float* arow = src + (((size_t)pb * d.H + hh) * d.W + ww) * Cs;
            for (int c0 = 0; c0 < Cs; c0 += CBK) {
                F4 av = {0.f, 0.f, 0.f, 0.f};
                if (inb && c0 + lc4 < Cs) av = ld4(arow + c0 + lc4);
                F4 bv = {0.f, 0.f, 0.f, 0.f};
                if (nvalid && c0 + lc4 < Cs) bv = ld4(static_cast<const float*>(d.w) + (size_t)nrow * A.Ktot + kbase + tap * Cs + c0 + lc4);
                __syncthreads();
                As[lc4 + 0][lp] = av.x; As[lc4 + 1][lp] = av.y; As[lc4 + 2][lp] = av.z; As[lc4 + 3][lp] = av.w;
                Bs[lc4 + 0][lp] = bv.x; Bs[lc4 + 1][lp] = bv.y; Bs[lc4 + 2][lp] = bv.z; Bs[lc4 + 3][lp] = bv.w;
                __syncthreads();
#pragma unroll
                for (int k = 0; k < CBK; ++k) {
                    const F4 a4 = ld4(&As[k][ty * 4]);
                    const F4 b4 = ld4(&Bs[k][tx * 4]);
                    const float a[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
                }
            }
        }
        kbase += ntaps * Cs;
    }

    // epilogue
    const int n = n0 + tx * 4;
    if (n >= d.Cout) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = m0 + ty * 4 + i;
        if (q >= A.M) continue;
        const int w = q % d.W, h = (q / d.W) % d.H, b = q / (d.W * d.H);
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[j] = acc[i][j];
            if (d.bias) v[j] += d.bias[n + j];
            if (d.emb) v[j] += d.emb[(size_t)b * d.emb_stride + n + j];
        }
        if (d.res_mode == 1) {
            const F4 r = ld4(d.res + (size_t)q * d.Cout + n);
            v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
        } else if (d.res_mode == 2) {
            const int H2 = d.H * 2, W2 = d.W * 2;
            const float* rb = d.res + (((size_t)b * H2 + 2 * h) * W2 + 2 * w) * d.Cout + n;
            const F4 r0 = ld4(rb), r1 = ld4(rb + d.Cout), r2 = ld4(rb + (size_t)W2 * d.Cout),
                     r3 = ld4(rb + (size_t)W2 * d.Cout + d.Cout);
            v[0] += (r0.x + r1.x + r2.x + r3.x) * 0.25f;
            v[1] += (r0.y + r1.y + r2.y + r3.y) * 0.25f;
            v[2] += (r0.z + r1.z + r2.z + r3.z) * 0.25f;
            v[3] += (r0.w + r1.w + r2.w + r3.w) * 0.25f;
        } else if (d.res_mode == 3) {
            const int Hh = d.H / 2, Wh = d.W / 2;
            const F4 r = ld4(d.res + (((size_t)b * Hh + (h >> 1)) * Wh + (w >> 1)) * d.Cout + n);
            v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
        }
        F4 o; o.x = v[0]; o.y = v[1]; o.z = v[2]; o.w = v[3];
        st4(static_cast<float*>(d.out) + (size_t)q * d.Cout + n, o);
    }
}

// ---------------------------------------------------------------- stem: NCHW image -> NHWC features
// thread = (pixel, 4 output channels): the 9*Cin inputs are warp-broadcast loads, weights sit in shared memory
// as [tap*Cin][Cout] (conflict-free float4 reads), one coalesced float4 store per thread.  HBM-bound on the store.
__global__ void __launch_bounds__(256) conv_stem_kernel(float* __restrict__ out, const float* __restrict__ x,
                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                        int B, int Cin, int H, int W, int Cout) {
    extern __shared__ float sw[];  // [Cin*9][Cout] then bias[Cout]
    const int K = Cin * 9;
    for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) {
        const int co = i / K, k = i - co * K;      // OIHW: w[co][ci][ky][kx], k = ci*9 + ky*3 + kx
        sw[k * Cout + co] = w[i];
    }
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[K * Cout + i] = bias ? bias[i] : 0.f;
    __syncthreads();
    const int nc4 = Cout >> 2;
    const int ppb = blockDim.x / nc4;              // pixels per block iteration
    const int pl = threadIdx.x / nc4, c4 = (threadIdx.x - pl * nc4) * 4;
    if (pl >= ppb) return;
    const int npix = B * H * W;
    for (int p = blockIdx.x * ppb + pl; p < npix; p += gridDim.x * ppb) {
        const int xw = p % W, yh = (p / W) % H, b = p / (W * H);
        float a0 = sw[K * Cout + c4], a1 = sw[K * Cout + c4 + 1], a2 = sw[K * Cout + c4 + 2], a3 = sw[K * Cout + c4 + 3];
        // bias first, then (ci, ky, kx) accumulation
        for (int ci = 0; ci < Cin; ++ci) {
            const float* xp = x + ((size_t)b * Cin + ci) * H * W;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int yy = yh + ky - 1;
                if (yy < 0 || yy >= H) continue;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int xx = xw + kx - 1;
                    if (xx < 0 || xx >= W) continue;
                    const float v = xp[(size_t)yy * W + xx];
                    const F4 wv = ld4(&sw[(ci * 9 + ky * 3 + kx) * Cout + c4]);
                    a0 = fmaf(v, wv.x, a0); a1 = fmaf(v, wv.y, a1); a2 = fmaf(v, wv.z, a2); a3 = fmaf(v, wv.w, a3);
                }
            }
        }
        F4 o; o.x = a0; o.y = a1; o.z = a2; o.w = a3;
        st4(out + (size_t)p * Cout + c4, o);
    }
}

// stem, 4 horizontally adjacent pixels per thread (W % 4 == 0): the 3 x 6 input window is loaded once per input
// channel, index arithmetic is amortised over 16 outputs, four coalesced float4 stores per thread.
__global__ void __launch_bounds__(256) conv_stem4_kernel(float* __restrict__ out, const float* __restrict__ x,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         int B, int Cin, int H, int W, int Cout) {
    extern __shared__ float sw[];  // [Cin*9][Cout] then bias[Cout]
    const int K = Cin * 9;
    for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) {
        const int co = i / K, k = i - co * K;
        sw[k * Cout + co] = w[i];
    }
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[K * Cout + i] = bias ? bias[i] : 0.f;
    __syncthreads();
    const int nc4 = Cout >> 2;
    const int gpb = blockDim.x / nc4;              // pixel groups per block iteration
    const int gl = threadIdx.x / nc4, c4 = (threadIdx.x - gl * nc4) * 4;
    if (gl >= gpb) return;
    const int W4 = W >> 2;
    const int ngroups = B * H * W4;
    for (int g = blockIdx.x * gpb + gl; g < ngroups; g += gridDim.x * gpb) {
        const int xg = g % W4, yh = (g / W4) % H, b = g / (W4 * H);
        const int x0 = xg * 4;
        float acc[4][4];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[p][e] = sw[K * Cout + c4 + e];
        for (int ci = 0; ci < Cin; ++ci) {
            const float* xp = x + ((size_t)b * Cin + ci) * H * W;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int yy = yh + ky - 1;
                if (yy < 0 || yy >= H) continue;
                float win[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const int xx = x0 + j - 1;
                    win[j] = (xx >= 0 && xx < W) ? xp[(size_t)yy * W + xx] : 0.f;
                }
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const F4 wv = ld4(&sw[(ci * 9 + ky * 3 + kx) * Cout + c4]);
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const float v = win[p + kx];
                        acc[p][0] = fmaf(v, wv.x, acc[p][0]);
                        acc[p][1] = fmaf(v, wv.y, acc[p][1]);
                        acc[p][2] = fmaf(v, wv.z, acc[p][2]);
                        acc[p][3] = fmaf(v, wv.w, acc[p][3]);
                    }
                }
            }
        }
        float* o = out + (((size_t)b * H + yh) * W + x0) * Cout + c4;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            F4 v; v.x = acc[p][0]; v.y = acc[p][1]; v.z = acc[p][2]; v.w = acc[p][3];
            st4(o + (size_t)p * Cout, v);
        }
    }
}

// ---------------------------------------------------------------- head: NHWC features -> NCHW image
// CTA = 8x32 output pixels; per 16-channel chunk the (10 x 34)-pixel halo tile is staged in shared memory as
// [pixel][20] floats (stride 20: 128-bit reads by 8 pixel-consecutive lanes touch all 32 banks once), weights of the
// chunk as [co][tap][16].  Each thread owns one output pixel (all COUT <= 4 channels, compile-time).  The input is
// read once from HBM (+33% halo); inner loop = 2 LDS.128 + 4*COUT FFMA per 4 channels.
constexpr int HT_H = 8, HT_W = 32, HT_C = 16, HT_P = 20;
// GN = true: `a` is the raw tensor and GroupNorm + SiLU (UNet.py:384-386) are applied while the tile is staged (each
// element once per CTA), from per-channel statistics -- no normalised copy of the last feature map is written at all.
constexpr int HT_MAXC = 512;
template <int COUT, bool GN>
__global__ void __launch_bounds__(256) conv_head_kernel(float* __restrict__ out, const float* __restrict__ a,
                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                        int B, int Cin, int H, int W, const double* __restrict__ stats,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        int groups, float eps) {
    __shared__ __align__(16) float st[(HT_H + 2) * (HT_W + 2) * HT_P];
    __shared__ __align__(16) float swt[COUT * 9 * HT_C];
    __shared__ __align__(16) float s_sc[GN ? HT_MAXC : 4], s_bi[GN ? HT_MAXC : 4];
    __shared__ float s_mean[64], s_rstd[64];
    const int tiles_x = (W + HT_W - 1) / HT_W, tiles_y = (H + HT_H - 1) / HT_H;
    const int tile = blockIdx.x;
    const int b = tile / (tiles_x * tiles_y), tr = tile - b * tiles_x * tiles_y;
    if (GN) {
        // group statistics exactly as gn_apply_kernel: thread (g, r) adds up replica r of group g
        const int gs = Cin / groups;
        const int g = threadIdx.x / ANODDPM_STAT_REPLICAS, r = threadIdx.x % ANODDPM_STAT_REPLICAS;
        double s = 0.0, q = 0.0;
        if (g < groups) {
            for (int c = g * gs; c < (g + 1) * gs; ++c) {
                const double* sp = stats + (size_t)r * B * Cin * 2 + ((size_t)b * Cin + c) * 2;
                s += sp[0];
                q += sp[1];
            }
        }
#pragma unroll
        for (int o = 1; o < ANODDPM_STAT_REPLICAS; o <<= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (g < groups && r == 0) {
            const double n = (double)H * W * gs;
            const double m = s / n;
            double var = q / n - m * m;
            if (var < 0) var = 0;
            s_mean[g] = (float)m;
            s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
        }
        __syncthreads();
        for (int c = threadIdx.x; c < Cin; c += blockDim.x) {
            const int gg = c / gs;
            const float sc = s_rstd[gg] * gamma[c];
            s_sc[c] = sc;
            s_bi[c] = beta[c] - sc * s_mean[gg];
        }
    }
    const int y0 = (tr / tiles_x) * HT_H, x0 = (tr % tiles_x) * HT_W;
    const int py = threadIdx.x / HT_W, px = threadIdx.x % HT_W;
    float acc[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[co] = 0.f;
    for (int c0 = 0; c0 < Cin; c0 += HT_C) {
        __syncthreads();
        for (int i = threadIdx.x; i < (HT_H + 2) * (HT_W + 2) * (HT_C / 4); i += blockDim.x) {
            const int pix = i >> 2, q = (i & 3) * 4;
            const int ly = pix / (HT_W + 2), lx = pix - ly * (HT_W + 2);
            const int gy = y0 + ly - 1, gx = x0 + lx - 1;
            F4 v = {0.f, 0.f, 0.f, 0.f};
            if (gy >= 0 && gy < H && gx >= 0 && gx < W && c0 + q < Cin) {
                v = ld4(a + (((size_t)b * H + gy) * W + gx) * Cin + c0 + q);
                if (GN) {   // the zero padding applies to the ACTIVATED tensor: out-of-image taps stay 0
                    const F4 sc = ld4(&s_sc[c0 + q]), bi = ld4(&s_bi[c0 + q]);
                    v.x = silu(v.x * sc.x + bi.x);
                    v.y = silu(v.y * sc.y + bi.y);
                    v.z = silu(v.z * sc.z + bi.z);
                    v.w = silu(v.w * sc.w + bi.w);
                }
            }
            st4(&st[pix * HT_P + q], v);
        }
        for (int i = threadIdx.x; i < COUT * 9 * HT_C; i += blockDim.x) {
            const int co = i / (9 * HT_C), r = i - co * 9 * HT_C, tap = r / HT_C, c = r - tap * HT_C;
            swt[i] = (c0 + c < Cin) ? w[(size_t)co * 9 * Cin + tap * Cin + c0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const float* sp = &st[((py + tap / 3) * (HT_W + 2) + px + tap % 3) * HT_P];
#pragma unroll
            for (int c = 0; c < HT_C; c += 4) {
                const F4 v = ld4(sp + c);
#pragma unroll
                for (int co = 0; co < COUT; ++co) {
                    const F4 ww = ld4(&swt[(co * 9 + tap) * HT_C + c]);
                    acc[co] = fmaf(v.x, ww.x, acc[co]);
                    acc[co] = fmaf(v.y, ww.y, acc[co]);
                    acc[co] = fmaf(v.z, ww.z, acc[co]);
                    acc[co] = fmaf(v.w, ww.w, acc[co]);
                }
            }
        }
    }
    const int gy = y0 + py, gx = x0 + px;
    if (gy < H && gx < W) {
#pragma unroll
        for (int co = 0; co < COUT; ++co)
            out[(((size_t)b * COUT + co) * H + gy) * W + gx] = acc[co] + (bias ? bias[co] : 0.f);
    }
}

// ---------------------------------------------------------------- attention (UNet.py:137-153)
// CTA: one (b, head), QB queries, 256 threads.  smem: q tile [QB][ch] (pre-scaled by 1/sqrt(ch)), scores [QB][T].
// K and V of the (b, head) are streamed once per CTA, so QB = 32 re-reads them T/32 times in total.
template <int QB>
__global__ void __launch_bounds__(256) attention_kernel(float* __restrict__ out, const float* __restrict__ qkv, int T,
                                                        int C, int heads, int rtf32) {
    extern __shared__ float sm[];
    constexpr int NW = 8;
    const int ch = C / heads;
    float* sq = sm;            // QB * ch
    float* ss = sm + QB * ch;  // QB * T
    const int b = blockIdx.z, h = blockIdx.y, t0 = blockIdx.x * QB;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t row = (size_t)3 * C;
    const float* base = qkv + (size_t)b * T * row + (size_t)h * 3 * ch;
    const float scale = 1.0f / sqrtf((float)ch);
    for (int i = tid; i < QB * ch; i += blockDim.x) {
        const int qi = i / ch, c = i - qi * ch;
        sq[i] = (t0 + qi < T) ? base[(size_t)(t0 + qi) * row + c] * scale : 0.f;
    }
    __syncthreads();
    // scores: warp per key, lanes over channels (coalesced K reads), QB dot products per key
    for (int s = warp; s < T; s += NW) {
        const float* kr = base + (size_t)s * row + ch;
        float part[QB];
#pragma unroll
        for (int qi = 0; qi < QB; ++qi) part[qi] = 0.f;
        for (int c = lane; c < ch; c += 32) {
            const float kv = kr[c];
#pragma unroll
            for (int qi = 0; qi < QB; ++qi) part[qi] = fmaf(sq[qi * ch + c], kv, part[qi]);
        }
#pragma unroll
        for (int qi = 0; qi < QB; ++qi) {
            float v = part[qi];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) ss[qi * T + s] = v;
        }
    }
    __syncthreads();
    // softmax rows
    for (int qi = warp; qi < QB; qi += NW) {
        float* r = ss + qi * T;
        float mx = -INFINITY;
        for (int s = lane; s < T; s += 32) mx = fmaxf(mx, r[s]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int s = lane; s < T; s += 32) {
            const float e = expf(r[s] - mx);
            r[s] = e;
            sum += e;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.0f / sum;
        for (int s = lane; s < T; s += 32) r[s] *= inv;
    }
    __syncthreads();
    // out = P . V : thread per channel (coalesced V reads), probabilities broadcast from smem
    for (int c = tid; c < ch; c += blockDim.x) {
        float acc[QB];
#pragma unroll
        for (int qi = 0; qi < QB; ++qi) acc[qi] = 0.f;
        const float* vr = base + 2 * ch + c;
#pragma unroll 4
        for (int s = 0; s < T; ++s) {
            const float v = vr[(size_t)s * row];
#pragma unroll
            for (int qi = 0; qi < QB; ++qi) acc[qi] = fmaf(ss[qi * T + s], v, acc[qi]);
        }
#pragma unroll
        for (int qi = 0; qi < QB; ++qi)
            if (t0 + qi < T) {
                const size_t o = ((size_t)b * T + t0 + qi) * C + h * ch + c;
                if (rtf32 == 2) reinterpret_cast<__half*>(out)[o] = __float2half_rn(fminf(fmaxf(acc[qi], -65504.f), 65504.f));
                else out[o] = rtf32 == 1 ? round_tf32(acc[qi]) : acc[qi];
            }
    }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int anoddpm_conv_tc(const anoddpm_conv_desc* d, void* stream);   // conv_tc.cu  (impl 1)
int anoddpm_conv_tc2(const anoddpm_conv_desc* d, void* stream);  // conv_tc2.cu (impl 2)
int anoddpm_conv_tc3(const anoddpm_conv_desc* d, void* stream);  // conv_tc3.cu (impl 3)

extern "C" int anoddpm_gather_rows(float* out, const float* table, const int64_t* idx, int B, int S, int T, void* stream) {
    ADP_REQUIRE(out && table && idx && B > 0 && S > 0 && S % 4 == 0 && T > 0 && al16(out) && al16(table),
                "anoddpm_gather_rows: bad arguments (S must be a multiple of 4)");
    gather_rows_kernel<<<adp_ceil_div((long long)B * (S / 4), 256), 256, 0, (cudaStream_t)stream>>>(out, table, (const long long*)idx,
                                                                                                  B, S / 4, T);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}

extern "C" int anoddpm_timestep_embedding(float* emb, const int64_t* t, const float* freq, int B, int dim,
                                          void* stream) {
    ADP_REQUIRE(emb && t && freq && B > 0 && dim > 0 && dim % 2 == 0, "anoddpm_timestep_embedding: bad arguments");
    const int n = B * (dim / 2);
    tembed_kernel<<<adp_ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(emb, (const long long*)t, freq, B, dim);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}

extern "C" int anoddpm_linear(float* y, int ldy, const float* x, int ldx, const float* w, const float* bias, int B,
                              int N, int K, int act, void* stream) {
    ADP_REQUIRE(y && x && w && B > 0 && N > 0 && K > 0 && act >= 0 && act <= 3, "anoddpm_linear: bad arguments");
    const long long warps = (long long)adp_ceil_div(B, LB) * N;
    linear_kernel<<<adp_ceil_div(warps * 32, 256), 256, 0, (cudaStream_t)stream>>>(y, ldy, x, ldx, w, bias, B, N, K, act);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}

static int gn_vec(int C1, int C2, int groups, const void* p1, const void* p2) {
    const int C = C1 + C2, gs = C / groups;
    (void)gs;
    return (C1 % 4 == 0 && C2 % 4 == 0 && al16(p1) && (!p2 || al16(p2))) ? 4 : 1;
}

extern "C" int anoddpm_gn_stats(double* stats, const float* x, int C, int B, int HW, void* stream) {
    ADP_REQUIRE(stats && x && B > 0 && B <= 65535 && HW > 0 && C > 0, "anoddpm_gn_stats: bad arguments");
    ADP_REQUIRE((size_t)C * 2 * sizeof(double) <= 48 * 1024, "anoddpm_gn_stats: too many channels (%d)", C);
    int ppc = 2048;  // pixels per CTA
    while (ppc > 64 && (long long)adp_ceil_div(HW, ppc) * B < 148 * 2) ppc >>= 1;
    dim3 grid(adp_ceil_div(HW, ppc), B);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)C * 2 * sizeof(double);
    if (C % 4 == 0 && al16(x))
        gn_stats_kernel<4><<<grid, 256, smem, st>>>(stats, x, C, HW, ppc);
    else
        gn_stats_kernel<1><<<grid, 256, smem, st>>>(stats, x, C, HW, ppc);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}

extern "C" int anoddpm_gn_apply(void* out_v, void* raw_out_v, const float* x1, int C1, const float* x2, int C2,
                                const double* stats1, const double* stats2, const float* gamma, const float* beta, int B, int H, int W,
                                int groups, float eps, int do_silu, int mode, int rtf32, void* stream) {
    float* out = static_cast<float*>(out_v);
    float* raw_out = static_cast<float*>(raw_out_v);
    ADP_REQUIRE(out && x1 && stats1 && gamma && beta && B > 0 && B <= 65535 && H > 0 && W > 0,
                "anoddpm_gn_apply: bad arguments");
    ADP_REQUIRE((C2 == 0) || stats2 != nullptr, "anoddpm_gn_apply: second source without statistics");
    ADP_REQUIRE((C2 == 0) == (x2 == nullptr), "anoddpm_gn_apply: x2/C2 mismatch");
    const int C = C1 + C2;
    ADP_REQUIRE(groups > 0 && groups * ANODDPM_STAT_REPLICAS <= 256 && C % groups == 0, "anoddpm_gn_apply: bad groups (at most 32)");
    ADP_REQUIRE(mode >= 0 && mode <= 2, "anoddpm_gn_apply: bad mode %d", mode);
    ADP_REQUIRE(mode != 1 || (H % 2 == 0 && W % 2 == 0), "anoddpm_gn_apply: avg-pool needs even H, W");
    ADP_REQUIRE(raw_out == nullptr || mode == 0 || (mode == 1 && C2 == 0), "anoddpm_gn_apply: raw_out needs mode 0, or mode 1 with one source");
    ADP_REQUIRE(rtf32 >= 0 && rtf32 <= 2, "anoddpm_gn_apply: out_mode must be 0 (fp32), 1 (tf32-rounded fp32) or 2 (fp16)");
    const int V = (gn_vec(C1, C2, groups, x1, x2) == 4 && al16(out) && (!raw_out || al16(raw_out))) ? 4 : 1;
    const int Ho = mode == 1 ? H / 2 : (mode == 2 ? H * 2 : H), Wo = mode == 1 ? W / 2 : (mode == 2 ? W * 2 : W);
    const int npo = Ho * Wo;
    const int npl = mode == 2 ? H * W : npo;      // pixels the kernel loops over (input pixels in the upsample mode)
    int wpc = 512;   // loop pixels per CTA (>= 6 waves at full resolution: small tail)
    while (wpc > 32 && (long long)adp_ceil_div(npl, wpc) * B < 148 * 4) wpc >>= 1;
    dim3 grid((unsigned)adp_ceil_div(npl, wpc), B);
    cudaStream_t st = (cudaStream_t)stream;
#define ADP_GN_APPLY(VV, NN)                                                                                       \
    gn_apply_kernel<VV, NN><<<grid, 256, 0, st>>>(out, raw_out, x1, C1, x2, C2, stats1, stats2, gamma, beta, H, W, groups, eps, \
                                                  do_silu, mode, rtf32, wpc)
    if (V == 4) { if (mode == 1) ADP_GN_APPLY(4, 4); else ADP_GN_APPLY(4, 1); }
    else { if (mode == 1) ADP_GN_APPLY(1, 4); else ADP_GN_APPLY(1, 1); }
#undef ADP_GN_APPLY
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}

extern "C" int anoddpm_gn_apply_f16in(void* out, const void* x_f16, int C, const double* stats, const float* gamma,
                                      const float* beta, const float* shift_c, const float* shift_bc, int shift_stride,
                                      int B, int H, int W, int groups, float eps, int do_silu, int out_mode, void* stream) {
    ADP_REQUIRE(out && x_f16 && stats && gamma && beta && B > 0 && B <= 65535 && H > 0 && W > 0,
                "anoddpm_gn_apply_f16in: bad arguments");
    ADP_REQUIRE(groups > 0 && groups * ANODDPM_STAT_REPLICAS <= 256 && C % groups == 0 && C % 8 == 0,
                "anoddpm_gn_apply_f16in: needs C %% 8 == 0 and at most 32 groups dividing C (C=%d groups=%d)", C, groups);
    ADP_REQUIRE(out_mode >= 0 && out_mode <= 2, "anoddpm_gn_apply_f16in: out_mode must be 0, 1 or 2");
    ADP_REQUIRE(al16(out) && al16(x_f16), "anoddpm_gn_apply_f16in: pointers must be 16-byte aligned");
    ADP_REQUIRE(!shift_bc || shift_stride >= C, "anoddpm_gn_apply_f16in: shift stride smaller than C");
    const int HW = H * W;
    int wpc = 512;
    while (wpc > 32 && (long long)adp_ceil_div(HW, wpc) * B < 148 * 4) wpc >>= 1;
    dim3 grid((unsigned)adp_ceil_div(HW, wpc), B);
    cudaStream_t st = (cudaStream_t)stream;
    const __half* x = static_cast<const __half*>(x_f16);
#define ADP_GN_H(OM) gn_apply_h_kernel<OM><<<grid, 256, 0, st>>>(out, x, C, stats, gamma, beta, shift_c, shift_bc, shift_stride, \
                                                                HW, groups, eps, do_silu, wpc)
    if (out_mode == 2) ADP_GN_H(2); else if (out_mode == 1) ADP_GN_H(1); else ADP_GN_H(0);
#undef ADP_GN_H
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}

extern "C" int anoddpm_gn_table(float* table, const double* stats1, int C1, const double* stats2, int C2,
                                const float* gamma, const float* beta, const float* shift_c, const float* shift_bc,
                                int shift_stride, int B, int HW, int groups, float eps, void* stream) {
    ADP_REQUIRE(table && stats1 && gamma && beta && B > 0 && HW > 0 && C1 > 0 && C2 >= 0, "anoddpm_gn_table: bad arguments");
    ADP_REQUIRE((C2 == 0) == (stats2 == nullptr), "anoddpm_gn_table: second source / channel mismatch");
    ADP_REQUIRE(groups > 0 && groups * ANODDPM_STAT_REPLICAS <= 256 && (C1 + C2) % groups == 0,
                "anoddpm_gn_table: at most 32 groups dividing C (C=%d groups=%d)", C1 + C2, groups);
    ADP_REQUIRE(!shift_bc || shift_stride >= C1 + C2, "anoddpm_gn_table: shift stride smaller than C");
    gn_table_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float2*>(table), stats1, C1, stats2, C2, gamma, beta,
                                                         shift_c, shift_bc, shift_stride, HW, groups, eps);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}

extern "C" int anoddpm_conv(const anoddpm_conv_desc* d, void* stream) {
    ADP_REQUIRE(d && d->out && d->a && d->w, "anoddpm_conv: null pointer");
    ADP_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->Cout > 0 && d->Ca > 0, "anoddpm_conv: bad shape");
    ADP_REQUIRE(d->ksize == 1 || d->ksize == 3, "anoddpm_conv: ksize must be 1 or 3");
    ADP_REQUIRE((d->Cr1 == 0) == (d->r1 == nullptr) && (d->Cr2 == 0) == (d->r2 == nullptr),
                "anoddpm_conv: skip source / channel mismatch");
    ADP_REQUIRE(d->Ca % 4 == 0 && d->Cr1 % 4 == 0 && d->Cr2 % 4 == 0 && d->Cout % 4 == 0,
                "anoddpm_conv: channel counts must be multiples of 4 (Ca=%d Cr1=%d Cr2=%d Cout=%d)", d->Ca, d->Cr1,
                d->Cr2, d->Cout);
    ADP_REQUIRE(d->res_mode >= 0 && d->res_mode <= 3 && ((d->res_mode == 0) == (d->res == nullptr)),
                "anoddpm_conv: residual mode / pointer mismatch");
    ADP_REQUIRE(d->res_mode != 3 || (d->H % 2 == 0 && d->W % 2 == 0), "anoddpm_conv: upsampled residual needs even H, W");
    ADP_REQUIRE(al16(d->out) && al16(d->a) && al16(d->w) && (!d->r1 || al16(d->r1)) && (!d->r2 || al16(d->r2)) &&
                    (!d->res || al16(d->res)),
                "anoddpm_conv: pointers must be 16-byte aligned");
    ADP_REQUIRE(!d->operand_f16 || d->impl == 2 || d->impl == 3, "anoddpm_conv: fp16 operands need impl 2 or 3");
    ADP_REQUIRE(!d->out_f16 || d->impl == 3, "anoddpm_conv: fp16 output needs impl 3");
    ADP_REQUIRE((d->Ca2 == 0) == (d->a2 == nullptr) && d->Ca2 >= 0, "anoddpm_conv: second main source / channel mismatch");
    ADP_REQUIRE((!d->a2 && !d->gn_table && !d->out16) || d->impl == 3,
                "anoddpm_conv: a2 / gn_table / out16 (operand-side GroupNorm) need impl 3");
    if (d->impl == 3) return anoddpm_conv_tc3(d, stream);
    ADP_REQUIRE(d->gn_stats == nullptr || d->impl == 2, "anoddpm_conv: fused GroupNorm statistics need impl 2 or 3");
    if (d->impl == 1) return anoddpm_conv_tc(d, stream);
    if (d->impl == 2) return anoddpm_conv_tc2(d, stream);
    ADP_REQUIRE(d->impl == 0, "anoddpm_conv: unknown impl %d", d->impl);
    ConvArgs A;
    A.d = *d;
    A.M = d->B * d->H * d->W;
    A.Ktot = d->ksize * d->ksize * d->Ca + d->Cr1 + d->Cr2;
    dim3 grid(adp_ceil_div(A.M, CBM), adp_ceil_div(d->Cout, CBN));
    conv_simt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}

extern "C" int anoddpm_conv_stem(float* out, const float* x, const float* w, const float* bias, int B, int Cin, int H,
                                 int W, int Cout, void* stream) {
    ADP_REQUIRE(out && x && w && B > 0 && Cin > 0 && H > 0 && W > 0 && Cout > 0, "anoddpm_conv_stem: bad arguments");
    ADP_REQUIRE(Cout % 4 == 0 && Cout <= 1024 && al16(out), "anoddpm_conv_stem: Cout must be a multiple of 4 (<= 1024)");
    ADP_REQUIRE((long long)B * H * W < (1LL << 31), "anoddpm_conv_stem: too many pixels");
    const size_t smem = ((size_t)Cout * Cin * 9 + Cout) * sizeof(float);
    ADP_REQUIRE(smem <= 48 * 1024, "anoddpm_conv_stem: weights (%zu B) exceed shared memory", smem);
    const int ppb = 256 / (Cout / 4) > 0 ? 256 / (Cout / 4) : 1;
    int blocks = adp_ceil_div((long long)B * H * W, ppb);
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (W % 4 == 0) {
        int blocks4 = adp_ceil_div((long long)B * H * (W / 4), ppb);
        if (blocks4 > 148 * 16) blocks4 = 148 * 16;
        conv_stem4_kernel<<<blocks4, 256, smem, (cudaStream_t)stream>>>(out, x, w, bias, B, Cin, H, W, Cout);
    } else {
        conv_stem_kernel<<<blocks, 256, smem, (cudaStream_t)stream>>>(out, x, w, bias, B, Cin, H, W, Cout);
    }
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}

extern "C" int anoddpm_conv_head(float* out, const float* a, const float* w, const float* bias, int B, int Cin, int H,
                                 int W, int Cout, void* stream) {
    ADP_REQUIRE(out && a && w && B > 0 && Cin > 0 && H > 0 && W > 0, "anoddpm_conv_head: bad arguments");
    ADP_REQUIRE(Cout >= 1 && Cout <= 4, "anoddpm_conv_head: Cout must be 1..4 (got %d)", Cout);
    ADP_REQUIRE(Cin % 4 == 0 && al16(a), "anoddpm_conv_head: Cin must be a multiple of 4");
    const long long tiles = (long long)B * adp_ceil_div(H, HT_H) * adp_ceil_div(W, HT_W);
    ADP_REQUIRE(tiles < (1LL << 31), "anoddpm_conv_head: too many tiles");
    cudaStream_t st = (cudaStream_t)stream;
    switch (Cout) {
        case 1: conv_head_kernel<1, false><<<(unsigned)tiles, 256, 0, st>>>(out, a, w, bias, B, Cin, H, W, nullptr, nullptr, nullptr, 1, 0.f); break;
        case 2: conv_head_kernel<2, false><<<(unsigned)tiles, 256, 0, st>>>(out, a, w, bias, B, Cin, H, W, nullptr, nullptr, nullptr, 1, 0.f); break;
        case 3: conv_head_kernel<3, false><<<(unsigned)tiles, 256, 0, st>>>(out, a, w, bias, B, Cin, H, W, nullptr, nullptr, nullptr, 1, 0.f); break;
        default: conv_head_kernel<4, false><<<(unsigned)tiles, 256, 0, st>>>(out, a, w, bias, B, Cin, H, W, nullptr, nullptr, nullptr, 1, 0.f); break;
    }
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}

extern "C" int anoddpm_conv_head_gn(float* out, const float* x, const double* stats, const float* gamma, const float* beta,
                                    const float* w, const float* bias, int B, int Cin, int H, int W, int Cout, int groups,
                                    float eps, void* stream) {
    ADP_REQUIRE(out && x && stats && gamma && beta && w && B > 0 && Cin > 0 && H > 0 && W > 0,
                "anoddpm_conv_head_gn: bad arguments");
    ADP_REQUIRE(Cout >= 1 && Cout <= 4, "anoddpm_conv_head_gn: Cout must be 1..4 (got %d)", Cout);
    ADP_REQUIRE(Cin % 4 == 0 && Cin <= HT_MAXC && al16(x), "anoddpm_conv_head_gn: Cin must be a multiple of 4, at most %d",
                HT_MAXC);
    ADP_REQUIRE(groups > 0 && groups * ANODDPM_STAT_REPLICAS <= 256 && Cin % groups == 0,
                "anoddpm_conv_head_gn: bad groups (at most 32, dividing Cin)");
    const long long tiles = (long long)B * adp_ceil_div(H, HT_H) * adp_ceil_div(W, HT_W);
    ADP_REQUIRE(tiles < (1LL << 31), "anoddpm_conv_head_gn: too many tiles");
    cudaStream_t st = (cudaStream_t)stream;
    switch (Cout) {
        case 1: conv_head_kernel<1, true><<<(unsigned)tiles, 256, 0, st>>>(out, x, w, bias, B, Cin, H, W, stats, gamma, beta, groups, eps); break;
        case 2: conv_head_kernel<2, true><<<(unsigned)tiles, 256, 0, st>>>(out, x, w, bias, B, Cin, H, W, stats, gamma, beta, groups, eps); break;
        case 3: conv_head_kernel<3, true><<<(unsigned)tiles, 256, 0, st>>>(out, x, w, bias, B, Cin, H, W, stats, gamma, beta, groups, eps); break;
        default: conv_head_kernel<4, true><<<(unsigned)tiles, 256, 0, st>>>(out, x, w, bias, B, Cin, H, W, stats, gamma, beta, groups, eps); break;
    }
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}

extern "C" int anoddpm_attention(void* out_v, const float* qkv, int B, int T, int C, int heads, int rtf32,
                                 void* stream) {
    float* out = static_cast<float*>(out_v);
    ADP_REQUIRE(out && qkv && B > 0 && T > 0 && C > 0 && heads > 0 && C % heads == 0, "anoddpm_attention: bad arguments");
    ADP_REQUIRE(B <= 65535 && heads <= 65535, "anoddpm_attention: batch/heads too large");
    const int ch = C / heads;
    // queries per CTA: as many as fit in shared memory while the grid still covers the GPU
    // (16 keeps >= 4 CTAs resident per SM and the grid balanced for T <= 256; 32 only pays for long sequences)
    int qb = T >= 1024 ? 32 : 16;
    while (qb > 8 && ((size_t)qb * (ch + T) * sizeof(float) > 200 * 1024 ||
                      (long long)adp_ceil_div(T, qb) * heads * B < 2 * 148)) qb >>= 1;
    const size_t smem = (size_t)qb * (ch + T) * sizeof(float);
    ADP_REQUIRE(smem <= 200 * 1024, "anoddpm_attention: T=%d ch=%d needs %zu B of shared memory", T, ch, smem);
    static bool attr_dev[64] = {};   // the attribute is per device
    int dev_ = 0;
    ADP_CUDA(cudaGetDevice(&dev_));
    if (!attr_dev[dev_ & 63]) {
        ADP_CUDA(cudaFuncSetAttribute(attention_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ADP_CUDA(cudaFuncSetAttribute(attention_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        ADP_CUDA(cudaFuncSetAttribute(attention_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_dev[dev_ & 63] = true;
    }
    dim3 grid(adp_ceil_div(T, qb), heads, B);
    cudaStream_t st = (cudaStream_t)stream;
    if (qb == 32) attention_kernel<32><<<grid, 256, smem, st>>>(out, qkv, T, C, heads, rtf32);
    else if (qb == 16) attention_kernel<16><<<grid, 256, smem, st>>>(out, qkv, T, C, heads, rtf32);
    else attention_kernel<8><<<grid, 256, smem, st>>>(out, qkv, T, C, heads, rtf32);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}
