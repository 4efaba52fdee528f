// anoddpm_b200/csrc/edge_convs.cu -- the two convolutions of the UNet that are not GEMM-shaped (reference UNet.py:279-281
// stem Conv2d(1 -> C, 3x3), UNet.py:384-388 head GroupNorm -> SiLU -> Conv2d(C -> 1, 3x3)), second generation: both are pure
// HBM streams over a (B, H, W, C) fp32 tensor and are written to run as such.
//   stem: out[b,p,:] = sum_tap x[b,p+tap] * w[:,tap] + bias.  One pass that ALSO emits what the first ResBlock needs: the
//         per-channel GroupNorm statistics of the output (fp64 atomics into the replica layout) and an fp16 copy (the raw
//         operand of the first conv with operand-side GroupNorm) -- so the stem output is never re-read by a statistics or
//         normalisation pass.
//   head: out[b,p] = bias + sum_tap sum_c act(GN(x))[b,p+tap,c] * w[tap][c].  Cout = 1 makes the conv SEPARABLE BY TAP:
//         phase 1 normalises + activates every input pixel of an 18x18 halo tile once and reduces its C channels against the
//         nine tap weight vectors (9 dot products per pixel, weights reused across 4 pixels in registers); phase 2 adds the nine
//         shifted partial maps.  The feature map is read once (plus the halo from L2).
#include "common.cuh"
#include "anoddpm_b200.h"
#include <cuda_fp16.h>

namespace {

struct alignas(16) F4e { float x, y, z, w; };
__device__ __forceinline__ F4e ld4e(const float* p) { return *reinterpret_cast<const F4e*>(p); }
__device__ __forceinline__ float silu_e(float y) { return __fdividef(y, 1.0f + __expf(-y)); }

// ------------------------------------------------------------------------------------------------ head (Cout = 1)
constexpr int HD_T = 16;                  // 16 x 16 output pixels per CTA
constexpr int HD_IN = HD_T + 2;           // 18 x 18 input pixels
constexpr int HD_NPIX = HD_IN * HD_IN;    // 324

__global__ void __launch_bounds__(256) conv_head_sep_kernel(float* __restrict__ out, const float* __restrict__ x,
                                                            const double* __restrict__ stats, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, const float* __restrict__ w,
                                                            const float* __restrict__ bias, int B, int C, int H, int W,
                                                            int groups, float eps) {
    extern __shared__ __align__(16) float sm_head[];
    float* s_w = sm_head;                 // [9][C]
    float* s_sc = s_w + 9 * C;            // [C]
    float* s_bi = s_sc + C;               // [C]
    float* s_t = s_bi + C;                // [9][324] partial maps
    __shared__ float s_mean[64], s_rstd[64];
    const int tiles_x = W / HD_T, tiles_y = H / HD_T;
    const int tile = blockIdx.x;
    const int b = tile / (tiles_x * tiles_y), tr = tile - b * tiles_x * tiles_y;
    const int y0 = (tr / tiles_x) * HD_T, x0 = (tr % tiles_x) * HD_T;
    {   // group statistics exactly as gn_apply_kernel: thread (g, r) adds up replica r of group g
        const int gs = C / groups;
        const int g = threadIdx.x / ANODDPM_STAT_REPLICAS, r = threadIdx.x % ANODDPM_STAT_REPLICAS;
        double s = 0.0, q = 0.0;
        if (g < groups) {
            for (int c = g * gs; c < (g + 1) * gs; ++c) {
                const double* sp = stats + (size_t)r * B * C * 2 + ((size_t)b * C + c) * 2;
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
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const int gg = c / gs;
            const float sc = s_rstd[gg] * gamma[c];
            s_sc[c] = sc;
            s_bi[c] = beta[c] - sc * s_mean[gg];
        }
        for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) s_w[i] = w[i];      // packed [tap][c]
    }
    __syncthreads();
    // phase 1: item = (input pixel, 32-channel slice); a thread keeps its slice and walks pixels four at a time
    const int nsl = C / 32;                        // slices (C % 32 == 0); adjacent threads = adjacent slices of one pixel
    const int sl = threadIdx.x % nsl, lane_px = threadIdx.x / nsl, npl = 256 / nsl;
    const unsigned full = 0xffffffffu;
    const int n_it = (HD_NPIX + 4 * npl - 1) / (4 * npl);      // the same trip count for every thread (warp shuffles below)
    for (int it = 0; it < n_it; ++it) {
        const int p0 = lane_px + it * 4 * npl;
        float acc[4][9];
        const float* src[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int p = p0 + u * npl;
            const int ly = p / HD_IN, lx = p - ly * HD_IN;
            const int gy = y0 + ly - 1, gx = x0 + lx - 1;
            ok[u] = p < HD_NPIX && gy >= 0 && gy < H && gx >= 0 && gx < W;
            src[u] = x + (((size_t)b * H + (ok[u] ? gy : 0)) * W + (ok[u] ? gx : 0)) * C + sl * 32;
#pragma unroll
            for (int t = 0; t < 9; ++t) acc[u][t] = 0.f;
        }
#pragma unroll 2
        for (int q = 0; q < 8; ++q) {
            const int c = sl * 32 + q * 4;
            const F4e sc = ld4e(s_sc + c), bi = ld4e(s_bi + c);
            F4e v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                v[u] = ok[u] ? ld4e(src[u] + q * 4) : F4e{0.f, 0.f, 0.f, 0.f};
                v[u].x = silu_e(v[u].x * sc.x + bi.x);
                v[u].y = silu_e(v[u].y * sc.y + bi.y);
                v[u].z = silu_e(v[u].z * sc.z + bi.z);
                v[u].w = silu_e(v[u].w * sc.w + bi.w);
            }
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const F4e ww = ld4e(s_w + t * C + c);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    acc[u][t] = fmaf(v[u].x, ww.x, acc[u][t]);
                    acc[u][t] = fmaf(v[u].y, ww.y, acc[u][t]);
                    acc[u][t] = fmaf(v[u].z, ww.z, acc[u][t]);
                    acc[u][t] = fmaf(v[u].w, ww.w, acc[u][t]);
                }
            }
        }
        // add up the slices of each pixel (adjacent lanes), slice 0 stores the nine partial sums; out-of-image pixels are
        // the zero padding of the ACTIVATED tensor
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                float a = ok[u] ? acc[u][t] : 0.f;
                for (int o = 1; o < nsl; o <<= 1) a += __shfl_xor_sync(full, a, o);
                acc[u][t] = a;
            }
            const int p = p0 + u * npl;
            if (sl == 0 && p < HD_NPIX) {
#pragma unroll
                for (int t = 0; t < 9; ++t) s_t[t * HD_NPIX + p] = acc[u][t];
            }
        }
    }
    __syncthreads();
    // phase 2: thread = output pixel; tap (ky, kx) reads the partial map of tap at the shifted position
    const int py = threadIdx.x / HD_T, px = threadIdx.x % HD_T;
    float o = bias ? bias[0] : 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) o += s_t[t * HD_NPIX + (py + t / 3) * HD_IN + px + t % 3];
    out[((size_t)b * H + y0 + py) * W + x0 + px] = o;
}

// ------------------------------------------------------------------------------------------------ stem (Cin = 1)
// thread = (4 output channels, pixel lane); a CTA walks 8-pixel row segments; per-channel statistics are kept per thread and
// reduced once per CTA.  grid-stride over (b, y, x-segment) with all segments of a CTA belonging to one sample.
__global__ void __launch_bounds__(256) conv_stem_fused_kernel(float* __restrict__ out, __half* __restrict__ out16,
                                                              double* __restrict__ stats, const float* __restrict__ x,
                                                              const float* __restrict__ w, const float* __restrict__ bias, int B,
                                                              int H, int W, int Cout, int segs_per_cta) {
    extern __shared__ __align__(16) float sm_stem[];     // [9][Cout] weights (tap-major), [Cout] bias, then reduction scratch
    float* s_w = sm_stem;
    float* s_b = s_w + 9 * Cout;
    float* s_red = s_b + Cout;                            // [256][8]
    for (int i = threadIdx.x; i < 9 * Cout; i += blockDim.x) {
        const int co = i / 9, t = i - co * 9;
        s_w[t * Cout + co] = w[i];                         // w is [Cout][1][3][3]
    }
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) s_b[i] = bias ? bias[i] : 0.f;
    __syncthreads();
    const int nc4 = Cout >> 2, lanes = 256 / nc4;          // pixel lanes (Cout = 128: 8 lanes)
    const int pl = threadIdx.x / nc4, c4 = (threadIdx.x - pl * nc4) * 4;
    const int segs_x = W / 8, segs_img = segs_x * H;
    const long long seg0 = (long long)blockIdx.x * segs_per_cta;
    const int b = (int)(seg0 / segs_img);                  // segs_img % segs_per_cta == 0: one sample per CTA
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    F4e wv[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) wv[t] = ld4e(s_w + t * Cout + c4);
    const F4e bv = ld4e(s_b + c4);
    if (pl < lanes) {
        for (int s = 0; s < segs_per_cta; ++s) {
            const int sg = (int)((seg0 + s) % segs_img);
            const int y = sg / segs_x, xs = (sg - y * segs_x) * 8;
            // this thread's pixel of the segment: lanes pixels in parallel, 8 / lanes rounds
            for (int px = pl; px < 8; px += lanes) {
                const int xx = xs + px;
                float win[9];
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const int yy = y + t / 3 - 1, xq = xx + t % 3 - 1;
                    win[t] = (yy >= 0 && yy < H && xq >= 0 && xq < W) ? __ldg(x + ((size_t)b * H + yy) * W + xq) : 0.f;
                }
                F4e a = bv;
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    a.x = fmaf(win[t], wv[t].x, a.x);
                    a.y = fmaf(win[t], wv[t].y, a.y);
                    a.z = fmaf(win[t], wv[t].z, a.z);
                    a.w = fmaf(win[t], wv[t].w, a.w);
                }
                const size_t o = (((size_t)b * H + y) * W + xx) * Cout + c4;
                *reinterpret_cast<F4e*>(out + o) = a;
                if (out16) {
                    const __half2 h0 = __floats2half2_rn(fminf(fmaxf(a.x, -65504.f), 65504.f), fminf(fmaxf(a.y, -65504.f), 65504.f));
                    const __half2 h1 = __floats2half2_rn(fminf(fmaxf(a.z, -65504.f), 65504.f), fminf(fmaxf(a.w, -65504.f), 65504.f));
                    uint2 u;
                    u.x = *reinterpret_cast<const uint32_t*>(&h0);
                    u.y = *reinterpret_cast<const uint32_t*>(&h1);
                    *reinterpret_cast<uint2*>(out16 + o) = u;
                }
                s1[0] += a.x; s1[1] += a.y; s1[2] += a.z; s1[3] += a.w;
                s2[0] += a.x * a.x; s2[1] += a.y * a.y; s2[2] += a.z * a.z; s2[3] += a.w * a.w;
            }
        }
    }
    if (stats) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { s_red[threadIdx.x * 8 + j] = s1[j]; s_red[threadIdx.x * 8 + 4 + j] = s2[j]; }
        __syncthreads();
        if (pl == 0) {
            double t1[4] = {0, 0, 0, 0}, t2[4] = {0, 0, 0, 0};
            for (int l = 0; l < lanes; ++l)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    t1[j] += (double)s_red[(l * nc4 + threadIdx.x) * 8 + j];
                    t2[j] += (double)s_red[(l * nc4 + threadIdx.x) * 8 + 4 + j];
                }
            double* st = stats + (size_t)(blockIdx.x % ANODDPM_STAT_REPLICAS) * B * Cout * 2 + ((size_t)b * Cout + c4) * 2;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                atomicAdd(st + 2 * j, t1[j]);
                atomicAdd(st + 2 * j + 1, t2[j]);
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------ input pipeline (N4)
// The deterministic part of the reference's MRI transform (dataset.py:585-593: ToPILImage -> CenterCrop(crop) -> Resize(bilinear)
// -> ToTensor -> Normalize(0.5, 0.5)) for a batch of slices read straight out of cached volumes, BIT-EXACT:
//   * ToPILImage turns a float32 array into an 8-bit 'L' image: q = uint8(trunc(v * 255));
//   * CenterCrop pads with 0 to at least crop x crop (left / top = floor(d / 2)) and crops at round-half-even((size - crop) / 2);
//   * Resize = PIL's 8-bit resampler: triangle filter, half-pixel centres, support scaled by the downscale factor
//     (antialiasing), coefficients normalised, quantised to 22-bit fixed point, horizontal pass to uint8, then vertical pass:
//     out = clip8((2^21 + sum pix * kk) >> 22);
//   * ToTensor (/ 255) and Normalize: (v - 0.5) / 0.5 in fp32.
// thread = output pixel; slice b = rows of `row_stride` elements starting at vols + vol_offset[b].
constexpr int RS_MAXT = 32;
__device__ __forceinline__ int resample_coeffs(int out_i, int in_size, int out_size, int* kk, int& xmin_out) {
    const double scale = (double)in_size / out_size;
    const double fs = scale > 1.0 ? scale : 1.0, support = fs, ss = 1.0 / fs;
    const double center = (out_i + 0.5) * scale;
    int xmin = (int)(center - support + 0.5), xmax = (int)(center + support + 0.5);
    if (xmin < 0) xmin = 0;
    if (xmax > in_size) xmax = in_size;
    const int n = xmax - xmin;
    double k[RS_MAXT], ww = 0.0;
    for (int x = 0; x < n; ++x) {
        double w = 1.0 - fabs((x + xmin - center + 0.5) * ss);
        if (w < 0.0) w = 0.0;
        k[x] = w;
        ww += w;
    }
    for (int x = 0; x < n; ++x) kk[x] = (int)(0.5 + (k[x] / ww) * 4194304.0);      // 1 << 22
    xmin_out = xmin;
    return n;
}
__global__ void __launch_bounds__(128) slice_resize_kernel(float* __restrict__ out, const float* __restrict__ vols,
                                                           const long long* __restrict__ vol_offset, int B, int Hs, int Ws,
                                                           long long row_stride, int crop, int Ho, int Wo) {
    const long long idx = (long long)blockIdx.x * 128 + threadIdx.x;
    if (idx >= (long long)B * Ho * Wo) return;
    const int ox = (int)(idx % Wo), oy = (int)((idx / Wo) % Ho), b = (int)(idx / ((long long)Wo * Ho));
    const float* src = vols + vol_offset[b];
    const int pad_l = Ws < crop ? (crop - Ws) / 2 : 0, pad_t = Hs < crop ? (crop - Hs) / 2 : 0;
    const int Wp = Ws < crop ? crop : Ws, Hp = Hs < crop ? crop : Hs;
    const int crop_l = (int)rint((Wp - crop) / 2.0), crop_t = (int)rint((Hp - crop) / 2.0);
    int kx[RS_MAXT], ky[RS_MAXT], x0, y0;
    const int nx = resample_coeffs(ox, crop, Wo, kx, x0), ny = resample_coeffs(oy, crop, Ho, ky, y0);
    long long acc = 1LL << 21;
    for (int j = 0; j < ny; ++j) {
        const int yy = y0 + j + crop_t - pad_t;                   // row of the original slice (outside: CenterCrop's zero padding)
        long long h = 1LL << 21;
        if (yy >= 0 && yy < Hs) {
            for (int i = 0; i < nx; ++i) {
                const int xx = x0 + i + crop_l - pad_l;
                if (xx >= 0 && xx < Ws) {
                    const float v = src[(long long)yy * row_stride + xx] * 255.0f;
                    h += (long long)(unsigned char)(int)v * kx[i];          // uint8(trunc(v * 255))
                }
            }
        }
        long long hq = h >> 22;                                    // horizontal pass result, clipped to 8 bits
        hq = hq < 0 ? 0 : (hq > 255 ? 255 : hq);
        acc += hq * ky[j];
    }
    long long q = acc >> 22;
    q = q < 0 ? 0 : (q > 255 ? 255 : q);
    out[idx] = ((float)q / 255.0f - 0.5f) / 0.5f;
}

}  // namespace

extern "C" int anoddpm_conv_head_gn_supported(int C, int H, int W, int Cout) {
    return Cout == 1 && C % 32 == 0 && C <= 512 && (C / 32 == 1 || C / 32 == 2 || C / 32 == 4 || C / 32 == 8 || C / 32 == 16) &&
           H % 16 == 0 && W % 16 == 0;
}

// the separable head: same contract as anoddpm_conv_head_gn, for Cout = 1 and 16-pixel-aligned images
int anoddpm_conv_head_sep(float* out, const float* x, const double* stats, const float* gamma, const float* beta, const float* w,
                          const float* bias, int B, int C, int H, int W, int groups, float eps, void* stream) {
    const long long tiles = (long long)B * (H / HD_T) * (W / HD_T);
    ADP_REQUIRE(tiles < (1LL << 31), "anoddpm_conv_head_gn: too many tiles");
    const size_t smem = ((size_t)11 * C + 9 * HD_NPIX) * sizeof(float);
    static bool attr_dev[64] = {};
    int dev_ = 0;
    ADP_CUDA(cudaGetDevice(&dev_));
    if (!attr_dev[dev_ & 63]) {
        ADP_CUDA(cudaFuncSetAttribute(conv_head_sep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr_dev[dev_ & 63] = true;
    }
    conv_head_sep_kernel<<<(unsigned)tiles, 256, smem, (cudaStream_t)stream>>>(out, x, stats, gamma, beta, w, bias, B, C, H, W, groups, eps);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}

extern "C" int anoddpm_conv_stem_fused(float* out_nhwc, void* out16_nhwc, double* stats, const float* x_nchw, const float* w_oihw,
                                       const float* bias, int B, int Cin, int H, int W, int Cout, void* stream) {
    ADP_REQUIRE(out_nhwc && x_nchw && w_oihw && B > 0 && H > 0 && W > 0, "anoddpm_conv_stem_fused: bad arguments");
    ADP_REQUIRE(Cin == 1 && Cout % 4 == 0 && Cout <= 1024 && 256 % (Cout / 4) == 0 && W % 8 == 0,
                "anoddpm_conv_stem_fused: needs Cin == 1, Cout/4 dividing 256, W %% 8 == 0 (Cin=%d Cout=%d W=%d)", Cin, Cout, W);
    const int segs_img = (W / 8) * H;
    int spc = 64;                                   // segments per CTA; must divide the segments of one image
    while (spc > 1 && (segs_img % spc || (long long)B * segs_img / spc < 148 * 8)) spc >>= 1;
    const long long ctas = (long long)B * segs_img / spc;
    const size_t smem = ((size_t)10 * Cout + 256 * 8) * sizeof(float);
    conv_stem_fused_kernel<<<(unsigned)ctas, 256, smem, (cudaStream_t)stream>>>(out_nhwc, static_cast<__half*>(out16_nhwc), stats,
                                                                                 x_nchw, w_oihw, bias, B, H, W, Cout, spc);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}

extern "C" int anoddpm_slice_resize_norm(float* out, const float* vols, const int64_t* vol_offset, int B, int Hs, int Ws,
                                         int64_t row_stride, int crop, int Ho, int Wo, void* stream) {
    ADP_REQUIRE(out && vols && vol_offset && B > 0 && Hs > 0 && Ws > 0 && crop > 0 && Ho > 0 && Wo > 0 && row_stride >= Ws,
                "anoddpm_slice_resize_norm: bad arguments");
    const double sc = (double)crop / (Ho < Wo ? Ho : Wo);
    ADP_REQUIRE(2.0 * (sc > 1.0 ? sc : 1.0) + 2.0 <= RS_MAXT, "anoddpm_slice_resize_norm: downscale factor %.1f needs more than %d taps", sc, RS_MAXT);
    const long long n = (long long)B * Ho * Wo;
    slice_resize_kernel<<<adp_ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(out, vols, (const long long*)vol_offset, B, Hs, Ws,
                                                                                row_stride, crop, Ho, Wo);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}
