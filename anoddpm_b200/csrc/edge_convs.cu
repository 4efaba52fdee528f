// anoddpm_b200/csrc/edge_convs.cu -- edge kernels of the path: the fused stem and the device input pipeline.
//   stem (reference UNet.py:279-281, Conv2d(1 -> C, 3x3)): out[b,p,:] = sum_tap x[b,p+tap] * w[:,tap] + bias in one pass that
//         ALSO emits what the first ResBlock needs: the per-channel GroupNorm statistics of the output (fp64 atomics into the
//         replica layout) and an fp16 copy (the raw operand of the first conv with operand-side GroupNorm) -- so the stem
//         output is never re-read by a statistics or normalisation pass.
//   head (reference UNet.py:384-388, GroupNorm -> SiLU -> Conv2d(C -> 1, 3x3)): Cout = 1 makes the convolution separable by
//         tap.  Pass 1 (head_proj_kernel) streams the last feature map ONCE, with no halo: per pixel it normalises + activates
//         the C channels and reduces them against the nine tap weight vectors (lane = 4 channels, weights in registers, eight
//         pixels per warp iteration, reduce-scatter butterfly over the lanes) -> nine fp32 maps t[k][b][y][x] (L2-resident,
//         9 x 4 B per pixel against 4 C B read).  Pass 2 (head_gather_kernel) adds the nine shifted maps.  (A first
//         tap-separable attempt with 18x18 halo tiles and a 32-channel slice per thread was measured slower than unet_ops.cu's
//         conv_head_kernel -- 63 % lane utilisation, 27 % halo re-reads -- and dropped; this one has neither.)
#include "common.cuh"
#include "anoddpm_b200.h"
#include <cuda_fp16.h>

namespace {

struct alignas(16) F4e { float x, y, z, w; };
__device__ __forceinline__ F4e ld4e(const float* p) { return *reinterpret_cast<const F4e*>(p); }
__device__ __forceinline__ float silu_e(float y) { return __fdividef(y, 1.0f + __expf(-y)); }


// ------------------------------------------------------------------------------------------------ head (Cout = 1), two passes
// pass 1: t[k][pix] = sum_c SiLU(x[pix][c] * scale[b][c] + shift[b][c]) * w[k][c];  pix = linear (b, y, x) index.
// CPL = float4 slices per lane (C = 128 * CPL); PB = pixels per warp iteration (8 for CPL 1, 4 for CPL 2).
template <int CPL, int PB>
__global__ void __launch_bounds__(256) head_proj_kernel(float* __restrict__ t, const float* __restrict__ x,
                                                        const float2* __restrict__ tab, const float* __restrict__ w,
                                                        long long HW, long long npix, int pix_per_cta) {
    constexpr int C = 128 * CPL;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long cta0 = (long long)blockIdx.x * pix_per_cta;      // HW % pix_per_cta == 0: one sample per CTA
    const int b = (int)(cta0 / HW);
    float wk[CPL][9][4], sc[CPL][4], sh[CPL][4];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
        const int c = 128 * j + 4 * lane;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const F4e v = ld4e(w + k * C + c);
            wk[j][k][0] = v.x; wk[j][k][1] = v.y; wk[j][k][2] = v.z; wk[j][k][3] = v.w;
        }
        const F4e t0 = ld4e(reinterpret_cast<const float*>(tab + (size_t)b * C + c));
        const F4e t1 = ld4e(reinterpret_cast<const float*>(tab + (size_t)b * C + c + 2));
        sc[j][0] = t0.x; sh[j][0] = t0.y; sc[j][1] = t0.z; sh[j][1] = t0.w;
        sc[j][2] = t1.x; sh[j][2] = t1.y; sc[j][3] = t1.z; sh[j][3] = t1.w;
    }
    const int per_warp = pix_per_cta / 8;
    const long long p_begin = cta0 + (long long)warp * per_warp;
    for (int i0 = 0; i0 < per_warp; i0 += PB) {
        const long long p0 = p_begin + i0;
        F4e v[PB][CPL];
#pragma unroll
        for (int u = 0; u < PB; ++u)
#pragma unroll
            for (int j = 0; j < CPL; ++j)
                v[u][j] = (p0 + u < npix) ? ld4e(x + (size_t)(p0 + u) * C + 128 * j + 4 * lane) : F4e{0.f, 0.f, 0.f, 0.f};
        float acc[PB][9];
#pragma unroll
        for (int u = 0; u < PB; ++u) {
#pragma unroll
            for (int k = 0; k < 9; ++k) acc[u][k] = 0.f;
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                const float a0 = silu_e(fmaf(v[u][j].x, sc[j][0], sh[j][0])), a1 = silu_e(fmaf(v[u][j].y, sc[j][1], sh[j][1]));
                const float a2 = silu_e(fmaf(v[u][j].z, sc[j][2], sh[j][2])), a3 = silu_e(fmaf(v[u][j].w, sc[j][3], sh[j][3]));
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    acc[u][k] = fmaf(a0, wk[j][k][0], acc[u][k]);
                    acc[u][k] = fmaf(a1, wk[j][k][1], acc[u][k]);
                    acc[u][k] = fmaf(a2, wk[j][k][2], acc[u][k]);
                    acc[u][k] = fmaf(a3, wk[j][k][3], acc[u][k]);
                }
            }
        }
        // reduce-scatter over the lanes: each halving level keeps one half of the pixels and adds the partner's other half
        // (lane bit 4 <-> pixel bit log2(PB)-1, ...), then the remaining levels all-reduce the nine sums of the lane's pixel
        int mask = 16;
#pragma unroll
        for (int half = PB / 2; half >= 1; half >>= 1, mask >>= 1) {
            const bool up = (lane & mask) != 0;
#pragma unroll
            for (int i = 0; i < half; ++i)
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const float send = up ? acc[i][k] : acc[i + half][k];
                    const float keep = up ? acc[i + half][k] : acc[i][k];
                    acc[i][k] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
                }
        }
#pragma unroll
        for (; mask >= 1; mask >>= 1)
#pragma unroll
            for (int k = 0; k < 9; ++k) acc[0][k] += __shfl_xor_sync(0xffffffffu, acc[0][k], mask);
        // the 32 / PB lanes of a pixel hold the same nine sums: they share the nine stores
        constexpr int LPP = 32 / PB;                        // lanes per pixel
        const int u = lane / LPP, jl = lane % LPP;
        const long long p = p0 + u;
        if (p < npix) {
#pragma unroll
            for (int k = 0; k < 9; ++k)
                if (k % LPP == jl) t[(size_t)k * npix + p] = acc[0][k];
        }
    }
}

// pass 2: out[b][y][x] = bias + sum over the nine taps of t[tap][b][y + dy][x + dx] (taps outside the image contribute zero:
// the convolution pads the ACTIVATED tensor)
__global__ void __launch_bounds__(256) head_gather_kernel(float* __restrict__ out, const float* __restrict__ t,
                                                          const float* __restrict__ bias, int H, int W, long long npix) {
    const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
    if (p >= npix) return;
    const int xq = (int)(p % W), yq = (int)((p / W) % H);
    float o = bias ? bias[0] : 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int dy = k / 3 - 1, dx = k % 3 - 1;
        if ((unsigned)(yq + dy) < (unsigned)H && (unsigned)(xq + dx) < (unsigned)W) o += t[(size_t)k * npix + p + (long long)dy * W + dx];
    }
    out[p] = o;
}

// ------------------------------------------------------------------------------------------------ stem (Cin = 1)
// thread = (4 output channels, pixel lane) x PPT consecutive pixels; a CTA walks (8 PPT)-pixel row segments; per-channel statistics
// are kept per thread and reduced once per CTA.  All segments of a CTA belong to one sample.  PPT = 4 keeps four pixels' stores
// (3 KB per warp) in flight per iteration of the load -> FMA -> store chain: one pixel per iteration left the kernel latency-bound
// at 1.85 TB/s.
template <int PPT>
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
    constexpr int SEG = 8 * PPT;                           // pixels per segment
    const int nc4 = Cout >> 2, lanes = 256 / nc4;          // pixel lanes (Cout = 128: 8 lanes)
    const int pl = threadIdx.x / nc4, c4 = (threadIdx.x - pl * nc4) * 4;
    const int segs_x = W / SEG, segs_img = segs_x * H;
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
            const int y = sg / segs_x, xs = (sg - y * segs_x) * SEG;
            // this thread's PPT consecutive pixels of the segment
            for (int px = pl * PPT; px < SEG; px += lanes * PPT) {
                const int xx = xs + px;
                float win[3][PPT + 2];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const int yy = y + r - 1;
#pragma unroll
                    for (int c = 0; c < PPT + 2; ++c) {
                        const int xq = xx + c - 1;
                        win[r][c] = (yy >= 0 && yy < H && xq >= 0 && xq < W) ? __ldg(x + ((size_t)b * H + yy) * W + xq) : 0.f;
                    }
                }
#pragma unroll
                for (int i = 0; i < PPT; ++i) {
                    F4e a = bv;
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        const float v = win[t / 3][i + t % 3];
                        a.x = fmaf(v, wv[t].x, a.x);
                        a.y = fmaf(v, wv[t].y, a.y);
                        a.z = fmaf(v, wv[t].z, a.z);
                        a.w = fmaf(v, wv[t].w, a.w);
                    }
                    const size_t o = (((size_t)b * H + y) * W + xx + i) * Cout + c4;
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


extern "C" int anoddpm_conv_head_sep_supported(int C, int Cout) { return Cout == 1 && (C == 128 || C == 256); }

extern "C" long long anoddpm_conv_head_sep_workspace_bytes(int B, int H, int W) { return 9LL * B * H * W * (long long)sizeof(float); }

extern "C" int anoddpm_conv_head_sep(float* out_nchw, const float* x_nhwc, const float* gn_table, const float* w_packed,
                                     const float* bias, float* workspace, int B, int C, int H, int W, void* stream) {
    ADP_REQUIRE(out_nchw && x_nhwc && gn_table && w_packed && workspace && B > 0 && H > 0 && W > 0,
                "anoddpm_conv_head_sep: bad arguments");
    ADP_REQUIRE(anoddpm_conv_head_sep_supported(C, 1), "anoddpm_conv_head_sep: C must be 128 or 256 (got %d)", C);
    const long long HW = (long long)H * W, npix = (long long)B * HW;
    int ppc = 512;                                  // pixels per CTA (64 per warp); must divide the pixels of one image
    while (ppc > 64 && HW % ppc) ppc >>= 1;
    ADP_REQUIRE(HW % ppc == 0, "anoddpm_conv_head_sep: H*W must be a multiple of 64 (got %lld)", HW);
    const long long ctas = npix / ppc;
    ADP_REQUIRE(ctas < (1LL << 31), "anoddpm_conv_head_sep: too many tiles");
    cudaStream_t st = (cudaStream_t)stream;
    const float2* tab = reinterpret_cast<const float2*>(gn_table);
    if (C == 128) head_proj_kernel<1, 8><<<(unsigned)ctas, 256, 0, st>>>(workspace, x_nhwc, tab, w_packed, HW, npix, ppc);
    else head_proj_kernel<2, 4><<<(unsigned)ctas, 256, 0, st>>>(workspace, x_nhwc, tab, w_packed, HW, npix, ppc);
    ADP_LAUNCH_CHECK();
    head_gather_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(out_nchw, workspace, bias, H, W, npix);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}

extern "C" int anoddpm_conv_stem_fused(float* out_nhwc, void* out16_nhwc, double* stats, const float* x_nchw, const float* w_oihw,
                                       const float* bias, int B, int Cin, int H, int W, int Cout, void* stream) {
    ADP_REQUIRE(out_nhwc && x_nchw && w_oihw && B > 0 && H > 0 && W > 0, "anoddpm_conv_stem_fused: bad arguments");
    ADP_REQUIRE(Cin == 1 && Cout % 4 == 0 && Cout <= 1024 && 256 % (Cout / 4) == 0 && W % 8 == 0,
                "anoddpm_conv_stem_fused: needs Cin == 1, Cout/4 dividing 256, W %% 8 == 0 (Cin=%d Cout=%d W=%d)", Cin, Cout, W);
    const int ppt = (W % 32 == 0) ? 4 : 1;          // pixels per thread: segments of 32 (or 8) pixels
    const int segs_img = (W / (8 * ppt)) * H;
    int spc = 64 / ppt;                             // segments per CTA; must divide the segments of one image
    while (spc > 1 && (segs_img % spc || (long long)B * segs_img / spc < 148 * 8)) spc >>= 1;
    const long long ctas = (long long)B * segs_img / spc;
    const size_t smem = ((size_t)10 * Cout + 256 * 8) * sizeof(float);
    if (ppt == 4)
        conv_stem_fused_kernel<4><<<(unsigned)ctas, 256, smem, (cudaStream_t)stream>>>(out_nhwc, static_cast<__half*>(out16_nhwc), stats,
                                                                                        x_nchw, w_oihw, bias, B, H, W, Cout, spc);
    else
        conv_stem_fused_kernel<1><<<(unsigned)ctas, 256, smem, (cudaStream_t)stream>>>(out_nhwc, static_cast<__half*>(out16_nhwc), stats,
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
