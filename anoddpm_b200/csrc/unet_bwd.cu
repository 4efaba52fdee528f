// anoddpm_b200/csrc/unet_bwd.cu -- backward kernels of the UNet denoiser that are NOT tensor-core shaped (training step,
// reference diffusion_training.py:99-107 loss.backward(); UNet.py:137-217, 384-411): GroupNorm(+SiLU)(+resample) backward,
// gradient cast + bias column sums, a strided batched fp32 GEMM (attention core backward, the timestep MLPs), softmax /
// softmax-backward rows, stem weight gradient, head data / weight gradients.  The convolution gradients themselves run on
// tcgen05: data gradients through anoddpm_conv with flipped weight packs, weight gradients through conv_wgrad.cu.
// All activation gradients carry the loss scale S (fp16 operands need it); parameter gradients are multiplied by
// alpha = 1 / S where they are produced.
#include "common.cuh"
#include "anoddpm_b200.h"
#include <cuda_fp16.h>

namespace {

__device__ __forceinline__ float sigmoid_f(float y) { return 1.0f / (1.0f + __expf(-y)); }
struct alignas(16) F4 { float x, y, z, w; };
__device__ __forceinline__ F4 ld4(const float* p) { return *reinterpret_cast<const F4*>(p); }
__device__ __forceinline__ void st4(float* p, F4 v) { *reinterpret_cast<F4*>(p) = v; }

struct GnBwdArgs {
    const void* x1;            // fp32 [B][H][W][C1], or fp16 when in16 (then x = half + shift)
    const float* x2;           // optional fp32 [B][H][W][C2]
    int C1, C2;
    const double* st1;         // [REPLICAS][B][C1][2]
    const double* st2;
    const float* gamma;
    const float* beta;
    const float* shift_c;      // in16: per-channel offset (conv bias) or NULL
    const float* shift_bc;     // in16: per-(sample, channel) offset (timestep-embedding row) or NULL
    int shift_stride;
    const void* da;            // gradient wrt the OUTPUT of resample(act(GN(x))): [B][Ho][Wo][C], fp16 or fp32
    int da16;
    float* sums;               // [B][C][2]: sum_p g, sum_p g * xhat  (g = gradient wrt the GroupNorm output)
    void* dx1;                 // fp32 (or fp16 when in16) [B][H][W][C1]
    float* dx2;
    const void* add;           // optional addend (gradient that reaches x on the skip path)
    int add_mode;              // 0 none, 1 same resolution, 2 add is H/2 x W/2 (x was avg-pooled on the skip path: 0.25 * add[p/2]),
                               // 3 add is 2H x 2W (x was nearest-upsampled on the skip path: sum of the 4 children)
    int add16;
    float* dsum;               // optional [B][C]: += sum over pixels of dx (gradient of a per-(sample, channel) additive shift)
    float* dgamma;             // += alpha * sum_b sums[b][c][1]
    float* dbeta;              // += alpha * sum_b sums[b][c][0]
    float alpha;
    int B, H, W, groups, silu, mode;   // mode as the forward: 0 same, 1 output avg-pooled 2x2, 2 output nearest-upsampled x2
    float eps;
    int pix_per_cta;
};

// group mean / rstd of sample b from the per-channel fp64 sums (thread (g, r) adds up replica r of group g)
__device__ __forceinline__ void group_stats(const GnBwdArgs& P, int b, float* s_mean, float* s_rstd) {
    const int C = P.C1 + P.C2, gs = C / P.groups, nb = P.B;
    const int g = threadIdx.x / ANODDPM_STAT_REPLICAS, r = threadIdx.x % ANODDPM_STAT_REPLICAS;
    double s = 0.0, q = 0.0;
    if (g < P.groups) {
        for (int c = g * gs; c < (g + 1) * gs; ++c) {
            const double* st = c < P.C1 ? P.st1 + (size_t)r * nb * P.C1 * 2 + ((size_t)b * P.C1 + c) * 2
                                        : P.st2 + (size_t)r * nb * P.C2 * 2 + ((size_t)b * P.C2 + (c - P.C1)) * 2;
            s += st[0];
            q += st[1];
        }
    }
#pragma unroll
    for (int o = 1; o < ANODDPM_STAT_REPLICAS; o <<= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (g < P.groups && r == 0) {
        const double n = (double)P.H * P.W * gs;
        const double m = s / n;
        double var = q / n - m * m;
        if (var < 0) var = 0;
        s_mean[g] = (float)m;
        s_rstd[g] = (float)(1.0 / sqrt(var + (double)P.eps));
    }
}

__device__ __forceinline__ void load4(const void* p, bool f16, size_t idx, float v[4]) {
    if (f16) {
        const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(p) + idx);
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else {
        const F4 t = ld4(reinterpret_cast<const float*>(p) + idx);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
}

// gradient wrt the pre-resample activation at full-resolution pixel (y, x): mode 0 da[p]; 1 0.25 * da[p/2]; 2 sum of 4 children
__device__ __forceinline__ void load_dpre(const GnBwdArgs& P, int b, int y, int x, int C, int c, float d[4]) {
    if (P.mode == 0) {
        load4(P.da, P.da16, (((size_t)b * P.H + y) * P.W + x) * C + c, d);
    } else if (P.mode == 1) {
        const int Ho = P.H >> 1, Wo = P.W >> 1;
        load4(P.da, P.da16, (((size_t)b * Ho + (y >> 1)) * Wo + (x >> 1)) * C + c, d);
#pragma unroll
        for (int j = 0; j < 4; ++j) d[j] *= 0.25f;
    } else {
        const int Ho = P.H * 2, Wo = P.W * 2;
        float t[4];
        d[0] = d[1] = d[2] = d[3] = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            load4(P.da, P.da16, (((size_t)b * Ho + 2 * y + (i >> 1)) * Wo + 2 * x + (i & 1)) * C + c, t);
#pragma unroll
            for (int j = 0; j < 4; ++j) d[j] += t[j];
        }
    }
}

// x (with the in16 shift added), for channel vector c of the concatenated tensor
__device__ __forceinline__ void load_x(const GnBwdArgs& P, bool in16, int b, size_t pix, int c, const float sh[4], float v[4]) {
    if (in16) {
        load4(P.x1, true, pix * P.C1 + c, v);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += sh[j];
    } else if (c < P.C1) {
        load4(P.x1, false, pix * P.C1 + c, v);
    } else {
        load4(P.x2, false, pix * P.C2 + (c - P.C1), v);
    }
}

// grid (pixel chunks, B); thread = 4 consecutive channels; rows of 256 / (C/4) pixels in flight
template <bool IN16, bool APPLY>
__global__ void __launch_bounds__(256) gn_bwd_kernel(const GnBwdArgs P) {
    __shared__ float s_mean[64], s_rstd[64], s_a[64], s_b[64];
    __shared__ float s_red[256 * 8];
    const int b = blockIdx.y, C = P.C1 + P.C2, gs = C / P.groups, ncv = C / 4, HW = P.H * P.W;
    group_stats(P, b, s_mean, s_rstd);
    __syncthreads();
    if (APPLY) {
        if (threadIdx.x < P.groups) {
            const int g = threadIdx.x;
            float a = 0.f, bb = 0.f;
            for (int c = g * gs; c < (g + 1) * gs; ++c) {
                a += P.gamma[c] * P.sums[((size_t)b * C + c) * 2];
                bb += P.gamma[c] * P.sums[((size_t)b * C + c) * 2 + 1];
            }
            const float n = (float)HW * gs;
            s_a[g] = a / n;
            s_b[g] = bb / n;
        }
        if (blockIdx.x == 0 && P.dgamma) {     // parameter gradients: one CTA per sample adds its (b, c) sums
            for (int c = threadIdx.x; c < C; c += blockDim.x) {
                atomicAdd(P.dbeta + c, P.alpha * P.sums[((size_t)b * C + c) * 2]);
                atomicAdd(P.dgamma + c, P.alpha * P.sums[((size_t)b * C + c) * 2 + 1]);
            }
        }
        __syncthreads();
    }
    const int tpr = ncv < 256 ? ncv : 256, rows = 256 / tpr;
    const int r = threadIdx.x / tpr, cvl = threadIdx.x - r * tpr;
    const int p0 = blockIdx.x * P.pix_per_cta, p1 = min(HW, p0 + P.pix_per_cta);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};      // reduce: S1[4], S2[4]; apply: column sums of dx [4]
    if (r < rows && cvl < ncv) {
        const int c = cvl * 4;
        float ga[4], be[4], sh[4], mu[4], rs[4], ag[4], bg[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int g = (c + j) / gs;
            ga[j] = P.gamma[c + j];
            be[j] = P.beta[c + j];
            mu[j] = s_mean[g];
            rs[j] = s_rstd[g];
            sh[j] = 0.f;
            if (IN16) {
                if (P.shift_c) sh[j] += P.shift_c[c + j];
                if (P.shift_bc) sh[j] += P.shift_bc[(size_t)b * P.shift_stride + c + j];
            }
            if (APPLY) { ag[j] = s_a[g]; bg[j] = s_b[g]; }
        }
        constexpr int U = APPLY ? 2 : 4;            // pixels in flight per thread (the passes are HBM-latency bound otherwise)
        for (int pb = p0 + r; pb < p1; pb += rows * U) {
            float xv[U][4], d[U][4], ad[U][4];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int p = pb + u * rows;
                if (p >= p1) continue;
                const int y = p / P.W, x = p - y * P.W;
                const size_t pix = (size_t)b * HW + p;
                load_x(P, IN16, b, pix, c, sh, xv[u]);
                load_dpre(P, b, y, x, C, c, d[u]);
                if (APPLY && P.add_mode) {
                    if (P.add_mode == 1) {
                        load4(P.add, P.add16, pix * C + c, ad[u]);
                    } else if (P.add_mode == 2) {
                        load4(P.add, P.add16, (((size_t)b * (P.H >> 1) + (y >> 1)) * (P.W >> 1) + (x >> 1)) * C + c, ad[u]);
#pragma unroll
                        for (int j = 0; j < 4; ++j) ad[u][j] *= 0.25f;
                    } else {
                        float t[4];
                        ad[u][0] = ad[u][1] = ad[u][2] = ad[u][3] = 0.f;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            load4(P.add, P.add16, (((size_t)b * (2 * P.H) + 2 * y + (i >> 1)) * (2 * P.W) + 2 * x + (i & 1)) * C + c, t);
#pragma unroll
                            for (int j = 0; j < 4; ++j) ad[u][j] += t[j];
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int p = pb + u * rows;
                if (p >= p1) continue;
                const size_t pix = (size_t)b * HW + p;
                float dx[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float xh = (xv[u][j] - mu[j]) * rs[j];
                    float g = d[u][j];
                    if (P.silu) {
                        const float yy = xh * ga[j] + be[j];
                        const float sg = __fdividef(1.0f, 1.0f + __expf(-yy));
                        g *= sg * (1.0f + yy * (1.0f - sg));
                    }
                    if (APPLY) dx[j] = rs[j] * (g * ga[j] - ag[j] - xh * bg[j]) + (P.add_mode ? ad[u][j] : 0.f);
                    else { acc[j] += g; acc[4 + j] += g * xh; }
                }
                if (APPLY) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[j] += dx[j];
                    if (IN16) {
                        const __half2 h0 = __floats2half2_rn(dx[0], dx[1]), h1 = __floats2half2_rn(dx[2], dx[3]);
                        uint2 uu;
                        uu.x = *reinterpret_cast<const uint32_t*>(&h0);
                        uu.y = *reinterpret_cast<const uint32_t*>(&h1);
                        *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(P.dx1) + pix * P.C1 + c) = uu;
                    } else {
                        F4 o;
                        o.x = dx[0]; o.y = dx[1]; o.z = dx[2]; o.w = dx[3];
                        if (c < P.C1) st4(reinterpret_cast<float*>(P.dx1) + pix * P.C1 + c, o);
                        else st4(P.dx2 + pix * P.C2 + (c - P.C1), o);
                    }
                }
            }
        }
    }
    // CTA reduction over the pixel rows, then one atomic per (channel, quantity)
    constexpr int NQ = APPLY ? 4 : 8;
    if (!APPLY || P.dsum) {
#pragma unroll
        for (int j = 0; j < NQ; ++j) s_red[threadIdx.x * 8 + j] = acc[j];
        __syncthreads();
        if (r == 0 && cvl < ncv) {
            float tot[NQ];
#pragma unroll
            for (int j = 0; j < NQ; ++j) tot[j] = 0.f;
            for (int rr = 0; rr < rows; ++rr)
#pragma unroll
                for (int j = 0; j < NQ; ++j) tot[j] += s_red[(rr * tpr + cvl) * 8 + j];
            const int c = cvl * 4;
            if (APPLY) {
#pragma unroll
                for (int j = 0; j < 4; ++j) atomicAdd(P.dsum + (size_t)b * C + c + j, tot[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    atomicAdd(P.sums + ((size_t)b * C + c + j) * 2, tot[j]);
                    atomicAdd(P.sums + ((size_t)b * C + c + j) * 2 + 1, tot[4 + j]);
                }
            }
        }
    }
}

// y16 = half(x * scale), colsum[c] += alpha * sum_n x[n][c]   (x fp32 [N][C]); thread = 4 channels
__global__ void __launch_bounds__(256) cast_colsum_kernel(__half* __restrict__ y, const float* __restrict__ x,
                                                          float* __restrict__ colsum, long long N, int C, float scale,
                                                          float alpha, int rows_per_cta) {
    __shared__ float s_red[256 * 4];
    const int ncv = C / 4, tpr = ncv < 256 ? ncv : 256, rows = 256 / tpr;
    const int r = threadIdx.x / tpr, cvl = threadIdx.x - r * tpr;
    const long long n0 = (long long)blockIdx.x * rows_per_cta, n1 = min(N, n0 + rows_per_cta);
    for (int cv = cvl; cv < ncv; cv += tpr) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (r < rows) {
            for (long long n = n0 + r; n < n1; n += rows) {
                const F4 v = ld4(x + n * C + cv * 4);
                acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
                if (y) {
                    const __half2 h0 = __floats2half2_rn(v.x * scale, v.y * scale), h1 = __floats2half2_rn(v.z * scale, v.w * scale);
                    uint2 u;
                    u.x = *reinterpret_cast<const uint32_t*>(&h0);
                    u.y = *reinterpret_cast<const uint32_t*>(&h1);
                    *reinterpret_cast<uint2*>(y + n * C + cv * 4) = u;
                }
            }
        }
        if (colsum) {
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 4; ++j) s_red[threadIdx.x * 4 + j] = acc[j];
            __syncthreads();
            if (r == 0) {
                float tot[4] = {0.f, 0.f, 0.f, 0.f};
                for (int rr = 0; rr < rows; ++rr)
#pragma unroll
                    for (int j = 0; j < 4; ++j) tot[j] += s_red[(rr * tpr + cvl) * 4 + j];
#pragma unroll
                for (int j = 0; j < 4; ++j) atomicAdd(colsum + cv * 4 + j, alpha * tot[j]);
            }
        }
    }
}

// the same for an fp16 source: colsum[c] += alpha * sum_n x16[n][c]
__global__ void __launch_bounds__(256) colsum16_kernel(const __half* __restrict__ x, float* __restrict__ colsum, long long N,
                                                       int C, float alpha, int rows_per_cta) {
    __shared__ float s_red[256 * 4];
    const int ncv = C / 4, tpr = ncv < 256 ? ncv : 256, rows = 256 / tpr;
    const int r = threadIdx.x / tpr, cvl = threadIdx.x - r * tpr;
    const long long n0 = (long long)blockIdx.x * rows_per_cta, n1 = min(N, n0 + rows_per_cta);
    for (int cv = cvl; cv < ncv; cv += tpr) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        if (r < rows) {
            for (long long n = n0 + r; n < n1; n += rows) {
                const uint2 u = *reinterpret_cast<const uint2*>(x + n * C + cv * 4);
                const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
                const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
                acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) s_red[threadIdx.x * 4 + j] = acc[j];
        __syncthreads();
        if (r == 0) {
            float tot[4] = {0.f, 0.f, 0.f, 0.f};
            for (int rr = 0; rr < rows; ++rr)
#pragma unroll
                for (int j = 0; j < 4; ++j) tot[j] += s_red[(rr * tpr + cvl) * 4 + j];
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(colsum + cv * 4 + j, alpha * tot[j]);
        }
    }
}

// ---------------------------------------------------------------- strided batched fp32 GEMM (CUDA cores)
// C[z][m][n] (+)= alpha * sum_k A[z](m, k) * B[z](k, n) with element strides; 64 x 64 x 16 tiles, 4 x 4 per thread.
struct GemmArgs {
    float* c;
    const float* a;
    const float* b;
    int M, N, K;
    long long a_ms, a_ks, b_ks, b_ns, ldc;
    int n2;                                   // batch z = z1 * n2 + z2
    long long a_s1, a_s2, b_s1, b_s2, c_s1, c_s2;
    float alpha;
    int accumulate;
};

__global__ void __launch_bounds__(256) gemm_batched_kernel(const GemmArgs G) {
    __shared__ float As[16][65], Bs[16][65];
    const int z1 = blockIdx.z / G.n2, z2 = blockIdx.z - z1 * G.n2;
    const float* A = G.a + z1 * G.a_s1 + z2 * G.a_s2;
    const float* Bm = G.b + z1 * G.b_s1 + z2 * G.b_s2;
    float* Cm = G.c + z1 * G.c_s1 + z2 * G.c_s2;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < G.K; k0 += 16) {
        // tile loads: 1024 elements each, 4 per thread; the fast thread index follows the unit-stride dimension
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = threadIdx.x + 256 * i;
            int mm, kk;
            if (G.a_ks == 1) { kk = e & 15; mm = e >> 4; } else { mm = e & 63; kk = e >> 6; }
            const int m = m0 + mm, k = k0 + kk;
            As[kk][mm] = (m < G.M && k < G.K) ? A[m * G.a_ms + k * G.a_ks] : 0.f;
            int nn, kb;
            if (G.b_ks == 1) { kb = e & 15; nn = e >> 4; } else { nn = e & 63; kb = e >> 6; }
            const int n = n0 + nn, k2 = k0 + kb;
            Bs[kb][nn] = (n < G.N && k2 < G.K) ? Bm[k2 * G.b_ks + n * G.b_ns] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { av[i] = As[kk][ty * 4 + i]; bv[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= G.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= G.N) continue;
            float* p = Cm + m * G.ldc + n;
            *p = G.accumulate ? *p + G.alpha * acc[i][j] : G.alpha * acc[i][j];
        }
    }
}

// rows of length L (L <= 1024): p = softmax(s) in place; one warp per row
__global__ void __launch_bounds__(256) softmax_rows_kernel(float* __restrict__ s, long long rows, int L) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float* p = s + row * L;
    float mx = -INFINITY;
    for (int i = lane; i < L; i += 32) mx = fmaxf(mx, p[i]);
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int i = lane; i < L; i += 32) { const float e = expf(p[i] - mx); p[i] = e; sum += e; }
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    for (int i = lane; i < L; i += 32) p[i] *= inv;
}
// dl = p * (dp - sum_j dp_j p_j), written over dp
__global__ void __launch_bounds__(256) softmax_bwd_rows_kernel(float* __restrict__ dp, const float* __restrict__ p, long long rows,
                                                               int L) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float* d = dp + row * L;
    const float* q = p + row * L;
    float dot = 0.f;
    for (int i = lane; i < L; i += 32) dot += d[i] * q[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    for (int i = lane; i < L; i += 32) d[i] = q[i] * (d[i] - dot);
}

// dx = dy * f'(x) for y = SiLU(x) (elementwise, small tensors of the timestep MLPs)
__global__ void __launch_bounds__(256) silu_bwd_kernel(float* __restrict__ dx, const float* __restrict__ dy,
                                                       const float* __restrict__ x, long long n) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float v = x[i], sg = sigmoid_f(v);
    dx[i] = dy[i] * sg * (1.0f + v * (1.0f - sg));
}

// ---------------------------------------------------------------- stem (Cin = 1) weight gradient
// dw[co][tap] += alpha * sum_{b,p} dy[b,p,co] * x[b, p + tap]; x NCHW with one channel, dy NHWC fp16.
// grid: strips of rows; thread = (pixel lane, co); each thread keeps its 9 tap sums in registers
__global__ void __launch_bounds__(256) stem_wgrad_kernel(float* __restrict__ dw, const __half* __restrict__ dy,
                                                         const float* __restrict__ x, int B, int H, int W, int Cout,
                                                         float alpha, int rows_per_cta) {
    const int lanes = 256 / Cout;                    // pixel lanes (Cout <= 256 divides 256)
    const int co = threadIdx.x % Cout, pl = threadIdx.x / Cout;
    const long long rows_total = (long long)B * H;
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows_total, r0 + rows_per_cta);
    float acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.f;
    if (pl < lanes) {
        for (long long rr = r0; rr < r1; ++rr) {
            const int b = (int)(rr / H), y = (int)(rr - (long long)b * H);
            const float* xb = x + (size_t)b * H * W;
            for (int xx = pl; xx < W; xx += lanes) {
                const float g = __half2float(dy[(((size_t)b * H + y) * W + xx) * Cout + co]);
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const int yy = y + t / 3 - 1, xq = xx + t % 3 - 1;
                    const float v = (yy >= 0 && yy < H && xq >= 0 && xq < W) ? __ldg(xb + (size_t)yy * W + xq) : 0.f;
                    acc[t] = fmaf(g, v, acc[t]);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < 9; ++t) atomicAdd(dw + (size_t)co * 9 + t, alpha * acc[t]);
    }
}

// ---------------------------------------------------------------- head (Cout = 1): data and weight gradients
// forward: out[b,p] = sum_{tap,ci} a[b,p+tap,ci] * w[tap][ci] + bias.   dy: [B][H][W] fp32 (scaled), a: NHWC fp16.
// da[b,q,ci] = sum_tap dy[b, q - tap] * w[tap][ci]  (fp16 out);   thread = (pixel, 8 channels)
__global__ void __launch_bounds__(256) head_dgrad_kernel(__half* __restrict__ da, const float* __restrict__ dy,
                                                         const float* __restrict__ w, int B, int H, int W, int C) {
    const int ncv = C / 8;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long total = (long long)B * H * W * ncv;
    if (idx >= total) return;
    const int cv = (int)(idx % ncv);
    const long long pix = idx / ncv;
    const int x = (int)(pix % W), y = (int)((pix / W) % H), b = (int)(pix / ((long long)W * H));
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        // output pixel p = q - (tap - centre) received a[q] through tap t
        const int yy = y - (t / 3 - 1), xx = x - (t % 3 - 1);
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
        const float g = dy[((size_t)b * H + yy) * W + xx];
        const float* wt = w + (size_t)t * C + cv * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(g, __ldg(wt + j), acc[j]);
    }
    uint4 o;
    uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __half2 h = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
        op[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(da + pix * C + cv * 8) = o;
}
// dw[tap][ci] += alpha * sum_{b,p} dy[b,p] * a[b,p+tap,ci];  dbias += alpha * sum dy.   thread = (pixel lane, ci)
__global__ void __launch_bounds__(256) head_wgrad_kernel(float* __restrict__ dw, float* __restrict__ dbias,
                                                         const float* __restrict__ dy, const __half* __restrict__ a, int B,
                                                         int H, int W, int C, float alpha, int rows_per_cta) {
    const int lanes = 256 / C;
    const int ci = threadIdx.x % C, pl = threadIdx.x / C;
    const long long rows_total = (long long)B * H;
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows_total, r0 + rows_per_cta);
    float acc[9], bsum = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.f;
    if (pl < lanes) {
        for (long long rr = r0; rr < r1; ++rr) {
            const int b = (int)(rr / H), y = (int)(rr - (long long)b * H);
            const float* db = dy + (size_t)b * H * W;
            for (int xx = pl; xx < W; xx += lanes) {
                // activation pixel q = (y, xx) feeds output pixel p = q - (tap - centre) through tap t
                const float v = __half2float(a[(((size_t)b * H + y) * W + xx) * C + ci]);
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const int yy = y - (t / 3 - 1), xq = xx - (t % 3 - 1);
                    const float g = (yy >= 0 && yy < H && xq >= 0 && xq < W) ? __ldg(db + (size_t)yy * W + xq) : 0.f;
                    acc[t] = fmaf(g, v, acc[t]);
                }
                if (ci == 0) bsum += db[(size_t)y * W + xx];
            }
        }
#pragma unroll
        for (int t = 0; t < 9; ++t) atomicAdd(dw + (size_t)t * C + ci, alpha * acc[t]);
        if (ci == 0 && dbias) atomicAdd(dbias, alpha * bsum);
    }
}

// ---------------------------------------------------------------- weight packs for the tensor-core convolutions
// OIHW (or OI1) fp32 parameter -> fp16 pack, saturating round to nearest.  One thread per (row, column) pair moves the
// ksize^2 taps of that pair.   normal:     dst[co][col_offset + tap * Cin + ci]              (forward operand, K-major)
//                              transposed: dst[ci][col_offset + (T - 1 - tap) * Cout + co]   (data-gradient operand: the
// spatially flipped, in/out-transposed filter).  The fast thread index follows the destination's contiguous dimension.
template <bool TRANSPOSED>
__global__ void __launch_bounds__(256) pack_w_kernel(__half* __restrict__ dst, const float* __restrict__ w, int Cout, int Cin,
                                                     int T, long long row_stride, int col_offset) {
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (long long)Cout * Cin) return;
    int co, ci;
    if (TRANSPOSED) { co = (int)(idx % Cout); ci = (int)(idx / Cout); }
    else { ci = (int)(idx % Cin); co = (int)(idx / Cin); }
    const float* src = w + ((size_t)co * Cin + ci) * T;
    for (int t = 0; t < T; ++t) {
        const float v = fminf(fmaxf(src[t], -65504.f), 65504.f);
        if (TRANSPOSED) dst[(size_t)ci * row_stride + col_offset + (size_t)(T - 1 - t) * Cout + co] = __float2half_rn(v);
        else dst[(size_t)co * row_stride + col_offset + (size_t)t * Cin + ci] = __float2half_rn(v);
    }
}

bool al16b(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" int anoddpm_gn_silu_bwd(void* dx1, float* dx2, float* dgamma, float* dbeta, float* dsum, float* sums_ws,
                                   const void* x1, int x1_f16, int C1, const float* x2, int C2, const double* stats1,
                                   const double* stats2, const float* gamma, const float* beta, const float* shift_c,
                                   const float* shift_bc, int shift_stride, const void* da, int da_f16, const void* add,
                                   int add_mode, int add_f16, int B, int H, int W, int groups, float eps, int silu, int mode,
                                   float alpha, void* stream) {
    ADP_REQUIRE(dx1 && sums_ws && x1 && stats1 && gamma && beta && da, "anoddpm_gn_silu_bwd: null pointer");
    ADP_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && C1 > 0 && C2 >= 0, "anoddpm_gn_silu_bwd: bad shape");
    const int C = C1 + C2;
    ADP_REQUIRE(C % 4 == 0 && C1 % 4 == 0 && C <= 1024 && C % groups == 0 && groups * ANODDPM_STAT_REPLICAS <= 256,
                "anoddpm_gn_silu_bwd: needs C %% 4 == 0, C <= 1024, at most 32 groups dividing C (C=%d)", C);
    ADP_REQUIRE((C2 == 0) == (x2 == nullptr) && (C2 == 0) == (stats2 == nullptr) && (C2 == 0 || dx2),
                "anoddpm_gn_silu_bwd: second source mismatch");
    ADP_REQUIRE(!x1_f16 || (C2 == 0 && mode == 0), "anoddpm_gn_silu_bwd: fp16 input: one source, same resolution");
    ADP_REQUIRE(mode >= 0 && mode <= 2 && add_mode >= 0 && add_mode <= 3, "anoddpm_gn_silu_bwd: bad mode");
    ADP_REQUIRE(mode != 1 || (H % 2 == 0 && W % 2 == 0), "anoddpm_gn_silu_bwd: pooled mode needs even H, W");
    ADP_REQUIRE((add_mode == 0) == (add == nullptr), "anoddpm_gn_silu_bwd: addend / mode mismatch");
    ADP_REQUIRE(al16b(dx1) && al16b(x1) && al16b(da), "anoddpm_gn_silu_bwd: pointers must be 16-byte aligned");
    GnBwdArgs P;
    P.x1 = x1; P.x2 = x2; P.C1 = C1; P.C2 = C2; P.st1 = stats1; P.st2 = stats2; P.gamma = gamma; P.beta = beta;
    P.shift_c = shift_c; P.shift_bc = shift_bc; P.shift_stride = shift_stride; P.da = da; P.da16 = da_f16;
    P.sums = sums_ws; P.dx1 = dx1; P.dx2 = dx2; P.add = add; P.add_mode = add_mode; P.add16 = add_f16; P.dsum = dsum;
    P.dgamma = dgamma; P.dbeta = dbeta; P.alpha = alpha; P.B = B; P.H = H; P.W = W; P.groups = groups; P.silu = silu;
    P.mode = mode; P.eps = eps;
    const int HW = H * W;
    int ppc = 1024;
    while (ppc > 32 && (long long)adp_ceil_div(HW, ppc) * B < 148 * 4) ppc >>= 1;
    P.pix_per_cta = ppc;
    dim3 grid((unsigned)adp_ceil_div(HW, ppc), B);
    cudaStream_t st = (cudaStream_t)stream;
    ADP_CUDA(cudaMemsetAsync(sums_ws, 0, (size_t)B * C * 2 * sizeof(float), st));
    if (x1_f16) {
        gn_bwd_kernel<true, false><<<grid, 256, 0, st>>>(P);
        gn_bwd_kernel<true, true><<<grid, 256, 0, st>>>(P);
    } else {
        gn_bwd_kernel<false, false><<<grid, 256, 0, st>>>(P);
        gn_bwd_kernel<false, true><<<grid, 256, 0, st>>>(P);
    }
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}

extern "C" int anoddpm_cast_colsum(void* y_f16, const float* x, float* colsum, int64_t N, int C, float scale, float alpha,
                                   void* stream) {
    ADP_REQUIRE(x && N > 0 && C > 0 && C % 4 == 0 && (y_f16 || colsum), "anoddpm_cast_colsum: bad arguments");
    ADP_REQUIRE(C <= 1024, "anoddpm_cast_colsum: C <= 1024 (C=%d)", C);
    int rpc = 2048;
    while (rpc > 32 && adp_ceil_div(N, rpc) < 148 * 4) rpc >>= 1;
    cast_colsum_kernel<<<adp_ceil_div(N, rpc), 256, 0, (cudaStream_t)stream>>>(static_cast<__half*>(y_f16), x, colsum, N, C, scale,
                                                                               alpha, rpc);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}

extern "C" int anoddpm_colsum_f16(float* colsum, const void* x_f16, int64_t N, int C, float alpha, void* stream) {
    ADP_REQUIRE(x_f16 && colsum && N > 0 && C > 0 && C % 4 == 0, "anoddpm_colsum_f16: bad arguments");
    ADP_REQUIRE(C <= 1024, "anoddpm_colsum_f16: C <= 1024 (C=%d)", C);
    int rpc = 2048;
    while (rpc > 32 && adp_ceil_div(N, rpc) < 148 * 4) rpc >>= 1;
    colsum16_kernel<<<adp_ceil_div(N, rpc), 256, 0, (cudaStream_t)stream>>>(static_cast<const __half*>(x_f16), colsum, N, C, alpha, rpc);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}

extern "C" int anoddpm_gemm_batched(float* c, const float* a, const float* b, int M, int N, int K, int64_t a_ms, int64_t a_ks,
                                    int64_t b_ks, int64_t b_ns, int64_t ldc, int n1, int n2, int64_t a_s1, int64_t a_s2,
                                    int64_t b_s1, int64_t b_s2, int64_t c_s1, int64_t c_s2, float alpha, int accumulate,
                                    void* stream) {
    ADP_REQUIRE(c && a && b && M > 0 && N > 0 && K > 0 && n1 > 0 && n2 > 0 && (long long)n1 * n2 <= 65535,
                "anoddpm_gemm_batched: bad arguments");
    GemmArgs G;
    G.c = c; G.a = a; G.b = b; G.M = M; G.N = N; G.K = K; G.a_ms = a_ms; G.a_ks = a_ks; G.b_ks = b_ks; G.b_ns = b_ns; G.ldc = ldc;
    G.n2 = n2; G.a_s1 = a_s1; G.a_s2 = a_s2; G.b_s1 = b_s1; G.b_s2 = b_s2; G.c_s1 = c_s1; G.c_s2 = c_s2; G.alpha = alpha;
    G.accumulate = accumulate;
    dim3 grid(adp_ceil_div(N, 64), adp_ceil_div(M, 64), n1 * n2);
    gemm_batched_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(G);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}

extern "C" int anoddpm_softmax_rows(float* s, int64_t rows, int L, void* stream) {
    ADP_REQUIRE(s && rows > 0 && L > 0, "anoddpm_softmax_rows: bad arguments");
    softmax_rows_kernel<<<adp_ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(s, rows, L);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}

extern "C" int anoddpm_softmax_bwd_rows(float* dp, const float* p, int64_t rows, int L, void* stream) {
    ADP_REQUIRE(dp && p && rows > 0 && L > 0, "anoddpm_softmax_bwd_rows: bad arguments");
    softmax_bwd_rows_kernel<<<adp_ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(dp, p, rows, L);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}

extern "C" int anoddpm_silu_bwd(float* dx, const float* dy, const float* x, int64_t n, void* stream) {
    ADP_REQUIRE(dx && dy && x && n > 0, "anoddpm_silu_bwd: bad arguments");
    silu_bwd_kernel<<<adp_ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(dx, dy, x, n);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}

extern "C" int anoddpm_stem_wgrad(float* dw, const void* dy_f16, const float* x, int B, int H, int W, int Cout, float alpha,
                                  void* stream) {
    ADP_REQUIRE(dw && dy_f16 && x && B > 0 && H > 0 && W > 0, "anoddpm_stem_wgrad: bad arguments");
    ADP_REQUIRE(Cout > 0 && Cout <= 256 && 256 % Cout == 0, "anoddpm_stem_wgrad: Cout must divide 256 (got %d)", Cout);
    const long long rows = (long long)B * H;
    int rpc = 64;
    while (rpc > 1 && adp_ceil_div(rows, rpc) < 148 * 4) rpc >>= 1;
    stem_wgrad_kernel<<<adp_ceil_div(rows, rpc), 256, 0, (cudaStream_t)stream>>>(dw, static_cast<const __half*>(dy_f16), x, B, H, W,
                                                                                 Cout, alpha, rpc);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}

extern "C" int anoddpm_head_bwd(void* da_f16, float* dw, float* dbias, const float* dy, const void* a_f16, const float* w,
                                int B, int H, int W, int C, float alpha, void* stream) {
    ADP_REQUIRE(da_f16 && dw && dy && a_f16 && w && B > 0 && H > 0 && W > 0, "anoddpm_head_bwd: bad arguments");
    ADP_REQUIRE(C % 8 == 0 && C <= 256 && 256 % C == 0, "anoddpm_head_bwd: C must be a multiple of 8 dividing 256 (got %d)", C);
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)B * H * W * (C / 8);
    head_dgrad_kernel<<<adp_ceil_div(total, 256), 256, 0, st>>>(static_cast<__half*>(da_f16), dy, w, B, H, W, C);
    const long long rows = (long long)B * H;
    int rpc = 64;
    while (rpc > 1 && adp_ceil_div(rows, rpc) < 148 * 4) rpc >>= 1;
    head_wgrad_kernel<<<adp_ceil_div(rows, rpc), 256, 0, st>>>(dw, dbias, dy, static_cast<const __half*>(a_f16), B, H, W, C, alpha, rpc);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}

extern "C" int anoddpm_pack_conv_w(void* dst_f16, const float* w, int Cout, int Cin, int ksize, int64_t row_stride,
                                   int col_offset, int transposed, void* stream) {
    ADP_REQUIRE(dst_f16 && w && Cout > 0 && Cin > 0 && (ksize == 1 || ksize == 3) && row_stride > 0 && col_offset >= 0,
                "anoddpm_pack_conv_w: bad arguments");
    const long long n = (long long)Cout * Cin;
    const int T = ksize * ksize;
    cudaStream_t st = (cudaStream_t)stream;
    if (transposed) pack_w_kernel<true><<<adp_ceil_div(n, 256), 256, 0, st>>>(static_cast<__half*>(dst_f16), w, Cout, Cin, T, row_stride, col_offset);
    else pack_w_kernel<false><<<adp_ceil_div(n, 256), 256, 0, st>>>(static_cast<__half*>(dst_f16), w, Cout, Cin, T, row_stride, col_offset);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}
