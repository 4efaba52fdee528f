// anoddpm_b200/csrc/attention_tc.cu -- QKV self-attention on tcgen05 tensor cores (sm_100a).
//
// Reference: UNet.py:137-153 (QKVAttention): per (sample, head)  W = softmax_s((q*s)^T (k*s)),  a = W v,  s = ch^-1/4,
// on qkv[B][3C][T] split per head into q | k | v.  Here qkv is the NHWC output of the qkv 1x1 conv: [B][T][3C].
//
// Two launches:
//   1. attn_pack_kernel : fp32 qkv -> fp16 operands  Q[BH][T][ch] (x s), K[BH][T][ch] (x s), Vt[BH][ch][T]
//                         (round to nearest; V transposed so that every MMA operand is K-major).
//   2. attention_tc_kernel : one CTA per (128 queries, sample*head):
//        GEMM1  S[128][T]  = Q K^T        tcgen05.mma kind::f16, fp32 accumulators in TMEM (T <= 256 columns),
//                                          operands streamed by TMA in 64-channel K-blocks through a 3-stage ring;
//        softmax            thread = query row: two passes over its TMEM row (max; exp2, sum), un-normalised
//                           probabilities written as fp16 into shared memory in the K-major SWIZZLE_128B layout;
//        GEMM2  O[128][ch] = P Vt^T       same pipeline, K-blocks of 64 keys (Vt tiles loaded while the softmax runs);
//        epilogue           tcgen05.ld, x 1/sum, store fp32 / TF32-rounded fp32 / fp16 into out[B][T][C].
//   Warp roles (192 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2..5 softmax + epilogue.
// Contract (anoddpm_attention_tc_supported): T, ch = C/heads multiples of 64; T <= 256 and ch <= 256 run this kernel,
// longer sequences (T a multiple of 256) and wider heads the tiled kernel below.
#include "common.cuh"
#include "anoddpm_b200.h"
#include <cuda.h>
#include <cuda_fp16.h>

namespace {

constexpr int kThreads = 192;
constexpr int kStages = 3;
constexpr int kQBytes = 128 * 128;          // Q K-block / P K-block: 128 rows x 128 B
constexpr int kStageBytes = kQBytes + 256 * 128;   // + up to 256 rows of K (or Vt)
constexpr int kOCol = 256;                  // TMEM column of the O accumulator (S occupies [0, T))

struct AttnArgs {
    void* out;
    int T, C, ch, heads;
    int out_mode;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float round_tf32(float v) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return __uint_as_float(u);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format): 8-row atoms of 128 B, SBO = 1024 B.
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------- pack
// grid (ch/64, T/64, B*heads), 256 threads: one 64-token x 64-channel tile of q, k and v each.
__global__ void __launch_bounds__(256) attn_pack_kernel(__half* __restrict__ Q, __half* __restrict__ K,
                                                        __half* __restrict__ Vt, const float* __restrict__ qkv, int T,
                                                        int C, int heads, float scale) {
    __shared__ __half tile[64][66];
    const int ch = C / heads;
    const int bh = blockIdx.z, b = bh / heads, h = bh - b * heads;
    const int t0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t row = (size_t)3 * C;
    const float* src = qkv + ((size_t)b * T + t0) * row + (size_t)h * 3 * ch + c0;
    for (int r = warp; r < 64; r += 8) {
        const float* p = src + (size_t)r * row + 2 * lane;
        const float2 q2 = *reinterpret_cast<const float2*>(p);
        const float2 k2 = *reinterpret_cast<const float2*>(p + ch);
        const float2 v2 = *reinterpret_cast<const float2*>(p + 2 * ch);
        const size_t o = ((size_t)bh * T + t0 + r) * ch + c0 + 2 * lane;
        *reinterpret_cast<__half2*>(Q + o) = __floats2half2_rn(q2.x * scale, q2.y * scale);
        *reinterpret_cast<__half2*>(K + o) = __floats2half2_rn(k2.x * scale, k2.y * scale);
        tile[2 * lane][r] = __float2half_rn(v2.x);
        tile[2 * lane + 1][r] = __float2half_rn(v2.y);
    }
    __syncthreads();
    for (int c = warp; c < 64; c += 8) {
        const __half2 v = *reinterpret_cast<const __half2*>(&tile[c][2 * lane]);
        *reinterpret_cast<__half2*>(Vt + ((size_t)bh * ch + c0 + c) * T + t0 + 2 * lane) = v;
    }
}

// ---------------------------------------------------------------------------------------------------- attention
__global__ void __launch_bounds__(kThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, const AttnArgs P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;                                   // [kStages][kStageBytes]
    uint8_t* sP = smem + kStages * kStageBytes;             // [T/64][128 rows][128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * kQBytes);
    uint64_t* full = bars;                 // [kStages]
    uint64_t* empty = bars + kStages;      // [kStages]
    uint64_t* s_full = bars + 2 * kStages; // S accumulated
    uint64_t* p_full = s_full + 1;         // P staged in shared memory (128 arrivals)
    uint64_t* o_full = s_full + 2;         // O accumulated
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 3);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * 128, bh = blockIdx.y;
    const int T = P.T, ch = P.ch;
    const int nkb1 = ch / 64, nkb2 = T / 64;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
        for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(s_full, 1);
        mbar_init(p_full, 128);
        mbar_init(o_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (elect_one()) {
            int it = 0;
            for (int kb = 0; kb < nkb1; ++kb, ++it) {
                const int s = it % kStages;
                mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);
                mbar_expect_tx(&full[s], (uint32_t)(kQBytes + T * 128));
                tma_load_3d(ring + s * kStageBytes, &map_q, &full[s], kb * 64, m0, bh);            // rows >= T: zeros
                tma_load_3d(ring + s * kStageBytes + kQBytes, &map_k, &full[s], kb * 64, 0, bh);
            }
            for (int kb = 0; kb < nkb2; ++kb, ++it) {
                const int s = it % kStages;
                mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);
                mbar_expect_tx(&full[s], (uint32_t)(ch * 128));
                tma_load_3d(ring + s * kStageBytes + kQBytes, &map_v, &full[s], kb * 64, 0, bh);
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        const uint32_t idesc1 = (1u << 4) | ((uint32_t)(T >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc2 = (1u << 4) | ((uint32_t)(ch >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        int it = 0;
        for (int kb = 0; kb < nkb1; ++kb, ++it) {
            const int s = it % kStages;
            mbar_wait(&full[s], (it / kStages) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint64_t ad = make_sdesc(smem_u32(ring + s * kStageBytes));
                const uint64_t bd = make_sdesc(smem_u32(ring + s * kStageBytes + kQBytes));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(tmem_base, ad + 2 * k, bd + 2 * k, idesc1, (kb | k) ? 1u : 0u);
                umma_commit(&empty[s]);
                if (kb == nkb1 - 1) umma_commit(s_full);
            }
            __syncwarp();
        }
        mbar_wait(p_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kb = 0; kb < nkb2; ++kb, ++it) {
            const int s = it % kStages;
            mbar_wait(&full[s], (it / kStages) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint64_t ad = make_sdesc(smem_u32(sP + kb * kQBytes));
                const uint64_t bd = make_sdesc(smem_u32(ring + s * kStageBytes + kQBytes));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16(tmem_base + kOCol, ad + 2 * k, bd + 2 * k, idesc2, (kb | k) ? 1u : 0u);
                umma_commit(&empty[s]);
                if (kb == nkb2 - 1) umma_commit(o_full);
            }
            __syncwarp();
        }
    } else {
        // ---------------- softmax + epilogue: thread = query row ----------------
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
        mbar_wait(s_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float mx = -INFINITY;
        for (int c = 0; c < T; c += 32) {
            uint32_t v[32];
            tmem_ld32(trow + (uint32_t)c, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
        }
        const float kLog2e = 1.4426950408889634f;
        const float mxs = mx * kLog2e;
        float sum = 0.f;
        for (int c = 0; c < T; c += 32) {
            uint32_t v[32];
            tmem_ld32(trow + (uint32_t)c, v);
            uint32_t h2[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const float p0 = exp2f(fmaf(__uint_as_float(v[j]), kLog2e, -mxs));
                const float p1 = exp2f(fmaf(__uint_as_float(v[j + 1]), kLog2e, -mxs));
                const __half2 hp = __floats2half2_rn(p0, p1);
                const float2 back = __half22float2(hp);     // normalise by the sum of what the MMA will actually see
                sum += back.x + back.y;
                h2[j >> 1] = *reinterpret_cast<const uint32_t*>(&hp);
            }
            // row r of K-block c/64: 16-byte chunk j of the 128-byte row lands at chunk j ^ (r & 7) (SWIZZLE_128B)
            uint8_t* prow = sP + (c >> 6) * kQBytes + r * 128;
            const int j0 = (c & 63) >> 3;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                *reinterpret_cast<uint4*>(prow + (((j0 + q) ^ (r & 7)) << 4)) =
                    make_uint4(h2[4 * q], h2[4 * q + 1], h2[4 * q + 2], h2[4 * q + 3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(p_full);
        const float inv = 1.0f / sum;
        mbar_wait(o_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int t = m0 + r;
        const int b = bh / P.heads, h = bh - b * P.heads;
        const size_t o = ((size_t)b * T + t) * P.C + (size_t)h * ch;
        for (int c = 0; c < ch; c += 32) {
            uint32_t v[32];
            tmem_ld32(trow + (uint32_t)(kOCol + c), v);
            if (t >= T) continue;
            if (P.out_mode == 2) {
                __half* op = reinterpret_cast<__half*>(P.out) + o + c;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float a0 = fminf(fmaxf(__uint_as_float(v[j + 2 * e]) * inv, -65504.f), 65504.f);
                        const float a1 = fminf(fmaxf(__uint_as_float(v[j + 2 * e + 1]) * inv, -65504.f), 65504.f);
                        const __half2 hh = __floats2half2_rn(a0, a1);
                        w[e] = *reinterpret_cast<const uint32_t*>(&hh);
                    }
                    *reinterpret_cast<uint4*>(op + j) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            } else {
                float* op = reinterpret_cast<float*>(P.out) + o + c;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float a[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        a[e] = __uint_as_float(v[j + e]) * inv;
                        if (P.out_mode == 1) a[e] = round_tf32(a[e]);
                    }
                    *reinterpret_cast<float4*>(op + j) = make_float4(a[0], a[1], a[2], a[3]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------- attention, tiled
// The same attention for long sequences and wide heads (T > 256 keys or ch > 256 channels: what the reference's constructor
// defaults instantiate -- attention at 32x32 maps = 1024 tokens, a single head of 256 / 512 channels; UNet.py:232, 106-112).
// S no longer fits the 512 TMEM columns next to O, so the keys are processed in tiles of TK = 256:
//   pass 1  for every key tile: S_j = Q K_j^T into TMEM columns [0, TK); thread-per-row running MAXIMUM only;
//   pass 2  for every key tile: S_j again, P_j = exp2((S_j - max) log2 e) rounded to fp16 into shared memory (row sums of
//           the rounded values), O += P_j V_j into TMEM columns [256, 256 + chs).
// (max-first instead of an online rescale of O: no TMEM read-modify-write of the accumulator; the price is QK^T twice.)
// Heads wider than 256 channels split their OUTPUT channels over gridDim.z CTAs (each recomputes S over the full head).
// T <= 256 (one tile) skips pass 1 and takes the maximum inside the tile like attention_tc_kernel.
__global__ void __launch_bounds__(kThreads, 1)
attention_tc_tiled_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                          const __grid_constant__ CUtensorMap map_v, const AttnArgs P, int TK, int chs) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;                                   // [kStages][kStageBytes]
    uint8_t* sP = smem + kStages * kStageBytes;             // [TK/64][128 rows][128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * kQBytes);
    uint64_t* full = bars;                 // [kStages]
    uint64_t* empty = bars + kStages;      // [kStages]
    uint64_t* s_full = bars + 2 * kStages; // S_j accumulated
    uint64_t* s_empty = s_full + 1;        // S_j read by all 128 rows (pass 1)
    uint64_t* p_full = s_full + 2;         // P_j staged in shared memory (128 arrivals)
    uint64_t* p_empty = s_full + 3;        // P_j consumed by GEMM 2
    uint64_t* o_full = s_full + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 5);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * 128, bh = blockIdx.y, split = blockIdx.z;
    const int T = P.T, ch = P.ch;
    const int ntiles = (T + TK - 1) / TK;
    const int npass1 = ntiles > 1 ? ntiles : 0;
    const int nkb1 = ch / 64, nkb2 = TK / 64;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v) : "memory");
        for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(s_full, 1);
        mbar_init(s_empty, 128);
        mbar_init(p_full, 128);
        mbar_init(p_empty, 1);
        mbar_init(o_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (elect_one()) {
            int it = 0;
            for (int g = 0; g < npass1 + ntiles; ++g) {
                const bool pass2 = g >= npass1;
                const int j = pass2 ? g - npass1 : g;
                for (int kb = 0; kb < nkb1; ++kb, ++it) {
                    const int s = it % kStages;
                    mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);
                    mbar_expect_tx(&full[s], (uint32_t)(kQBytes + TK * 128));
                    tma_load_3d(ring + s * kStageBytes, &map_q, &full[s], kb * 64, m0, bh);            // rows >= T: zeros
                    tma_load_3d(ring + s * kStageBytes + kQBytes, &map_k, &full[s], kb * 64, j * TK, bh);
                }
                if (pass2) {
                    for (int kb = 0; kb < nkb2; ++kb, ++it) {
                        const int s = it % kStages;
                        mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);
                        mbar_expect_tx(&full[s], (uint32_t)(chs * 128));
                        tma_load_3d(ring + s * kStageBytes + kQBytes, &map_v, &full[s], j * TK + kb * 64, split * chs, bh);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        const uint32_t idesc1 = (1u << 4) | ((uint32_t)(TK >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc2 = (1u << 4) | ((uint32_t)(chs >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        int it = 0, n_se = 0, n_pf = 0;
        for (int g = 0; g < npass1 + ntiles; ++g) {
            const bool pass2 = g >= npass1;
            const int j = pass2 ? g - npass1 : g;
            // S may be overwritten once every row has read the previous pass-1 tile (pass-2 tiles hand S over through p_full)
            if (g > 0 && g <= npass1) {
                mbar_wait(s_empty, n_se & 1);
                ++n_se;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            for (int kb = 0; kb < nkb1; ++kb, ++it) {
                const int s = it % kStages;
                mbar_wait(&full[s], (it / kStages) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint64_t ad = make_sdesc(smem_u32(ring + s * kStageBytes));
                    const uint64_t bd = make_sdesc(smem_u32(ring + s * kStageBytes + kQBytes));
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16(tmem_base, ad + 2 * k, bd + 2 * k, idesc1, (kb | k) ? 1u : 0u);
                    umma_commit(&empty[s]);
                    if (kb == nkb1 - 1) umma_commit(s_full);
                }
                __syncwarp();
            }
            if (pass2) {
                mbar_wait(p_full, n_pf & 1);
                ++n_pf;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = 0; kb < nkb2; ++kb, ++it) {
                    const int s = it % kStages;
                    mbar_wait(&full[s], (it / kStages) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
                        const uint64_t ad = make_sdesc(smem_u32(sP + kb * kQBytes));
                        const uint64_t bd = make_sdesc(smem_u32(ring + s * kStageBytes + kQBytes));
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_f16(tmem_base + kOCol, ad + 2 * k, bd + 2 * k, idesc2, (j | kb | k) ? 1u : 0u);
                        umma_commit(&empty[s]);
                        if (kb == nkb2 - 1) {
                            umma_commit(p_empty);
                            if (j == ntiles - 1) umma_commit(o_full);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ---------------- softmax + epilogue: thread = query row ----------------
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const float kLog2e = 1.4426950408889634f;
        float mx = -INFINITY, sum = 0.f;
        int n_sf = 0, n_pe = 0;
        for (int g = 0; g < npass1; ++g, ++n_sf) {            // pass 1: row maximum over all key tiles
            mbar_wait(s_full, n_sf & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int c = 0; c < TK; c += 32) {
                uint32_t v[32];
                tmem_ld32(trow + (uint32_t)c, v);
#pragma unroll
                for (int q = 0; q < 32; ++q) mx = fmaxf(mx, __uint_as_float(v[q]));
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(s_empty);
        }
        for (int j = 0; j < ntiles; ++j, ++n_sf) {            // pass 2: probabilities and P V
            mbar_wait(s_full, n_sf & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (npass1 == 0) {
                for (int c = 0; c < TK; c += 32) {
                    uint32_t v[32];
                    tmem_ld32(trow + (uint32_t)c, v);
#pragma unroll
                    for (int q = 0; q < 32; ++q) mx = fmaxf(mx, __uint_as_float(v[q]));
                }
            }
            const float mxs = mx * kLog2e;
            if (j > 0) {                                        // GEMM 2 of the previous tile must have read sP
                mbar_wait(p_empty, n_pe & 1);
                ++n_pe;
            }
            for (int c = 0; c < TK; c += 32) {
                uint32_t v[32];
                tmem_ld32(trow + (uint32_t)c, v);
                uint32_t h2[16];
#pragma unroll
                for (int q = 0; q < 32; q += 2) {
                    const float p0 = exp2f(fmaf(__uint_as_float(v[q]), kLog2e, -mxs));
                    const float p1 = exp2f(fmaf(__uint_as_float(v[q + 1]), kLog2e, -mxs));
                    const __half2 hp = __floats2half2_rn(p0, p1);
                    const float2 back = __half22float2(hp);
                    sum += back.x + back.y;
                    h2[q >> 1] = *reinterpret_cast<const uint32_t*>(&hp);
                }
                uint8_t* prow = sP + (c >> 6) * kQBytes + r * 128;
                const int j0 = (c & 63) >> 3;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<uint4*>(prow + (((j0 + q) ^ (r & 7)) << 4)) =
                        make_uint4(h2[4 * q], h2[4 * q + 1], h2[4 * q + 2], h2[4 * q + 3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(p_full);
        }
        const float inv = 1.0f / sum;
        mbar_wait(o_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int t = m0 + r;
        const int b = bh / P.heads, h = bh - b * P.heads;
        const size_t o = ((size_t)b * T + t) * P.C + (size_t)h * ch + (size_t)split * chs;
        for (int c = 0; c < chs; c += 32) {
            uint32_t v[32];
            tmem_ld32(trow + (uint32_t)(kOCol + c), v);
            if (t >= T) continue;
            if (P.out_mode == 2) {
                __half* op = reinterpret_cast<__half*>(P.out) + o + c;
#pragma unroll
                for (int q = 0; q < 32; q += 8) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float a0 = fminf(fmaxf(__uint_as_float(v[q + 2 * e]) * inv, -65504.f), 65504.f);
                        const float a1 = fminf(fmaxf(__uint_as_float(v[q + 2 * e + 1]) * inv, -65504.f), 65504.f);
                        const __half2 hh = __floats2half2_rn(a0, a1);
                        w[e] = *reinterpret_cast<const uint32_t*>(&hh);
                    }
                    *reinterpret_cast<uint4*>(op + q) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            } else {
                float* op = reinterpret_cast<float*>(P.out) + o + c;
#pragma unroll
                for (int q = 0; q < 32; q += 4) {
                    float a[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        a[e] = __uint_as_float(v[q + e]) * inv;
                        if (P.out_mode == 1) a[e] = round_tf32(a[e]);
                    }
                    *reinterpret_cast<float4*>(op + q) = make_float4(a[0], a[1], a[2], a[3]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_attn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// fp16 [n2][n1][n0] (n0 contiguous), box (64, rows, 1), 128-byte swizzle
int make_map3(CUtensorMap* m, const void* p, int n0, int n1, int n2, int rows) {
    EncodeTiledFn enc = get_encode_attn();
    if (!enc) { anoddpm_set_error("cuTensorMapEncodeTiled entry point not available"); return ANODDPM_ERR_CUDA; }
    cuuint64_t dims[3] = {(cuuint64_t)n0, (cuuint64_t)n1, (cuuint64_t)n2};
    cuuint64_t strides[2] = {(cuuint64_t)n0 * 2, (cuuint64_t)n0 * n1 * 2};
    cuuint32_t box[3] = {64u, (cuuint32_t)rows, 1u};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(p), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        anoddpm_set_error("cuTensorMapEncodeTiled(attention %d x %d x %d, box rows %d) failed: %d", n0, n1, n2, rows,
                          (int)r);
        return ANODDPM_ERR_CUDA;
    }
    return ANODDPM_OK;
}

}  // namespace

extern "C" int anoddpm_attention_tc_supported(int T, int C, int heads) {
    if (T <= 0 || C <= 0 || heads <= 0 || C % heads) return 0;
    const int ch = C / heads;
    if (T % 64 || ch % 64) return 0;
    if (T <= 256 && ch <= 256) return 1;                       // one CTA holds S and O in TMEM (attention_tc_kernel)
    // tiled kernel: key tiles of 256, output channels of wide heads split over CTAs
    if (T > 256 && T % 256) return 0;
    if (ch > 1024) return 0;
    const int nsplit = (ch + 255) / 256;
    if (ch % nsplit || (ch / nsplit) % 64) return 0;
    return 1;
}

extern "C" long long anoddpm_attention_tc_workspace_bytes(int B, int T, int C) {
    return 3LL * B * T * C * 2;
}

extern "C" int anoddpm_attention_tc(void* out, const float* qkv, void* workspace, int B, int T, int C, int heads,
                                    int out_mode, void* stream) {
    ADP_REQUIRE(out && qkv && workspace && B > 0, "anoddpm_attention_tc: bad arguments");
    ADP_REQUIRE(anoddpm_attention_tc_supported(T, C, heads), "anoddpm_attention_tc: T=%d C=%d heads=%d not supported "
                "(T and C/heads multiples of 64; T <= 256 or a multiple of 256)", T, C, heads);
    ADP_REQUIRE(out_mode >= 0 && out_mode <= 2, "anoddpm_attention_tc: out_mode must be 0, 1 or 2");
    ADP_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 127) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(qkv) & 15) == 0, "anoddpm_attention_tc: misaligned pointers");
    ADP_REQUIRE((long long)B * heads <= 65535, "anoddpm_attention_tc: batch x heads too large");
    const int ch = C / heads, BH = B * heads;
    __half* Q = static_cast<__half*>(workspace);
    __half* K = Q + (size_t)B * T * C;
    __half* Vt = K + (size_t)B * T * C;
    cudaStream_t st = (cudaStream_t)stream;
    const float scale = 1.0f / sqrtf(sqrtf((float)ch));
    attn_pack_kernel<<<dim3(ch / 64, T / 64, BH), 256, 0, st>>>(Q, K, Vt, qkv, T, C, heads, scale);
    ADP_LAUNCH_CHECK();
    const bool tiled = T > 256 || ch > 256;
    const int TK = T < 256 ? T : 256;
    const int nsplit = (ch + 255) / 256, chs = ch / nsplit;
    CUtensorMap mq, mk, mv;
    int rc;
    if ((rc = make_map3(&mq, Q, ch, T, BH, 128))) return rc;
    if ((rc = make_map3(&mk, K, ch, T, BH, tiled ? TK : T))) return rc;
    if ((rc = make_map3(&mv, Vt, T, ch, BH, tiled ? chs : ch))) return rc;
    AttnArgs P;
    P.out = out; P.T = T; P.C = C; P.ch = ch; P.heads = heads; P.out_mode = out_mode;
    const size_t smem = (size_t)kStages * kStageBytes + 4 * kQBytes + 256 + 1024;
    static bool attr_dev[64] = {};   // the attribute is per device
    int dev_ = 0;
    ADP_CUDA(cudaGetDevice(&dev_));
    if (!attr_dev[dev_ & 63]) {
        ADP_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ADP_CUDA(cudaFuncSetAttribute(attention_tc_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_dev[dev_ & 63] = true;
    }
    if (tiled) attention_tc_tiled_kernel<<<dim3((T + 127) / 128, BH, nsplit), kThreads, smem, st>>>(mq, mk, mv, P, TK, chs);
    else attention_tc_kernel<<<dim3((T + 127) / 128, BH), kThreads, smem, st>>>(mq, mk, mv, P);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}
