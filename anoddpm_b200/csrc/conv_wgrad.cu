// anoddpm_b200/csrc/conv_wgrad.cu -- weight gradient of the 3x3 / 1x1 convolutions on tcgen05 (sm_100a).
//
//   dW[co][tap][ci] = sum over (b, y, x) of  dY[b, y, x, co] * A[b, y + dy - 1, x + dx - 1, ci]          (tap = 3 dy + dx)
//
// the backward-by-weights of nn.Conv2d(3x3, padding 1) / Conv2d(1x1) / Conv1d(1) as the reference's autograd computes it for
// loss.backward() (diffusion_training.py:101; UNet.py:170-200, 115-118).  The reduction runs over PIXELS, and both operands
// live in HBM as NHWC fp16 with the channels contiguous -- i.e. they are "MN-major" GEMM operands (M = co, N = ci, K =
// pixels).  tcgen05.mma takes MN-major shared-memory operands directly (instruction-descriptor bits 15 / 16), so no
// transposed copy of either tensor is ever made:
//   * a TMA box (64 channels, TW pixels, TH rows) with SWIZZLE_128B lands as [pixel rows][128 B of channels] = the canonical
//     MN-major layout ((8,8,m),(8,k)) : ((1,8,LBO),(64,SBO)) in 16-bit elements: 64 channels contiguous, consecutive K
//     (pixels) 128 B apart, 8-pixel groups SBO apart, the next 64 channels LBO (= one box) apart;
//   * the three horizontal taps of one vertical tap dy are three descriptors into ONE activation box with a one-pixel halo
//     (TW + 2 pixels per row): tap dx starts dx rows (128 B) further -- the MMA swizzles on absolute shared-memory address
//     bits, so a start that is not atom-aligned still addresses the rows TMA wrote (profiles/r01_umma_desc_probe.md);
//     `dx_boxes = 1` loads one aligned box per tap instead (fallback / cross-check);
//   * zero padding comes from TMA's out-of-bounds fill.
// One CTA = (128 output channels) x (128 input channels) x (one vertical tap: 3 accumulators of 128 TMEM columns; 1 for a
// 1x1 conv) x (a strided subset of the pixel tiles, split-K); fp32 partial sums are added to the packed gradient
// [Cout][taps][Cin] with red.global.add.f32, scaled by `alpha` (1 / loss scale).
// Warps: 0 = TMA producer, 1 = MMA issuer, 2..5 = epilogue (TMEM lane quarter = warp % 4).
#include "common.cuh"
#include "anoddpm_b200.h"
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

namespace {

constexpr int kRow = 128;          // bytes per smem row: 64 fp16 channels of one pixel
constexpr int kThreadsW = 192;
constexpr int kMaxStagesW = 4;

struct WgArgs {
    float* dw;                 // [Cout][taps][Cin] fp32, accumulated
    int B, H, W, Cout, Cin;
    int ksize, taps;           // 3 -> 9 taps (3 per CTA), 1 -> 1 tap
    int TW, TH;                // pixel tile
    int tiles_x, tiles_y, n_tiles;
    int splits;                // CTAs that share one (co, ci, dy) unit: tile = split, split + splits, ...
    int n_co, n_ci;            // 128-channel blocks
    int stages;
    int dy_box_bytes, a_box_bytes, stage_bytes;
    int dx_boxes;
    float alpha;
};

__device__ __forceinline__ uint32_t smem_u32w(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init_w(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32w(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx_w(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32w(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_w(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP_W:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE_W;\n\t"
        "bra WAIT_LOOP_W;\n\t"
        "WAIT_DONE_W:\n\t"
        "}\n" ::"r"(smem_u32w(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ bool elect_one_w() {
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
__device__ __forceinline__ void tma_load_4d_w(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                              int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32w(smem)), "l"(map), "r"(smem_u32w(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_w(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32w(bar))
                 : "memory");
}
// MN-major SWIZZLE_128B descriptor: start address, LBO = byte distance between 64-channel blocks, SBO = byte distance
// between 8-pixel (K) groups
__device__ __forceinline__ uint64_t make_mn_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__global__ void __launch_bounds__(kThreadsW, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_a, const WgArgs P) {
    extern __shared__ __align__(1024) uint8_t smem_raw_w[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_w) + 1023) & ~uintptr_t(1023));
    const int S = P.stages;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * P.stage_bytes);
    uint64_t* full = bars;                     // [S]
    uint64_t* empty = bars + kMaxStagesW;      // [S]
    uint64_t* acc_full = bars + 2 * kMaxStagesW;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // unit = (co block, ci block, vertical tap, split)
    int u = blockIdx.x;
    const int split = u % P.splits; u /= P.splits;
    const int ndy = P.ksize == 3 ? 3 : 1;
    const int dy = u % ndy; u /= ndy;
    const int cib = u % P.n_ci;
    const int cob = u / P.n_ci;
    const int ndx = P.ksize == 3 ? 3 : 1;
    const int halo = P.ksize == 3 ? 1 : 0;
    const int n_my = P.n_tiles > split ? (P.n_tiles - split + P.splits - 1) / P.splits : 0;

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < S; ++i) { mbar_init_w(&full[i], 1); mbar_init_w(&empty[i], 1); }
        mbar_init_w(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32w(tmem_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dy) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    const int a_boxes = P.dx_boxes ? ndx : 1;                 // activation boxes per 64-channel block
    if (warp == 0) {
        if (elect_one_w()) {
            for (int i = 0; i < n_my; ++i) {
                const int tile = split + i * P.splits;
                const int b = tile / (P.tiles_x * P.tiles_y), tr = tile - b * (P.tiles_x * P.tiles_y);
                const int y0 = (tr / P.tiles_x) * P.TH, x0 = (tr % P.tiles_x) * P.TW;
                const int s = i % S;
                mbar_wait_w(&empty[s], ((i / S) & 1) ^ 1);
                uint8_t* st = smem + (size_t)s * P.stage_bytes;
                mbar_expect_tx_w(&full[s], (uint32_t)(2 * P.dy_box_bytes + 2 * a_boxes * P.a_box_bytes));
                for (int h = 0; h < 2; ++h)
                    tma_load_4d_w(st + h * P.dy_box_bytes, &map_dy, &full[s], cob * 128 + 64 * h, x0, y0, b);
                uint8_t* sa = st + 2 * P.dy_box_bytes;
                for (int h = 0; h < 2; ++h)
                    for (int k = 0; k < a_boxes; ++k)
                        tma_load_4d_w(sa + (h * a_boxes + k) * P.a_box_bytes, &map_a, &full[s], cib * 128 + 64 * h,
                                      x0 - halo + (P.dx_boxes ? k : 0), y0 + dy - halo, b);
            }
        }
    } else if (warp == 1) {
        // instruction descriptor: D fp32, A/B fp16, both MN-major, M = 128, N = 128
        const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const int a_row_px = P.dx_boxes ? P.TW : P.TW + 2 * halo;     // pixels per image row of an activation box
        // a K step = 16 pixels: two 8-pixel groups -- the two halves of one 16-pixel row (TW = 16) or two image rows (TW = 8)
        const uint32_t sbo_dy = 1024;
        const uint32_t sbo_a = P.TW == 16 ? 1024 : (uint32_t)a_row_px * kRow;
        const int ksteps = P.TH * P.TW / 16;
        for (int i = 0; i < n_my; ++i) {
            const int s = i % S;
            mbar_wait_w(&full[s], (i / S) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one_w()) {
                const uint32_t dy_base = smem_u32w(smem + (size_t)s * P.stage_bytes);
                const uint32_t a_base = dy_base + 2 * P.dy_box_bytes;
                for (int ks = 0; ks < ksteps; ++ks) {
                    // first pixel of this K step inside the tile
                    const int py = P.TW == 16 ? ks : 2 * ks;
                    const uint64_t ad = make_mn_desc(dy_base + (uint32_t)(py * P.TW) * kRow, (uint32_t)P.dy_box_bytes, sbo_dy);
                    for (int dx = 0; dx < ndx; ++dx) {
                        const uint32_t start = P.dx_boxes ? a_base + (uint32_t)dx * P.a_box_bytes + (uint32_t)(py * a_row_px) * kRow
                                                           : a_base + (uint32_t)(py * a_row_px + dx) * kRow;
                        const uint64_t bd = make_mn_desc(start, (uint32_t)(a_boxes * P.a_box_bytes), sbo_a);
                        umma_f16(tmem_base + (uint32_t)(dx * 128), ad, bd, idesc, (i | ks) ? 1u : 0u);
                    }
                }
                umma_commit_w(&empty[s]);
                if (i == n_my - 1) umma_commit_w(acc_full);
            }
            __syncwarp();
        }
    } else if (n_my > 0) {
        // ---------------- epilogue: TMEM lane = output channel, column = input channel of tap dx ----------------
        const int quarter = warp & 3;
        mbar_wait_w(acc_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int co = cob * 128 + quarter * 32 + lane;
        for (int dx = 0; dx < ndx; ++dx) {
            const int tap = P.ksize == 3 ? dy * 3 + dx : 0;
            float* dst = P.dw + ((size_t)co * P.taps + tap) * P.Cin + cib * 128;
#pragma unroll 1
            for (int cb = 0; cb < 128; cb += 16) {
                uint32_t r[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(dx * 128 + cb);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (P.splits == 1) {     // this CTA holds the whole sum: plain 16-byte read-modify-writes (dw is a running gradient)
#pragma unroll
                    for (int e = 0; e < 16; e += 4) {
                        float4 v = *reinterpret_cast<float4*>(dst + cb + e);
                        v.x += __uint_as_float(r[e]) * P.alpha; v.y += __uint_as_float(r[e + 1]) * P.alpha;
                        v.z += __uint_as_float(r[e + 2]) * P.alpha; v.w += __uint_as_float(r[e + 3]) * P.alpha;
                        *reinterpret_cast<float4*>(dst + cb + e) = v;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 16; e += 4)     // split-K partial sums: 16-byte vector reductions in L2
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + cb + e),
                                     "f"(__uint_as_float(r[e]) * P.alpha), "f"(__uint_as_float(r[e + 1]) * P.alpha),
                                     "f"(__uint_as_float(r[e + 2]) * P.alpha), "f"(__uint_as_float(r[e + 3]) * P.alpha)
                                     : "memory");
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

typedef CUresult (*EncodeTiledFnW)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFnW get_encode_w() {
    static EncodeTiledFnW fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFnW>(p);
    }
    return fn;
}

// NHWC fp16 map: dims (C, W, H, B), box (64 ch, px, rows, 1), SWIZZLE_128B
int make_map_w(CUtensorMap* m, const void* p, int C, int W, int H, int B, int px, int rows) {
    EncodeTiledFnW enc = get_encode_w();
    if (!enc) { anoddpm_set_error("cuTensorMapEncodeTiled entry point not available"); return ANODDPM_ERR_CUDA; }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64u, (cuuint32_t)px, (cuuint32_t)rows, 1u};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(p), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        anoddpm_set_error("cuTensorMapEncodeTiled(wgrad C=%d W=%d H=%d B=%d px=%d rows=%d) failed: %d", C, W, H, B, px, rows, (int)r);
        return ANODDPM_ERR_CUDA;
    }
    return ANODDPM_OK;
}

}  // namespace

extern "C" int anoddpm_conv_wgrad_supported(int B, int H, int W, int Cin, int Cout, int ksize) {
    if (B <= 0 || (ksize != 1 && ksize != 3)) return 0;
    if (Cin % 128 || Cout % 128) return 0;
    if (H % 8 || W % 8) return 0;
    if (W > 8 && W % 16) return 0;
    return 1;
}

extern "C" int anoddpm_conv_wgrad(float* dw, const void* dy16, const void* a16, int B, int H, int W, int Cin, int Cout,
                                  int ksize, float alpha, void* stream) {
    ADP_REQUIRE(dw && dy16 && a16, "anoddpm_conv_wgrad: null pointer");
    ADP_REQUIRE(anoddpm_conv_wgrad_supported(B, H, W, Cin, Cout, ksize),
                "anoddpm_conv_wgrad: needs ksize 1|3, channels %% 128 == 0, H %% 8 == 0, W == 8 or W %% 16 == 0 "
                "(B=%d H=%d W=%d Cin=%d Cout=%d ksize=%d)", B, H, W, Cin, Cout, ksize);
    WgArgs P;
    P.dw = dw; P.B = B; P.H = H; P.W = W; P.Cout = Cout; P.Cin = Cin;
    P.ksize = ksize; P.taps = ksize * ksize;
    P.TW = W >= 16 ? 16 : 8;
    P.TH = 8;
    P.tiles_x = W / P.TW; P.tiles_y = H / P.TH;
    P.n_tiles = B * P.tiles_x * P.tiles_y;
    P.n_co = Cout / 128; P.n_ci = Cin / 128;
    { const char* e = getenv("ANODDPM_WGRAD_DX_BOXES"); P.dx_boxes = e ? atoi(e) : 0; }
    if (ksize == 1) P.dx_boxes = 0;
    const int halo = ksize == 3 ? 1 : 0;
    const int ndx = ksize == 3 ? 3 : 1;
    P.dy_box_bytes = P.TH * P.TW * kRow;
    const int a_px = P.dx_boxes ? P.TW : P.TW + 2 * halo;
    P.a_box_bytes = P.TH * a_px * kRow;
    const int a_boxes = P.dx_boxes ? ndx : 1;
    P.stage_bytes = 2 * P.dy_box_bytes + 2 * a_boxes * P.a_box_bytes;     // multiples of 1024 (TH = 8 rows of 128 B per pixel column)
    int stages = (int)((227 * 1024 - 1024 - 256) / P.stage_bytes);
    if (stages > kMaxStagesW) stages = kMaxStagesW;
    ADP_REQUIRE(stages >= 2, "anoddpm_conv_wgrad: stage of %d bytes leaves fewer than 2 stages", P.stage_bytes);
    P.stages = stages;
    const int units = P.n_co * P.n_ci * (ksize == 3 ? 3 : 1);
    // one CTA per SM (204 KB of shared memory): ONE wave -- units * splits <= 148 -- so no partial second wave; each CTA
    // then runs a long K loop and pays its epilogue (147 K reductions) once
    int splits = 148 / units;
    if (splits > P.n_tiles) splits = P.n_tiles;
    if (splits < 1) splits = 1;
    P.splits = splits;
    P.alpha = alpha;
    CUtensorMap mdy, ma;
    int rc;
    if ((rc = make_map_w(&mdy, dy16, Cout, W, H, B, P.TW, P.TH))) return rc;
    if ((rc = make_map_w(&ma, a16, Cin, W, H, B, a_px, P.TH))) return rc;
    const size_t smem = (size_t)stages * P.stage_bytes + 256 + 1024;
    static bool attr_dev[64] = {};
    int dev_ = 0;
    ADP_CUDA(cudaGetDevice(&dev_));
    if (!attr_dev[dev_ & 63]) {
        ADP_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_dev[dev_ & 63] = true;
    }
    conv_wgrad_kernel<<<units * splits, kThreadsW, smem, (cudaStream_t)stream>>>(mdy, ma, P);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}
