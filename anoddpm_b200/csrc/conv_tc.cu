// anoddpm_b200/csrc/conv_tc.cu -- tensor-core convolution for sm_100a: implicit GEMM on
// tcgen05.mma kind::tf32 (fp32 accumulators in TMEM) fed by TMA.
//
//   M = B*H*W output pixels (tile 128 = one UMMA M), N = Cout (tile <= 128), K = 9*Ca (+ Cr1 + Cr2).
//   A operand  : NHWC activations.  One K-block = 32 channels of one filter tap: a 4-D TMA box
//                (32 ch, TW, TH, TB) at coordinates shifted by the tap offset; out-of-image rows are
//                zero-filled by TMA, which IS the conv zero padding -- no im2col buffer, no halo code.
//                The box lands in shared memory as [128 pixels][128 B] with the 128-byte swizzle = the
//                canonical K-major SWIZZLE_128B UMMA layout.
//   B operand  : packed weights [Cout][Ktot] (K-major), 2-D TMA box (32, BN), same swizzle.
//   Skip 1x1   : extra K-blocks read the raw (un-normalised) ResBlock input(s) at the centre tap, so
//                conv3x3(h) + skip_connection(x) (+ torch.cat of the up path) is ONE accumulation in TMEM.
//   Epilogue   : tcgen05.ld -> registers -> + bias + timestep-embedding row + residual (same-res,
//                avg-pooled or nearest-upsampled) -> NHWC fp32 store.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected
// lane), warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4).  Two CTAs are resident per SM so one
// CTA's prologue/epilogue overlaps the other's main loop.
//
// Replaces nn.Conv2d / nn.Conv1d call sites of reference UNet.py:115-117,172,193,200 (cuDNN TF32 on GPU).
#include "common.cuh"
#include "anoddpm_b200.h"
#include <cuda.h>

namespace {

constexpr int kBM = 128;          // pixels per tile (UMMA M)
constexpr int kBK = 32;           // fp32 elements per K-block (= 128 B swizzle span)
constexpr int kUmmaK = 8;         // tf32: 32 B per MMA along K
constexpr int kStages = 3;
constexpr int kThreads = 192;
constexpr int kABytes = kBM * kBK * 4;  // 16 KiB

struct TcArgs {
    float* out;
    const float* bias;
    const float* emb;
    const float* res;
    int M, H, W, Cout;
    int emb_stride, res_mode;
    int TW, TH;          // spatial tile; TB = 128 / (TW*TH)
    int nkb_a, chunks_a, ks;   // K-blocks of the main source (= taps * Ca/32), chunks per tap, kernel size
    int nkb_r1, nkb_r2;
    int BN;              // N tile (UMMA N)
    int tmem_cols;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void tma_load_4d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
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

__global__ void __launch_bounds__(kThreads, 2)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_r1,
               const __grid_constant__ CUtensorMap map_r2, const __grid_constant__ CUtensorMap map_w, const TcArgs P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment is required by SWIZZLE_128B; align manually in case the runtime base is weaker
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int b_bytes = P.BN * kBK * 4;
    uint8_t* sA = smem;
    uint8_t* sB = smem + kStages * kABytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + kStages * b_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kStages;
    uint64_t* tmem_full = bars + 2 * kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * P.BN;
    const int nkb = P.nkb_a + P.nkb_r1 + P.nkb_r2;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)P.tmem_cols)
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
            const int w0 = m0 % P.W, h0 = (m0 / P.W) % P.H, b0 = m0 / (P.W * P.H);
            const int pad = P.ks / 2;
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStages;
                mbar_wait(&empty[s], ((kb / kStages) & 1) ^ 1);
                mbar_expect_tx(&full[s], (uint32_t)(kABytes + b_bytes));
                const CUtensorMap* am;
                int c0, dy = 0, dx = 0;
                if (kb < P.nkb_a) {
                    const int tap = kb / P.chunks_a;
                    c0 = (kb - tap * P.chunks_a) * kBK;
                    dy = tap / P.ks - pad;
                    dx = tap % P.ks - pad;
                    am = &map_a;
                } else if (kb < P.nkb_a + P.nkb_r1) {
                    c0 = (kb - P.nkb_a) * kBK;
                    am = &map_r1;
                } else {
                    c0 = (kb - P.nkb_a - P.nkb_r1) * kBK;
                    am = &map_r2;
                }
                tma_load_4d(sA + s * kABytes, am, &full[s], c0, w0 + dx, h0 + dy, b0);
                tma_load_2d(sB + s * b_bytes, &map_w, &full[s], kb * kBK, n0);
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(P.BN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % kStages;
            mbar_wait(&full[s], (kb / kStages) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint64_t ad = make_sdesc(smem_u32(sA + s * kABytes));
                const uint64_t bd = make_sdesc(smem_u32(sB + s * b_bytes));
#pragma unroll
                for (int k = 0; k < kBK / kUmmaK; ++k)  // +32 B along K inside the swizzle atom = +2 in the address field
                    umma_tf32(tmem_base, ad + 2 * k, bd + 2 * k, idesc, (kb | k) ? 1u : 0u);
                umma_commit(&empty[s]);                  // frees the smem stage when these MMAs retire
                if (kb == nkb - 1) umma_commit(tmem_full);
            }
            __syncwarp();
        }
    } else {
        // ---------------- epilogue: TMEM -> registers -> global (NHWC) ----------------
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int q = m0 + row;
        const bool valid = q < P.M;
        int w = 0, h = 0, b = 0;
        if (valid) { w = q % P.W; h = (q / P.W) % P.H; b = q / (P.W * P.H); }
        for (int cb = 0; cb < P.BN; cb += 32) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)cb;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (!valid) continue;
            const int nb = n0 + cb;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const int n = nb + j;
                if (n >= P.Cout) break;
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    v[e] = __uint_as_float(r[j + e]);
                    if (P.bias) v[e] += P.bias[n + e];
                    if (P.emb) v[e] += P.emb[(size_t)b * P.emb_stride + n + e];
                }
                if (P.res_mode == 1) {
                    const float4 t = *reinterpret_cast<const float4*>(P.res + (size_t)q * P.Cout + n);
                    v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
                } else if (P.res_mode == 2) {
                    const int H2 = P.H * 2, W2 = P.W * 2;
                    const float* rb = P.res + (((size_t)b * H2 + 2 * h) * W2 + 2 * w) * P.Cout + n;
                    const float4 r0 = *reinterpret_cast<const float4*>(rb);
                    const float4 r1 = *reinterpret_cast<const float4*>(rb + P.Cout);
                    const float4 r2 = *reinterpret_cast<const float4*>(rb + (size_t)W2 * P.Cout);
                    const float4 r3 = *reinterpret_cast<const float4*>(rb + (size_t)W2 * P.Cout + P.Cout);
                    v[0] += (r0.x + r1.x + r2.x + r3.x) * 0.25f;
                    v[1] += (r0.y + r1.y + r2.y + r3.y) * 0.25f;
                    v[2] += (r0.z + r1.z + r2.z + r3.z) * 0.25f;
                    v[3] += (r0.w + r1.w + r2.w + r3.w) * 0.25f;
                } else if (P.res_mode == 3) {
                    const int Hh = P.H / 2, Wh = P.W / 2;
                    const float4 t = *reinterpret_cast<const float4*>(
                        P.res + (((size_t)b * Hh + (h >> 1)) * Wh + (w >> 1)) * P.Cout + n);
                    v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
                }
                *reinterpret_cast<float4*>(P.out + (size_t)q * P.Cout + n) = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)P.tmem_cols)
                     : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
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

// NHWC activation map: dims (C, W, H, B), box (32, TW, TH, TB)
int make_act_map(CUtensorMap* m, const float* p, int C, int W, int H, int B, int TW, int TH, int TB) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { anoddpm_set_error("cuTensorMapEncodeTiled entry point not available"); return ANODDPM_ERR_CUDA; }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TB};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(p), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        anoddpm_set_error("cuTensorMapEncodeTiled(act C=%d W=%d H=%d B=%d box %d,%d,%d) failed: %d", C, W, H, B, TW, TH,
                          TB, (int)r);
        return ANODDPM_ERR_CUDA;
    }
    return ANODDPM_OK;
}

int make_w_map(CUtensorMap* m, const float* p, int Ktot, int Cout, int BN) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { anoddpm_set_error("cuTensorMapEncodeTiled entry point not available"); return ANODDPM_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)Ktot * 4};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        anoddpm_set_error("cuTensorMapEncodeTiled(weights K=%d Cout=%d BN=%d) failed: %d", Ktot, Cout, BN, (int)r);
        return ANODDPM_ERR_CUDA;
    }
    return ANODDPM_OK;
}

}  // namespace

// Returns 1 if the tensor-core tiling covers this problem, 0 otherwise (caller uses the exact kernel).
extern "C" int anoddpm_conv_tc_supported(const anoddpm_conv_desc* d) {
    if (!d) return 0;
    if (d->Ca % kBK || d->Cr1 % kBK || d->Cr2 % kBK) return 0;
    if (d->Cout % 16) return 0;
    const int TW = d->W < 128 ? d->W : 128;
    if (128 % TW || d->W % TW) return 0;
    int TH = 128 / TW;
    if (TH > d->H) TH = d->H;
    if (d->H % TH || 128 % (TW * TH)) return 0;
    const int TB = 128 / (TW * TH);
    if (TB > 1 && TH != d->H) return 0;
    if (TB > 256) return 0;
    return 1;
}

int anoddpm_conv_tc(const anoddpm_conv_desc* d, void* stream) {
    ADP_REQUIRE(anoddpm_conv_tc_supported(d), "anoddpm_conv(impl=1): shape not covered by the tensor-core tiling "
                "(Ca=%d Cr1=%d Cr2=%d Cout=%d H=%d W=%d); use impl=0", d->Ca, d->Cr1, d->Cr2, d->Cout, d->H, d->W);
    TcArgs P;
    P.out = static_cast<float*>(d->out); P.bias = d->bias; P.emb = d->emb; P.res = d->res;
    P.M = d->B * d->H * d->W; P.H = d->H; P.W = d->W; P.Cout = d->Cout;
    P.emb_stride = d->emb_stride; P.res_mode = d->res_mode;
    P.TW = d->W < 128 ? d->W : 128;
    P.TH = 128 / P.TW; if (P.TH > d->H) P.TH = d->H;
    const int TB = 128 / (P.TW * P.TH);
    P.ks = d->ksize;
    P.chunks_a = d->Ca / kBK;
    P.nkb_a = d->ksize * d->ksize * P.chunks_a;
    P.nkb_r1 = d->Cr1 / kBK;
    P.nkb_r2 = d->Cr2 / kBK;
    P.BN = d->Cout < 128 ? d->Cout : 128;
    P.tmem_cols = P.BN <= 32 ? 32 : (P.BN <= 64 ? 64 : 128);
    const int Ktot = (P.nkb_a + P.nkb_r1 + P.nkb_r2) * kBK;

    CUtensorMap ma, m1, m2, mw;
    int rc;
    if ((rc = make_act_map(&ma, static_cast<const float*>(d->a), d->Ca, d->W, d->H, d->B, P.TW, P.TH, TB))) return rc;
    m1 = ma; m2 = ma;
    if (d->r1 && (rc = make_act_map(&m1, static_cast<const float*>(d->r1), d->Cr1, d->W, d->H, d->B, P.TW, P.TH, TB))) return rc;
    if (d->r2 && (rc = make_act_map(&m2, static_cast<const float*>(d->r2), d->Cr2, d->W, d->H, d->B, P.TW, P.TH, TB))) return rc;
    if ((rc = make_w_map(&mw, static_cast<const float*>(d->w), Ktot, d->Cout, P.BN))) return rc;

    const size_t smem = (size_t)kStages * (kABytes + P.BN * kBK * 4) + (2 * kStages + 1) * 8 + 16 + 1024;
    static bool attr_dev[64] = {};   // the attribute is per device
    int dev_ = 0;
    ADP_CUDA(cudaGetDevice(&dev_));
    if (!attr_dev[dev_ & 63]) {
        ADP_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        attr_dev[dev_ & 63] = true;
    }
    dim3 grid(adp_ceil_div(P.M, kBM), adp_ceil_div(d->Cout, P.BN));
    conv_tc_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(ma, m1, m2, mw, P);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}
