// anoddpm_b200/csrc/conv_tc2.cu -- persistent, CTA-pair tensor-core convolution for sm_100a (impl = 2).
//
// Same contract and operand layouts as conv_tc.cu (implicit GEMM, NHWC activations via 4-D TMA boxes at
// tap-shifted coordinates, packed K-major weights, kind::tf32, fp32 accumulators in TMEM), restructured for
// tensor-pipe utilisation:
//   * a cluster of 2 CTAs (one per SM of a TPC) computes a 256-pixel x BN tile with
//     tcgen05.mma.cta_group::2 (UMMA M = 256): each CTA stages its own 128 pixel rows of A and only HALF of
//     the weight tile, so shared-memory operand bandwidth per SM drops from 128 to 96 (BN=128) / 64 (BN=256)
//     bytes per clock -- the single-CTA M=128,N=128 shape is pinned at the 128 B/clk smem limit;
//   * persistent: grid = min(#tiles, 74) clusters, static round-robin tile schedule, deep smem ring shared
//     across tiles, so the TMA pipeline never drains between tiles;
//   * two TMEM accumulator stages: the epilogue of tile i (TMEM -> registers -> +bias/+emb/+residual ->
//     global) overlaps the MMAs of tile i+1.
// Warp roles per CTA (256 threads): 0 = TMA producer, 1 = MMA issuer (leader CTA only), 2 = TMEM allocator,
// 4..7 = epilogue.  Barrier protocol follows the canonical sm_100 2-CTA GEMM: `full` barriers live in the
// leader CTA (both producers signal it), `empty` / `tmem_full` are multicast to both CTAs by tcgen05.commit,
// `tmem_empty` lives in the leader and collects all 2 x 128 epilogue threads.
#include "common.cuh"
#include "anoddpm_b200.h"
#include <cuda.h>

namespace {

constexpr int kBM = 128;
constexpr int kRowBytes = 128;   // one K-block row: 32 fp32 (TF32 MMA) or 64 fp16 elements
constexpr int kKSteps = 4;       // MMAs per K-block: 32 B of K each (8 tf32 / 16 f16 elements)
constexpr int kThreads = 256;
constexpr int kABytes = kBM * kRowBytes;
constexpr int kAccStages = 2;
constexpr int kMaxStages = 8;
constexpr int kEpiTileBytes = 4 * 32 * 33 * 4;   // per epilogue warp a 32 x 33 fp32 transpose tile

struct Tc2Args {
    float* out;
    const float* bias;
    const float* emb;
    const float* res;
    int M, H, W, Cout;
    int emb_stride, res_mode;
    int nkb_a, chunks_a, ks;
    int nkb_r1, nkb_r2;
    int BN;            // pair-tile N (UMMA N); each CTA stages BN/2 weight rows
    int tmem_cols;
    int stages;
    int ke;                    // elements per 128-byte K-block row: 32 (fp32 operands) or 64 (fp16 operands)
    int n_tiles_n, n_tiles;   // tiles along N, total pair tiles
    double* gn_stats;          // optional [REPLICAS][B][Cout][2]: per-channel (sum, sum of squares) of the output
    long long stat_replica_stride;
    // split-K (small maps: fewer tiles than CTA pairs): a work item = (tile, K range); partial accumulators go to `ws_part`
    // ([splits][M_pad][ws_ld] fp32) and the LAST warp to finish a 32-row block (arrival counter) adds the partials in split
    // order (deterministic) and runs the normal epilogue on the sum
    int splits;
    float* ws_part;
    int* ws_cnt;               // [n_tiles][2 ranks][4 warps], zero on entry, reset by the last arriver
    int ws_ld;                 // n_tiles_n * BN
    long long ws_split_stride; // M_pad * ws_ld
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// arrive (no tx) on the barrier at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t"
        ".reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(cta)
        : "memory");
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
// 2-CTA TMA loads: data lands in THIS CTA's smem, the transaction bytes are signalled on the LEADER's barrier
// (peer bit of the barrier address cleared).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma2_load_4d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                             int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
        : "memory");
}
template <bool F16>
__device__ __forceinline__ void umma2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                      uint32_t accumulate) {
    if (F16) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
            "}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// commit all prior MMAs of this thread; arrive on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

template <bool F16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_r1,
                const __grid_constant__ CUtensorMap map_r2, const __grid_constant__ CUtensorMap map_w,
                const Tc2Args P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int b_bytes = (P.BN / 2) * kRowBytes;
    const int S = P.stages;
    uint8_t* sA = smem;
    uint8_t* sB = smem + S * kABytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + S * b_bytes);
    uint64_t* full = bars;                       // [S]   leader only
    uint64_t* empty = bars + kMaxStages;         // [S]   both CTAs (multicast commit)
    uint64_t* tmem_full = bars + 2 * kMaxStages; // [2]   both CTAs (multicast commit)
    uint64_t* tmem_empty = tmem_full + kAccStages;  // [2] leader only, 2 x 128 arrivals
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + kAccStages);
    float* epi_tiles = reinterpret_cast<float*>(tmem_slot + 4);   // [4 epilogue warps][32][33]: accumulator chunk transpose

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int nkb = P.nkb_a + P.nkb_r1 + P.nkb_r2;
    const int n_items = P.n_tiles * P.splits;

    cluster_sync_all();   // both CTAs resident before the pair-wide TMEM allocation
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(&full[i], 2); mbar_init(&empty[i], 1); }
        for (int i = 0; i < kAccStages; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 2 * 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)P.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------- TMA producer (both CTAs; own A rows + own half of the weight tile) ----------------
        if (elect_one()) {
            const int pad = P.ks / 2;
            int kbc = 0;
            for (int item = cluster_id; item < n_items; item += n_clusters) {
                const int tile = item / P.splits, sp = item - tile * P.splits;
                const int kb0 = (int)((long long)sp * nkb / P.splits), kb1 = (int)((long long)(sp + 1) * nkb / P.splits);
                const int mp = tile / P.n_tiles_n, nt = tile - mp * P.n_tiles_n;
                const int m0 = (mp * 2 + (int)rank) * kBM;
                const int n0 = nt * P.BN + (int)rank * (P.BN / 2);
                const int w0 = m0 % P.W, h0 = (m0 / P.W) % P.H, b0 = m0 / (P.W * P.H);
                for (int kb = kb0; kb < kb1; ++kb, ++kbc) {
                    const int s = kbc % S;
                    mbar_wait(&empty[s], ((kbc / S) & 1) ^ 1);
                    if (leader) mbar_expect_tx(&full[s], (uint32_t)(2 * (kABytes + b_bytes)));
                    else mbar_arrive_remote(&full[s], 0);
                    const CUtensorMap* am;
                    int c0, dy = 0, dx = 0;
                    if (kb < P.nkb_a) {
                        const int tap = kb / P.chunks_a;
                        c0 = (kb - tap * P.chunks_a) * P.ke;
                        dy = tap / P.ks - pad;
                        dx = tap % P.ks - pad;
                        am = &map_a;
                    } else if (kb < P.nkb_a + P.nkb_r1) {
                        c0 = (kb - P.nkb_a) * P.ke;
                        am = &map_r1;
                    } else {
                        c0 = (kb - P.nkb_a - P.nkb_r1) * P.ke;
                        am = &map_r2;
                    }
                    tma2_load_4d(sA + s * kABytes, am, &full[s], c0, w0 + dx, h0 + dy, b0);
                    tma2_load_2d(sB + s * b_bytes, &map_w, &full[s], kb * P.ke, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (leader CTA only) ----------------
        if (leader) {
            constexpr uint32_t fmt = F16 ? 0u : 2u;   // a/b format: F16 = 0, TF32 = 2 ; c format F32 = 1
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(P.BN >> 3) << 17) |
                                   ((uint32_t)(256 >> 4) << 24);
            int kbc = 0, it = 0;
            for (int item = cluster_id; item < n_items; item += n_clusters, ++it) {
                const int sp = item % P.splits;
                const int kb0 = (int)((long long)sp * nkb / P.splits), kb1 = (int)((long long)(sp + 1) * nkb / P.splits);
                const int acc = it & 1;
                mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * P.BN);
                for (int kb = kb0; kb < kb1; ++kb, ++kbc) {
                    const int s = kbc % S;
                    mbar_wait(&full[s], (kbc / S) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
                        const uint64_t ad = make_sdesc(smem_u32(sA + s * kABytes));
                        const uint64_t bd = make_sdesc(smem_u32(sB + s * b_bytes));
#pragma unroll
                        for (int k = 0; k < kKSteps; ++k)
                            umma2<F16>(tmem_d, ad + 2 * k, bd + 2 * k, idesc, ((kb - kb0) | k) ? 1u : 0u);
                        umma2_commit(&empty[s]);
                        if (kb == kb1 - 1) umma2_commit(&tmem_full[acc]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 4) {
        // ---------------- epilogue (both CTAs): TMEM -> registers -> global (NHWC) ----------------
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        float* tbuf = epi_tiles + quarter * (32 * 33);
        int it = 0;
        for (int item = cluster_id; item < n_items; item += n_clusters, ++it) {
            const int tile = item / P.splits, sp = item - tile * P.splits;
            const int acc = it & 1;
            const int mp = tile / P.n_tiles_n, nt = tile - mp * P.n_tiles_n;
            const int m0 = (mp * 2 + (int)rank) * kBM;
            const int n0 = nt * P.BN;
            const int q = m0 + row;
            // the shared tail of both modes: bias / embedding / residual, store, statistics of 32 accumulator columns
            // Through a per-warp 32 x 33 shared-memory tile the accumulator chunk changes from lane = row (TMEM) to lane = column:
            // every global access of the tail is then one contiguous 128-byte row segment (the lane-per-row form touched 32
            // lines with 16 bytes each per instruction and left the 1x1 convolutions of the attention blocks store-bound at
            // 0.7 TB/s), bias / embedding are per-lane scalars, and a column's statistics are a plain register sum over the rows.
            const int HW = P.H * P.W;
            const int q0 = m0 + quarter * 32;            // first row of this warp
            const bool wvalid = q0 < P.M;
            const int b0 = wvalid ? q0 / HW : 0;
            const int rb = (b0 + 1) * HW - q0;           // rows >= rb belong to sample b0 + 1
            const int rows = P.M - q0 < 32 ? P.M - q0 : 32;
            auto finish = [&](uint32_t (&r)[32], int cb) {
                const int n = n0 + cb + lane;
                const bool cvalid = n < P.Cout;
#pragma unroll
                for (int j = 0; j < 32; ++j) tbuf[lane * 33 + j] = __uint_as_float(r[j]);
                __syncwarp();
                float add0 = 0.f, add1 = 0.f;
                if (cvalid) {
                    if (P.bias) add0 = add1 = P.bias[n];
                    if (P.emb) {
                        add0 += P.emb[(size_t)b0 * P.emb_stride + n];
                        if (rb < rows) add1 += P.emb[(size_t)(b0 + 1) * P.emb_stride + n];
                    }
                }
                float s1 = 0.f, s2 = 0.f;
                if (cvalid) {
                    // residual values of the 32 rows first: 32 independent loads in flight instead of one latency per row
                    float rv[32];
                    if (P.res_mode == 1) {
#pragma unroll
                        for (int rr = 0; rr < 32; ++rr) rv[rr] = rr < rows ? __ldg(P.res + (size_t)(q0 + rr) * P.Cout + n) : 0.f;
                    } else if (P.res_mode == 2 || P.res_mode == 3) {
#pragma unroll
                        for (int rr = 0; rr < 32; ++rr) {
                            float t = 0.f;
                            if (rr < rows) {
                                const int br = rr < rb ? b0 : b0 + 1;
                                const int pix = q0 + rr - br * HW, hr = pix / P.W, wr = pix - hr * P.W;
                                if (P.res_mode == 2) {
                                    const int H2 = P.H * 2, W2 = P.W * 2;
                                    const float* rp = P.res + (((size_t)br * H2 + 2 * hr) * W2 + 2 * wr) * P.Cout + n;
                                    t = (rp[0] + rp[P.Cout] + rp[(size_t)W2 * P.Cout] + rp[(size_t)W2 * P.Cout + P.Cout]) * 0.25f;
                                } else {
                                    const int Hh = P.H / 2, Wh = P.W / 2;
                                    t = P.res[(((size_t)br * Hh + (hr >> 1)) * Wh + (wr >> 1)) * P.Cout + n];
                                }
                            }
                            rv[rr] = t;
                        }
                    } else {
#pragma unroll
                        for (int rr = 0; rr < 32; ++rr) rv[rr] = 0.f;
                    }
#pragma unroll
                    for (int rr = 0; rr < 32; ++rr) {
                        if (rr < rows) {
                            const float v = (tbuf[rr * 33 + lane] + (rr < rb ? add0 : add1)) + rv[rr];
                            P.out[(size_t)(q0 + rr) * P.Cout + n] = v;
                            s1 += v;
                            s2 += v * v;
                        }
                    }
                    if (P.gn_stats && wvalid) {   // the 32 rows of a warp belong to one sample (host checks H*W % 32 == 0)
                        double* st = P.gn_stats + (size_t)(cluster_id % ANODDPM_STAT_REPLICAS) * P.stat_replica_stride +
                                     ((size_t)b0 * P.Cout + n) * 2;
                        atomicAdd(st, (double)s1);
                        atomicAdd(st + 1, (double)s2);
                    }
                }
                __syncwarp();      // the tile is rewritten by the next chunk
            };
            auto load_acc = [&](uint32_t (&r)[32], int cb) {
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * P.BN + cb);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
                      "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
                      "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
                      "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (cb + 32 >= P.BN) {
                    // accumulator fully read: hand the TMEM stage back to the MMA issuer (leader's barrier)
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive_remote(&tmem_empty[acc], 0);
                }
            };
            mbar_wait(&tmem_full[acc], (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (P.splits == 1) {
                for (int cb = 0; cb < P.BN; cb += 32) {
                    uint32_t r[32];
                    load_acc(r, cb);
                    finish(r, cb);
                }
            } else {
                // partial accumulators of this K range -> workspace (thread = row: 128 contiguous bytes per 32 columns)
                float* wrow = P.ws_part + (size_t)sp * P.ws_split_stride + (size_t)q * P.ws_ld + n0;
                for (int cb = 0; cb < P.BN; cb += 32) {
                    uint32_t r[32];
                    load_acc(r, cb);
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        __stcg(reinterpret_cast<float4*>(wrow + cb + j), make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                                      __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
                }
                __threadfence();
                __syncwarp();
                int last = 0;
                if (lane == 0) {
                    int* cnt = P.ws_cnt + ((size_t)tile * 2 + rank) * 4 + quarter;
                    last = atomicAdd(cnt, 1) == P.splits - 1;
                    if (last) { *cnt = 0; __threadfence(); }      // ready for the next launch; orders the partial reads below
                }
                last = __shfl_sync(0xffffffffu, last, 0);
                if (last) {
                    const float* prow = P.ws_part + (size_t)q * P.ws_ld + n0;
                    for (int cb = 0; cb < P.BN; cb += 32) {
                        uint32_t r[32];
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 a4 = __ldcg(reinterpret_cast<const float4*>(prow + cb + j));
                            for (int s2 = 1; s2 < P.splits; ++s2) {
                                const float4 t4 = __ldcg(reinterpret_cast<const float4*>(prow + (size_t)s2 * P.ws_split_stride + cb + j));
                                a4.x += t4.x; a4.y += t4.y; a4.z += t4.z; a4.w += t4.w;
                            }
                            r[j] = __float_as_uint(a4.x); r[j + 1] = __float_as_uint(a4.y);
                            r[j + 2] = __float_as_uint(a4.z); r[j + 3] = __float_as_uint(a4.w);
                        }
                        finish(r, cb);
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();   // all MMAs committed + all remote barrier arrivals landed before either CTA tears down
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)P.tmem_cols)
                     : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode2() {
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

int make_act_map2(CUtensorMap* m, const void* p, int f16, int C, int W, int H, int B, int TW, int TH, int TB) {
    EncodeTiledFn enc = get_encode2();
    if (!enc) { anoddpm_set_error("cuTensorMapEncodeTiled entry point not available"); return ANODDPM_ERR_CUDA; }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t es = f16 ? 2 : 4;
    cuuint64_t strides[3] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es};
    cuuint32_t box[4] = {(cuuint32_t)(f16 ? 64 : 32), (cuuint32_t)TW, (cuuint32_t)TH, (cuuint32_t)TB};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(p), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        anoddpm_set_error("cuTensorMapEncodeTiled(act C=%d W=%d H=%d B=%d box %d,%d,%d) failed: %d", C, W, H, B, TW, TH,
                          TB, (int)r);
        return ANODDPM_ERR_CUDA;
    }
    return ANODDPM_OK;
}

int make_w_map2(CUtensorMap* m, const void* p, int f16, int Ktot, int Cout, int rows) {
    EncodeTiledFn enc = get_encode2();
    if (!enc) { anoddpm_set_error("cuTensorMapEncodeTiled entry point not available"); return ANODDPM_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)Cout};
    const cuuint64_t es = f16 ? 2 : 4;
    cuuint64_t strides[1] = {(cuuint64_t)Ktot * es};
    cuuint32_t box[2] = {(cuuint32_t)(f16 ? 64 : 32), (cuuint32_t)rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(p), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        anoddpm_set_error("cuTensorMapEncodeTiled(weights K=%d Cout=%d rows=%d) failed: %d", Ktot, Cout, rows, (int)r);
        return ANODDPM_ERR_CUDA;
    }
    return ANODDPM_OK;
}

// Pair-tile N for a problem of m_pairs 256-row tiles, Cout columns and nkb K-blocks.  The kernel is bound by the per-SM L2->SM
// operand stream (one 16 KB A tile + BN/2 weight rows of 128 B per K-block and CTA), so a small map is fastest when its tiles
// cover the 74 CTA pairs even if that means narrow tiles: 8x8 maps at bs 32 (8 row tiles) run 16 tiles with BN = 256 and 64
// with BN = 64.  cost = rounds x (K-blocks x bytes per K-block + a fixed per-tile term), ties to the wider tile.
int tc2_pick_bn(int Cout, int m_pairs, int nkb) {
    if (Cout < 128) return Cout;
    int best = 0;
    long long best_cost = 0;
    for (int bn = 256; bn >= 32; bn >>= 1) {
        if (Cout % bn) continue;
        const long long tiles = (long long)m_pairs * (Cout / bn);
        const long long rounds = (tiles + 73) / 74;
        const long long cost = rounds * ((long long)nkb * (kABytes + (bn / 2) * kRowBytes) + 48 * 1024 + (long long)bn * 512);
        if (!best || cost < best_cost) { best = bn; best_cost = cost; }
    }
    return best ? best : 128;
}

// K splits for a problem of n_tiles pair tiles and nkb K-blocks: fill the 74 CTA pairs, at least 8 K-blocks per split
int tc2_splits(int n_tiles, int nkb) {
    if (n_tiles * 2 > 74 || nkb < 16) return 1;
    int s = 74 / n_tiles;
    if (s > nkb / 8) s = nkb / 8;
    if (s > 8) s = 8;
    return s < 1 ? 1 : s;
}

}  // namespace

extern "C" int anoddpm_conv_tc_supported(const anoddpm_conv_desc* d);

// bytes of anoddpm_conv_desc::workspace that let impl 2 split the K loop of this problem (0: the problem is not split)
extern "C" long long anoddpm_conv_workspace_bytes(const anoddpm_conv_desc* d) {
    if (!d || d->impl != 2 || !anoddpm_conv_tc_supported(d)) return 0;
    const int ke = d->operand_f16 ? 64 : 32;
    const int m_pairs = adp_ceil_div((long long)d->B * d->H * d->W, 2 * kBM);
    const int nkb = (d->ksize * d->ksize * d->Ca + d->Cr1 + d->Cr2) / ke;
    const int BN = tc2_pick_bn(d->Cout, m_pairs, nkb);
    const int n_tiles_n = adp_ceil_div(d->Cout, BN);
    const int n_tiles = m_pairs * n_tiles_n;
    const int splits = tc2_splits(n_tiles, nkb);
    if (splits == 1) return 0;
    const long long cnt_bytes = (((long long)n_tiles * 8 * (long long)sizeof(int)) + 255) / 256 * 256;
    return cnt_bytes + (long long)splits * m_pairs * 2 * kBM * n_tiles_n * BN * (long long)sizeof(float);
}

int anoddpm_conv_tc2(const anoddpm_conv_desc* d, void* stream) {
    ADP_REQUIRE(anoddpm_conv_tc_supported(d), "anoddpm_conv(impl=2): shape not covered by the tensor-core tiling "
                "(Ca=%d Cr1=%d Cr2=%d Cout=%d H=%d W=%d); use impl=0", d->Ca, d->Cr1, d->Cr2, d->Cout, d->H, d->W);
    Tc2Args P;
    P.out = static_cast<float*>(d->out); P.bias = d->bias; P.emb = d->emb; P.res = d->res;
    P.M = d->B * d->H * d->W; P.H = d->H; P.W = d->W; P.Cout = d->Cout;
    P.emb_stride = d->emb_stride; P.res_mode = d->res_mode;
    ADP_REQUIRE(!d->gn_stats || (d->H * d->W) % 32 == 0, "anoddpm_conv(impl=2): fused statistics need H*W %% 32 == 0");
    P.gn_stats = d->gn_stats;
    P.stat_replica_stride = (long long)d->B * d->Cout * 2;
    const int TW = d->W < 128 ? d->W : 128;
    int TH = 128 / TW; if (TH > d->H) TH = d->H;
    const int TB = 128 / (TW * TH);
    P.ks = d->ksize;
    const int f16 = d->operand_f16 ? 1 : 0;
    P.ke = f16 ? 64 : 32;
    ADP_REQUIRE(d->Ca % P.ke == 0 && d->Cr1 % P.ke == 0 && d->Cr2 % P.ke == 0,
                "anoddpm_conv(impl=2): fp16 operands need channel counts that are multiples of 64");
    P.chunks_a = d->Ca / P.ke;
    P.nkb_a = d->ksize * d->ksize * P.chunks_a;
    P.nkb_r1 = d->Cr1 / P.ke;
    P.nkb_r2 = d->Cr2 / P.ke;
    const int m_pairs = adp_ceil_div(P.M, 2 * kBM);
    P.BN = tc2_pick_bn(d->Cout, m_pairs, P.nkb_a + P.nkb_r1 + P.nkb_r2);
    const int acc_cols = kAccStages * P.BN;
    P.tmem_cols = acc_cols <= 32 ? 32 : acc_cols <= 64 ? 64 : acc_cols <= 128 ? 128 : acc_cols <= 256 ? 256 : 512;
    const int Ktot = (P.nkb_a + P.nkb_r1 + P.nkb_r2) * P.ke;
    const int b_bytes = (P.BN / 2) * kRowBytes;
    int stages = (192 * 1024) / (kABytes + b_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    P.stages = stages;
    P.n_tiles_n = adp_ceil_div(d->Cout, P.BN);
    P.n_tiles = m_pairs * P.n_tiles_n;
    // split-K when the tiles do not fill the machine (8x8 maps at bs 32: 16 tiles for 74 CTA pairs) and the caller gave a
    // workspace: each (tile, K range) item is one CTA-pair job, the last arriver of a 32-row block reduces + finishes it
    P.splits = tc2_splits(P.n_tiles, P.nkb_a + P.nkb_r1 + P.nkb_r2);
    P.ws_ld = P.n_tiles_n * P.BN;
    P.ws_split_stride = (long long)m_pairs * 2 * kBM * P.ws_ld;
    const long long cnt_bytes = (((long long)P.n_tiles * 8 * (long long)sizeof(int)) + 255) / 256 * 256;
    if (P.splits > 1 && (!d->workspace || d->workspace_bytes < cnt_bytes + P.splits * P.ws_split_stride * (long long)sizeof(float)))
        P.splits = 1;
    P.ws_cnt = static_cast<int*>(d->workspace);
    P.ws_part = P.splits > 1 ? reinterpret_cast<float*>(static_cast<char*>(d->workspace) + cnt_bytes) : nullptr;

    CUtensorMap ma, m1, m2, mw;
    int rc;
    if ((rc = make_act_map2(&ma, d->a, f16, d->Ca, d->W, d->H, d->B, TW, TH, TB))) return rc;
    m1 = ma; m2 = ma;
    if (d->r1 && (rc = make_act_map2(&m1, d->r1, f16, d->Cr1, d->W, d->H, d->B, TW, TH, TB))) return rc;
    if (d->r2 && (rc = make_act_map2(&m2, d->r2, f16, d->Cr2, d->W, d->H, d->B, TW, TH, TB))) return rc;
    if ((rc = make_w_map2(&mw, d->w, f16, Ktot, d->Cout, P.BN / 2))) return rc;

    const size_t smem = (size_t)stages * (kABytes + b_bytes) + (2 * kMaxStages + 2 * kAccStages) * 8 + 16 + kEpiTileBytes + 1024;
    static bool attr_dev[64] = {};   // the attribute is per device
    int dev_ = 0;
    ADP_CUDA(cudaGetDevice(&dev_));
    if (!attr_dev[dev_ & 63]) {
        ADP_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ADP_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_dev[dev_ & 63] = true;
    }
    int sms = 148;
    const int n_items = P.n_tiles * P.splits;
    int n_clusters = n_items < sms / 2 ? n_items : sms / 2;
    if (f16) conv_tc2_kernel<true><<<dim3(2 * n_clusters), kThreads, smem, (cudaStream_t)stream>>>(ma, m1, m2, mw, P);
    else conv_tc2_kernel<false><<<dim3(2 * n_clusters), kThreads, smem, (cudaStream_t)stream>>>(ma, m1, m2, mw, P);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}
