// anoddpm_b200/csrc/conv_tc3.cu -- halo-reuse tensor-core 3x3 convolution for sm_100a (impl = 3).
//
// conv_tc2 is L2-bandwidth bound: every filter tap re-reads its own 128-pixel A tile and every tile re-reads the
// weights, ~96 B/clk/SM against a measured L2->SM ceiling of ~40-48 B/clk/SM (ncu: profiles/).  This kernel cuts the
// operand traffic per MMA cycle to ~39 B/clk/SM:
//   * spatial tiling: an M-subtile is 16 rows x 8 pixels of one image.  For each 32-channel chunk and each
//     horizontal tap dx, ONE TMA box (32 ch, 8 px, 18 rows) is loaded; the three vertical taps dy are UMMA
//     descriptors into the same box advanced by dy*8 rows (= dy*1024 B, swizzle-atom aligned) -> A traffic / 3;
//   * each CTA owns TWO subtiles (a 16x16 region) that share every weight tile -> B traffic / 2 per flop;
//   * CTA pairs (tcgen05.mma.cta_group::2, UMMA M = 256) split the weight tile in halves, persistent clusters,
//     3-stage smem ring (60 KB/stage), 2 TMEM accumulator stages x 2 subtiles x BN columns (512 columns).
// 1x1 "skip" sources (raw ResBlock input, centre tap only) are extra stages with a (32, 8, 16) box.
// Single-box main loop (default when there are no skip sources): tcgen05.mma swizzles on absolute shared-memory address
// bits (profiles/r01_umma_desc_probe.md), so ONE (K-block, 18 px, 18 rows) box per channel chunk serves all nine taps of
// both subtiles -- tap (dy, dx) of subtile s is the descriptor that starts at row 18 dy + 8 s + dx with 8-row groups
// 18 rows (2304 B) apart.  Two A buffers (41.5 KB each) + a 4-stage ring of (chunk, dy) weight stages; A-operand
// L2->SM traffic and TMA writes drop by 62 % against the box-per-(subtile, dx) loop.
// Epilogue (8 warps: TMEM lane quarter x subtile; two per scheduler -- the work is a chain of short-latency steps), in
// steps of (32 pixels x 16 channels) per warp through two 2 KB shared-memory buffers in the TMA SWIZZLE_64B layout:
// the residual box of step s+2 is TMA-loaded into the buffer once the store of step s has read it, the
// thread-per-pixel pass adds bias + timestep embedding + residual to the TMEM row and writes it back in place, a TMA
// store moves the buffer to the NHWC output, and a column pass over the buffer accumulates optional per-channel
// statistics of the OUTPUT (sum, sum of squares per (sample, channel)): fp32 in registers while the cluster stays on
// one sample (contiguous tile runs per cluster), flushed with fp64 atomics, so the GroupNorm(s) that consume the tensor
// -- alone or concatenated with another tensor -- need no statistics pass.
// Residual modes: 1 = same resolution, 3 = nearest-upsampled (half-resolution source); an avg-pooled residual is
// pooled by the preceding anoddpm_gn_apply pass (res_mode 2 is not taken by this kernel).
// Contract: ksize == 3, H % 16 == 0, W % 16 == 0, channels % 32 == 0, Cout % 32 == 0 (anoddpm_conv_tc3_supported).
#include "common.cuh"
#include "anoddpm_b200.h"
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

namespace {

constexpr int kRowBytes = 128;   // one K-block row: 32 fp32 (TF32 MMA) or 64 fp16 elements
constexpr int kKSteps = 4;       // MMAs per K-block: 32 B of K each (8 tf32 / 16 f16 elements)
constexpr int kThreads = 384;       // warps 0-3: TMA producer, MMA issuer, TMEM allocator, idle; warps 4-11: epilogue
constexpr int kSubRows = 18;                        // 16 output rows + 1 halo row above and below
constexpr int kSubBytes = kSubRows * 8 * kRowBytes;   // one (32 ch, 8 px, 18 rows) box = 18432 B
constexpr int kABytes = 2 * kSubBytes;              // two subtiles per CTA
constexpr int kBoxBytes = 18 * 18 * kRowBytes;      // single-box main loop: one (K-block, 18 px, 18 rows) box = both subtiles + halo
constexpr int kBoxStride = 41 * 1024;               // 1024-byte aligned buffer stride (41,472 B used)
constexpr int kAccStages = 2;
constexpr int kMaxStages = 4;
constexpr int kNAMax = 3;                           // GN variant: up to three A boxes in flight (ANODDPM_TC3_NA)
constexpr int kEpiWarps = 8;
constexpr int kStgBuf = 32 * 64;                    // one epilogue buffer: 32 pixel rows x 16 fp32 channels, SWIZZLE_64B
constexpr int kBeBytes = kEpiWarps * 128 * 4;       // per-warp (bias + embedding) row of the current tile
constexpr int kStg16Buf = 32 * 32;                  // fp16 copy of a step: 32 pixel rows x 16 halves, no swizzle
constexpr int kStg16Bytes = kEpiWarps * 2 * kStg16Buf;
constexpr int kTransformWarps = 8;                  // warps 12..19, two per scheduler: the transform is MUFU-bound (ex2 + rcp per element)

constexpr int kThreadsGN = 384 + 32 * kTransformWarps;   // + warps 12..: operand transform (GroupNorm + SiLU on the landed A box)

struct Tc3Args {
    void* out;
    const float* bias;
    const float* emb;
    const float* res;
    int M, H, W, Cout;
    int emb_stride, res_mode;
    int Ca, chunks_a;          // main 3x3 source: channels, 32-channel chunks
    int chunks_r1, chunks_r2;  // 1x1 skip sources
    int BN;                    // pair-tile N (UMMA N); each CTA stages BN/2 weight rows per tap
    int tmem_cols;
    int stages;
    int ke;                    // elements per 128-byte K-block row: 32 (fp32 operands) or 64 (fp16 operands)
    int n_tiles_n, n_tiles;    // tiles along N, total pair tiles
    int regions_x, regions_per_img, n_regions;   // 16x16 regions
    double* gn_stats;          // optional [REPLICAS][B][Cout][2]: per-channel (sum, sum of squares) of the output
    long long stat_replica_stride;   // B*Cout*2
    int tile_order;
    int single_box;            // main loop variant: ONE (K-block, 18 px, 18 rows) A box per chunk, nine taps by descriptor offsets
    // operand-side GroupNorm (GN kernel variant): the A boxes hold RAW fp16 activations; warps 12-15 rewrite each landed box
    // to SiLU(v * scale + shift) before the MMA warp may read it
    const float2* gn_table;    // [B][c_tab] (scale, shift)
    int c_tab;                 // Ca1 + Ca2
    int chunks_a1;             // chunks of the first main source; chunks_a - chunks_a1 come from the second (torch.cat)
    int has_out16;             // the epilogue also stores an fp16 copy of the fp32 output
    int na;                    // A boxes in flight: 2, or 3 in the GN variant when shared memory allows
    int dbg;                   // ANODDPM_TC3_DBG (attribution runs only): 1 = transform warps only forward the barrier,
                               // 2 = read + compute but no write-back
    int stg_nbuf;              // fp32 epilogue buffers per warp: 2, or 3 with a residual (the residual box of step s+2 is
                               // fetched once the store of step s-1 has read its buffer: no store latency on the step chain)
    int res_pf;                // L2 prefetch of the next tile's residual boxes
    int a_pf;                  // GN variant: L2 prefetch distance of the raw A boxes, in chunks (0 = off)
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
// the same with release semantics at cluster scope: the arriving thread's prior shared-memory writes (made visible to the
// async proxy by fence.proxy.async) are ordered before the waiter in the other CTA of the pair
__device__ __forceinline__ void mbar_arrive_remote_release(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t"
        ".reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(cta)
        : "memory");
}
// arrive on a barrier of this CTA (release at CTA scope)
__device__ __forceinline__ void mbar_arrive_local(uint64_t* bar) {
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
// explicit shared-window accesses: pointers derived from the re-aligned dynamic smem base are generic to the compiler
// (LD.E / ST.E through the L1TEX path); the epilogue buffers are hot enough for that to matter
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float lds32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128u(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts128u(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// two fp32 -> one half2 word (a in the low half), round to nearest, saturating at +-65504: one F2FP.SATFINITE
__device__ __forceinline__ uint32_t pack_half2_sat(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// plain (per-CTA) TMA load / store used by the epilogue
__device__ __forceinline__ void tma_load_4d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// L2 prefetch of a box (no shared memory, no barrier): the later load then pays L2 latency instead of HBM latency
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2),
                 "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* smem, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
                 "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
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
// the same with an explicit stride between 8-row groups.  The MMA swizzles on absolute shared-memory address bits
// (profiles/r01_umma_desc_probe.md): any 128-byte start row and any group stride address the rows TMA wrote there.
__device__ __forceinline__ uint64_t make_sdesc_sbo(uint32_t saddr, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

template <bool F16, bool OUT16, bool GN, bool T12 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GN ? kThreadsGN : kThreads, 1)
conv_tc3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a2,
                const __grid_constant__ CUtensorMap map_r1,
                const __grid_constant__ CUtensorMap map_r2, const __grid_constant__ CUtensorMap map_w,
                const __grid_constant__ CUtensorMap map_res, const __grid_constant__ CUtensorMap map_out,
                const __grid_constant__ CUtensorMap map_out16, const Tc3Args P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int b_bytes = (P.BN / 2) * kRowBytes;      // one weight tile (this CTA's half of the N rows)
    const int S = P.stages;
    // multi-box: [S][2 subtiles][kSubBytes] then [S][3 taps][b_bytes]; single-box: [2][kBoxStride] then [S][3 taps][b_bytes]
    uint8_t* sA = smem;
    uint8_t* sB = smem + (P.single_box ? P.na * kBoxStride : S * kABytes);
    uint8_t* stg_all = sB + S * 3 * b_bytes;       // [8 warps][2][kStgBuf], 1024-byte aligned (ring sizes are multiples)
    uint8_t* stg16_all = stg_all + kEpiWarps * P.stg_nbuf * kStgBuf;   // [8 warps][2][kStg16Buf] when the fp16 copy is requested
    float* be_all = reinterpret_cast<float*>(stg16_all + (P.has_out16 ? kStg16Bytes : 0));   // [8 warps][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(be_all) + kBeBytes);
    uint64_t* full = bars;                         // [S]   leader only
    uint64_t* empty = bars + kMaxStages;           // [S]   both CTAs (multicast commit)
    uint64_t* tmem_full = bars + 2 * kMaxStages;   // [2]   both CTAs (multicast commit)
    uint64_t* tmem_empty = tmem_full + kAccStages; // [2]   leader only, 2 x 256 arrivals
    uint64_t* res_bar = tmem_empty + kAccStages;   // [8 warps][3] residual box landed
    uint64_t* a_full = res_bar + 3 * kEpiWarps;    // [2]   single-box main loop: A box landed (leader only)
    uint64_t* a_empty = a_full + kNAMax;           // [3]   both CTAs (multicast commit)
    uint64_t* a_land = a_empty + kNAMax;           // [3]   GN: this CTA's raw A box landed (per CTA)
    uint64_t* t_done = a_land + kNAMax;            // [3]   GN: this CTA's transform warps are done with the box (per CTA)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_done + kNAMax);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // warp roles: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 3 relay (GN), then kE epilogue warps, then kT transform warps.
    // The conv1 form (GN + fp16 output) has a light epilogue and a transform-bound main loop: 4 epilogue warps (each drains
    // both subtiles of its TMEM lane quarter) leave room for 12 transform warps in the same 640 threads.
    constexpr int kE = T12 ? 4 : kEpiWarps;
    constexpr int kT = T12 ? 12 : kTransformWarps;
    static_assert(!T12 || (GN && OUT16), "12 transform warps: conv1 form only");
    const uint32_t rank = cluster_rank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    // tile sequence of this cluster: strided (tile_order 0: the clusters sweep the tensor together) or one contiguous run
    int tile_first, tile_step, tile_end;
    if (P.tile_order == 0) { tile_first = cluster_id; tile_step = n_clusters; tile_end = P.n_tiles; }
    else {
        const int q = P.n_tiles / n_clusters, rem = P.n_tiles - q * n_clusters;
        tile_first = cluster_id * q + min(cluster_id, rem);
        tile_step = 1;
        tile_end = tile_first + q + (cluster_id < rem ? 1 : 0);
    }
    const int nst_a = P.chunks_a * 3;                       // (chunk, dx) stages of the 3x3 source
    const int nst = nst_a + P.chunks_r1 + P.chunks_r2;      // + centre-tap stages of the 1x1 skip sources

    cluster_sync_all();
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
        if (P.res) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_res) : "memory");
        if (P.has_out16) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out16) : "memory");
        if (GN && P.chunks_a1 < P.chunks_a) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a2) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(&full[i], 2); mbar_init(&empty[i], 1); }
        for (int i = 0; i < kAccStages; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 2 * 32 * kE); }
        for (int i = 0; i < 3 * kEpiWarps; ++i) mbar_init(&res_bar[i], 1);
        // GN: a_full collects one arrival per CTA (the relay warp, once its CTA's transform warps are done with the box)
        for (int i = 0; i < kNAMax; ++i) {
            mbar_init(&a_full[i], 2); mbar_init(&a_empty[i], 1); mbar_init(&a_land[i], 1); mbar_init(&t_done[i], kT);
        }
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
        // ---------------- TMA producer (both CTAs) ----------------
        const bool elected = elect_one();
        if (GN) {
            // raw A boxes land on this CTA's OWN barrier (the transform warps of this CTA rewrite them, then signal the
            // leader's a_full); weight stages as in the single-box loop
            if (elected) {
                int bc = 0, ac = 0;
                const int nchunks = T12 ? P.chunks_a : P.chunks_a + P.chunks_r1 + P.chunks_r2;   // T12: no skip sources (host)
                // optional L2 prefetch cursor (ANODDPM_TC3_APF=n): the raw box of the chunk n positions ahead of the one being
                // loaded.  Off by default: neutral for conv1 forms, and in the conv2 forms (residual / two-store epilogue
                // competing for L2) the prefetched lines were re-fetched -- DRAM reads 2.15 GB instead of 1.68 GB, 2.5 % slower
                const int kPF = P.a_pf;
                const int n_my_tiles = tile_end > tile_first ? (tile_end - tile_first + tile_step - 1) / tile_step : 0;
                const long long q_end = (long long)n_my_tiles * nchunks;
                auto prefetch_chunk = [&](long long q) {
                    if (q >= q_end) return;
                    const int ti = (int)(q / nchunks), c = (int)(q - (long long)ti * nchunks);
                    if (c >= P.chunks_a) return;                 // skip sources: small centre boxes, not prefetched
                    const int tile = tile_first + ti * tile_step;
                    const int R = (tile / P.n_tiles_n) * 2 + (int)rank;
                    if (R >= P.n_regions) return;
                    const int b0 = R / P.regions_per_img, rr = R - b0 * P.regions_per_img;
                    const int y0 = (rr / P.regions_x) * 16, x0 = (rr % P.regions_x) * 16;
                    const bool first = c < P.chunks_a1;
                    tma_prefetch_4d(first ? &map_a : &map_a2, (first ? c : c - P.chunks_a1) * P.ke, x0 - 1, y0 - 1, b0);
                };
                for (int q = 0; q < kPF; ++q) prefetch_chunk(q);
                for (int tile = tile_first; tile < tile_end; tile += tile_step) {
                    const int pair = tile / P.n_tiles_n, nt = tile - pair * P.n_tiles_n;
                    const int R = pair * 2 + (int)rank;
                    const int b0 = R / P.regions_per_img, rr = R - b0 * P.regions_per_img;
                    const int y0 = (rr / P.regions_x) * 16, x0 = (rr % P.regions_x) * 16;
                    const int n0 = nt * P.BN + (int)rank * (P.BN / 2);
                    for (int c = 0; c < nchunks; ++c, ++ac) {
                        const int ab = ac % P.na;
                        uint8_t* a_dst = sA + ab * kBoxStride;
                        if (kPF) prefetch_chunk((long long)ac + kPF);
                        mbar_wait(&a_empty[ab], ((ac / P.na) & 1) ^ 1);
                        if (c < P.chunks_a) {
                            const bool first = c < P.chunks_a1;
                            mbar_expect_tx(&a_land[ab], (uint32_t)kBoxBytes);
                            tma_load_4d(a_dst, first ? &map_a : &map_a2, &a_land[ab], (first ? c : c - P.chunks_a1) * P.ke,
                                        x0 - 1, y0 - 1, b0);
                            for (int dy = 0; dy < 3; ++dy, ++bc) {
                                const int s = bc % S;
                                mbar_wait(&empty[s], ((bc / S) & 1) ^ 1);
                                uint8_t* b_dst = sB + s * 3 * b_bytes;
                                if (leader) mbar_expect_tx(&full[s], 2u * 3u * (uint32_t)b_bytes);
                                else mbar_arrive_remote(&full[s], 0);
#pragma unroll
                                for (int dx = 0; dx < 3; ++dx)
                                    tma2_load_2d(b_dst + dx * b_bytes, &map_w, &full[s], (dy * 3 + dx) * P.Ca + c * P.ke, n0);
                            }
                        } else {
                            const int cs = c - P.chunks_a;
                            const bool first = cs < P.chunks_r1;
                            const CUtensorMap* am = first ? &map_r1 : &map_r2;
                            const int c0 = (first ? cs : cs - P.chunks_r1) * P.ke;
                            // raw skip boxes need no transform: they signal the leader's a_full directly (one expect_tx arrival
                            // from the leader + one plain arrival from the peer = the two arrivals a relayed box delivers)
                            if (leader) mbar_expect_tx(&a_full[ab], 2u * 2u * (16 * 8 * kRowBytes));
                            else mbar_arrive_remote(&a_full[ab], 0);
                            tma2_load_4d(a_dst, am, &a_full[ab], c0, x0, y0, b0);
                            tma2_load_4d(a_dst + 16 * 8 * kRowBytes, am, &a_full[ab], c0, x0 + 8, y0, b0);
                            const int s = bc % S;
                            mbar_wait(&empty[s], ((bc / S) & 1) ^ 1);
                            if (leader) mbar_expect_tx(&full[s], 2u * (uint32_t)b_bytes);
                            else mbar_arrive_remote(&full[s], 0);
                            tma2_load_2d(sB + s * 3 * b_bytes, &map_w, &full[s], 9 * P.Ca + cs * P.ke, n0);
                            ++bc;
                        }
                    }
                }
            }
        } else if (elected && P.single_box) {
            // one (K-block, 18 px, 18 rows) box per chunk (2 buffers) + a ring of (chunk, dy) weight stages
            int bc = 0, ac = 0;
            for (int tile = tile_first; tile < tile_end; tile += tile_step) {
                const int pair = tile / P.n_tiles_n, nt = tile - pair * P.n_tiles_n;
                const int R = pair * 2 + (int)rank;
                const int b0 = R / P.regions_per_img, rr = R - b0 * P.regions_per_img;
                const int y0 = (rr / P.regions_x) * 16, x0 = (rr % P.regions_x) * 16;
                const int n0 = nt * P.BN + (int)rank * (P.BN / 2);
                const int nchunks = T12 ? P.chunks_a : P.chunks_a + P.chunks_r1 + P.chunks_r2;   // T12: no skip sources (host)
                for (int c = 0; c < nchunks; ++c, ++ac) {
                    const int ab = ac % P.na;
                    uint8_t* a_dst = sA + ab * kBoxStride;
                    mbar_wait(&a_empty[ab], ((ac / P.na) & 1) ^ 1);
                    if (c < P.chunks_a) {
                        if (leader) mbar_expect_tx(&a_full[ab], 2u * (uint32_t)kBoxBytes);
                        else mbar_arrive_remote(&a_full[ab], 0);
                        tma2_load_4d(a_dst, &map_a, &a_full[ab], c * P.ke, x0 - 1, y0 - 1, b0);
                        for (int dy = 0; dy < 3; ++dy, ++bc) {
                            const int s = bc % S;
                            mbar_wait(&empty[s], ((bc / S) & 1) ^ 1);
                            uint8_t* b_dst = sB + s * 3 * b_bytes;
                            if (leader) mbar_expect_tx(&full[s], 2u * 3u * (uint32_t)b_bytes);
                            else mbar_arrive_remote(&full[s], 0);
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx)
                                tma2_load_2d(b_dst + dx * b_bytes, &map_w, &full[s], (dy * 3 + dx) * P.Ca + c * P.ke, n0);
                        }
                    } else {
                        const int cs = c - P.chunks_a;
                        const bool first = cs < P.chunks_r1;
                        const CUtensorMap* am = first ? &map_r1 : &map_r2;
                        const int c0 = (first ? cs : cs - P.chunks_r1) * P.ke;
                        if (leader) mbar_expect_tx(&a_full[ab], 2u * 2u * (16 * 8 * kRowBytes));
                        else mbar_arrive_remote(&a_full[ab], 0);
                        tma2_load_4d(a_dst, am, &a_full[ab], c0, x0, y0, b0);
                        tma2_load_4d(a_dst + 16 * 8 * kRowBytes, am, &a_full[ab], c0, x0 + 8, y0, b0);
                        const int s = bc % S;
                        mbar_wait(&empty[s], ((bc / S) & 1) ^ 1);
                        if (leader) mbar_expect_tx(&full[s], 2u * (uint32_t)b_bytes);
                        else mbar_arrive_remote(&full[s], 0);
                        tma2_load_2d(sB + s * 3 * b_bytes, &map_w, &full[s], 9 * P.Ca + cs * P.ke, n0);
                        ++bc;
                    }
                }
            }
        } else if (elected) {
            int kbc = 0;
            for (int tile = tile_first; tile < tile_end; tile += tile_step) {
                const int pair = tile / P.n_tiles_n, nt = tile - pair * P.n_tiles_n;
                const int R = pair * 2 + (int)rank;                 // this CTA's 16x16 region (may be past the end)
                const int b0 = R / P.regions_per_img, rr = R - b0 * P.regions_per_img;
                const int y0 = (rr / P.regions_x) * 16, x0 = (rr % P.regions_x) * 16;
                const int n0 = nt * P.BN + (int)rank * (P.BN / 2);
                for (int st = 0; st < nst; ++st, ++kbc) {
                    const int s = kbc % S;
                    mbar_wait(&empty[s], ((kbc / S) & 1) ^ 1);
                    uint8_t* a_dst = sA + s * kABytes;
                    uint8_t* b_dst = sB + s * 3 * b_bytes;
                    if (st < nst_a) {
                        const int chunk = st / 3, dx = st - chunk * 3;
                        const uint32_t bytes = 2u * kSubBytes + 3u * (uint32_t)b_bytes;
                        if (leader) mbar_expect_tx(&full[s], 2u * bytes);
                        else mbar_arrive_remote(&full[s], 0);
                        tma2_load_4d(a_dst, &map_a, &full[s], chunk * P.ke, x0 + dx - 1, y0 - 1, b0);
                        tma2_load_4d(a_dst + kSubBytes, &map_a, &full[s], chunk * P.ke, x0 + 8 + dx - 1, y0 - 1, b0);
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy)
                            tma2_load_2d(b_dst + dy * b_bytes, &map_w, &full[s], (dy * 3 + dx) * P.Ca + chunk * P.ke, n0);
                    } else {
                        const int cs = st - nst_a;
                        const bool first = cs < P.chunks_r1;
                        const CUtensorMap* am = first ? &map_r1 : &map_r2;
                        const int c0 = (first ? cs : cs - P.chunks_r1) * P.ke;
                        const uint32_t bytes = 2u * (16 * 8 * kRowBytes) + (uint32_t)b_bytes;
                        if (leader) mbar_expect_tx(&full[s], 2u * bytes);
                        else mbar_arrive_remote(&full[s], 0);
                        tma2_load_4d(a_dst, am, &full[s], c0, x0, y0, b0);
                        tma2_load_4d(a_dst + kSubBytes, am, &full[s], c0, x0 + 8, y0, b0);
                        tma2_load_2d(b_dst, &map_w, &full[s], 9 * P.Ca + cs * P.ke, n0);
                    }
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
            if (P.single_box) {
                int bc = 0, ac = 0;
                const int nchunks = T12 ? P.chunks_a : P.chunks_a + P.chunks_r1 + P.chunks_r2;   // T12: no skip sources (host)
                for (int tile = tile_first; tile < tile_end; tile += tile_step, ++it) {
                    const int acc = it & 1;
                    mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t tmem_d0 = tmem_base + (uint32_t)((acc * 2 + 0) * P.BN);
                    const uint32_t tmem_d1 = tmem_base + (uint32_t)((acc * 2 + 1) * P.BN);
                    for (int c = 0; c < nchunks; ++c, ++ac) {
                        const int ab = ac % P.na;
                        mbar_wait(&a_full[ab], (ac / P.na) & 1);
                        const uint32_t a_base = smem_u32(sA + ab * kBoxStride);
                        const bool main_src = c < P.chunks_a;
                        const int ndy = main_src ? 3 : 1;
                        for (int dy = 0; dy < ndy; ++dy, ++bc) {
                            const int s = bc % S;
                            mbar_wait(&full[s], (bc / S) & 1);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            if (elect_one()) {
                                const uint32_t b_addr = smem_u32(sB + s * 3 * b_bytes);
                                if (main_src) {
                                    // tap (dy, dx) of subtile sub = rows 18 dy + 8 sub + dx .. of the box, 8-row groups 18 rows apart
                                    for (int dx = 0; dx < 3; ++dx) {
                                        const uint64_t bd = make_sdesc(b_addr + dx * b_bytes);
                                        const uint64_t ad0 = make_sdesc_sbo(a_base + (uint32_t)(18 * dy + dx) * kRowBytes, 18 * kRowBytes);
                                        const uint64_t ad1 = make_sdesc_sbo(a_base + (uint32_t)(18 * dy + 8 + dx) * kRowBytes, 18 * kRowBytes);
#pragma unroll
                                        for (int k = 0; k < kKSteps; ++k) {
                                            const uint32_t accum = (c | dy | dx | k) ? 1u : 0u;
                                            umma2<F16>(tmem_d0, ad0 + 2 * k, bd + 2 * k, idesc, accum);
                                            umma2<F16>(tmem_d1, ad1 + 2 * k, bd + 2 * k, idesc, accum);
                                        }
                                    }
                                } else {
                                    const uint64_t bd = make_sdesc(b_addr);
                                    const uint64_t ad0 = make_sdesc(a_base), ad1 = make_sdesc(a_base + 16 * 8 * kRowBytes);
#pragma unroll
                                    for (int k = 0; k < kKSteps; ++k) {
                                        umma2<F16>(tmem_d0, ad0 + 2 * k, bd + 2 * k, idesc, 1u);
                                        umma2<F16>(tmem_d1, ad1 + 2 * k, bd + 2 * k, idesc, 1u);
                                    }
                                }
                                umma2_commit(&empty[s]);
                                if (dy == ndy - 1) {
                                    umma2_commit(&a_empty[ab]);
                                    if (c == nchunks - 1) umma2_commit(&tmem_full[acc]);
                                }
                            }
                            __syncwarp();
                        }
                    }
                }
            } else
            for (int tile = tile_first; tile < tile_end; tile += tile_step, ++it) {
                const int acc = it & 1;
                mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d0 = tmem_base + (uint32_t)((acc * 2 + 0) * P.BN);
                const uint32_t tmem_d1 = tmem_base + (uint32_t)((acc * 2 + 1) * P.BN);
                for (int st = 0; st < nst; ++st, ++kbc) {
                    const int s = kbc % S;
                    mbar_wait(&full[s], (kbc / S) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
                        const uint64_t a0 = make_sdesc(smem_u32(sA + s * kABytes));
                        const uint64_t a1 = make_sdesc(smem_u32(sA + s * kABytes + kSubBytes));
                        const uint32_t b_addr = smem_u32(sB + s * 3 * b_bytes);
                        const int ntap = st < nst_a ? 3 : 1;
                        for (int dy = 0; dy < ntap; ++dy) {
                            // vertical tap = the same box advanced by dy rows of 8 pixels = dy * 1024 B (atom aligned)
                            const uint64_t bd = make_sdesc(b_addr + dy * b_bytes);
                            const uint64_t ad0 = a0 + (uint64_t)(dy * 64), ad1 = a1 + (uint64_t)(dy * 64);
#pragma unroll
                            for (int k = 0; k < kKSteps; ++k) {
                                const uint32_t accum = (st | dy | k) ? 1u : 0u;
                                umma2<F16>(tmem_d0, ad0 + 2 * k, bd + 2 * k, idesc, accum);
                                umma2<F16>(tmem_d1, ad1 + 2 * k, bd + 2 * k, idesc, accum);
                            }
                        }
                        umma2_commit(&empty[s]);
                        if (st == nst - 1) umma2_commit(&tmem_full[acc]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (GN && warp >= 4 + kE) {
        // ---------------- operand transform (both CTAs): GroupNorm + SiLU on the landed raw box, in place ----------------
        // thread = (16-byte chunk j = channels 8j..8j+7 of the K-block, row lane rl); box row r = pixel (r / 18, r % 18) of
        // the 18x18 halo region sits at r * 128 B, its chunk j at j ^ (r & 7) (SWIZZLE_128B on a 1024-byte aligned buffer).
        // Pixels outside the image keep TMA's zero fill: the convolution pads AFTER normalisation + activation.
        const int tid = (warp - (4 + kE)) * 32 + lane;
        const int j = tid & 7, rl = tid >> 3;
        constexpr int kRowLanes = kT * 4;                               // box rows per sweep of the transform warps (32 or 48)
        constexpr int kSweeps = (18 * 18 + kRowLanes - 1) / kRowLanes;  // 11
        constexpr int kU = 4;                                           // rows in flight per thread (the loop is latency-bound)
        const int nchunks = T12 ? P.chunks_a : P.chunks_a + P.chunks_r1 + P.chunks_r2;   // T12: no skip sources (host)
        int ac = 0;
        uint32_t land_ph = 0;      // phase bit per A slot: a_land / t_done complete once per MAIN chunk that uses the slot
        for (int tile = tile_first; tile < tile_end; tile += tile_step) {
            const int pair = tile / P.n_tiles_n;
            const int R = pair * 2 + (int)rank;
            const int b0 = R / P.regions_per_img, rr = R - b0 * P.regions_per_img;
            const int y0 = (rr / P.regions_x) * 16, x0 = (rr % P.regions_x) * 16;
            const bool live = R < P.n_regions;
            for (int c = 0; c < nchunks; ++c, ++ac) {
                if (c >= P.chunks_a) continue;      // skip-conv boxes go straight to the MMA issuer
                const int ab = ac % P.na;
                const bool work = live && P.dbg != 1;
                float4 t4[4];
                if (work) {   // table row of this thread's 8 channels: issued before the wait, it does not depend on the box
                    const float4* tp = reinterpret_cast<const float4*>(P.gn_table + (size_t)b0 * P.c_tab + c * 64 + 8 * j);
#pragma unroll
                    for (int q = 0; q < 4; ++q) t4[q] = __ldg(tp + q);
                }
                mbar_wait(&a_land[ab], (land_ph >> ab) & 1);
                land_ph ^= 1u << ab;
                if (work) {
                    const uint32_t a_base = smem_u32(sA + ab * kBoxStride);
                    // y = v * sc + sh;  SiLU(y) = y / (1 + 2^(v * nsc + nsh)) with (nsc, nsh) = -log2(e) * (sc, sh)
                    const float sc[8] = {t4[0].x, t4[0].z, t4[1].x, t4[1].z, t4[2].x, t4[2].z, t4[3].x, t4[3].z};
                    const float sh[8] = {t4[0].y, t4[0].w, t4[1].y, t4[1].w, t4[2].y, t4[2].w, t4[3].y, t4[3].w};
                    float nsc[8], nsh[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) { nsc[e] = sc[e] * -1.4426950408889634f; nsh[e] = sh[e] * -1.4426950408889634f; }
#pragma unroll 1
                    for (int i0 = 0; i0 < kSweeps; i0 += kU) {
                        uint4 v[kU];
                        uint32_t addr[kU];
                        bool ok[kU];
#pragma unroll
                        for (int u = 0; u < kU; ++u) {   // kU independent rows: all loads first, no divergent control flow
                            const int r = rl + kRowLanes * (i0 + u);
                            const int rc = min(r, 18 * 18 - 1);
                            const int py = (rc * 3641) >> 16, px = rc - py * 18;      // rc / 18, rc % 18 for rc < 324
                            ok[u] = r < 18 * 18 && (i0 + u) < kSweeps && (unsigned)(y0 - 1 + py) < (unsigned)P.H &&
                                    (unsigned)(x0 - 1 + px) < (unsigned)P.W;
                            addr[u] = a_base + (uint32_t)rc * kRowBytes + (uint32_t)((j ^ (rc & 7)) << 4);
                            v[u] = lds128u(addr[u]);
                        }
#pragma unroll
                        for (int u = 0; u < kU; ++u) {
                            uint32_t* vw = reinterpret_cast<uint32_t*>(&v[u]);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 fv = __half22float2(*reinterpret_cast<const __half2*>(&vw[e]));
                                const float ya = fmaf(fv.x, sc[2 * e], sh[2 * e]), yb = fmaf(fv.y, sc[2 * e + 1], sh[2 * e + 1]);
                                const float da = 1.0f + ex2_approx(fmaf(fv.x, nsc[2 * e], nsh[2 * e]));
                                const float db = 1.0f + ex2_approx(fmaf(fv.y, nsc[2 * e + 1], nsh[2 * e + 1]));
                                vw[e] = pack_half2_sat(ya * rcp_approx(da), yb * rcp_approx(db));
                            }
                        }
#pragma unroll
                        for (int u = 0; u < kU; ++u)
                            if (ok[u] && P.dbg != 2) sts128u(addr[u], v[u]);      // rows outside the image keep TMA's zero fill
                    }
                }
                // done with the box: a cheap CTA-local arrive; the relay warp (warp 3) turns the eight of them into ONE cluster-scope
                // release on the leader's a_full.  (Each transform warp used to do that release itself and sat ~10 % of its
                // samples in it -- profiles/r02b_ncu_conv_tc3_conv2.md.)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive_local(&t_done[ab]);
            }
        }
    } else if (GN && warp == 3) {
        // ---------------- relay (both CTAs): transform warps done -> release the box to the MMA issuer ----------------
        if (lane == 0) {
            const int nchunks = T12 ? P.chunks_a : P.chunks_a + P.chunks_r1 + P.chunks_r2;   // T12: no skip sources (host)
            int ac = 0;
            uint32_t done_ph = 0;
            for (int tile = tile_first; tile < tile_end; tile += tile_step) {
                for (int c = 0; c < nchunks; ++c, ++ac) {
                    if (c >= P.chunks_a) continue;
                    const int ab = ac % P.na;
                    mbar_wait(&t_done[ab], (done_ph >> ab) & 1);
                    done_ph ^= 1u << ab;
                    mbar_arrive_remote_release(&a_full[ab], 0);
                }
            }
        }
    } else if (warp >= 4 && warp < 4 + kE) {
        // ---------------- epilogue (both CTAs) ----------------
        const int ew = warp - 4;            // 0..7
        const int quarter = warp & 3;       // TMEM lanes 32*quarter.. : image rows 4*quarter..+3 of the subtile
        const int sub = ew >> 2;            // subtile (left / right 8 pixel columns of the 16x16 region)
        const int NB = P.stg_nbuf;
        uint8_t* stg = stg_all + ew * (NB * kStgBuf);
        float* be = be_all + ew * 128;
        const bool has_res = P.res != nullptr;
        auto tile_region = [&](int tile, int& b, int& ybase, int& xbase, int& n0) -> bool {
            const int pair = tile / P.n_tiles_n, nt = tile - pair * P.n_tiles_n;
            const int R = pair * 2 + (int)rank;
            b = R / P.regions_per_img;
            const int rr = R - b * P.regions_per_img;
            ybase = (rr / P.regions_x) * 16;
            xbase = (rr % P.regions_x) * 16;
            n0 = nt * P.BN;
            return tile < tile_end && R < P.n_regions;
        };
        if constexpr (OUT16) {
            // ---- fp16 output: the convolution sum only (no bias / embedding, no residual), statistics of sum + bias + emb.
            // Steps of (32 pixels x 32 channels): 64-byte fp16 rows in the same SWIZZLE_64B buffers.
            const int cq = (lane & 3) * 8;      // 8 channels per lane in the column pass
            float keep1[8], keep2[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) keep1[e] = keep2[e] = 0.f;
            int cur_b = -1, cur_n0 = 0, kept = 0;
            auto flush_stats = [&]() {
                if (cur_b >= 0 && 32 * (lane >> 2) < P.BN) {
                    double* st = P.gn_stats + (size_t)(cluster_id % ANODDPM_STAT_REPLICAS) * P.stat_replica_stride +
                                 ((size_t)cur_b * P.Cout + cur_n0 + 32 * (lane >> 2) + cq) * 2;
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        atomicAdd(st + 2 * e, (double)keep1[e]);
                        atomicAdd(st + 2 * e + 1, (double)keep2[e]);
                        keep1[e] = 0.f;
                        keep2[e] = 0.f;
                    }
                }
            };
            int it = 0, step = 0;
            for (int tile = tile_first; tile < tile_end; tile += tile_step, ++it) {
                const int acc = it & 1;
                int b, ybase, xbase, n0;
                const bool valid = tile_region(tile, b, ybase, xbase, n0);
                if (P.gn_stats && valid) {
                    if (b != cur_b || n0 != cur_n0 || kept == 8) {
                        flush_stats();
                        cur_b = b;
                        cur_n0 = n0;
                        kept = 0;
                    }
                    ++kept;
                }
                {   // (bias + timestep embedding) row of this tile: enters the statistics only
                    const int n = n0 + lane * 4;
                    float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (valid && lane * 4 < P.BN && n < P.Cout) {
                        if (P.bias) v4 = *reinterpret_cast<const float4*>(P.bias + n);
                        if (P.emb) {
                            const float4 e4 = *reinterpret_cast<const float4*>(P.emb + (size_t)b * P.emb_stride + n);
                            v4.x += e4.x; v4.y += e4.y; v4.z += e4.z; v4.w += e4.w;
                        }
                    }
                    __syncwarp();
                    sts128(smem_u32(be + lane * 4), v4);
                    __syncwarp();
                }
                mbar_wait(&tmem_full[acc], (it >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // with four epilogue warps (GN variant) each warp drains both subtiles of its TMEM lane quarter
#pragma unroll 1
                for (int sb = 0; sb < (kE == 4 ? 2 : 1); ++sb) {
                const int sub_ = kE == 4 ? sb : sub;
#pragma unroll 1
                for (int cb = 0; cb < P.BN; cb += 32) {
                    uint32_t r[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((acc * 2 + sub_) * P.BN + cb);
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
                    if (cb + 32 >= P.BN && (kE != 4 || sb == 1)) {
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        mbar_arrive_remote(&tmem_empty[acc], 0);
                    }
                    if (!valid) continue;
                    const int buf = step & 1;
                    const uint32_t sba = smem_u32(stg + buf * kStgBuf), bea = smem_u32(be + cb);
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    __syncwarp();
                    // pixel pass: thread = pixel row, 32 channels -> 4 chunks of 8 halves
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 w;
                        w.x = pack_half2_sat(__uint_as_float(r[8 * j]), __uint_as_float(r[8 * j + 1]));
                        w.y = pack_half2_sat(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3]));
                        w.z = pack_half2_sat(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5]));
                        w.w = pack_half2_sat(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7]));
                        sts128u(sba + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4), w);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_4d(&map_out, stg + buf * kStgBuf, n0 + cb, xbase + 8 * sub_, ybase + 4 * quarter, b);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    if (P.gn_stats) {
                        // column pass over the STORED halves (+ bias + embedding): lane owns channels cq..cq+7 of rows (lane >> 2) + 8k
                        const float4 b0 = lds128(bea + cq * 4), b1 = lds128(bea + cq * 4 + 16);
                        const float bev[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                        float s1[8], s2[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) s1[e] = s2[e] = 0.f;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int row = 8 * k + (lane >> 2);
                            const uint4 u = lds128u(sba + row * 64 + (((lane & 3) ^ ((row >> 1) & 3)) << 4));
                            const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&uw[e]));
                                const float v0 = f.x + bev[2 * e], v1 = f.y + bev[2 * e + 1];
                                s1[2 * e] += v0; s2[2 * e] += v0 * v0;
                                s1[2 * e + 1] += v1; s2[2 * e + 1] += v1 * v1;
                            }
                        }
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
#pragma unroll
                            for (int o = 4; o < 32; o <<= 1) {
                                s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], o);
                                s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], o);
                            }
                        }
                        if ((lane >> 2) == (cb >> 5)) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) { keep1[e] += s1[e]; keep2[e] += s2[e]; }
                        }
                    }
                    ++step;
                }
                }
            }
            if (P.gn_stats) flush_stats();
        } else {
        // residual box of step (tile, cb) -> buffer buf (one lane issues; the warp's 4 image rows x 8 px x 16 ch).
        // TMA, not register loads: measured, global loads issued by the epilogue warps (pixel-per-lane or coalesced, one
        // or two steps ahead) are 15-25% slower -- they queue behind the operand bursts of the main loop
        uint8_t* stg16 = stg16_all + ew * (2 * kStg16Buf);
        uint64_t* rbar = res_bar + ew * 3;
        const uint32_t res_bytes = P.res_mode == 3 ? 8u * 64u : (uint32_t)kStgBuf;
        // region of the current tile and of the next one, computed ONCE per tile: the per-step residual fetch below used to
        // re-derive them (three integer divisions by run-time values per step, 13 % of the kernel's stall samples)
        struct Reg { int b, y, x, n0; bool ok; };
        auto region_of = [&](int tile) {
            Reg g;
            g.ok = tile_region(tile, g.b, g.y, g.x, g.n0);
            return g;
        };
        Reg cur = region_of(tile_first), nxt = cur;
        auto issue_res = [&](const Reg& g, int cb, int buf) {
            if (!g.ok) return;
            mbar_expect_tx(&rbar[buf], res_bytes);
            const int y = g.y + 4 * quarter, x = g.x + 8 * sub;
            if (P.res_mode == 3) tma_load_4d(stg + buf * kStgBuf, &map_res, &rbar[buf], g.n0 + cb, x >> 1, y >> 1, g.b);
            else tma_load_4d(stg + buf * kStgBuf, &map_res, &rbar[buf], g.n0 + cb, x, y, g.b);
        };
        // L2 prefetch of a tile's residual boxes (lane = channel step): the TMA load of a step then pays L2, not HBM, latency
        auto prefetch_res = [&](int tile, int first_step) {
            int b, ybase, xbase, n0;
            if (lane < first_step || lane * 16 >= P.BN || !tile_region(tile, b, ybase, xbase, n0)) return;
            const int y = ybase + 4 * quarter, x = xbase + 8 * sub;
            if (P.res_mode == 3) tma_prefetch_4d(&map_res, n0 + lane * 16, x >> 1, y >> 1, b);
            else tma_prefetch_4d(&map_res, n0 + lane * 16, x, y, b);
        };
        if (has_res && lane == 0) {
            issue_res(cur, 0, 0);
            issue_res(cur, 16, 1);
        }
        if (has_res && P.res_pf) prefetch_res(tile_first, 2);
        int buf = 0;
        uint32_t phases = 0;
        // per-channel statistics: lane (c = lane & 15, hf = lane >> 4) owns channel c of every 16-channel step; it keeps the
        // running sums of steps 4 hf .. 4 hf + 3 (channels 16 (4 hf + k) + c) of the current (sample, channel tile) in registers
        // and flushes them with fp64 atomics when either changes
        const int sc_c = lane & 15, sc_h = lane >> 4;
        float keep1[4] = {0.f, 0.f, 0.f, 0.f}, keep2[4] = {0.f, 0.f, 0.f, 0.f};
        int cur_b = -1, cur_n0 = 0, kept = 0;
        auto flush_stats = [&]() {
            if (cur_b >= 0) {
                // replica = cluster id mod 8 spreads the CTAs that work on the same image over 8x more addresses
                double* st = P.gn_stats + (size_t)(cluster_id % ANODDPM_STAT_REPLICAS) * P.stat_replica_stride +
                             ((size_t)cur_b * P.Cout + cur_n0) * 2;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int ch = 16 * (4 * sc_h + k) + sc_c;
                    if (ch < P.BN) {
                        atomicAdd(st + 2 * ch, (double)keep1[k]);
                        atomicAdd(st + 2 * ch + 1, (double)keep2[k]);
                    }
                    keep1[k] = 0.f;
                    keep2[k] = 0.f;
                }
            }
        };
        int it = 0, step = 0;
        for (int tile = tile_first; tile < tile_end; tile += tile_step, ++it) {
            const int acc = it & 1;
            if (tile != tile_first) cur = nxt;
            nxt = region_of(tile + tile_step);
            const int b = cur.b, ybase = cur.y, xbase = cur.x, n0 = cur.n0;
            const bool valid = cur.ok;
            if (has_res && P.res_pf) prefetch_res(tile + tile_step, 0);
            if (P.gn_stats && valid) {
                if (b != cur_b || n0 != cur_n0 || kept == 8) {   // bounded fp32 chains: at most 8 tiles per flush
                    flush_stats();
                    cur_b = b;
                    cur_n0 = n0;
                    kept = 0;
                }
                ++kept;
            }
            // (bias + timestep embedding) row of this tile; lanes own 4 channels each
            {
                const int n = n0 + lane * 4;
                float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid && lane * 4 < P.BN && n < P.Cout) {
                    if (P.bias) v4 = *reinterpret_cast<const float4*>(P.bias + n);
                    if (P.emb) {
                        const float4 e4 = *reinterpret_cast<const float4*>(P.emb + (size_t)b * P.emb_stride + n);
                        v4.x += e4.x; v4.y += e4.y; v4.z += e4.z; v4.w += e4.w;
                    }
                }
                __syncwarp();
                sts128(smem_u32(be + lane * 4), v4);
                __syncwarp();
            }
            mbar_wait(&tmem_full[acc], (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int cb = 0; cb < P.BN; cb += 16) {
                uint32_t r[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((acc * 2 + sub) * P.BN + cb);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (cb + 16 >= P.BN) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive_remote(&tmem_empty[acc], 0);   // this warp's part of the accumulators is read
                }
                if (!valid) continue;
                const uint32_t sba = smem_u32(stg + buf * kStgBuf), bea = smem_u32(be + cb);
                // 1. pixel pass: thread = pixel row `lane` of the buffer (64 B); 16-byte chunk j sits at j ^ ((row >> 1) & 3)
                float4 rs[4];
                if (has_res) {
                    mbar_wait(&rbar[buf], (phases >> buf) & 1);
                    phases ^= 1u << buf;
                    const int srow = P.res_mode == 3 ? ((lane >> 4) * 4 + ((lane & 7) >> 1)) : lane;
#pragma unroll
                    for (int j = 0; j < 4; ++j) rs[j] = lds128(sba + srow * 64 + ((j ^ ((srow >> 1) & 3)) << 4));
                    __syncwarp();   // upsampled mode: rows are shared between lanes, all reads precede the writes
                } else {
                    // the TMA store that last read this buffer (two steps ago) must be done with it
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    __syncwarp();
                }
                uint32_t h16[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 be4 = lds128(bea + 16 * j);
                    float4 v4 = make_float4(__uint_as_float(r[4 * j]) + be4.x, __uint_as_float(r[4 * j + 1]) + be4.y,
                                            __uint_as_float(r[4 * j + 2]) + be4.z, __uint_as_float(r[4 * j + 3]) + be4.w);
                    if (has_res) { v4.x += rs[j].x; v4.y += rs[j].y; v4.z += rs[j].z; v4.w += rs[j].w; }
                    sts128(sba + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4), v4);
                    h16[2 * j] = pack_half2_sat(v4.x, v4.y);
                    h16[2 * j + 1] = pack_half2_sat(v4.z, v4.w);
                }
                if (P.has_out16) {   // fp16 copy of the same 32 px x 16 ch step: dense 32-byte rows (same group as the fp32 store)
                    const uint32_t s16 = smem_u32(stg16 + (step & 1) * kStg16Buf) + lane * 32;
                    sts128u(s16, make_uint4(h16[0], h16[1], h16[2], h16[3]));
                    sts128u(s16 + 16, make_uint4(h16[4], h16[5], h16[6], h16[7]));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                // 2. the buffer is the output box (16 ch, 8 px, 4 rows): one TMA store (+ one for the fp16 copy)
                if (lane == 0) {
                    tma_store_4d(&map_out, stg + buf * kStgBuf, n0 + cb, xbase + 8 * sub, ybase + 4 * quarter, b);
                    if (P.has_out16)
                        tma_store_4d(&map_out16, stg16 + (step & 1) * kStg16Buf, n0 + cb, xbase + 8 * sub, ybase + 4 * quarter, b);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                // 3. column pass: lane owns channel sc_c; the two half-warps read alternate rows (16 conflict-free scalar loads each,
                //    no shuffles but the final exchange between the halves) -> per-channel statistics
                if (P.gn_stats) {   // warp-uniform; all pixels of the CTA's region belong to sample b
                    float s1 = 0.f, s2 = 0.f;
                    const uint32_t cbase = sba + (uint32_t)sc_h * 64 + (uint32_t)(sc_c & 3) * 4;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {     // row 2 i + sc_h; its 16-byte chunk q sits at q ^ (i & 3)
                        const float t = lds32(cbase + (uint32_t)i * 128 + (uint32_t)(((sc_c >> 2) ^ (i & 3)) << 4));
                        s1 += t;
                        s2 = fmaf(t, t, s2);
                    }
                    s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
                    const int sti = cb >> 4;
                    const bool mine = (sti >> 2) == sc_h;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const bool sel = mine && (sti & 3) == k;
                        keep1[k] += sel ? s1 : 0.f;
                        keep2[k] += sel ? s2 : 0.f;
                    }
                }
                if (has_res) {
                    // fetch the residual of step + 2 into the buffer that will hold it: two buffers -> this step's own buffer,
                    // once its store has read it; three -> the buffer of step - 1, whose store had a whole step to finish
                    __syncwarp();
                    if (lane == 0) {
                        const int nb = NB == 3 ? (buf == 0 ? 2 : buf - 1) : buf;
                        if (NB == 3) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        if (cb + 32 < P.BN) issue_res(cur, cb + 32, nb);
                        else issue_res(nxt, cb + 32 - P.BN, nb);
                    }
                    buf = buf + 1 == NB ? 0 : buf + 1;
                } else {
                    buf ^= 1;
                }
                ++step;
            }
        }
        if (P.gn_stats) flush_stats();
        }   // !OUT16
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // buffers outlive the stores
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)P.tmem_cols)
                     : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode3() {
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

// NHWC activation map: dims (C, W, H, B), box (32 ch, 8 px, rows, 1)
int make_act_map3(CUtensorMap* m, const void* p, int f16, int C, int W, int H, int B, int rows, int px = 8) {
    EncodeTiledFn enc = get_encode3();
    if (!enc) { anoddpm_set_error("cuTensorMapEncodeTiled entry point not available"); return ANODDPM_ERR_CUDA; }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t es = f16 ? 2 : 4;
    cuuint64_t strides[3] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es};
    cuuint32_t box[4] = {(cuuint32_t)(f16 ? 64 : 32), (cuuint32_t)px, (cuuint32_t)rows, 1u};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(p), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        anoddpm_set_error("cuTensorMapEncodeTiled(act C=%d W=%d H=%d B=%d rows=%d) failed: %d", C, W, H, B, rows, (int)r);
        return ANODDPM_ERR_CUDA;
    }
    return ANODDPM_OK;
}

// NHWC epilogue map: dims (C, W, H, B), box (16 fp32 / 32 fp16 channels = 64 B, px, rows, 1), 64-byte swizzle
int make_epi_map3(CUtensorMap* m, const void* p, int C, int W, int H, int B, int px, int rows, int f16 = 0) {
    EncodeTiledFn enc = get_encode3();
    if (!enc) { anoddpm_set_error("cuTensorMapEncodeTiled entry point not available"); return ANODDPM_ERR_CUDA; }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t es = f16 ? 2 : 4;
    cuuint64_t strides[3] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es};
    cuuint32_t box[4] = {f16 ? 32u : 16u, (cuuint32_t)px, (cuuint32_t)rows, 1u};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(p), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        anoddpm_set_error("cuTensorMapEncodeTiled(epilogue C=%d W=%d H=%d B=%d) failed: %d", C, W, H, B, (int)r);
        return ANODDPM_ERR_CUDA;
    }
    return ANODDPM_OK;
}

// fp16 copy of the output: dims (C, W, H, B), box (16 halves = 32 B, 8 px, 4 rows, 1), no swizzle
int make_out16_map3(CUtensorMap* m, const void* p, int C, int W, int H, int B) {
    EncodeTiledFn enc = get_encode3();
    if (!enc) { anoddpm_set_error("cuTensorMapEncodeTiled entry point not available"); return ANODDPM_ERR_CUDA; }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {16u, 8u, 4u, 1u};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(p), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        anoddpm_set_error("cuTensorMapEncodeTiled(out16 C=%d W=%d H=%d B=%d) failed: %d", C, W, H, B, (int)r);
        return ANODDPM_ERR_CUDA;
    }
    return ANODDPM_OK;
}

int make_w_map3(CUtensorMap* m, const void* p, int f16, int Ktot, int Cout, int rows) {
    EncodeTiledFn enc = get_encode3();
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

}  // namespace

// 1 if the halo-reuse kernel covers the problem (3x3, 16x16 regions, 32-channel chunks).
extern "C" int anoddpm_conv_tc3_supported(const anoddpm_conv_desc* d) {
    if (!d) return 0;
    if (d->ksize != 3) return 0;
    if (d->H % 16 || d->W % 16) return 0;
    const int ke = d->operand_f16 ? 64 : 32;
    if (d->Ca % ke || d->Cr1 % ke || d->Cr2 % ke || d->Ca2 % ke) return 0;
    if ((d->gn_table || d->a2) && !d->operand_f16) return 0;   // operand-side GroupNorm rewrites fp16 boxes
    if (d->a2 && !d->gn_table) return 0;                       // a second main source is loaded by the GN variant only
    if (d->out16 && d->out_f16) return 0;
    if (d->Cout % 32) return 0;
    if (d->Cout > 128 && d->Cout % 128) return 0;
    if (d->res && d->res_mode == 2) return 0;   // avg-pooled residual: pool it in anoddpm_gn_apply (or use impl 2)
    if (d->out_f16 && d->res) return 0;         // fp16 output is for tensors that feed a GroupNorm only
    return 1;
}

int anoddpm_conv_tc3(const anoddpm_conv_desc* d, void* stream) {
    ADP_REQUIRE(anoddpm_conv_tc3_supported(d), "anoddpm_conv(impl=3): needs ksize 3, H,W %% 16 == 0, channels %% 32 == 0, res_mode "
                "!= 2 (Ca=%d Cr1=%d Cr2=%d Cout=%d H=%d W=%d res_mode=%d)", d->Ca, d->Cr1, d->Cr2, d->Cout, d->H, d->W,
                d->res ? d->res_mode : 0);
    ADP_REQUIRE(!d->res || d->res_mode == 1 || d->res_mode == 3, "anoddpm_conv(impl=3): res_mode must be 1 or 3");
    ADP_REQUIRE(!d->res || d->res_mode != 3 || (d->H % 2 == 0 && d->W % 2 == 0), "anoddpm_conv(impl=3): odd size");
    Tc3Args P;
    P.out = d->out; P.bias = d->bias; P.emb = d->emb; P.res = d->res;
    P.M = d->B * d->H * d->W; P.H = d->H; P.W = d->W; P.Cout = d->Cout;
    P.emb_stride = d->emb_stride; P.res_mode = d->res_mode;
    P.Ca = d->Ca + d->Ca2;                 // concatenated main source: weights are packed over Ca + Ca2 channels per tap
    const int f16 = d->operand_f16 ? 1 : 0;
    const bool gn = d->gn_table != nullptr;
    P.ke = f16 ? 64 : 32;
    P.chunks_a = P.Ca / P.ke;
    P.chunks_a1 = d->Ca / P.ke;
    P.chunks_r1 = d->Cr1 / P.ke;
    P.chunks_r2 = d->Cr2 / P.ke;
    P.gn_table = reinterpret_cast<const float2*>(d->gn_table);
    P.c_tab = P.Ca;
    P.has_out16 = d->out16 ? 1 : 0;
    { const char* e = getenv("ANODDPM_TC3_DBG"); P.dbg = e ? atoi(e) : 0; }
    P.BN = d->Cout < 128 ? d->Cout : 128;
    const int acc_cols = kAccStages * 2 * P.BN;
    P.tmem_cols = acc_cols <= 32 ? 32 : acc_cols <= 64 ? 64 : acc_cols <= 128 ? 128 : acc_cols <= 256 ? 256 : 512;
    const int Ktot = 9 * P.Ca + d->Cr1 + d->Cr2;
    const int b_bytes = (P.BN / 2) * kRowBytes;
    // single-box main loop unless the conv has fused 1x1 skip sources (their small chunks would each hold one of only two
    // A buffers for 8 MMAs: measured 4 % slower there, 4-5 % faster everywhere else); ANODDPM_TC3_SINGLEBOX=0/1 forces.
    // The operand-transform variant always runs the single-box loop.
    { const char* e = getenv("ANODDPM_TC3_SINGLEBOX"); P.single_box = e ? atoi(e) : ((d->Cr1 + d->Cr2) == 0 ? 1 : 0); }
    if (gn) P.single_box = 1;
    // three fp32 epilogue buffers per warp when a residual is fetched (ANODDPM_TC3_RBUF=2 restores two; measured neutral),
    // optional L2 prefetch of the next tile's residual boxes (ANODDPM_TC3_RESPF=1; measured 4-6 % SLOWER: off)
    { const char* e = getenv("ANODDPM_TC3_RBUF"); P.stg_nbuf = (d->res && !d->out_f16 && (!e || atoi(e) == 3)) ? 3 : 2; }
    { const char* e = getenv("ANODDPM_TC3_RESPF"); P.res_pf = e ? atoi(e) : 0; }
    { const char* e = getenv("ANODDPM_TC3_APF"); P.a_pf = e ? atoi(e) : 0; if (P.a_pf < 0 || P.a_pf > 8) P.a_pf = 0; }
    const size_t fixed = (size_t)kEpiWarps * P.stg_nbuf * kStgBuf + (P.has_out16 ? kStg16Bytes : 0) + kBeBytes + 512 + 1024;
    const size_t budget = 227 * 1024 - fixed;
    P.na = 2;
    if (gn) { const char* e = getenv("ANODDPM_TC3_NA"); P.na = e ? atoi(e) : 2; if (P.na < 2 || P.na > kNAMax) P.na = 2; }
    int stages = P.single_box ? (int)((budget - (size_t)P.na * kBoxStride) / (3 * b_bytes)) : (int)(budget / (kABytes + 3 * b_bytes));
    if (stages > kMaxStages) stages = kMaxStages;
    ADP_REQUIRE(stages >= 2, "anoddpm_conv(impl=3): shared memory budget leaves fewer than 2 stages");
    P.stages = stages;
    P.n_tiles_n = adp_ceil_div(d->Cout, P.BN);
    P.regions_x = d->W / 16;
    P.regions_per_img = (d->H / 16) * (d->W / 16);
    P.n_regions = d->B * P.regions_per_img;
    P.n_tiles = adp_ceil_div(P.n_regions, 2) * P.n_tiles_n;
    P.gn_stats = d->gn_stats;
    P.stat_replica_stride = (long long)d->B * d->Cout * 2;
    P.tile_order = 1;   // contiguous run of tiles per cluster (statistics stay in registers across a sample's tiles)

    CUtensorMap ma, ma2, m1, m2, mw;
    int rc;
    if ((rc = make_act_map3(&ma, d->a, f16, d->Ca, d->W, d->H, d->B, kSubRows, P.single_box ? 18 : 8))) return rc;
    ma2 = ma; m1 = ma; m2 = ma;
    if (d->a2 && (rc = make_act_map3(&ma2, d->a2, f16, d->Ca2, d->W, d->H, d->B, kSubRows, 18))) return rc;
    if (d->r1 && (rc = make_act_map3(&m1, d->r1, f16, d->Cr1, d->W, d->H, d->B, 16))) return rc;
    if (d->r2 && (rc = make_act_map3(&m2, d->r2, f16, d->Cr2, d->W, d->H, d->B, 16))) return rc;
    if ((rc = make_w_map3(&mw, d->w, f16, Ktot, d->Cout, P.BN / 2))) return rc;
    // epilogue store box: (16 ch fp32, 8 px, 4 rows) of the output
    CUtensorMap mres, mout, mout16;
    if ((rc = make_epi_map3(&mout, d->out, d->Cout, d->W, d->H, d->B, 8, 4, d->out_f16 ? 1 : 0))) return rc;
    mres = mout; mout16 = mout;
    if (d->out16 && (rc = make_out16_map3(&mout16, d->out16, d->Cout, d->W, d->H, d->B))) return rc;
    if (d->res) {   // residual boxes: (16, 8, 4) at the same resolution, (16, 4, 2) of a half-resolution source
        if (d->res_mode == 3) rc = make_epi_map3(&mres, d->res, d->Cout, d->W / 2, d->H / 2, d->B, 4, 2);
        else rc = make_epi_map3(&mres, d->res, d->Cout, d->W, d->H, d->B, 8, 4);
        if (rc) return rc;
    }

    const size_t ring = P.single_box ? (size_t)P.na * kBoxStride + (size_t)stages * 3 * b_bytes : (size_t)stages * (kABytes + 3 * b_bytes);
    const size_t smem = ring + fixed;
    static bool attr_dev[64] = {};   // the attribute is per device
    int dev_ = 0;
    ADP_CUDA(cudaGetDevice(&dev_));
    if (!attr_dev[dev_ & 63]) {
        ADP_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ADP_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ADP_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ADP_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ADP_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ADP_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ADP_CUDA(cudaFuncSetAttribute(conv_tc3_kernel<true, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_dev[dev_ & 63] = true;
    }
    const int sms = 148;
    const int n_clusters = P.n_tiles < sms / 2 ? P.n_tiles : sms / 2;
    const dim3 grid(2 * n_clusters);
    cudaStream_t st = (cudaStream_t)stream;
    if (gn) {
        // conv1 form: 4 epilogue + 12 transform warps (ANODDPM_TC3_T12=0: 8 + 8)
        static const int t12 = [] { const char* e = getenv("ANODDPM_TC3_T12"); return e ? atoi(e) : 1; }();
        if (d->out_f16 && t12 && d->Cr1 + d->Cr2 == 0) conv_tc3_kernel<true, true, true, true><<<grid, kThreadsGN, smem, st>>>(ma, ma2, m1, m2, mw, mres, mout, mout16, P);
        else if (d->out_f16) conv_tc3_kernel<true, true, true><<<grid, kThreadsGN, smem, st>>>(ma, ma2, m1, m2, mw, mres, mout, mout16, P);
        else conv_tc3_kernel<true, false, true><<<grid, kThreadsGN, smem, st>>>(ma, ma2, m1, m2, mw, mres, mout, mout16, P);
    } else if (d->out_f16) {
        if (f16) conv_tc3_kernel<true, true, false><<<grid, kThreads, smem, st>>>(ma, ma2, m1, m2, mw, mres, mout, mout16, P);
        else conv_tc3_kernel<false, true, false><<<grid, kThreads, smem, st>>>(ma, ma2, m1, m2, mw, mres, mout, mout16, P);
    } else {
        if (f16) conv_tc3_kernel<true, false, false><<<grid, kThreads, smem, st>>>(ma, ma2, m1, m2, mw, mres, mout, mout16, P);
        else conv_tc3_kernel<false, false, false><<<grid, kThreads, smem, st>>>(ma, ma2, m1, m2, mw, mres, mout, mout16, P);
    }
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}
