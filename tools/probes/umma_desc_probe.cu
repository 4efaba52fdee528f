// tools/probes/umma_desc_probe.cu -- which shared-memory descriptor variants does tcgen05.mma accept for a K-major
// SWIZZLE_128B A operand whose rows are NOT the canonical "128 consecutive rows from a 1024-byte aligned base"?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/probes/umma_desc_probe tools/probes/umma_desc_probe.cu
//
// Motivation: a 3x3 conv could load ONE (channels, 18 px, 18 rows) box per chunk and address the
// nine taps as descriptor offsets if (a) the start address may sit at any 128-byte row inside a swizzle atom (dx shift)
// and (b) the stride between 8-row groups (SBO) may differ from 1024 B (row pitch of 18 pixels).  The tile is written
// with the TMA swizzle (16-byte chunk index XOR (row address >> 7) & 7), then one M=128, N=64, K=64 MMA group is
// issued per variant and the result compared with the row mapping the variant is meant to realise.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Variant { int row0; int sbo; int base_offset; };   // first row, bytes between 8-row groups, descriptor base_offset field

__global__ void __launch_bounds__(128) probe_kernel(const __half* __restrict__ A, int a_rows, const __half* __restrict__ Bm,
                                                    float* __restrict__ D, Variant v) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                       // a_rows x 128 B, swizzled like a TMA box
    uint8_t* sB = smem + 64 * 1024;           // 64 x 128 B
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < a_rows * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(sA + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + r * 64 + c * 8);
    }
    for (int i = tid; i < 64 * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(sB + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(Bm + r * 64 + c * 8);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // f16 x f16 -> f32
        const uint32_t a_addr = smem_u32(sA) + v.row0 * 128;
        const uint64_t ad = (uint64_t)((a_addr >> 4) & 0x3FFF) | ((uint64_t)(v.sbo >> 4) << 32) | ((uint64_t)1 << 46) |
                            ((uint64_t)(v.base_offset & 7) << 49) | ((uint64_t)2 << 61);
        const uint64_t bd = (uint64_t)((smem_u32(sB) >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
                            ((uint64_t)2 << 61);
        for (int k = 0; k < 4; ++k) {
            const uint32_t acc = k ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
                "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra DONE;\n\tbra W;\n\tDONE:\n\t}\n" ::"r"(
            smem_u32(&bar))
        : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[32];
    for (int cb = 0; cb < 64; cb += 32) {
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + cb;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) D[tid * 64 + cb + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

int main() {
    const int a_rows = 400;
    std::vector<__half> A(a_rows * 64), B(64 * 64);
    std::vector<float> Af(a_rows * 64), Bf(64 * 64);
    srand(1);
    for (size_t i = 0; i < A.size(); ++i) { Af[i] = (float)((rand() % 17) - 8) / 8.0f; A[i] = __float2half(Af[i]); }
    for (size_t i = 0; i < B.size(); ++i) { Bf[i] = (float)((rand() % 13) - 6) / 4.0f; B[i] = __float2half(Bf[i]); }
    __half *dA, *dB; float* dD;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, 128 * 64 * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    const int smem = 64 * 1024 + 8 * 1024 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    // intended mapping of every variant: output row m = input row  row0 + (m / 8) * (sbo / 128) + (m % 8)
    const Variant vs[] = {{0, 1024, 0}, {1, 1024, 0}, {1, 1024, 1}, {3, 1024, 0}, {3, 1024, 3}, {8, 1024, 0},
                          {0, 1280, 0}, {0, 2304, 0}, {1, 2304, 0}, {1, 2304, 1}, {19, 2304, 0}, {19, 2304, 3}, {0, 1152, 0}, {0, 1152, 1}};
    std::vector<float> D(128 * 64);
    printf("| row0 | SBO | base_offset | max abs err vs intended rows | verdict |\n|---:|---:|---:|---:|---|\n");
    for (const Variant& v : vs) {
        cudaMemset(dD, 0xff, 128 * 64 * 4);
        probe_kernel<<<1, 128, smem>>>(dA, a_rows, dB, dD, v);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("| %d | %d | %d | - | CUDA error: %s |\n", v.row0, v.sbo, v.base_offset, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double worst = 0;
        for (int m = 0; m < 128; ++m) {
            const int row = v.row0 + (m / 8) * (v.sbo / 128) + (m % 8);
            for (int n = 0; n < 64; ++n) {
                double ref = 0;
                for (int k = 0; k < 64; ++k) ref += (double)Af[row * 64 + k] * Bf[n * 64 + k];
                const double d = fabs(ref - D[m * 64 + n]);
                if (!(d <= worst)) worst = d;
            }
        }
        printf("| %d | %d | %d | %.3g | %s |\n", v.row0, v.sbo, v.base_offset, worst, worst < 1e-3 ? "MATCH" : "differs");
    }
    return 0;
}
