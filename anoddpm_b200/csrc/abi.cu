// anoddpm_b200/csrc/abi.cu -- error channel and version of the C ABI.
#include "common.cuh"
#include "anoddpm_b200.h"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void anoddpm_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int anoddpm_abi_version(void) { return ANODDPM_ABI_VERSION; }
extern "C" const char* anoddpm_last_error(void) { return g_err; }
extern "C" const char* anoddpm_build_info(void) {
    return "sm_100a; conv: tcgen05.mma kind::f16 / kind::tf32 + TMA (impl 1-3), fp32 CUDA-core (impl 0); tcgen05 weight gradient; "
           "simplex fp64 -fmad=false";
}

// ---------------------------------------------------------------- FP64 FMA throughput probe (roofline denominator)
// The simplex generator is FP64-ALU bound; MEASURED_PEAKS.json has no FP64 figure, so bench.py measures one the same way
// (best of 10 + a sustained loop): every thread runs `iters` rounds of 8 independent dependent-FMA chains.
__global__ void __launch_bounds__(256) fp64_fma_probe_kernel(double* __restrict__ out, int iters, double a, double b) {
    double v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (double)(threadIdx.x + j) * 1e-3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fma(v[j], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];
    if (s == 123456.789) out[0] = s;          // never true: keeps the chains alive
}
extern "C" int anoddpm_fp64_fma_probe(double* out, int blocks, int iters, void* stream) {
    ADP_REQUIRE(out && blocks > 0 && iters > 0, "anoddpm_fp64_fma_probe: bad arguments");
    fp64_fma_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, iters, 0.999999, 1e-7);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}
