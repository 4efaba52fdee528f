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
    return "sm_100a; conv: tcgen05.mma kind::tf32 + TMA (impl=1), fp32 CUDA-core (impl=0); simplex fp64 -fmad=false";
}
