// anoddpm_b200/csrc/common.cuh -- shared host/device helpers for the C-ABI library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ANODDPM_OK 0
#define ANODDPM_ERR_ARG (-1)
#define ANODDPM_ERR_CUDA (-2)
#define ANODDPM_ERR_UNSUPPORTED (-3)

// thread-local error string, readable through anoddpm_last_error()
void anoddpm_set_error(const char* fmt, ...);

#define ADP_REQUIRE(cond, ...)              \
    do {                                    \
        if (!(cond)) {                      \
            anoddpm_set_error(__VA_ARGS__); \
            return ANODDPM_ERR_ARG;         \
        }                                   \
    } while (0)

#define ADP_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            anoddpm_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return ANODDPM_ERR_CUDA;                                                            \
        }                                                                                       \
    } while (0)

#define ADP_LAUNCH_CHECK() ADP_CUDA(cudaGetLastError())

static inline int adp_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
