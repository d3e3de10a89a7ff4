// Shared helpers for the virconv_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/virconv_b200.h"

namespace vc {

void set_error(const char* fmt, ...);
void count_launch();  // bumps the counter behind vc_launch_count()

#define VC_CHECK_ARG(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            vc::set_error(__VA_ARGS__);         \
            return VC_ERR_INVALID;              \
        }                                       \
    } while (0)

#define VC_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            vc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return VC_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

#define VC_LAUNCH_CHECK()          \
    do {                           \
        vc::count_launch();        \
        VC_CUDA(cudaGetLastError()); \
    } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

struct Geom {  // conv geometry, passed by value to kernels
    int ndim;
    int K;
    int shape[VC_MAX_NDIM];      // input spatial shape
    int oshape[VC_MAX_NDIM];     // output spatial shape
    int ksize[VC_MAX_NDIM];
    int stride[VC_MAX_NDIM];
    int pad[VC_MAX_NDIM];
    int dil[VC_MAX_NDIM];
};

__device__ __forceinline__ uint32_t mix64(uint64_t k) {  // murmur3 finaliser
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return (uint32_t)k;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    int sz = valid ? 16 : 0;  // src-size 0 => 16 bytes of zeros
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

}  // namespace vc
