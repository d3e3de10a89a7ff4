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
    int batch;                   // batch size: rows with a batch index outside [0, batch) or coordinates outside `shape` are ignored
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
// same, destination given as a 32-bit shared-space address (saves the generic->shared conversion per copy)
__device__ __forceinline__ void cp_async16_s(uint32_t smem_addr, const void* gsrc, bool valid) {
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_addr), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// Stage the tile's slice nbr[k0 .. k0+kcount) x [base, base+128) of a neighbour table into shared memory.  Loads are
// issued in batches of 8 independent LDGs per thread (a plain `nbr_s[i] = nbr[...]` loop serialises one global-memory
// round trip per iteration: ~20 dependent round trips for K = 27, which was most of a conv CTA's prologue).
template <int THREADS, int BATCH = 8>
__device__ __forceinline__ void stage_nbr_tile(const int32_t* __restrict__ nbr, int n_rows, int k0, int kcount, int base,
                                               int* nbr_s) {
    const int total = kcount * 128;
    for (int i0 = threadIdx.x; i0 < total; i0 += THREADS * BATCH) {
        int v[BATCH];
#pragma unroll
        for (int j = 0; j < BATCH; ++j) {
            const int i = i0 + j * THREADS;
            const int k = i >> 7, row = base + (i & 127);
            v[j] = (i < total && row < n_rows) ? __ldg(nbr + (size_t)(k0 + k) * n_rows + row) : -1;
        }
#pragma unroll
        for (int j = 0; j < BATCH; ++j) {
            const int i = i0 + j * THREADS;
            if (i < total) nbr_s[i] = v[j];
        }
    }
}

// ---- programmatic dependent launch (PDL) for the kernels of the main-stream chain ----
// Every step is ~160 small dependent kernels on one stream; a plain launch starts only after the previous grid has
// drained AND been flushed (2-3 us each, 0.4-0.6 ms per step).  Launched with the programmatic-stream-serialisation
// attribute a kernel may start (block scheduling, prologue) as soon as every CTA of its predecessor has issued
// `griddepcontrol.launch_dependents`, and blocks in `griddepcontrol.wait` until the predecessor has completed and its
// writes are visible.  Kernels using this call pdl_wait() before their first global read or write of dependent data;
// both instructions are no-ops for a normally launched grid.  vc_set_pdl(0) turns the attribute off.
extern int g_pdl;
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KP, typename... A>
static inline cudaError_t launch_chain(void (*kern)(KP...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at{};
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &at;
    cfg.numAttrs = g_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KP>(args)...);
}

#define VC_LAUNCH_CHAIN(...)                      \
    do {                                          \
        vc::count_launch();                       \
        VC_CUDA(vc::launch_chain(__VA_ARGS__));   \
    } while (0)

// ---- shared between rulebook.cu and executor.cu ----
// `n_dev` (optional): the row count lives in device memory and `n` / `n_out` are capacities (static mode, graph capturable)
int subm_rulebook_dev(const int32_t* indices, int n, const int* n_dev, int ndim, int batch_size, const int32_t* spatial_shape,
                      const int32_t* ksize, const int32_t* dilation, int32_t* nbr, int32_t* pair_num, void* ws, size_t ws_bytes,
                      cudaStream_t stream);
int conv_rulebook_count_dev(const int32_t* indices, int n, const int* n_dev, int ndim, int batch_size,
                            const int32_t* spatial_shape, const int32_t* ksize, const int32_t* stride, const int32_t* padding,
                            const int32_t* dilation, int32_t* n_out_dev, int cap_out, int* overflow, void* ws, size_t ws_bytes,
                            cudaStream_t stream);
int conv_rulebook_fill_phases(const int32_t* indices, int n, const int* n_dev, int ndim, int batch_size,
                              const int32_t* spatial_shape, const int32_t* ksize, const int32_t* stride, const int32_t* padding,
                              const int32_t* dilation, int n_out, int32_t* out_indices, int32_t* nbr_fwd, int32_t* nbr_bwd,
                              int32_t* pair_num, void* ws, size_t ws_bytes, cudaStream_t stream, int phases);

// ---- bn.cu / misc.cu internals used by the plan executor (device row counts, strided outputs) ----
int bn_apply_relu_dev(const float* x, const double* sums, int n_rows, const int* n_dev, int c, const float* gamma,
                      const float* beta, float* running_mean, float* running_var, long long* num_batches_tracked, float momentum,
                      float eps, int training, float* y, int y_ld, void* y_bf16, int yb_ld, float* stats_out, int relu,
                      int tail_zero, cudaStream_t stream);
int bn_relu_bwd_dev(const float* dy, int dy_ld, const float* x, const float* gamma, const float* stats, float* dx, void* dx_bf16,
                    float* dgamma, float* dbeta, int n, const int* n_dev, int c, int training, double* bsums, int tail_zero,
                    cudaStream_t stream);
int cat2_dev(const float* a, const float* b, float* out, void* out_bf16, int n, const int* n_dev, int ca, int cb, cudaStream_t stream);
int index2uv_dev(const int32_t* indices, int n, const int* n_dev, int batch_size, const float* params, const float* grid,
                 int stride, int u_max, int v_max, int32_t* uv_out, cudaStream_t stream);

// ---- shared between conv_tc.cu and executor.cu ----
struct TcPrepEntry {
    const float* w;   // parameter, spconv layout [C_out, K, C_in]
    void* img;        // bf16 UMMA image, K * cin * cout elements
    int cin, cout, K;
    int mode;         // 0 forward image, 1 dgrad image
    int mirror;       // dgrad: kernel offsets mirrored (submanifold table re-used as its own transpose)
    int layout;       // 0: round-1 core-matrix image (tc_scatter_kernel, legacy gather kernel); 1: swizzled + padded to 16 (conv_tc2.cu)
    int first;        // element prefix (filled by tc_prep_images)
};
static constexpr int TC_PREP_MAX = 64;
struct TcPrepTable {
    TcPrepEntry e[TC_PREP_MAX];
    int n, total;
};
int tc_prep_images(TcPrepTable& t, cudaStream_t stream);
int tc_scatter_with_image(int kc, int nr, const void* dout_bf16, const void* wimg, const int32_t* nbr, float* din, int n_out,
                          int K, int* err, cudaStream_t stream);
int tc_conv_with_image(int kc, int nr, const void* in_bf16, const void* wimg, const int32_t* nbr, long long pitch, float* out,
                       int n_rows, const int* n_dev, int K, double* bn_sums, int* err, cudaStream_t stream, const float* addend,
                       int* tile_counter = nullptr, int flags = 0);   // flags: bit 0 sparse offsets (strided dgrad), bit 1 early tables
bool tc_conv_ch_ok(int c);
// ---- conv_tc2.cu (persistent tensor-core conv) ----
extern int g_tc_variant;
size_t tc_image_bytes(int cin, int cout, int K, int layout);
bool tc2_ch_ok(int c);
int tc2_conv(int kc, int nr, const void* in_bf16, const void* wimg, const int32_t* nbr, long long pitch, float* out, int n_rows,
             const int* n_dev, int K, double* bn_sums, int* err, cudaStream_t stream, const float* addend,
             int* tile_counter = nullptr, int sparse_k = 0);

// ---- wgrad_tc3.cu (half-tile-stage variant of the persistent wgrad, A/B) ----
extern int g_wgrad_variant;
int tc3_wgrad(int cin, int cout, const void* in_bf16, const void* dout_bf16, const int32_t* nbr, long long pitch, float* scratch,
              int n_rows, const int* n_dev, int K, int* err, cudaStream_t stream, int* tile_counter);
// ---- wgrad_tc2.cu (persistent tensor-core wgrad) ----
struct WgradFinEntry {
    const float* scratch;   // [K][cin][cout] accumulated by tc2_wgrad
    float* dw;              // parameter layout [cout][K][cin]
    int cin, cout, K, first;
};
static constexpr int WGRAD_FIN_MAX = 64;
struct WgradFinTable {
    WgradFinEntry e[WGRAD_FIN_MAX];
    int n, total;
};
int wgrad_finalize(WgradFinTable& t, cudaStream_t stream);
int tc2_wgrad(int cin, int cout, const void* in_bf16, const void* dout_bf16, const int32_t* nbr, long long pitch, float* scratch,
              int n_rows, const int* n_dev, int K, int* err, cudaStream_t stream, int* tile_counter);
int wgrad2_passes(int cin, int cout, int K);

}  // namespace vc
