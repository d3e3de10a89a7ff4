// Round-2 experiment (stand-alone): are the row-major, 32/64/128-byte-swizzled tiles that the persistent conv kernel
// gathers (rows of C channels, 8-row swizzle atoms) valid MN-MAJOR tcgen05 operands — i.e. can the weight-gradient
// contraction  D[(g, ci), co] = sum_rows A_g[row, ci] * B[row, co]  run directly on the forward kernel's operand image,
// with G tiles stacked along M through the descriptor's leading-byte offset?  (cute/atom/mma_traits_sm100.hpp:
// Major-MN canonical layouts  Swizzle<b,4,3> o ((2^b, n), (8, k)) : ((1, LBO), (2^b, SBO))  in 16-byte units.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 profiles/exp_mnmajor.cu -o /tmp/exp_mn && /tmp/exp_mn
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x)                                                                                \
    do {                                                                                     \
        cudaError_t e__ = (x);                                                               \
        if (e__ != cudaSuccess) {                                                            \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); \
            exit(1);                                                                         \
        }                                                                                    \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar), done = 0;
    for (unsigned spin = 0; spin < (1u << 22); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return true;
    }
    return false;
}
__host__ __device__ inline uint32_t swz_off(int r, int c, int rowb) {
    uint32_t off = (uint32_t)r * rowb + (uint32_t)c * 16;
    uint32_t mask = rowb == 128 ? 7u : rowb == 64 ? 3u : rowb == 32 ? 1u : 0u;
    return off ^ (((off >> 7) & mask) << 4);
}
// MN-major, swizzled: LBO = byte distance between swizzle atoms along M/N, SBO = between 8-row groups along K
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo, int rowb) {
    uint64_t lt = rowb == 128 ? 2 : rowb == 64 ? 4 : 6;
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
           (1ULL << 46) | (lt << 61);
}
__host__ __device__ constexpr uint32_t idesc_mn(int m, int n) {   // D f32, A = B = bf16, both MN-major (bits 15, 16)
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
                 "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// A image: G tiles of [128 rows][CI] (row pitch CI*2 = rowb_a, swizzled) A_BYTES apart; B image: [128][CO] (rowb_b, swizzled)
template <int CI, int CO>
__global__ void __launch_bounds__(128) k_wgrad(const unsigned char* a_img, const unsigned char* b_img, float* out, int* status) {
    constexpr int RA = CI * 2, RB = CO * 2, G = 128 / CI, A_BYTES = 128 * RA;
    extern __shared__ __align__(1024) unsigned char sm[];
    unsigned char* A = sm;
    unsigned char* B = sm + G * A_BYTES;
    __shared__ __align__(8) uint64_t done;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(64u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        mbar_init(&done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < G * A_BYTES / 16; i += 128) reinterpret_cast<uint4*>(A)[i] = reinterpret_cast<const uint4*>(a_img)[i];
    for (int i = tid; i < 128 * RB / 16; i += 128) reinterpret_cast<uint4*>(B)[i] = reinterpret_cast<const uint4*>(b_img)[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_s;
    if (tid == 0) {
        constexpr uint32_t IDESC = idesc_mn(128, CO);
#pragma unroll
        for (int j = 0; j < 8; ++j)      // 16 rows (two 8-row groups) per MMA
            umma_f16(tmem, desc_mn(smem_u32(A) + j * 16 * RA, A_BYTES, 8 * RA, RA), desc_mn(smem_u32(B) + j * 16 * RB, 0, 8 * RB, RB),
                     IDESC, j > 0);
        umma_commit(&done);
    }
    bool ok = mbar_wait(&done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok) atomicExch(status, 1);
    const int r = warp * 32 + lane;
#pragma unroll
    for (int c0 = 0; c0 < CO; c0 += 16) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        for (int i = 0; i < 16; ++i) out[r * CO + c0 + i] = v[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u));
}

template <int CI, int CO>
static void run() {
    constexpr int RA = CI * 2, RB = CO * 2, G = 128 / CI, A_BYTES = 128 * RA;
    std::vector<float> a((size_t)G * 128 * CI), b((size_t)128 * CO);
    for (auto& x : a) x = (float)((rand() % 9) - 4) / 4.f;
    for (auto& x : b) x = (float)((rand() % 9) - 4) / 4.f;
    std::vector<unsigned char> ai((size_t)G * A_BYTES), bi((size_t)128 * RB);
    for (int g = 0; g < G; ++g)
        for (int r = 0; r < 128; ++r)
            for (int c = 0; c < CI; ++c) {
                __nv_bfloat16 v = __float2bfloat16_rn(a[((size_t)g * 128 + r) * CI + c]);
                memcpy(&ai[(size_t)g * A_BYTES + swz_off(r, c / 8, RA) + (c % 8) * 2], &v, 2);
            }
    for (int r = 0; r < 128; ++r)
        for (int c = 0; c < CO; ++c) {
            __nv_bfloat16 v = __float2bfloat16_rn(b[(size_t)r * CO + c]);
            memcpy(&bi[swz_off(r, c / 8, RB) + (c % 8) * 2], &v, 2);
        }
    unsigned char *da, *db;
    float* dout;
    int* dst;
    CK(cudaMalloc(&da, ai.size()));
    CK(cudaMalloc(&db, bi.size()));
    CK(cudaMalloc(&dout, 128 * CO * 4));
    CK(cudaMalloc(&dst, 4));
    CK(cudaMemset(dst, 0, 4));
    CK(cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice));
    const int smem = G * A_BYTES + 128 * RB + 1024;
    CK(cudaFuncSetAttribute(k_wgrad<CI, CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_wgrad<CI, CO><<<1, 128, smem>>>(da, db, dout, dst);
    CK(cudaDeviceSynchronize());
    std::vector<float> out(128 * CO);
    int st = 0;
    CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < CO; ++n) {
            const int g = m / CI, ci = m % CI;
            double ref = 0;
            for (int r = 0; r < 128; ++r) ref += (double)a[((size_t)g * 128 + r) * CI + ci] * b[(size_t)r * CO + n];
            maxerr = fmax(maxerr, fabs(ref - out[m * CO + n]));
            maxref = fmax(maxref, fabs(ref));
        }
    printf("[mn-major] CI=%2d (G=%d tiles along M, LBO %5d) CO=%2d: status %d max|err| %.3g (max|ref| %.3g) -> %s\n", CI, G, A_BYTES, CO,
           st, maxerr, maxref, (maxerr <= 1e-3 * maxref && !st) ? "OK" : "FAIL");
    CK(cudaFree(da)); CK(cudaFree(db)); CK(cudaFree(dout)); CK(cudaFree(dst));
}

int main() {
    srand(3);
    run<16, 16>(); run<16, 32>(); run<16, 64>();
    run<32, 16>(); run<32, 32>(); run<32, 64>();
    run<64, 16>(); run<64, 32>(); run<64, 64>();
    printf("done\n");
    return 0;
}
