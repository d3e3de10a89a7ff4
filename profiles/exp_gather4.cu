// Round-2 experiment (stand-alone; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo profiles/exp_gather4.cu):
//   1. semantics of TMA `cp.async.bulk.tensor.2d ... tile::gather4` on sm_100a: out-of-range row indices (-1, >= N)
//      are zero-filled and still counted on the mbarrier; the four rows land as a 4-row tile with the tensor map's
//      swizzle applied on shared-memory address bits (so two gather4 fill one 8-row UMMA swizzle atom);
//   2. a 128-row gathered tile written that way is a valid K-major tcgen05.mma operand (swizzle 32/64/128 B,
//      box wider than the tensor -> zero-filled channels: the C=8 layers run as K=16);
//   3. throughput of the gather alone, TMA gather4 vs 16-byte cp.async, persistent CTAs with a deep ring, for the
//      row widths of the backbone (32/64/128 B) at the fill fractions of its rulebooks.
// Nothing here is product code; the findings drive csrc/conv_tc.cu and are recorded in DESIGN.md.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e__ = (x);                                                                 \
        if (e__ != cudaSuccess) {                                                              \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__);   \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn g_encode = nullptr;

static CUtensorMapSwizzle swz_for(int box_bytes) {
    return box_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : box_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                       : box_bytes == 32   ? CU_TENSOR_MAP_SWIZZLE_32B
                                                                           : CU_TENSOR_MAP_SWIZZLE_NONE;
}
// [rows, c_real] bf16 matrix with row pitch c_real*2 bytes; box = {box_c, 1} (box_c may exceed c_real: zero fill)
static CUtensorMap make_map(void* ptr, long long rows, int c_real, int box_c) {
    CUtensorMap m;
    cuuint64_t gdim[2] = {(cuuint64_t)c_real, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)c_real * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_c, 1};
    cuuint32_t es[2] = {1, 1};
    CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swz_for(box_c * 2), CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        printf("cuTensorMapEncodeTiled failed: %d (rows %lld c %d box %d)\n", (int)r, rows, c_real, box_c);
        exit(1);
    }
    return m;
}

// ------------------------------------------------------------------ device helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar), done = 0;
    for (unsigned spin = 0; spin < (1u << 22); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return true;
    }
    return false;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(dst), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// byte offset inside a K-major operand tile whose rows are `rowb` bytes (= the swizzle span), 16-byte chunk c of row r
__host__ __device__ inline uint32_t swz_off(int r, int c, int rowb) {
    uint32_t off = (uint32_t)r * rowb + (uint32_t)c * 16;
    uint32_t mask = rowb == 128 ? 7u : rowb == 64 ? 3u : rowb == 32 ? 1u : 0u;
    return off ^ (((off >> 7) & mask) << 4);
}

// ------------------------------------------------------------------ test 1: semantics
__global__ void k_sem(const __grid_constant__ CUtensorMap map, int rowb, int4 i0, int4 i1, unsigned char* out, int* status) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar;
    for (int i = threadIdx.x; i < 8 * rowb; i += blockDim.x) sm[i] = 0xAB;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        fence_async_smem();
        mbar_expect_tx(&bar, 8 * rowb);
        tma_gather4(smem_u32(sm), &map, 0, i0.x, i0.y, i0.z, i0.w, &bar);
        tma_gather4(smem_u32(sm) + 4 * rowb, &map, 0, i1.x, i1.y, i1.z, i1.w, &bar);
    }
    bool ok = mbar_wait(&bar, 0);
    if (threadIdx.x == 0) *status = ok ? 0 : 1;
    __syncthreads();
    for (int i = threadIdx.x; i < 8 * rowb; i += blockDim.x) out[i] = sm[i];
}

// ------------------------------------------------------------------ test 2: gathered tile as a tcgen05 operand
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// K-major, swizzled: SBO = 8 rows * rowb; layout type 2 / 4 / 6 for 128 / 64 / 32-byte swizzle
__device__ __forceinline__ uint64_t umma_desc_sw(uint32_t saddr, int rowb) {
    uint64_t lt = rowb == 128 ? 2 : rowb == 64 ? 4 : 6;
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(((uint32_t)(8 * rowb) >> 4) & 0x3FFFu) << 32) | (1ULL << 46) | (lt << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
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
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// one tile: D[128, NR] = gather(A)[128, KC] x B^T, A rows through gather4 (use_tma) or swizzled cp.async
template <int KC, int NR>
__global__ void __launch_bounds__(128) k_mma(const __grid_constant__ CUtensorMap map, const __nv_bfloat16* feats, int c_real,
                                             const int* idx, const unsigned char* bimg, float* out, int use_tma, int* status) {
    constexpr int ROWB = KC * 2;
    extern __shared__ __align__(1024) unsigned char sm[];
    unsigned char* A = sm;                 // 128 * ROWB
    unsigned char* B = sm + 128 * ROWB;    // NR * ROWB  (multiple of 1024 for every case used here)
    __shared__ __align__(8) uint64_t full, done;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(64u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        mbar_init(&full, use_tma ? 1 : 129);
        mbar_init(&done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_s;
    if (use_tma) {
        if (tid == 0) {
            mbar_expect_tx(&full, 128 * ROWB + NR * ROWB);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(B)),
                         "l"(bimg), "r"((uint32_t)(NR * ROWB)), "r"(smem_u32(&full))
                         : "memory");
        }
        if (warp == 0) {
            const int4 q = *reinterpret_cast<const int4*>(idx + 4 * lane);
            tma_gather4(smem_u32(A) + lane * 4 * ROWB, &map, 0, q.x, q.y, q.z, q.w, &full);
        }
    } else {
        if (tid == 0) {
            mbar_expect_tx(&full, NR * ROWB);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(B)),
                         "l"(bimg), "r"((uint32_t)(NR * ROWB)), "r"(smem_u32(&full))
                         : "memory");
        }
        const int r = tid;
        const int src = idx[r];
        const bool v = src >= 0;       // (out-of-range >= N indices are not used on this path)
        for (int c = 0; c < ROWB / 16; ++c) {
            const bool vc = v && c * 8 < c_real;
            cp_async16(smem_u32(A) + swz_off(r, c, ROWB), feats + (size_t)(v ? src : 0) * c_real + (vc ? c * 8 : 0), vc);
        }
        cp_async_arrive_noinc(&full);
    }
    bool ok = true;
    if (tid == 0) {
        ok = mbar_wait(&full, 0);
        fence_async_smem();
        tc_fence_after();
        constexpr uint32_t IDESC = umma_idesc(128, NR);
#pragma unroll
        for (int m = 0; m < KC / 16; ++m)
            umma_f16(tmem, umma_desc_sw(smem_u32(A) + m * 32, ROWB), umma_desc_sw(smem_u32(B) + m * 32, ROWB), IDESC, m > 0);
        umma_commit(&done);
    }
    ok &= mbar_wait(&done, 0);
    tc_fence_after();
    if (!ok) atomicExch(status, 1);
    const int r = warp * 32 + lane;
#pragma unroll
    for (int c0 = 0; c0 < NR; c0 += 16) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        for (int i = 0; i < 16; ++i) out[r * NR + c0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u));
}

// ------------------------------------------------------------------ test 3: gather throughput
// persistent CTAs; producer warps fill a ring of 128-row stages, one consumer thread frees them (no MMA: the gather alone)
//   MODE 0: TMA gather4, every lane l < 32/NPW of each producer warp issues one gather4 per stage
//   MODE 1: TMA gather4, lane 0 of each producer warp issues all of the warp's gather4 (indices through shuffles)
//   MODE 2: 16-byte cp.async, full-sector lane mapping, completion through cp.async.mbarrier.arrive.noinc
template <int ROWB, int MODE, int NPW>
__global__ void __launch_bounds__(32 * (NPW + 1)) k_thr(const __grid_constant__ CUtensorMap map, const __nv_bfloat16* feats,
                                                        const int* idx, int stages_per_cta, int S, int* status,
                                                        unsigned long long* sink) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) uint64_t full[32], empty[32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int STAGE = 128 * ROWB;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], MODE == 2 ? 32 * NPW : 1);
            mbar_init(&empty[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int* my_idx = idx + (size_t)blockIdx.x * stages_per_cta * 128;
    bool ok = true;
    if (warp < NPW) {
        constexpr int RPW = 128 / NPW;   // rows per producer warp
        for (int g = 0; g < stages_per_cta; ++g) {
            const int s = g % S;
            if (g >= S) ok &= mbar_wait(&empty[s], ((g / S) - 1) & 1);
            const uint32_t base = smem_u32(sm) + s * STAGE;
            const int* ti = my_idx + (size_t)g * 128 + warp * RPW;
            if (MODE == 0) {
                if (warp == 0 && lane == 0) mbar_expect_tx(&full[s], STAGE);
                if (lane < RPW / 4) {
                    const int4 q = *reinterpret_cast<const int4*>(ti + 4 * lane);
                    tma_gather4(base + (warp * RPW + 4 * lane) * ROWB, &map, 0, q.x, q.y, q.z, q.w, &full[s]);
                }
            } else if (MODE == 1) {
                int4 q = make_int4(-1, -1, -1, -1);
                if (lane < RPW / 4) q = *reinterpret_cast<const int4*>(ti + 4 * lane);
                if (warp == 0 && lane == 0) mbar_expect_tx(&full[s], STAGE);
#pragma unroll
                for (int j = 0; j < RPW / 4; ++j) {
                    const int a = __shfl_sync(0xffffffffu, q.x, j), b = __shfl_sync(0xffffffffu, q.y, j);
                    const int c = __shfl_sync(0xffffffffu, q.z, j), d = __shfl_sync(0xffffffffu, q.w, j);
                    if (lane == 0) tma_gather4(base + (warp * RPW + 4 * j) * ROWB, &map, 0, a, b, c, d, &full[s]);
                }
            } else {
                constexpr int CPR = ROWB / 16;
                constexpr int CW = CPR < 4 ? CPR : 4;       // chunks of one row handled by adjacent lanes
                constexpr int RPI = 32 / CW;                // rows per warp instruction
                const int c_sub = lane % CW, r_sub = lane / CW;
#pragma unroll
                for (int it = 0; it < RPW / RPI; ++it) {
                    const int r = warp * RPW + it * RPI + r_sub;
                    const int src = ti[it * RPI + r_sub];
#pragma unroll
                    for (int cg = 0; cg < CPR / CW; ++cg) {
                        const int c = cg * CW + c_sub;
                        cp_async16(base + swz_off(r, c, ROWB), feats + (size_t)(src < 0 ? 0 : src) * (ROWB / 2) + c * 8, src >= 0);
                    }
                }
                cp_async_arrive_noinc(&full[s]);
            }
        }
    } else if (lane == 0) {
        unsigned long long acc = 0;
        for (int g = 0; g < stages_per_cta; ++g) {
            const int s = g % S;
            ok &= mbar_wait(&full[s], (g / S) & 1);
            acc += *reinterpret_cast<volatile unsigned long long*>(sm + s * STAGE + (g & 15) * 64);
            mbar_arrive(&empty[s]);
        }
        if (acc == 0x123456789ULL) *sink = acc;
    }
    if (!ok) atomicExch(status, 1);
}

// ------------------------------------------------------------------ host
static float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

static int g_status_host = 0;
static int* d_status = nullptr;

static void test_sem(int c_real, int box_c, long long N, __nv_bfloat16* d_feats, const std::vector<float>& h_feats) {
    const int rowb = box_c * 2;
    CUtensorMap map = make_map(d_feats, N, c_real, box_c);
    int4 i0 = make_int4(3, -1, (int)N, 7), i1 = make_int4((int)N - 1, 0, 100, 1 << 30);
    unsigned char* d_out;
    CK(cudaMalloc(&d_out, 8 * rowb));
    CK(cudaMemset(d_status, 0, 4));
    k_sem<<<1, 64, 8 * rowb>>>(map, rowb, i0, i1, d_out, d_status);
    CK(cudaDeviceSynchronize());
    std::vector<unsigned char> out(8 * rowb);
    CK(cudaMemcpy(out.data(), d_out, 8 * rowb, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&g_status_host, d_status, 4, cudaMemcpyDeviceToHost));
    int rows[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
    int bad = 0;
    for (int r = 0; r < 8; ++r)
        for (int c = 0; c < rowb / 16; ++c)
            for (int e = 0; e < 8; ++e) {
                const int ch = c * 8 + e;
                float want = 0.f;
                if (rows[r] >= 0 && rows[r] < N && ch < c_real) want = h_feats[(size_t)rows[r] * c_real + ch];
                __nv_bfloat16 got;
                memcpy(&got, &out[swz_off(r, c, rowb) + e * 2], 2);
                if (__bfloat162float(got) != want) ++bad;
            }
    printf("[sem] c_real=%2d box=%2d (row %3d B): barrier %s, %d mismatching elements of %d -> %s\n", c_real, box_c, rowb,
           g_status_host ? "TIMED OUT" : "completed", bad, 8 * rowb / 2, (bad == 0 && !g_status_host) ? "OK" : "FAIL");
    CK(cudaFree(d_out));
}

template <int KC, int NR>
static void test_mma(int c_real, long long N, __nv_bfloat16* d_feats, const std::vector<float>& h_feats, int use_tma) {
    constexpr int ROWB = KC * 2;
    CUtensorMap map = make_map(d_feats, N, c_real, KC);
    std::vector<int> idx(128);
    for (int r = 0; r < 128; ++r) idx[r] = (r % 5 == 3) ? -1 : (int)((r * 7919LL + 13) % N);
    std::vector<float> w((size_t)NR * KC);
    for (auto& x : w) x = bf16_round((float)((rand() % 17) - 8) / 8.f);
    std::vector<unsigned char> bimg((size_t)NR * ROWB, 0);
    for (int n = 0; n < NR; ++n)
        for (int kk = 0; kk < KC; ++kk) {
            __nv_bfloat16 b = __float2bfloat16_rn(kk < c_real ? w[(size_t)n * KC + kk] : 0.f);
            memcpy(&bimg[swz_off(n, kk / 8, ROWB) + (kk % 8) * 2], &b, 2);
        }
    int* d_idx;
    unsigned char* d_b;
    float* d_out;
    CK(cudaMalloc(&d_idx, 512));
    CK(cudaMalloc(&d_b, bimg.size()));
    CK(cudaMalloc(&d_out, 128 * NR * 4));
    CK(cudaMemcpy(d_idx, idx.data(), 512, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_b, bimg.data(), bimg.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(d_status, 0, 4));
    const int smem = 128 * ROWB + NR * ROWB + 1024;
    CK(cudaFuncSetAttribute(k_mma<KC, NR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_mma<KC, NR><<<1, 128, smem>>>(map, d_feats, c_real, d_idx, d_b, d_out, use_tma, d_status);
    CK(cudaDeviceSynchronize());
    std::vector<float> out(128 * NR);
    CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&g_status_host, d_status, 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (int r = 0; r < 128; ++r)
        for (int n = 0; n < NR; ++n) {
            double ref = 0;
            if (idx[r] >= 0)
                for (int kk = 0; kk < c_real && kk < KC; ++kk) ref += (double)h_feats[(size_t)idx[r] * c_real + kk] * w[(size_t)n * KC + kk];
            maxerr = fmax(maxerr, fabs(ref - out[r * NR + n]));
            maxref = fmax(maxref, fabs(ref));
        }
    printf("[mma] KC=%2d (c_real %2d) NR=%2d %s: status %d, max|err| %.3g (max|ref| %.3g) -> %s\n", KC, c_real, NR,
           use_tma ? "gather4 " : "cp.async", g_status_host, maxerr, maxref, (maxerr <= 1e-3 * maxref && !g_status_host) ? "OK" : "FAIL");
    CK(cudaFree(d_idx));
    CK(cudaFree(d_b));
    CK(cudaFree(d_out));
}

template <int ROWB, int MODE, int NPW>
static void run_thr(const char* label, long long N, __nv_bfloat16* d_feats, const int* d_idx, long long n_valid, int ctas,
                    int stages_per_cta, int S) {
    CUtensorMap map = make_map(d_feats, N, ROWB / 2, ROWB / 2);
    const int smem = S * 128 * ROWB + 1024;
    auto kern = k_thr<ROWB, MODE, NPW>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    unsigned long long* d_sink;
    CK(cudaMalloc(&d_sink, 8));
    CK(cudaMemset(d_status, 0, 4));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(a));
        kern<<<ctas, 32 * (NPW + 1), smem>>>(map, d_feats, d_idx, stages_per_cta, S, d_status, d_sink);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (rep > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpy(&g_status_host, d_status, 4, cudaMemcpyDeviceToHost));
    const double rows_total = (double)ctas * stages_per_cta * 128;
    printf("[thr] %-34s row %3d B ring %2d x %5d B, %4d CTAs: %8.1f us  %7.1f GB/s useful  %6.1f Grow-slots/s  status %d\n", label, ROWB,
           S, 128 * ROWB, ctas, best * 1e3, n_valid * (double)ROWB / (best * 1e-3) / 1e9, rows_total / (best * 1e-3) / 1e9, g_status_host);
    CK(cudaFree(d_sink));
}

int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
        printf("cuTensorMapEncodeTiled not available\n");
        return 1;
    }
    g_encode = (EncodeFn)fn;
    CK(cudaMalloc(&d_status, 4));
    srand(1);
    const long long N = 140000;

    for (int c_real : {8, 16, 32, 64}) {
        std::vector<float> h((size_t)N * c_real);
        std::vector<__nv_bfloat16> hb(h.size());
        for (size_t i = 0; i < h.size(); ++i) {
            h[i] = (float)((int)(rand() % 63) - 31) / 16.f;
            hb[i] = __float2bfloat16_rn(h[i]);
        }
        __nv_bfloat16* d;
        CK(cudaMalloc(&d, hb.size() * 2 + 256));
        CK(cudaMemcpy(d, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
        test_sem(c_real, c_real < 16 ? 16 : c_real, N, d, h);
        if (c_real == 8) {
            test_mma<16, 16>(8, N, d, h, 1);
            test_mma<16, 16>(8, N, d, h, 0);
        } else if (c_real == 16) {
            test_mma<16, 16>(16, N, d, h, 1);
            test_mma<16, 32>(16, N, d, h, 1);
            test_mma<16, 32>(16, N, d, h, 0);
        } else if (c_real == 32) {
            test_mma<32, 32>(32, N, d, h, 1);
            test_mma<32, 64>(32, N, d, h, 1);
            test_mma<32, 16>(32, N, d, h, 0);
        } else {
            test_mma<64, 64>(64, N, d, h, 1);
            test_mma<64, 32>(64, N, d, h, 1);
            test_mma<64, 32>(64, N, d, h, 0);
        }
        CK(cudaFree(d));
    }

    // throughput: 148 * k persistent CTAs, each `spc` stages of 128 row slots
    for (int rowb : {32, 64, 128}) {
        const int c = rowb / 2;
        std::vector<__nv_bfloat16> hb((size_t)N * c);
        for (size_t i = 0; i < hb.size(); ++i) hb[i] = __float2bfloat16_rn((float)(i % 7));
        __nv_bfloat16* d;
        CK(cudaMalloc(&d, hb.size() * 2 + 256));
        CK(cudaMemcpy(d, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
        for (int pattern = 0; pattern < 3; ++pattern) {
            // 0: 43 % fill, neighbourhood-local rows; 1: 100 % fill local; 2: 43 % fill, rows uniformly random
            const double fill = pattern == 1 ? 1.0 : 0.43;
            const int max_ctas = 148 * 4, spc = 200;
            std::vector<int> idx((size_t)max_ctas * spc * 128);
            long long nv[3] = {0, 0, 0};   // valid rows for 148 / 296 / 592 CTAs
            for (int cta = 0; cta < max_ctas; ++cta)
                for (int g = 0; g < spc; ++g)
                    for (int r = 0; r < 128; ++r) {
                        const size_t p = ((size_t)cta * spc + g) * 128 + r;
                        const bool v = (rand() / (double)RAND_MAX) < fill;
                        long long row;
                        if (pattern == 2) row = (long long)(rand() % N);
                        else {
                            row = ((long long)cta * 231 + (g / 27) * 128 + r + (rand() % 600) - 300) % N;
                            if (row < 0) row += N;
                        }
                        idx[p] = v ? (int)row : -1;
                        if (v) {
                            if (cta < 148) ++nv[0];
                            if (cta < 296) ++nv[1];
                            ++nv[2];
                        }
                    }
            int* d_idx;
            CK(cudaMalloc(&d_idx, idx.size() * 4));
            CK(cudaMemcpy(d_idx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice));
            printf("---- row %d B, pattern %d (%s)\n", rowb, pattern,
                   pattern == 0 ? "43% fill, local" : pattern == 1 ? "100% fill, local" : "43% fill, random rows");
#define THR(ROWB_, MODE_, NPW_, LABEL, CPS, S)                                                               \
    if (rowb == ROWB_) run_thr<ROWB_, MODE_, NPW_>(LABEL, N, d, d_idx, nv[CPS == 1 ? 0 : CPS == 2 ? 1 : 2], 148 * CPS, spc, S);
#define THR_ALL(ROWB_)                                                           \
    THR(ROWB_, 0, 1, "gather4 1 warp x 32 lanes", 1, 8)                          \
    THR(ROWB_, 0, 4, "gather4 4 warps x 8 lanes", 1, 8)                          \
    THR(ROWB_, 1, 4, "gather4 4 warps, lane 0 issues", 1, 8)                     \
    THR(ROWB_, 1, 1, "gather4 1 warp, lane 0 issues", 1, 8)                      \
    THR(ROWB_, 2, 4, "cp.async 4 warps", 1, 8)                                   \
    THR(ROWB_, 0, 4, "gather4 4 warps x 8 lanes", 1, (ROWB_ == 128 ? 12 : 16))   \
    THR(ROWB_, 2, 4, "cp.async 4 warps", 1, (ROWB_ == 128 ? 12 : 16))            \
    THR(ROWB_, 0, 4, "gather4 4 warps x 8 lanes", 2, 6)                          \
    THR(ROWB_, 2, 4, "cp.async 4 warps", 2, 6)                                   \
    THR(ROWB_, 0, 4, "gather4 4 warps x 8 lanes", 4, 3)                          \
    THR(ROWB_, 2, 4, "cp.async 4 warps", 4, 3)                                   \
    THR(ROWB_, 0, 4, "gather4 4 warps x 8 lanes", 4, 2)                          \
    THR(ROWB_, 2, 4, "cp.async 4 warps", 4, 2)
            THR_ALL(32)
            THR_ALL(64)
            THR_ALL(128)
            CK(cudaFree(d_idx));
        }
        CK(cudaFree(d));
    }
    printf("done\n");
    return 0;
}
