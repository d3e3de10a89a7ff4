// bf16 tensor-core sparse convolution (forward and dgrad) — tcgen05 / TMEM, sm_100a only.
//
// Same output-stationary tiling as conv_f32.cu (one CTA = 128 output rows, loop over the kernel offsets that
// touch the tile), but the inner [128 x C_in] x [C_in x C_out] contraction of every offset runs on the 5th-gen
// tensor cores:
//   * operands are bf16.  The gathered neighbour rows land in shared memory through 16-byte cp.async chunks
//     (zero-fill for missing neighbours) directly in the UMMA canonical K-major, no-swizzle layout
//     (8-row x 16-byte core matrices: chunk (r, c) at ((r/8)*C/8 + c)*128 + (r%8)*16), the weight slice of the
//     offset is a pre-formatted image copied linearly — no register staging, no re-layout pass;
//   * one elected thread issues C_in/16 `tcgen05.mma.cta_group::1.kind::f16` (M=128, N=C_out, K=16) per offset,
//     accumulating all offsets of the tile in TMEM (fp32, C_out columns); `tcgen05.commit` onto an mbarrier
//     frees the operand stage; a 3-deep ring keeps 2 gathers in flight behind the MMA;
//   * epilogue: `tcgen05.ld` 32x32b (thread = output row) -> smem -> coalesced fp32 stores + per-tile BN sums.
// Replaces spconv `ops.indice_conv` behind spconv_backbone.py:89,92-93,113,563-564 in the bf16 (benchmark)
// precision mode; the fp32 kernels in conv_f32.cu stay the 1e-4 parity path.
// Algorithmic bytes per launch (e=2 for the gathered operand, 4 for the fp32 output):
//   N_in*C_in*2 + N_out*C_out*4 + P*8 + K*C_in*C_out*2;  FLOPs 2*P*C_in*C_out.
#include <cuda_bf16.h>

#include "common.cuh"

namespace vc {

static constexpr int TCM = 128;        // rows per tile == UMMA M
static constexpr int TC_THREADS = 128; // 4 warps: warp w owns TMEM lanes [32w, 32w+32)
static constexpr int TC_STAGES = 3;
static constexpr int MAXK_TC = 32;
static constexpr unsigned SPIN_LIMIT = 1u << 24;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// bounded spin: a wedged pipeline must not hang the GPU (sets *err and returns false instead)
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* err) {
    uint32_t addr = smem_u32(bar), done = 0;
    for (unsigned spin = 0; spin < SPIN_LIMIT; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return true;
    }
    if (err) atomicExch(err, 1);
    return false;
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
//   [0,14) start>>4   [16,30) LBO>>4 (between the two 16-byte K chunks of one MMA)   [32,46) SBO>>4 (between 8-row groups)
//   [46,48) version = 1   [61,64) layout type = 0
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ULL << 46);
}
// instruction descriptor (InstrDescriptor): D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1), K-major both,
// N>>3 at bit 17, M>>4 at bit 24
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
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

template <int KC, int NR>
struct TcCfg {
    static constexpr int CPR = KC / 8;                       // 16-byte chunks per gathered row
    static constexpr int A_BYTES = TCM * KC * 2;
    static constexpr int B_BYTES = NR * KC * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = NR < 32 ? 32 : NR;      // power of two >= 32
    static constexpr int EPI_BYTES = TCM * (NR + 1) * 4;     // epilogue staging, aliases the stage ring
    static constexpr int RING_BYTES = TC_STAGES * STAGE_BYTES > EPI_BYTES ? TC_STAGES * STAGE_BYTES : EPI_BYTES;
    static constexpr size_t smem(int K) { return (size_t)RING_BYTES + (size_t)K * TCM * 4; }
};

// f32 -> bf16 cast of a feature matrix (the gathered operand)
__global__ void cast_bf16_kernel(const float4* __restrict__ in, uint2* __restrict__ out, size_t n4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) {
        float4 v = in[i];
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&a);
        o.y = *reinterpret_cast<uint32_t*>(&b);
        out[i] = o;
    }
}

// weight images: per offset k a [NR rows][KC] bf16 matrix in the UMMA K-major core-matrix layout.
//   mode 0 (forward): B[n=co][kk=ci] = w[co][k][ci]
//   mode 1 (dgrad)  : B[n=ci][kk=co] = w[co][k'][ci],  k' = mirror ? K-1-k : k
__global__ void prep_weights_tc_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ img, int cin, int cout,
                                       int K, int mode, int mirror) {
    int NRr = mode == 0 ? cout : cin, KCc = mode == 0 ? cin : cout;
    int total = K * NRr * KCc;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int kk = i % KCc, n = (i / KCc) % NRr, k = i / (KCc * NRr);
    float v;
    if (mode == 0) {
        v = w[((size_t)n * K + k) * cin + kk];
    } else {
        int ks = mirror ? (K - 1 - k) : k;
        v = w[((size_t)kk * K + ks) * cin + n];
    }
    size_t off = (size_t)k * NRr * KCc + ((size_t)((n >> 3) * (KCc >> 3) + (kk >> 3)) * 64) + (n & 7) * 8 + (kk & 7);
    img[off] = __float2bfloat16_rn(v);
}

template <int KC, int NR>
__global__ void __launch_bounds__(TC_THREADS)
tc_gather_gemm_kernel(const __nv_bfloat16* __restrict__ in, const __nv_bfloat16* __restrict__ wimg,
                      const int32_t* __restrict__ nbr, float* __restrict__ out, int n_out, int K,
                      float* __restrict__ bn_partial, int* __restrict__ err) {
    using C = TcCfg<KC, NR>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* ring = smem_raw;                                          // [stages][A | B]
    int* nbr_s = reinterpret_cast<int*>(smem_raw + C::RING_BYTES);           // [K][128]
    __shared__ __align__(8) uint64_t mma_done[TC_STAGES];
    __shared__ uint32_t tmem_base_s;
    __shared__ int klist[MAXK_TC];
    __shared__ unsigned kmask;
    __shared__ float red[4][2][NR];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int base = blockIdx.x * TCM;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)C::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < TC_STAGES; ++s) mbar_init(&mma_done[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        kmask = 0u;
    }
    // stage the tile's slice of the neighbour table
    for (int i = tid; i < K * TCM; i += TC_THREADS) {
        int k = i / TCM, r = i % TCM, row = base + r;
        nbr_s[i] = (row < n_out) ? __ldg(nbr + (size_t)k * n_out + row) : -1;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    for (int k = warp; k < K; k += TC_THREADS / 32) {
        bool any = false;
#pragma unroll
        for (int j = 0; j < TCM / 32; ++j) any |= nbr_s[k * TCM + j * 32 + lane] >= 0;
        if (__any_sync(0xffffffffu, any) && lane == 0) atomicOr(&kmask, 1u << k);
    }
    __syncthreads();
    const unsigned km = kmask;
    const int nk = __popc(km);
    if (tid == 0) {
        int c = 0;
        for (int k = 0; k < K; ++k)
            if (km >> k & 1u) klist[c++] = k;
    }
    __syncthreads();

    // gather mapping: per warp instruction 8 rows x (up to) 4 chunks -> conflict-free smem writes, full sectors
    constexpr int CW = C::CPR < 4 ? C::CPR : 4;        // chunks of one row covered by one instruction
    constexpr int RPI = 8 * (4 / CW);                  // rows per instruction
    const int rl = lane & 7, xq = lane >> 3;
    const int c_sub = xq % CW, r_sub = xq / CW;

    auto issue_stage = [&](int t) {
        const int k = klist[t];
        unsigned char* A = ring + (t % TC_STAGES) * C::STAGE_BYTES;
        unsigned char* B = A + C::A_BYTES;
        const int* nk_ = nbr_s + k * TCM;
#pragma unroll
        for (int it = 0; it < 32 / RPI; ++it) {
            const int r = warp * 32 + it * RPI + r_sub * 8 + rl;
            const int src = nk_[r];
            const __nv_bfloat16* srow = in + (size_t)(src < 0 ? 0 : src) * KC;
#pragma unroll
            for (int cg = 0; cg < C::CPR / CW; ++cg) {
                const int c = cg * CW + c_sub;
                cp_async16(A + ((r >> 3) * C::CPR + c) * 128 + (r & 7) * 16, srow + c * 8, src >= 0);
            }
        }
        const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(wimg) + (size_t)k * C::B_BYTES;
        for (int q = tid; q < C::B_BYTES / 16; q += TC_THREADS) cp_async16(B + q * 16, wsrc + q * 16, true);
    };

    bool ok = true;
    if (nk > 0) {
#pragma unroll
        for (int t = 0; t < TC_STAGES - 1; ++t) {
            if (t < nk) issue_stage(t);
            cp_async_commit();
        }
        constexpr uint32_t IDESC = umma_idesc(TCM, NR);
        for (int t = 0; t < nk; ++t) {
            const int tn = t + TC_STAGES - 1;
            if (tn < nk) {
                if (tn >= TC_STAGES) ok &= mbar_wait(&mma_done[tn % TC_STAGES], (uint32_t)((tn / TC_STAGES - 1) & 1), err);
                issue_stage(tn);
            }
            cp_async_commit();
            cp_async_wait<TC_STAGES - 1>();     // this thread's chunks of stage t have landed
            fence_async_smem();                 // generic-proxy writes -> visible to the tensor core (async proxy)
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                const uint32_t a0 = smem_u32(ring + (t % TC_STAGES) * C::STAGE_BYTES);
                const uint32_t b0 = a0 + C::A_BYTES;
#pragma unroll
                for (int m = 0; m < KC / 16; ++m) {
                    const uint64_t ad = umma_desc(a0 + m * 256, 128, C::CPR * 128);
                    const uint64_t bd = umma_desc(b0 + m * 256, 128, C::CPR * 128);
                    umma_f16(tmem_base, ad, bd, IDESC, (t > 0 || m > 0) ? 1u : 0u);
                }
                umma_commit(&mma_done[t % TC_STAGES]);
            }
        }
        ok &= mbar_wait(&mma_done[(nk - 1) % TC_STAGES], (uint32_t)(((nk - 1) / TC_STAGES) & 1), err);
        tc_fence_after();
    }
    __syncthreads();    // every thread is past the last operand use: the ring can be reused as epilogue staging

    float* stg = reinterpret_cast<float*>(ring);    // [128][NR+1]
    {
        const int r = warp * 32 + lane;
#pragma unroll
        for (int c0 = 0; c0 < NR; c0 += 16) {
            float v[16];
            if (nk > 0) {
                tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) stg[r * (NR + 1) + c0 + i] = v[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    // coalesced fp32 stores: consecutive threads -> consecutive float4 of the [128, NR] tile
    for (int q = tid; q < TCM * NR / 4; q += TC_THREADS) {
        const int r = q / (NR / 4), c4 = q % (NR / 4);
        if (base + r < n_out) {
            const float* s = stg + r * (NR + 1) + c4 * 4;
            *reinterpret_cast<float4*>(out + (size_t)(base + r) * NR + c4 * 4) = make_float4(s[0], s[1], s[2], s[3]);
        }
    }
    if (bn_partial != nullptr) {
        // column sums over the tile's valid rows: thread (ch, quarter) sums 32 rows
        const int rows_valid = min(TCM, n_out - base);
        for (int idx = tid; idx < NR * 4; idx += TC_THREADS) {
            const int ch = idx % NR, qd = idx / NR;
            float s = 0.f, q2 = 0.f;
            for (int r = qd * 32; r < min(qd * 32 + 32, rows_valid); ++r) {
                float x = stg[r * (NR + 1) + ch];
                s += x;
                q2 = fmaf(x, x, q2);
            }
            red[qd][0][ch] = s;
            red[qd][1][ch] = q2;
        }
        __syncthreads();
        for (int idx = tid; idx < 2 * NR; idx += TC_THREADS) {
            const int which = idx / NR, ch = idx % NR;
            bn_partial[((size_t)blockIdx.x * 2 + which) * NR + ch] =
                red[0][which][ch] + red[1][which][ch] + red[2][which][ch] + red[3][which][ch];
        }
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS));
    }
    (void)ok;
}

template <int KC, int NR>
static int launch_tc(const __nv_bfloat16* in, const __nv_bfloat16* wimg, const int32_t* nbr, float* out, int n_out, int K,
                     float* bn_partial, int* err, cudaStream_t stream) {
    size_t smem = TcCfg<KC, NR>::smem(K);
    auto kern = tc_gather_gemm_kernel<KC, NR>;
    VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<cdiv(n_out, TCM), TC_THREADS, smem, stream>>>(in, wimg, nbr, out, n_out, K, bn_partial, err);
    VC_LAUNCH_CHECK();
    return VC_OK;
}

static bool tc_ch_ok(int c) { return c == 16 || c == 32 || c == 64; }

static int dispatch_tc(int kc, int nr, const __nv_bfloat16* in, const __nv_bfloat16* wimg, const int32_t* nbr, float* out,
                       int n_out, int K, float* bn_partial, int* err, cudaStream_t stream) {
#define VC_TC_CASE(A, B) \
    if (kc == A && nr == B) return launch_tc<A, B>(in, wimg, nbr, out, n_out, K, bn_partial, err, stream);
    VC_TC_CASE(16, 16) VC_TC_CASE(16, 32) VC_TC_CASE(16, 64)
    VC_TC_CASE(32, 16) VC_TC_CASE(32, 32) VC_TC_CASE(32, 64)
    VC_TC_CASE(64, 16) VC_TC_CASE(64, 32) VC_TC_CASE(64, 64)
#undef VC_TC_CASE
    set_error("tensor-core conv: unsupported channel pair (%d, %d), need 16/32/64", kc, nr);
    return VC_ERR_UNSUPPORTED;
}

}  // namespace vc

using namespace vc;

extern "C" int vc_cast_f32_bf16(const float* in, void* out, long long n, vc_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    VC_CHECK_ARG(n >= 0 && n % 4 == 0, "cast: element count %lld must be a multiple of 4", n);
    if (n == 0) return VC_OK;
    VC_CHECK_ARG(in && out, "null pointer");
    size_t n4 = (size_t)n / 4;
    int blocks = (int)((n4 + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    cast_bf16_kernel<<<blocks, 256, 0, stream>>>((const float4*)in, (uint2*)out, n4);
    VC_LAUNCH_CHECK();
    return VC_OK;
}

extern "C" size_t vc_conv_tc_ws_bytes(int cin, int cout, int K) { return (size_t)K * cin * cout * 2; }

static int tc_common(const void* feats_bf16, const float* w, const int32_t* nbr, float* out, int n_rows, int cin, int cout,
                     int K, int mode, int mirror, float* bn_partial, void* ws, size_t ws_bytes, int32_t* err, cudaStream_t stream) {
    VC_CHECK_ARG(n_rows >= 0 && K >= 1 && K <= MAXK_TC, "bad n=%d or K=%d", n_rows, K);
    if (!tc_ch_ok(cin) || !tc_ch_ok(cout)) {
        set_error("tensor-core conv: unsupported channels cin=%d cout=%d (need 16/32/64)", cin, cout);
        return VC_ERR_UNSUPPORTED;
    }
    if (n_rows == 0) return VC_OK;
    VC_CHECK_ARG(feats_bf16 && w && nbr && out && ws, "null pointer");
    if (ws_bytes < vc_conv_tc_ws_bytes(cin, cout, K)) {
        set_error("tensor-core conv workspace %zu < %zu", ws_bytes, vc_conv_tc_ws_bytes(cin, cout, K));
        return VC_ERR_WORKSPACE;
    }
    __nv_bfloat16* img = (__nv_bfloat16*)ws;
    int total = K * cin * cout;
    prep_weights_tc_kernel<<<cdiv(total, 256), 256, 0, stream>>>(w, img, cin, cout, K, mode, mirror);
    VC_LAUNCH_CHECK();
    int kc = mode == 0 ? cin : cout, nr = mode == 0 ? cout : cin;
    return dispatch_tc(kc, nr, (const __nv_bfloat16*)feats_bf16, img, nbr, out, n_rows, K, bn_partial, err, stream);
}

extern "C" int vc_conv_fwd_tc(const void* in_bf16, const float* w, const int32_t* nbr, float* out, int n_out, int cin,
                              int cout, int K, float* bn_partial, void* ws, size_t ws_bytes, int32_t* err_flag,
                              vc_stream_t stream_) {
    return tc_common(in_bf16, w, nbr, out, n_out, cin, cout, K, 0, 0, bn_partial, ws, ws_bytes, err_flag, (cudaStream_t)stream_);
}

extern "C" int vc_conv_dgrad_tc(const void* dout_bf16, const float* w, const int32_t* nbr_t, float* din, int n_in, int cin,
                                int cout, int K, int mirror, void* ws, size_t ws_bytes, int32_t* err_flag,
                                vc_stream_t stream_) {
    return tc_common(dout_bf16, w, nbr_t, din, n_in, cin, cout, K, 1, mirror, nullptr, ws, ws_bytes, err_flag,
                     (cudaStream_t)stream_);
}
