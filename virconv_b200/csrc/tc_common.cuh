// tcgen05 / TMEM / mbarrier helpers shared by the tensor-core kernels (conv_tc.cu, conv_tc2.cu, wgrad_tc2.cu) — sm_100a.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace vc {

static constexpr int TCM = 128;        // rows per tile == UMMA M
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
    if (err) atomicCAS(err, 0, 0x300);
    return false;
}
// time-bounded wait (2 s of %globaltimer): used by the persistent kernels, whose roles abandon their loops on a timeout
// `code` says which wait of which kernel gave up (0x1xx persistent conv, 0x2xx persistent wgrad; the first one sticks)
__device__ __forceinline__ bool mbar_wait_t(uint64_t* bar, uint32_t parity, int* err, int code = 1) {
    uint32_t addr = smem_u32(bar), done = 0;
    unsigned long long t0 = 0;
    for (unsigned spin = 0;; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return true;
        if ((spin & 255u) == 255u) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 2000000000ULL) break;
        }
    }
    if (err) atomicCAS(err, 0, code);
    return false;
}
// one non-blocking probe of a phase
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// the arrival is triggered when all cp.async operations issued so far by this thread have landed; it is one of the
// arrivals the barrier was initialised with (no increment of the pending count)
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// linear bulk copy global -> shared through the TMA engine, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
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
// K-major operand whose rows are exactly one swizzle span wide (ROWB = 32 / 64 / 128 bytes): 8-row swizzle atoms of
// 8*ROWB bytes follow each other (SBO), the K step inside a row is taken by advancing the start address by 32 bytes;
// layout type 6 / 4 / 2 = SWIZZLE_32B / 64B / 128B.  Verified on the device by profiles/exp_gather4.cu.
template <int ROWB>
__device__ __forceinline__ uint64_t umma_desc_sw(uint32_t saddr) {
    constexpr uint64_t LT = ROWB == 128 ? 2 : ROWB == 64 ? 4 : 6;
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(((uint32_t)(8 * ROWB) >> 4) & 0x3FFFu) << 32) | (1ULL << 46) | (LT << 61);
}
// byte offset of 16-byte chunk c of row r inside such a tile (the TMA / UMMA swizzle: address bits [4,4+b) ^= bits [7,7+b))
template <int ROWB>
__host__ __device__ __forceinline__ uint32_t swz_off(int r, int c) {
    constexpr uint32_t MASK = ROWB == 128 ? 7u : ROWB == 64 ? 3u : ROWB == 32 ? 1u : 0u;
    const uint32_t off = (uint32_t)r * ROWB + (uint32_t)c * 16;
    return off ^ (((off >> 7) & MASK) << 4);
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
// The same two instructions for a CONVERGED warp: every lane executes the asm, one elected lane issues.  Keeping the issue out
// of a divergent `if (lane == 0)` matters: tcgen05 operands live in uniform registers, and inside divergent code the compiler
// wraps every UTCHMMA / UTCBAR in an ELECT + R2UR.BROADCAST "waterfall" loop (~110 cycles per MMA measured, which made the
// MMA thread the bottleneck of the persistent kernels: profiles/trace_tc2_r2_e.txt)
__device__ __forceinline__ void umma_f16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(smem_u32(bar))
        : "memory");
}
// ---- lean issue path for the MMA warp of the persistent kernels -----------------------------------------------------------
// The MMA warp is ONE instruction stream per SM: at ~200 SASS instructions per ring stage (inlined timed wait, 64-bit
// descriptor arithmetic in vector registers, five R2UR per MMA, address re-derivation for every barrier) it needed ~900 cycles
// per 16 KB stage and was the bottleneck of the whole kernel (ncu source view, profiles/ncu_conv2_r2.txt).  These helpers keep
// the hot path short: barrier addresses are precomputed shared-window offsets, the wait spins inside one asm block (the
// time-bounded wait is the cold fall-back), and all MMAs of one kernel offset go out in one asm block with one elect.

// up to `tries` probes inside one asm block; 1 when the phase completed
__device__ __forceinline__ uint32_t mbar_spin(uint32_t bar_addr, uint32_t parity, uint32_t tries) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .u32 c;\n\t"
        "mov.u32 c, 0;\n"
        "VC_SPIN:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "@p bra VC_SPIN_DONE;\n\t"
        "add.u32 c, c, 1;\n\t"
        "setp.lt.u32 q, c, %3;\n\t"
        "@q bra VC_SPIN;\n"
        "VC_SPIN_DONE:\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar_addr), "r"(parity), "r"(tries)
        : "memory");
    return ok;
}
// time-bounded wait on a precomputed shared-window address (cold path behind mbar_spin)
__device__ __forceinline__ bool mbar_wait_t_addr(uint32_t addr, uint32_t parity, int* err, int code) {
    uint32_t done = 0;
    unsigned long long t0 = 0;
    for (unsigned spin = 0;; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return true;
        if ((spin & 255u) == 255u) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 2000000000ULL) break;
        }
    }
    if (err) atomicCAS(err, 0, code);
    return false;
}
__device__ __forceinline__ void umma_commit_elect_addr(uint32_t bar_addr) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(bar_addr)
        : "memory");
}
// high words of the shared-memory descriptors (constants of the layout): SBO | version | layout type
template <int ROWB>
__host__ __device__ constexpr uint32_t umma_desc_hi() {
    return (uint32_t)((8 * ROWB) >> 4) | (1u << 14) | ((ROWB == 128 ? 2u : ROWB == 64 ? 4u : 6u) << 29);
}
// NK MMAs of one series from a converged warp (one elect): descriptor low words a_lo / b_lo (start address >> 4, plus the LBO
// field where the layout has one) advance by STEP_A / STEP_B per MMA; the first MMA accumulates iff acc_first != 0
template <int NK, int STEP_A, int STEP_B>
__device__ __forceinline__ void umma_series(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t a_hi, uint32_t b_hi, uint32_t idesc,
                                            uint32_t acc_first) {
    static_assert(NK == 1 || NK == 2 || NK == 4 || NK == 8, "series length");
#define VC_MMA_NEXT(I)                                                  \
    "add.u32 al, %1, " #I "*%7;\n\tadd.u32 bl, %2, " #I "*%8;\n\t"      \
    "mov.b64 da, {al, %3};\n\tmov.b64 db, {bl, %4};\n\t"               \
    "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, t;\n\t"
    if constexpr (NK == 1) {
        asm volatile(
            "{\n\t.reg .pred q, p;\n\t.reg .b64 da, db;\n\t"
            "elect.sync _|q, 0xffffffff;\n\t"
            "setp.ne.b32 p, %6, 0;\n\t"
            "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %4};\n\t"
            "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
            ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(a_hi), "r"(b_hi), "r"(idesc), "r"(acc_first), "n"(STEP_A), "n"(STEP_B)
            : "memory");
    } else if constexpr (NK == 2) {
        asm volatile(
            "{\n\t.reg .pred q, p, t;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\t"
            "elect.sync _|q, 0xffffffff;\n\t"
            "setp.ne.b32 p, %6, 0;\n\tsetp.eq.u32 t, 0, 0;\n\t"
            "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %4};\n\t"
            "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
            VC_MMA_NEXT(1) "}"
            ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(a_hi), "r"(b_hi), "r"(idesc), "r"(acc_first), "n"(STEP_A), "n"(STEP_B)
            : "memory");
    } else if constexpr (NK == 4) {
        asm volatile(
            "{\n\t.reg .pred q, p, t;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\t"
            "elect.sync _|q, 0xffffffff;\n\t"
            "setp.ne.b32 p, %6, 0;\n\tsetp.eq.u32 t, 0, 0;\n\t"
            "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %4};\n\t"
            "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
            VC_MMA_NEXT(1) VC_MMA_NEXT(2) VC_MMA_NEXT(3) "}"
            ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(a_hi), "r"(b_hi), "r"(idesc), "r"(acc_first), "n"(STEP_A), "n"(STEP_B)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred q, p, t;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\t"
            "elect.sync _|q, 0xffffffff;\n\t"
            "setp.ne.b32 p, %6, 0;\n\tsetp.eq.u32 t, 0, 0;\n\t"
            "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %4};\n\t"
            "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
            VC_MMA_NEXT(1) VC_MMA_NEXT(2) VC_MMA_NEXT(3) VC_MMA_NEXT(4) VC_MMA_NEXT(5) VC_MMA_NEXT(6) VC_MMA_NEXT(7) "}"
            ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(a_hi), "r"(b_hi), "r"(idesc), "r"(acc_first), "n"(STEP_A), "n"(STEP_B)
            : "memory");
    }
#undef VC_MMA_NEXT
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

__host__ __device__ constexpr int tc_pad16(int c) { return c < 16 ? 16 : c; }

}  // namespace vc
