// Weight gradient of the sparse convolutions on tensor cores — PERSISTENT variant (round 2), tcgen05 / TMEM, sm_100a only.
//
//   dW[k] (C_in x C_out) = sum over output rows o of  in[nbr[k,o], :]^T (x) dout[o, :]
//
// (The plan executor's default weight-gradient kernel is the half-tile-stage variant of this one, wgrad_tc3.cu, launched on
//  half the SMs beside the main stream; this file serves the per-operator C ABI — vc_conv_wgrad_tc — and the A/B switch.)
// Same machinery as the persistent forward kernel (conv_tc2.cu): one CTA per SM walks 128-row output tiles (dynamic tile
// counter), a loader warp stages the tile's neighbour-table slice (cp.async into 2-4 shared buffers, one to three tiles
// ahead), two groups of 8 producer warps gather the rows with 16-byte cp.async into 32/64/128-byte-swizzled tiles, one
// converged warp issues the MMAs.  What differs:
//   * the gathered tile is used MN-MAJOR (the reduction runs over the 128 rows): the row-major swizzled tile the forward
//     kernel builds is exactly the canonical Major-MN layout (profiles/exp_mnmajor.cu), and G = 128 / C_in kernel offsets
//     are stacked along the UMMA M dimension through the descriptor's leading-byte offset (one ring stage = G tiles);
//   * the second operand is the tile of dout itself (contiguous rows, loaded once per tile, double buffered);
//   * the accumulators of ALL offset groups stay resident in TMEM (<= 512 columns) across all tiles of the CTA; at the
//     end each CTA adds its [K, C_in, C_out] partial into ONE fp32 scratch image with 16-byte vector reductions.  Round 1
//     wrote 128 full partials per layer and re-read them in a second kernel (218 MB + 218 MB per step for 1.7 MB of
//     gradients); here the reduction traffic is (#CTAs x gradient size) of L2 atomics and the only extra launch per STEP
//     is the transposing finalize of all layers (wgrad_finalize_kernel).
//   * C = 8 layers run with channels padded to 16 (zero-filled by the gather), so stage 1 no longer needs the fp32 kernel.
// Replaces spconv `ops.indice_conv_backward` (filter gradient) behind spconv_backbone.py:89,92-93,113,563-564.
// Algorithmic bytes per launch: (N_in*C_in + N_out*C_out)*2 + P*8 + K*C_in*C_out*4;  FLOPs 2*P*C_in*C_out.
#include "tc_common.cuh"

namespace vc {
namespace {

#ifndef VC_W_GROUPS
#define VC_W_GROUPS 2            // producer groups of 8 warps; consecutive ring stages go to consecutive groups
#endif
#ifndef VC_P_SKIP
#define VC_P_SKIP 1              // 1: missing neighbours cost a shared-memory zero store, not a (zero-fill) cp.async
#endif
constexpr int W_WARP_LOADER = 4, W_WARP_MMA = 5, W_WARP_PROD0 = 6, W_PROD_WARPS = 8, W_GROUPS = VC_W_GROUPS;
constexpr int W_THREADS = 32 * (W_WARP_PROD0 + W_GROUPS * W_PROD_WARPS);   // 704 with two groups
constexpr int W_PROD_THREADS = 32 * W_PROD_WARPS * W_GROUPS;
constexpr int W_MAX_STAGES = 8;
constexpr int W_NTB = 4;
constexpr int W_ROWS_PER_PROD = TCM / W_PROD_WARPS;             // 16
constexpr int W_SMEM_BUDGET = 227 * 1024 - 4096;

template <int CI, int CO>
struct WCfg2 {
    static constexpr int RA = CI * 2, RB = CO * 2;           // row bytes of a gathered tile / of the dout tile (= swizzle spans)
    static constexpr int CPA = CI / 8, CPB = CO / 8;         // 16-byte chunks per row
    static constexpr int GW = 128 / CI;                      // kernel offsets stacked along M (one accumulator group)
    static constexpr int A_BYTES = TCM * RA;
    static constexpr int STAGE = GW * A_BYTES;               // 32 KB for every C_in
    static constexpr int B_BYTES = TCM * RB;
};

struct WArgs {
    const __nv_bfloat16* in;    // [n_in, in_c] gathered rows
    int in_c;                   // real channels (8 or CI)
    const __nv_bfloat16* dout;  // [n_out, out_c]
    int out_c;                  // real channels (8 or CO)
    const int32_t* nbr;         // [K][pitch]
    long long pitch;
    float* scratch;             // [K][in_c][out_c] fp32, zero before the launch; every CTA adds its partial
    const int* n_dev;
    int n_host;
    int* tile_counter;          // [passes] or NULL
    int K, S, groups_per_pass, tmem_cols;
    int* err;
};

#define W_WAIT(bar, parity, code)                                     \
    do {                                                              \
        if (!mbar_wait_t((bar), (parity), a.err, (code))) goto done;  \
    } while (0)

__device__ __forceinline__ void w_sts_zero16(uint32_t saddr) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(saddr), "r"(0u) : "memory");
}

template <int ROWB>
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo) {
    // Major-MN, swizzled (cute/atom/mma_traits_sm100.hpp): LBO = distance between swizzle atoms along M/N, SBO = between
    // 8-row groups along K
    constexpr uint64_t LT = ROWB == 128 ? 2 : ROWB == 64 ? 4 : 6;
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)(((uint32_t)(8 * ROWB) >> 4) & 0x3FFFu) << 32) | (1ULL << 46) | (LT << 61);
}

template <int CI, int CO>
__global__ void __launch_bounds__(W_THREADS, 1) tc_wgrad_persist_kernel(const WArgs a) {
    using C = WCfg2<CI, CO>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int S = a.S, K = a.K;
    unsigned char* ring = smem_raw;                                              // [S][GW x A tile]
    unsigned char* dout_s = smem_raw + (size_t)S * C::STAGE;                     // [2][B tile]
    int* nbr_s = reinterpret_cast<int*>(dout_s + 2 * C::B_BYTES);                // [W_NTB][kcount][128]
    __shared__ __align__(8) uint64_t full_bar[W_MAX_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[W_MAX_STAGES];
    __shared__ __align__(8) uint64_t tbl_full[W_NTB], tbl_empty[W_NTB];
    __shared__ __align__(8) uint64_t dout_full[2], dout_empty[2];
    __shared__ __align__(8) uint64_t final_bar, meta_bar;
    __shared__ int tile_s[W_NTB];
    __shared__ uint32_t tmem_base_s;
    __shared__ int started_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // this pass's slice of the kernel offsets (passes only when the accumulators of all groups exceed 512 TMEM columns)
    const int n_groups_total = (K + C::GW - 1) / C::GW;
    const int g_begin = blockIdx.y * a.groups_per_pass;
    const int g_count = min(a.groups_per_pass, n_groups_total - g_begin);
    const int k_begin = g_begin * C::GW;
    const int k_count = min(g_count * C::GW, K - k_begin);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)a.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == W_WARP_MMA * 32) {
        for (int s = 0; s < S; ++s) {
            // every producer thread of the stage's group (cp.async.mbarrier.arrive.noinc: fires when the thread's copies have
            // landed) + one release arrive per warp for its zero stores
            mbar_init(&full_bar[s], 32 * W_PROD_WARPS + (VC_P_SKIP ? W_PROD_WARPS : 0));
            mbar_init(&empty_bar[s], 1);                       // tcgen05.commit
        }
        for (int b = 0; b < W_NTB; ++b) {
            mbar_init(&tbl_full[b], 1);                        // loader
            mbar_init(&tbl_empty[b], W_GROUPS * W_PROD_WARPS + 1);   // producers + MMA warp
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&dout_full[b], W_PROD_THREADS);          // every producer thread (cp.async arrive)
            mbar_init(&dout_empty[b], 1);                      // tcgen05.commit
        }
        mbar_init(&final_bar, 1);
        mbar_init(&meta_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        started_s = 0;
    }
    pdl_wait();
    pdl_launch_dependents();
    const int n = a.n_dev != nullptr ? min(__ldg(a.n_dev), a.n_host) : a.n_host;
    // a pipeline wait that timed out in an EARLIER launch left the (sticky) error flag set: do nothing, so that whatever went
    // wrong costs one 2-second timeout, not one per launch
    const bool dead = a.err != nullptr && *reinterpret_cast<volatile int*>(a.err) != 0;
    const int n_tiles = dead ? 0 : (n + TCM - 1) / TCM;
    int* counter = a.tile_counter != nullptr ? a.tile_counter + blockIdx.y : nullptr;
    int ntb = W_NTB;            // table buffers in use == tiles a CTA holds claimed at once (see conv_tc2.cu)
    if (counter != nullptr) {
        int d = n_tiles / (3 * (int)gridDim.x);
        d = d < 1 ? 1 : (d > W_NTB - 1 ? W_NTB - 1 : d);
        ntb = d + 1;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == W_WARP_LOADER) {
        // ------------------------------------------------------------ tile scheduler + neighbour-table loader
        const bool vec_ok = (reinterpret_cast<uintptr_t>(a.nbr) & 15u) == 0 && (a.pitch & 3) == 0;
        const int lag = ntb - 1;
        int tq0 = -1, tq1 = -1, tq2 = -1, tq3 = -1;
        bool stop = false;
        int claimed = 0;
        for (int it = 0;; ++it) {
            const int tb = it % ntb;
            int tile = -1;
            if (!stop) {
                if (counter != nullptr) {
                    tile = lane == 0 ? atomicAdd(counter, 1) : 0;
                    tile = __shfl_sync(0xffffffffu, tile, 0);
                } else {
                    tile = blockIdx.x + claimed * gridDim.x;
                }
                ++claimed;
                if (tile >= n_tiles) {
                    tile = -1;
                    stop = true;
                }
            }
            switch (it & 3) {
                case 0: tq0 = tile; break;
                case 1: tq1 = tile; break;
                case 2: tq2 = tile; break;
                default: tq3 = tile; break;
            }
            if (it >= ntb) W_WAIT(&tbl_empty[tb], (uint32_t)(((it / ntb) - 1) & 1), 0x201);
            if (tile >= 0) {
                int* dst = nbr_s + (size_t)tb * k_count * TCM;
                const int base = tile * TCM;
                if (vec_ok && (long long)base + TCM <= a.pitch) {
                    uint32_t d = smem_u32(dst) + lane * 16;
                    const int32_t* sp = a.nbr + (size_t)k_begin * a.pitch + base + lane * 4;
                    for (int k = 0; k < k_count; ++k, d += TCM * 4, sp += a.pitch) cp_async16_s(d, sp, true);
                } else {
                    for (int k = 0; k < k_count; ++k) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int row = base + q * 32 + lane;
                            if (row < n) {
                                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_u32(dst + k * TCM + q * 32 + lane)),
                                             "l"(a.nbr + (size_t)(k_begin + k) * a.pitch + row));
                            } else {
                                dst[k * TCM + q * 32 + lane] = -1;
                            }
                        }
                    }
                }
            }
            cp_async_commit();
            if (it >= lag) {
                const int pi = it - lag;
                if (lag == 3) cp_async_wait<3>(); else if (lag == 2) cp_async_wait<2>(); else cp_async_wait<1>();
                __syncwarp();
                const int ptb = pi % ntb;
                const int ptile = (pi & 3) == 0 ? tq0 : (pi & 3) == 1 ? tq1 : (pi & 3) == 2 ? tq2 : tq3;
                if (ptile >= 0 && ptile * TCM + TCM > n) {
                    int* dst = nbr_s + (size_t)ptb * k_count * TCM;
                    const int r0 = ptile * TCM + lane * 4;
                    for (int k = 0; k < k_count; ++k) {
                        int4 v = reinterpret_cast<const int4*>(dst + k * TCM)[lane];
                        if (r0 + 0 >= n) v.x = -1;
                        if (r0 + 1 >= n) v.y = -1;
                        if (r0 + 2 >= n) v.z = -1;
                        if (r0 + 3 >= n) v.w = -1;
                        reinterpret_cast<int4*>(dst + k * TCM)[lane] = v;
                    }
                }
                if (lane == 0) tile_s[ptb] = ptile;
                __syncwarp();
                if (lane == 0) mbar_arrive(&tbl_full[ptb]);
                if (ptile < 0) break;
            }
        }
    } else if (warp >= W_WARP_PROD0) {
        // ------------------------------------------------------------ gather producers
        const int grp = (warp - W_WARP_PROD0) / W_PROD_WARPS, pw = (warp - W_WARP_PROD0) % W_PROD_WARPS;
        const int ptid = tid - W_WARP_PROD0 * 32;
        constexpr int CW = C::CPA < 4 ? C::CPA : 4;
        constexpr int RPI = 32 / CW;
        constexpr int NIT = W_ROWS_PER_PROD / RPI;
        constexpr int NCG = C::CPA / CW;
        const int c_sub = lane % CW, r_sub = lane / CW;
        int rows[NIT];
        uint32_t dst_off[NIT][NCG];
        bool ch_ok[NCG];
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
            rows[i] = pw * W_ROWS_PER_PROD + i * RPI + r_sub;
#pragma unroll
            for (int cg = 0; cg < NCG; ++cg) dst_off[i][cg] = swz_off<C::RA>(rows[i], cg * CW + c_sub);
        }
#pragma unroll
        for (int cg = 0; cg < NCG; ++cg) ch_ok[cg] = (cg * CW + c_sub) * 8 < a.in_c;
        const uint32_t ring_s = smem_u32(ring);
        int s = 0, wr = 0, turn = 0;
        for (int it = 0;; ++it) {
            const int tb = it % ntb, db = it & 1;
            W_WAIT(&tbl_full[tb], (uint32_t)((it / ntb) & 1), 0x211);
            const int tile = tile_s[tb];
            if (tile < 0) break;
            const int* tbl = nbr_s + (size_t)tb * k_count * TCM;
            const int base = tile * TCM;
            // the tile of dout (second operand of every group of this tile): contiguous rows, swizzled image, loaded by all
            // producer threads together
            if (it >= 2) W_WAIT(&dout_empty[db], (uint32_t)(((it >> 1) - 1) & 1), 0x212);
            {
                const uint32_t b_s = smem_u32(dout_s) + (uint32_t)db * C::B_BYTES;
                for (int q = ptid; q < TCM * C::CPB; q += W_PROD_THREADS) {
                    const int r = q / C::CPB, c = q % C::CPB;
                    const bool v = base + r < n && c * 8 < a.out_c;
                    cp_async16_s(b_s + swz_off<C::RB>(r, c), v ? a.dout + (size_t)(base + r) * a.out_c + c * 8 : a.dout, v);
                }
                cp_async_arrive_noinc(&dout_full[db]);
            }
            for (int g = 0; g < g_count; ++g) {
                if (turn == grp) {
                    int src[C::GW][NIT];
#pragma unroll
                    for (int j = 0; j < C::GW; ++j) {
                        const int kk = g * C::GW + j;
#pragma unroll
                        for (int i = 0; i < NIT; ++i) src[j][i] = kk < k_count ? tbl[kk * TCM + rows[i]] : -1;
                    }
                    if (wr > 0) W_WAIT(&empty_bar[s], (uint32_t)((wr - 1) & 1), 0x213);
                    const uint32_t st_s = ring_s + (uint32_t)s * C::STAGE;
#pragma unroll
                    for (int j = 0; j < C::GW; ++j) {
                        const uint32_t a_s = st_s + (uint32_t)j * C::A_BYTES;
#pragma unroll
                        for (int i = 0; i < NIT; ++i) {
                            const bool v = src[j][i] >= 0;
                            const __nv_bfloat16* srow = a.in + (size_t)(v ? src[j][i] : 0) * a.in_c + c_sub * 8;
#pragma unroll
                            for (int cg = 0; cg < NCG; ++cg) {
                                const bool vc = v && ch_ok[cg];
#if VC_P_SKIP
                                if (vc) cp_async16_s(a_s + dst_off[i][cg], srow + cg * CW * 8, true);
                                else w_sts_zero16(a_s + dst_off[i][cg]);
#else
                                cp_async16_s(a_s + dst_off[i][cg], vc ? srow + cg * CW * 8 : a.in, vc);
#endif
                            }
                        }
                    }
                    cp_async_arrive_noinc(&full_bar[s]);
#if VC_P_SKIP
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full_bar[s]);     // release: publishes the warp's zero stores
#endif
                }
                if (++turn == W_GROUPS) turn = 0;
                if (++s == S) {
                    s = 0;
                    ++wr;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&tbl_empty[tb]);
        }
    } else if (warp == W_WARP_MMA) {
        // ------------------------------------------------------------ MMA issuer
        constexpr uint32_t IDESC = umma_idesc(TCM, CO) | (1u << 15) | (1u << 16);     // both operands MN-major
        // Major-MN descriptors: low word = start >> 4 | (LBO >> 4) << 16, high word constant (tc_common.cuh umma_series)
        constexpr uint32_t A_HI = umma_desc_hi<C::RA>(), B_HI = umma_desc_hi<C::RB>();
        constexpr uint32_t A_LBO = (uint32_t)(C::A_BYTES >> 4) << 16;
        const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]), doutf0 = smem_u32(&dout_full[0]),
                       doute0 = smem_u32(&dout_empty[0]);
        const uint32_t ring_a = smem_u32(ring), dout_a = smem_u32(dout_s);
        int s = 0;
        uint32_t ph = 0;
        int n_done = 0;
        for (int it = 0;; ++it) {
            const int tb = it % ntb, db = it & 1;
            W_WAIT(&tbl_full[tb], (uint32_t)((it / ntb) & 1), 0x221);
            if (tile_s[tb] < 0) break;
            __syncwarp();
            if (lane == 0) mbar_arrive(&tbl_empty[tb]);
            if (!mbar_spin(doutf0 + 8u * db, (uint32_t)((it >> 1) & 1), 4096u) &&
                !mbar_wait_t_addr(doutf0 + 8u * db, (uint32_t)((it >> 1) & 1), a.err, 0x222))
                goto done;
            const uint32_t b_lo = (dout_a + (uint32_t)db * C::B_BYTES) >> 4;
            for (int g = 0; g < g_count; ++g) {
                if (!mbar_spin(full0 + 8u * s, ph, 4096u) && !mbar_wait_t_addr(full0 + 8u * s, ph, a.err, 0x223)) goto done;
                fence_async_smem();     // generic-proxy (cp.async, st.shared) writes -> visible to the tensor core's async proxy
                tc_fence_after();
                // all lanes converged; the 8 MMAs of the stage (16 rows = two 8-row groups each) in one asm block
                umma_series<8, C::RA, C::RB>(tmem_base + (uint32_t)(g * CO), ((ring_a + (uint32_t)s * C::STAGE) >> 4) | A_LBO, b_lo, A_HI, B_HI,
                                             IDESC, it == 0 ? 0u : 1u);
                umma_commit_elect_addr(empty0 + 8u * s);
                if (g == g_count - 1) umma_commit_elect_addr(doute0 + 8u * db);
                if (++s == S) {
                    s = 0;
                    ph ^= 1u;
                }
            }
            ++n_done;
        }
        if (n_done > 0) umma_commit_elect(&final_bar);     // arrives when every MMA of this CTA has completed
        if (lane == 0) {
            started_s = n_done;
            if (n_done == 0) mbar_arrive(&final_bar);       // (a CTA that was handed no tile has nothing in flight)
            mbar_arrive(&meta_bar);                         // publishes started_s (release)
        }
    } else {
        // ------------------------------------------------------------ epilogue (once, at the end): TMEM -> vector reductions
        W_WAIT(&meta_bar, 0u, 0x231);
        W_WAIT(&final_bar, 0u, 0x232);
        tc_fence_after();
        const int row = warp * 32 + lane;
        const int j = row / CI, ci = row % CI;
        const int oc = a.out_c, ic = a.in_c;
        if (started_s > 0) {
            for (int g = 0; g < g_count; ++g) {
                const int kk = g * C::GW + j;
                const bool mine = kk < k_count && ci < ic;
#pragma unroll
                for (int c0 = 0; c0 < CO; c0 += 16) {
                    float v[16];
                    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(g * CO + c0), v);
                    if (mine && c0 < oc) {
                        float* dst = a.scratch + ((size_t)(k_begin + kk) * ic + ci) * oc + c0;
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            if (c0 + i < oc)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(v[i]), "f"(v[i + 1]),
                                             "f"(v[i + 2]), "f"(v[i + 3])
                                             : "memory");
                    }
                }
            }
        }
    }
done:
    cp_async_wait<0>();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols));
    }
}

int g_wgrad2_ctas = 0;      // 0: one CTA per SM

int wg_num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

template <int CI, int CO>
int launch_wgrad2(const WArgs& a0, int n_cap, cudaStream_t stream) {
    using C = WCfg2<CI, CO>;
    WArgs a = a0;
    const int n_groups = (a.K + C::GW - 1) / C::GW;
    const int max_groups = 512 / CO;
    int passes = (n_groups + max_groups - 1) / max_groups;
    int gpp = (n_groups + passes - 1) / passes;
    int cols = 32;
    while (cols < gpp * CO) cols <<= 1;
    const int kcount = gpp * C::GW < a.K ? gpp * C::GW : a.K;
    a.groups_per_pass = gpp;
    a.tmem_cols = cols;
    const size_t fixed = 2 * (size_t)C::B_BYTES + (size_t)W_NTB * kcount * TCM * 4;
    int S = (int)((W_SMEM_BUDGET - fixed) / C::STAGE);
    if (S > W_MAX_STAGES) S = W_MAX_STAGES;
    if (S < 2) {
        set_error("tensor-core wgrad: no room for the operand ring (K=%d, %d->%d)", a.K, CI, CO);
        return VC_ERR_UNSUPPORTED;
    }
    a.S = S;
    const size_t smem = (size_t)S * C::STAGE + fixed;
    auto kern = tc_wgrad_persist_kernel<CI, CO>;
    static bool attr_done = false;
    if (!attr_done) {
        VC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, W_SMEM_BUDGET));
        attr_done = true;
    }
    const int tiles = cdiv(n_cap, TCM);
    int cap = g_wgrad2_ctas > 0 ? g_wgrad2_ctas : wg_num_sms();
    const int grid = tiles < cap ? (tiles < 1 ? 1 : tiles) : cap;
    VC_LAUNCH_CHAIN(kern, dim3(grid, passes), dim3(W_THREADS), smem, stream, a);
    return VC_OK;
}

}  // namespace

// scratch [K][cin][cout] (one per layer) -> parameter layout dW [cout][K][cin], every layer of a step in ONE launch
__global__ void __launch_bounds__(256) wgrad_finalize_kernel(WgradFinTable t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t.total) return;
    int lo = 0, hi = t.n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (t.e[mid].first <= i) lo = mid; else hi = mid - 1;
    }
    const WgradFinEntry& e = t.e[lo];
    const int j = i - e.first;                       // index in the parameter layout: coalesced writes
    const int ci = j % e.cin, k = (j / e.cin) % e.K, co = j / (e.cin * e.K);
    e.dw[j] = __ldg(e.scratch + ((size_t)k * e.cin + ci) * e.cout + co);
}

int wgrad_finalize(WgradFinTable& t, cudaStream_t stream) {
    if (t.n == 0) return VC_OK;
    int total = 0;
    for (int i = 0; i < t.n; ++i) {
        t.e[i].first = total;
        total += t.e[i].K * t.e[i].cin * t.e[i].cout;
    }
    t.total = total;
    wgrad_finalize_kernel<<<cdiv(total, 256), 256, 0, stream>>>(t);
    VC_LAUNCH_CHECK();
    return VC_OK;
}

// scratch: [K][cin][cout] fp32, ZERO on entry (the kernel accumulates); tile_counter: `wgrad2_passes(cin, cout, K)` ints, zero
int tc2_wgrad(int cin, int cout, const void* in_bf16, const void* dout_bf16, const int32_t* nbr, long long pitch, float* scratch,
              int n_rows, const int* n_dev, int K, int* err, cudaStream_t stream, int* tile_counter) {
    if (n_rows == 0) return VC_OK;
    if (!tc2_ch_ok(cin) || !tc2_ch_ok(cout) || K < 1 || K > MAXK_TC) {
        set_error("tensor-core wgrad: unsupported shape (%d -> %d channels, K=%d)", cin, cout, K);
        return VC_ERR_UNSUPPORTED;
    }
    WArgs a;
    a.in = (const __nv_bfloat16*)in_bf16; a.in_c = cin; a.dout = (const __nv_bfloat16*)dout_bf16; a.out_c = cout; a.nbr = nbr;
    a.pitch = pitch; a.scratch = scratch; a.n_dev = n_dev; a.n_host = n_rows; a.tile_counter = tile_counter; a.K = K; a.S = 0;
    a.groups_per_pass = 0; a.tmem_cols = 0; a.err = err;
    const int ci = tc_pad16(cin), co = tc_pad16(cout);
#define VC_W_CASE(A, B) \
    if (ci == A && co == B) return launch_wgrad2<A, B>(a, n_rows, stream);
    VC_W_CASE(16, 16) VC_W_CASE(16, 32) VC_W_CASE(16, 64)
    VC_W_CASE(32, 16) VC_W_CASE(32, 32) VC_W_CASE(32, 64)
    VC_W_CASE(64, 16) VC_W_CASE(64, 32) VC_W_CASE(64, 64)
#undef VC_W_CASE
    return VC_ERR_UNSUPPORTED;
}

int wgrad2_passes(int cin, int cout, int K) {
    const int ci = tc_pad16(cin), co = tc_pad16(cout);
    const int n_groups = (K + 128 / ci - 1) / (128 / ci);
    const int max_groups = 512 / co;
    return (n_groups + max_groups - 1) / max_groups;
}

}  // namespace vc

extern "C" int vc_conv_wgrad_tc2_config(int max_ctas) {
    VC_CHECK_ARG(max_ctas >= 0 && max_ctas <= 1024, "wgrad CTA cap out of range (%d)", max_ctas);
    vc::g_wgrad2_ctas = max_ctas;
    return VC_OK;
}
