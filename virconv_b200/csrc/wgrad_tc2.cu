// Weight gradient of the sparse convolutions on tensor cores — PERSISTENT variant (round 2), tcgen05 / TMEM, sm_100a only.
//
//   dW[k] (C_in x C_out) = sum over output rows o of  in[nbr[k,o], :]^T (x) dout[o, :]
//
// Same machinery as the persistent forward kernel (conv_tc2.cu): one CTA per SM walks 128-row output tiles (dynamic tile
// counter), a loader warp stages the tile's neighbour-table slice (L2 prefetch several tiles ahead, cp.async into a
// 4-deep shared ring), 8 producer warps gather the rows with 16-byte cp.async into 32/64/128-byte-swizzled tiles, one
// thread issues the MMAs.  What differs:
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

constexpr int W_WARP_LOADER = 4, W_WARP_MMA = 5, W_WARP_PROD0 = 6, W_PROD_WARPS = 8;
constexpr int W_THREADS = 32 * (W_WARP_PROD0 + W_PROD_WARPS);   // 448
constexpr int W_PROD_THREADS = 32 * W_PROD_WARPS;
constexpr int W_MAX_STAGES = 8;
constexpr int W_NTB = 4, W_AHEAD = 2, W_PREF = 4;
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

#define W_WAIT(bar, parity)                                   \
    do {                                                      \
        if (!mbar_wait_t((bar), (parity), a.err)) goto done;  \
    } while (0)

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
    __shared__ __align__(8) uint64_t dout_empty[2];
    __shared__ __align__(8) uint64_t final_bar, meta_bar;
    __shared__ int glist_s[W_NTB][MAXK_TC];
    __shared__ int ng_s[W_NTB], tile_s[W_NTB];
    __shared__ uint32_t tmem_base_s;
    __shared__ unsigned started_s;

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
            mbar_init(&full_bar[s], W_PROD_THREADS);           // every producer thread (cp.async arrive)
            mbar_init(&empty_bar[s], 1);                       // tcgen05.commit
        }
        for (int b = 0; b < W_NTB; ++b) {
            mbar_init(&tbl_full[b], 1);                        // loader
            mbar_init(&tbl_empty[b], W_PROD_WARPS + 1);        // producers + MMA warp
        }
        mbar_init(&dout_empty[0], 1);
        mbar_init(&dout_empty[1], 1);
        mbar_init(&final_bar, 1);
        mbar_init(&meta_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        started_s = 0u;
    }
    pdl_wait();
    pdl_launch_dependents();
    const int n = a.n_dev != nullptr ? min(__ldg(a.n_dev), a.n_host) : a.n_host;
    const int n_tiles = (n + TCM - 1) / TCM;
    int* counter = a.tile_counter != nullptr ? a.tile_counter + blockIdx.y : nullptr;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == W_WARP_LOADER) {
        // ------------------------------------------------------------ tile scheduler + neighbour-table loader
        const bool vec_ok = (reinterpret_cast<uintptr_t>(a.nbr) & 15u) == 0 && (a.pitch & 3) == 0;
        constexpr int PQ = W_AHEAD + W_PREF + 1;
        int my_tile[PQ];
        int claimed = 0, issued = 0;
        bool stop = false;
        auto claim = [&]() {
            int tile;
            if (counter != nullptr) {
                tile = lane == 0 ? atomicAdd(counter, 1) : 0;
                tile = __shfl_sync(0xffffffffu, tile, 0);
            } else {
                tile = blockIdx.x + claimed * gridDim.x;
            }
            if (tile >= n_tiles) {
                tile = -1;
                stop = true;
            } else {
                const int base = tile * TCM;
                if (vec_ok && (long long)base + TCM <= a.pitch) {
                    if (lane < k_count)
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.nbr + (size_t)(k_begin + lane) * a.pitch + base),
                                     "r"(TCM * 4)
                                     : "memory");
                } else {
                    for (int i = lane; i < k_count * 5; i += 32) {
                        const int k = i / 5, seg = i % 5;
                        const long long row = (long long)base + seg * 32;
                        if (row < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.nbr + (size_t)(k_begin + k) * a.pitch + row));
                    }
                }
            }
            my_tile[claimed % PQ] = tile;
            ++claimed;
        };
        auto issue = [&](int it) -> bool {
            const int tb = it % W_NTB;
            if (it >= W_NTB && !mbar_wait_t(&tbl_empty[tb], (uint32_t)(((it / W_NTB) - 1) & 1), a.err)) return false;
            const int tile = my_tile[it % PQ];
            if (tile >= 0) {
                int* dst = nbr_s + (size_t)tb * k_count * TCM;
                const int base = tile * TCM;
                if (vec_ok && (long long)base + TCM <= a.pitch) {
                    const uint32_t d0 = smem_u32(dst) + lane * 16;
                    const int32_t* s0 = a.nbr + (size_t)k_begin * a.pitch + base + lane * 4;
                    for (int k = 0; k < k_count; ++k) cp_async16_s(d0 + k * (TCM * 4), s0 + (size_t)k * a.pitch, true);
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
            return true;
        };
        for (int it = 0;; ++it) {
            while (!stop && claimed <= it + W_AHEAD + W_PREF - 1) claim();
            while (issued < claimed && issued <= it + W_AHEAD - 1) {
                if (!issue(issued)) goto done;
                ++issued;
            }
            const int tb = it % W_NTB;
            const int tile = my_tile[it % PQ];
            if (issued - it - 1 >= 1) cp_async_wait<1>(); else cp_async_wait<0>();
            __syncwarp();
            if (tile < 0) {
                if (lane == 0) {
                    tile_s[tb] = -1;
                    mbar_arrive(&tbl_full[tb]);
                }
                break;
            }
            int* dst = nbr_s + (size_t)tb * k_count * TCM;
            const int base = tile * TCM;
            const bool partial = base + TCM > n;
            unsigned gm = 0u;
            for (int k = 0; k < k_count; ++k) {
                int4 v = reinterpret_cast<const int4*>(dst + k * TCM)[lane];
                if (partial) {
                    const int r0 = base + lane * 4;
                    bool ch = false;
                    if (r0 + 0 >= n && v.x != -1) { v.x = -1; ch = true; }
                    if (r0 + 1 >= n && v.y != -1) { v.y = -1; ch = true; }
                    if (r0 + 2 >= n && v.z != -1) { v.z = -1; ch = true; }
                    if (r0 + 3 >= n && v.w != -1) { v.w = -1; ch = true; }
                    if (ch) reinterpret_cast<int4*>(dst + k * TCM)[lane] = v;
                }
                const bool any = (v.x >= 0) | (v.y >= 0) | (v.z >= 0) | (v.w >= 0);
                if (__any_sync(0xffffffffu, any)) gm |= 1u << (k / C::GW);
            }
            if (lane == 0) {
                int c = 0;
                for (int g = 0; g < g_count; ++g)
                    if (gm >> g & 1u) glist_s[tb][c++] = g;
                ng_s[tb] = c;
                tile_s[tb] = tile;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&tbl_full[tb]);
        }
    } else if (warp >= W_WARP_PROD0) {
        // ------------------------------------------------------------ gather producers
        const int pw = warp - W_WARP_PROD0, ptid = tid - W_WARP_PROD0 * 32;
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
        int s = 0;
        uint32_t ph = 0;
        bool wrapped = false;
        for (int it = 0;; ++it) {
            const int tb = it % W_NTB, db = it & 1;
            W_WAIT(&tbl_full[tb], (uint32_t)((it / W_NTB) & 1));
            const int tile = tile_s[tb];
            if (tile < 0) break;
            const int ng = ng_s[tb];
            const int* tbl = nbr_s + (size_t)tb * k_count * TCM;
            const int base = tile * TCM;
            // the tile of dout (second operand of every group of this tile): contiguous rows, swizzled image; its completion
            // is covered by the first stage's `full` barrier (cp.async.mbarrier.arrive tracks ALL earlier copies of a thread)
            if (it >= 2) W_WAIT(&dout_empty[db], (uint32_t)(((it >> 1) - 1) & 1));
            {
                const uint32_t b_s = smem_u32(dout_s) + (uint32_t)db * C::B_BYTES;
                for (int q = ptid; q < TCM * C::CPB; q += W_PROD_THREADS) {
                    const int r = q / C::CPB, c = q % C::CPB;
                    const bool v = base + r < n && c * 8 < a.out_c;
                    cp_async16_s(b_s + swz_off<C::RB>(r, c), v ? a.dout + (size_t)(base + r) * a.out_c + c * 8 : a.dout, v);
                }
            }
            for (int t = 0; t < ng; ++t) {
                const int g = glist_s[tb][t];
                int src[C::GW][NIT];
#pragma unroll
                for (int j = 0; j < C::GW; ++j) {
                    const int kk = g * C::GW + j;
#pragma unroll
                    for (int i = 0; i < NIT; ++i) src[j][i] = kk < k_count ? tbl[kk * TCM + rows[i]] : -1;
                }
                if (wrapped) W_WAIT(&empty_bar[s], ph);
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
                            cp_async16_s(a_s + dst_off[i][cg], vc ? srow + cg * CW * 8 : a.in, vc);
                        }
                    }
                }
                cp_async_arrive_noinc(&full_bar[s]);
                if (++s == S) {
                    s = 0;
                    if (wrapped) ph ^= 1u;
                    wrapped = true;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&tbl_empty[tb]);
        }
    } else if (warp == W_WARP_MMA) {
        // ------------------------------------------------------------ MMA issuer
        constexpr uint32_t IDESC = umma_idesc(TCM, CO) | (1u << 15) | (1u << 16);     // both operands MN-major
        int s = 0;
        uint32_t ph = 0;
        unsigned started = 0u;
        for (int it = 0;; ++it) {
            const int tb = it % W_NTB, db = it & 1;
            W_WAIT(&tbl_full[tb], (uint32_t)((it / W_NTB) & 1));
            if (tile_s[tb] < 0) break;
            const int ng = ng_s[tb];
            int gl[MAXK_TC / 2];
            for (int t = 0; t < ng; ++t) gl[t] = glist_s[tb][t];
            __syncwarp();
            if (lane == 0) mbar_arrive(&tbl_empty[tb]);
            for (int t = 0; t < ng; ++t) {
                W_WAIT(&full_bar[s], ph);
                fence_async_smem();
                tc_fence_after();
                if (lane == 0) {
                    const int g = gl[t];
                    const uint32_t a0 = smem_u32(ring) + (uint32_t)s * C::STAGE;
                    const uint32_t b0 = smem_u32(dout_s) + (uint32_t)db * C::B_BYTES;
                    const bool first = !(started >> g & 1u);
#pragma unroll
                    for (int j = 0; j < 8; ++j)      // 16 rows (two 8-row groups) per MMA
                        umma_f16(tmem_base + (uint32_t)(g * CO), umma_desc_mn<C::RA>(a0 + j * 16 * C::RA, C::A_BYTES),
                                 umma_desc_mn<C::RB>(b0 + j * 16 * C::RB, 0), IDESC, (first && j == 0) ? 0u : 1u);
                    started |= 1u << g;
                    umma_commit(&empty_bar[s]);
                    if (t == ng - 1) umma_commit(&dout_empty[db]);
                }
                __syncwarp();
                if (++s == S) {
                    s = 0;
                    ph ^= 1u;
                }
            }
            if (ng == 0 && lane == 0) umma_commit(&dout_empty[db]);
            __syncwarp();
        }
        if (lane == 0) {
            started_s = started;
            umma_commit(&final_bar);           // arrives when every MMA of this CTA has completed
            mbar_arrive(&meta_bar);            // publishes started_s (release)
        }
    } else {
        // ------------------------------------------------------------ epilogue (once, at the end): TMEM -> vector reductions
        W_WAIT(&meta_bar, 0u);
        W_WAIT(&final_bar, 0u);
        tc_fence_after();
        const unsigned st = started_s;
        const int row = warp * 32 + lane;
        const int j = row / CI, ci = row % CI;
        const int oc = a.out_c, ic = a.in_c;
        for (int g = 0; g < g_count; ++g) {
            if (!(st >> g & 1u)) continue;               // warp-uniform: tcgen05.ld is warp-collective
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
