// Rulebook (neighbour table) construction for submanifold and regular sparse convolution — sm_100a.
//
// Replaces spconv `ops.get_indice_pairs` behind spconv_backbone.py:89,92-93,113,563-564.
// Canonical form: oracle/rulebook.py.
//
// Data layout in HBM
//   indices   [N, 1+ndim] int32 (b, z, y, x) | (b, u, v)
//   hash      2^h x uint64, entry = (linear_key << 24) | row, EMPTY = ~0   (ws of vc_subm_rulebook)
//   nbr       [K, N_out] int32 — row o of offset-plane k is written by thread o: fully coalesced
//   bitmap    1 bit per OUTPUT cell (batch-major linear index), + per-word exclusive popcount rank
//             (ws of vc_conv_rulebook_*) — gives the sorted, duplicate-free output set with no sort.
// Roofline: all kernels here are HBM/L2-latency bound integer work (no FLOPs); algorithmic bytes per
// input voxel are stated in DESIGN.md §kernels.
#include "common.cuh"

namespace vc {

static constexpr unsigned long long HASH_EMPTY = ~0ULL;
static constexpr int ROW_BITS = 24;
static constexpr int MAXK = 32;  // >= 27 kernel offsets

__device__ __forceinline__ void decode_k(const Geom& g, int k, int* off) {
    for (int d = g.ndim - 1; d >= 0; --d) {
        off[d] = k % g.ksize[d];
        k /= g.ksize[d];
    }
}

// a row takes part in a rulebook only if its batch index and coordinates lie inside the grid (a padding row with b = -1 or
// a wrong `batch_size` must not write outside the bitmap / alias hash keys — ADVICE r1)
__device__ __forceinline__ bool index_ok(const Geom& g, int b, const int* c) {
    bool ok = b >= 0 && b < g.batch;
    for (int d = 0; d < g.ndim; ++d) ok &= (c[d] >= 0) & (c[d] < g.shape[d]);
    return ok;
}

__device__ __forceinline__ void load_index(const int32_t* __restrict__ idx, int row, int ndim, int& b, int* c) {
    if (ndim == 3) {
        int4 v = __ldg(reinterpret_cast<const int4*>(idx) + row);
        b = v.x; c[0] = v.y; c[1] = v.z; c[2] = v.w;
    } else {
        const int32_t* p = idx + (size_t)row * (1 + ndim);
        b = __ldg(p);
        for (int d = 0; d < ndim; ++d) c[d] = __ldg(p + 1 + d);
    }
}

// ------------------------------------------------------------------------------------------------
// hash table
// ------------------------------------------------------------------------------------------------
// (every kernel below that takes a row count `n` also takes `n_dev`: when non-NULL the count is read from device memory
// and `n` is the capacity the grid / buffers were sized for — the plan executor's static mode, CUDA-graph capturable)
__global__ void hash_insert_kernel(const int32_t* __restrict__ idx, int n, const int* __restrict__ n_dev, Geom g,
                                   unsigned long long* table, uint32_t mask) {
    if (n_dev != nullptr) n = min(n, __ldg(n_dev));
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    int b, c[VC_MAX_NDIM];
    load_index(idx, row, g.ndim, b, c);
    if (!index_ok(g, b, c)) return;
    unsigned long long key = (unsigned long long)b;
    for (int d = 0; d < g.ndim; ++d) key = key * (unsigned long long)g.shape[d] + (unsigned long long)c[d];
    unsigned long long packed = (key << ROW_BITS) | (unsigned long long)row;
    uint32_t slot = mix64(key) & mask;
    while (true) {
        unsigned long long cur = table[slot];
        if (cur == HASH_EMPTY) {
            cur = atomicCAS(&table[slot], HASH_EMPTY, packed);
            if (cur == HASH_EMPTY) return;
        }
        if ((cur >> ROW_BITS) == key) {  // duplicate coordinate: lowest row index wins
            atomicMin(&table[slot], packed);
            return;
        }
        slot = (slot + 1) & mask;
    }
}

__device__ __forceinline__ int hash_lookup(const unsigned long long* __restrict__ table, uint32_t mask,
                                           unsigned long long key) {
    uint32_t slot = mix64(key) & mask;
    while (true) {
        unsigned long long cur = __ldg(table + slot);
        if (cur == HASH_EMPTY) return -1;
        if ((cur >> ROW_BITS) == key) return (int)(cur & ((1ULL << ROW_BITS) - 1));
        slot = (slot + 1) & mask;
    }
}

// One thread per (output row, kernel offset): grid (ceil(n/256), K).  Every probe chain is independent, so the
// hash look-ups of all K offsets are in flight together (the per-row loop this replaced was latency bound);
// nbr[k, row] writes stay fully coalesced.  Pair counts: warp ballot -> shared counter -> one atomic per block.
__global__ void __launch_bounds__(256) subm_probe_kernel(const int32_t* __restrict__ idx, int n,
                                                         const int* __restrict__ n_dev, Geom g,
                                                         const unsigned long long* __restrict__ table,
                                                         uint32_t mask, int32_t* __restrict__ nbr,
                                                         int32_t* __restrict__ pair_num) {
    __shared__ int cnt;
    const int pitch = n;                  // table rows are [K][capacity]
    if (n_dev != nullptr) n = min(n, __ldg(n_dev));
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    int res = -1;
    if (row < n) {
        int centre = 0;
        for (int d = 0; d < g.ndim; ++d) centre = centre * g.ksize[d] + g.ksize[d] / 2;
        if (k == centre) {
            res = row;  // identity (spconv Native: out = features @ W[centre])
        } else {
            int b, c[VC_MAX_NDIM];
            load_index(idx, row, g.ndim, b, c);
            int off[VC_MAX_NDIM];
            decode_k(g, k, off);
            bool ok = index_ok(g, b, c);
            unsigned long long key = (unsigned long long)b;
            for (int d = 0; d < g.ndim; ++d) {
                int v = c[d] + (off[d] - g.ksize[d] / 2) * g.dil[d];
                ok &= (v >= 0) & (v < g.shape[d]);
                key = key * (unsigned long long)g.shape[d] + (unsigned long long)v;
            }
            if (ok) res = hash_lookup(table, mask, key);
        }
        nbr[(size_t)k * pitch + row] = res;
    } else if (row < pitch) {
        nbr[(size_t)k * pitch + row] = -1;    // static mode: rows between the count and the capacity have no neighbours
    }
    unsigned m = __ballot_sync(0xffffffffu, res >= 0);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&cnt, __popc(m));
    __syncthreads();
    if (pair_num != nullptr && threadIdx.x == 0 && cnt) atomicAdd(&pair_num[k], cnt);
}

// ------------------------------------------------------------------------------------------------
// regular conv: bitmap over output cells
// ------------------------------------------------------------------------------------------------
static constexpr int SCAN_WORDS = 2048;  // words per scan block (256 threads x 8)

// returns output linear cell (batch major) for (input coords c, offset k), or -1
__device__ __forceinline__ long long out_cell(const Geom& g, int b, const int* c, int k) {
    if (!index_ok(g, b, c)) return -1;
    int off[VC_MAX_NDIM];
    decode_k(g, k, off);
    long long lin = b;
    for (int d = 0; d < g.ndim; ++d) {
        int num = c[d] + g.pad[d] - off[d] * g.dil[d];
        if (num < 0) return -1;
        int o = num / g.stride[d];
        if (o * g.stride[d] != num || o >= g.oshape[d]) return -1;
        lin = lin * g.oshape[d] + o;
    }
    return lin;
}

__global__ void conv_mark_kernel(const int32_t* __restrict__ idx, int n, const int* __restrict__ n_dev, Geom g,
                                 uint32_t* bitmap) {   // grid (rows, K)
    if (n_dev != nullptr) n = min(n, __ldg(n_dev));
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    int b, c[VC_MAX_NDIM];
    load_index(idx, row, g.ndim, b, c);
    long long lin = out_cell(g, b, c, blockIdx.y);
    if (lin < 0) return;
    uint32_t bit = 1u << (lin & 31);
    uint32_t* w = bitmap + (lin >> 5);
    if (!(*w & bit)) atomicOr(w, bit);
}

// local exclusive scan of popcounts inside blocks of SCAN_WORDS words; block totals to block_sum
__global__ void __launch_bounds__(256) scan_local_kernel(const uint32_t* __restrict__ bitmap, long long n_words,
                                                         uint32_t* __restrict__ word_rank,
                                                         uint32_t* __restrict__ block_sum) {
    __shared__ uint32_t warp_tot[8];
    long long base = (long long)blockIdx.x * SCAN_WORDS + threadIdx.x * 8;
    uint32_t cnt[8];
    uint32_t tsum = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        long long w = base + i;
        cnt[i] = (w < n_words) ? __popc(bitmap[w]) : 0;
        tsum += cnt[i];
    }
    uint32_t incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += v;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < (threadIdx.x >> 5); ++w) woff += warp_tot[w];
    uint32_t run = woff + incl - tsum;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        long long w = base + i;
        if (w < n_words) word_rank[w] = run;
        run += cnt[i];
    }
    if (threadIdx.x == 255) block_sum[blockIdx.x] = woff + incl;
}

// single block: exclusive scan of block sums in place; total -> n_out
// (cap > 0: the output buffers hold `cap` rows; a larger total is clamped and reported through *overflow)
__global__ void __launch_bounds__(1024) scan_blocks_kernel(uint32_t* block_sum, int n_blocks, int32_t* n_out, int cap,
                                                           int* overflow) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n_blocks; base += 1024) {
        int i = base + threadIdx.x;
        uint32_t v = (i < n_blocks) ? block_sum[i] : 0;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint32_t woff = 0;
        for (int w = 0; w < (threadIdx.x >> 5); ++w) woff += warp_tot[w];
        uint32_t carry = carry_s;
        if (i < n_blocks) block_sum[i] = carry + woff + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + woff + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int total = (int)carry_s;
        if (cap > 0 && total > cap) {
            if (overflow != nullptr) atomicMax(overflow, total);
            total = cap;
        }
        *n_out = total;
    }
}

__device__ __forceinline__ int cell_rank(const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ word_rank,
                                         const uint32_t* __restrict__ block_sum, long long lin) {
    long long w = lin >> 5;
    uint32_t word = __ldg(bitmap + w);
    return (int)(__ldg(block_sum + (w / SCAN_WORDS)) + __ldg(word_rank + w) + __popc(word & ((1u << (lin & 31)) - 1u)));
}

__global__ void conv_emit_indices_kernel(const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ word_rank,
                                         const uint32_t* __restrict__ block_sum, long long n_words, Geom g,
                                         int32_t* __restrict__ out_idx, int cap) {
    long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint32_t word = bitmap[w];
    if (!word) return;
    int row = (int)(block_sum[w / SCAN_WORDS] + word_rank[w]);
    while (word) {
        int bit = __ffs(word) - 1;
        word &= word - 1;
        if (row >= cap) return;           // capacity overflow (flagged by scan_blocks_kernel)
        long long lin = (w << 5) + bit;
        int c[VC_MAX_NDIM];
        for (int d = g.ndim - 1; d >= 0; --d) {
            c[d] = (int)(lin % g.oshape[d]);
            lin /= g.oshape[d];
        }
        int32_t* o = out_idx + (size_t)row * (1 + g.ndim);
        o[0] = (int32_t)lin;
        for (int d = 0; d < g.ndim; ++d) o[1 + d] = c[d];
        ++row;
    }
}

__global__ void __launch_bounds__(256) conv_tables_kernel(const int32_t* __restrict__ idx, int n, const int* __restrict__ n_dev,
                                                          int n_out, Geom g,
                                                          const uint32_t* __restrict__ bitmap,
                                                          const uint32_t* __restrict__ word_rank,
                                                          const uint32_t* __restrict__ block_sum,
                                                          int32_t* __restrict__ nbr_fwd, int32_t* __restrict__ nbr_bwd,
                                                          int32_t* __restrict__ pair_num) {   // grid (rows, K)
    __shared__ int cnt;
    const int pitch_in = n;               // nbr_bwd rows are [K][input capacity], nbr_fwd rows [K][n_out = output capacity]
    if (n_dev != nullptr) n = min(n, __ldg(n_dev));
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    int orow = -1;
    if (row < n) {
        int b, c[VC_MAX_NDIM];
        load_index(idx, row, g.ndim, b, c);
        long long lin = out_cell(g, b, c, k);
        if (lin >= 0) {
            orow = cell_rank(bitmap, word_rank, block_sum, lin);
            if (orow < n_out) nbr_fwd[(size_t)k * n_out + orow] = row;  // unique writer: (o,k) determines the input cell
            else orow = -1;                                             // beyond the output capacity (overflow flagged)
        }
        nbr_bwd[(size_t)k * pitch_in + row] = orow;
    } else if (row < pitch_in) {
        nbr_bwd[(size_t)k * pitch_in + row] = -1;   // static mode: rows between the count and the capacity
    }
    unsigned m = __ballot_sync(0xffffffffu, orow >= 0);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&cnt, __popc(m));
    __syncthreads();
    if (pair_num != nullptr && threadIdx.x == 0 && cnt) atomicAdd(&pair_num[k], cnt);
}

// one block per offset: order-preserving compaction of nbr[k, :] into spconv-style pairs
__global__ void __launch_bounds__(1024) pairs_kernel(const int32_t* __restrict__ nbr, int K, int n,
                                                     int32_t* __restrict__ pairs, int32_t* __restrict__ pair_num) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    int k = blockIdx.x;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int32_t* src = nbr + (size_t)k * n;
    int32_t* pin = pairs + (size_t)k * n;
    int32_t* pout = pairs + (size_t)(K + k) * n;
    for (int base = 0; base < n; base += 1024) {
        int i = base + threadIdx.x;
        int v = (i < n) ? src[i] : -1;
        int f = v >= 0;
        unsigned m = __ballot_sync(0xffffffffu, f);
        int lane = threadIdx.x & 31;
        int wpre = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) warp_tot[threadIdx.x >> 5] = __popc(m);
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < (threadIdx.x >> 5); ++w) woff += warp_tot[w];
        int carry = carry_s;
        if (f) {
            pin[carry + woff + wpre] = v;
            pout[carry + woff + wpre] = i;
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + woff + wpre + f;
        __syncthreads();
    }
    int total = carry_s;
    for (int i = total + threadIdx.x; i < n; i += 1024) {
        pin[i] = -1;
        pout[i] = -1;
    }
    if (threadIdx.x == 0) pair_num[k] = total;
}

static int make_geom(Geom& g, int ndim, int batch_size, const int32_t* shape, const int32_t* ksize, const int32_t* stride,
                     const int32_t* pad, const int32_t* dil) {
    VC_CHECK_ARG(ndim >= 1 && ndim <= VC_MAX_NDIM, "ndim %d not in [1,%d]", ndim, VC_MAX_NDIM);
    VC_CHECK_ARG(batch_size > 0, "batch size %d", batch_size);
    memset(&g, 0, sizeof(g));
    g.ndim = ndim;
    g.batch = batch_size;
    g.K = 1;
    for (int d = 0; d < ndim; ++d) {
        g.shape[d] = shape[d];
        g.ksize[d] = ksize[d];
        g.stride[d] = stride ? stride[d] : 1;
        g.pad[d] = pad ? pad[d] : 0;
        g.dil[d] = dil ? dil[d] : 1;
        VC_CHECK_ARG(g.shape[d] > 0 && g.ksize[d] > 0 && g.stride[d] > 0 && g.dil[d] > 0 && g.pad[d] >= 0,
                     "bad geometry in dim %d", d);
        g.oshape[d] = (g.shape[d] + 2 * g.pad[d] - g.dil[d] * (g.ksize[d] - 1) - 1) / g.stride[d] + 1;
        VC_CHECK_ARG(g.oshape[d] > 0, "empty output shape in dim %d", d);
        g.K *= g.ksize[d];
    }
    VC_CHECK_ARG(g.K <= MAXK, "kernel volume %d > %d", g.K, MAXK);
    return VC_OK;
}

static uint32_t table_slots(int n) {
    uint32_t s = 1024;
    while (s < 2u * (uint32_t)(n > 0 ? n : 1)) s <<= 1;
    return s;
}

}  // namespace vc

using namespace vc;

extern "C" size_t vc_subm_rulebook_ws_bytes(int n) { return (size_t)table_slots(n) * sizeof(unsigned long long); }

int vc::subm_rulebook_dev(const int32_t* indices, int n, const int* n_dev, int ndim, int batch_size,
                          const int32_t* spatial_shape, const int32_t* ksize, const int32_t* dilation, int32_t* nbr,
                          int32_t* pair_num, void* ws, size_t ws_bytes, cudaStream_t stream) {
    Geom g;
    int rc = make_geom(g, ndim, batch_size, spatial_shape, ksize, nullptr, nullptr, dilation);
    if (rc) return rc;
    VC_CHECK_ARG(n >= 0 && n < (1 << ROW_BITS), "row count %d out of range", n);
    double cells = (double)batch_size;
    for (int d = 0; d < ndim; ++d) cells *= g.shape[d];
    VC_CHECK_ARG(cells < 1099511627776.0, "grid too large for 40-bit keys");
    if (pair_num) VC_CUDA(cudaMemsetAsync(pair_num, 0, sizeof(int32_t) * g.K, stream));
    if (n == 0) return VC_OK;
    VC_CHECK_ARG(indices && nbr && ws, "null pointer");
    uint32_t slots = table_slots(n);
    if (ws_bytes < (size_t)slots * 8) {
        set_error("subm rulebook workspace %zu < %zu", ws_bytes, (size_t)slots * 8);
        return VC_ERR_WORKSPACE;
    }
    unsigned long long* table = (unsigned long long*)ws;
    VC_CUDA(cudaMemsetAsync(table, 0xFF, (size_t)slots * 8, stream));
    hash_insert_kernel<<<cdiv(n, 256), 256, 0, stream>>>(indices, n, n_dev, g, table, slots - 1);
    VC_LAUNCH_CHECK();
    subm_probe_kernel<<<dim3(cdiv(n, 256), g.K), 256, 0, stream>>>(indices, n, n_dev, g, table, slots - 1, nbr, pair_num);
    VC_LAUNCH_CHECK();
    return VC_OK;
}

extern "C" int vc_subm_rulebook(const int32_t* indices, int n, int ndim, int batch_size, const int32_t* spatial_shape,
                                const int32_t* ksize, const int32_t* dilation, int32_t* nbr, int32_t* pair_num,
                                void* ws, size_t ws_bytes, vc_stream_t stream_) {
    return vc::subm_rulebook_dev(indices, n, nullptr, ndim, batch_size, spatial_shape, ksize, dilation, nbr, pair_num, ws, ws_bytes,
                                 (cudaStream_t)stream_);
}

extern "C" int vc_conv_out_shape(int ndim, const int32_t* spatial_shape, const int32_t* ksize, const int32_t* stride,
                                 const int32_t* padding, const int32_t* dilation, int32_t* out_shape) {
    Geom g;
    int rc = make_geom(g, ndim, 1, spatial_shape, ksize, stride, padding, dilation);
    if (rc) return rc;
    for (int d = 0; d < ndim; ++d) out_shape[d] = g.oshape[d];
    return VC_OK;
}

namespace {
struct ConvWs {
    long long n_words;
    int n_blocks;
    uint32_t *bitmap, *word_rank, *block_sum;
    size_t bytes;
};
ConvWs conv_ws_layout(int ndim, int batch_size, const int32_t* oshape, void* ws) {
    ConvWs w;
    long long cells = batch_size;
    for (int d = 0; d < ndim; ++d) cells *= oshape[d];
    w.n_words = (cells + 31) / 32;
    w.n_blocks = (int)((w.n_words + SCAN_WORDS - 1) / SCAN_WORDS);
    size_t a = ((size_t)w.n_words * 4 + 255) / 256 * 256;
    size_t b = ((size_t)w.n_blocks * 4 + 255) / 256 * 256;
    char* p = (char*)ws;
    w.bitmap = (uint32_t*)p;
    w.word_rank = (uint32_t*)(p + a);
    w.block_sum = (uint32_t*)(p + 2 * a);
    w.bytes = 2 * a + b;
    return w;
}
}  // namespace

extern "C" size_t vc_conv_rulebook_ws_bytes(int ndim, int batch_size, const int32_t* out_shape) {
    return conv_ws_layout(ndim, batch_size, out_shape, nullptr).bytes;
}

// cap_out > 0: the caller's output buffers hold cap_out rows (static mode): the count is clamped, *overflow gets the real total
int vc::conv_rulebook_count_dev(const int32_t* indices, int n, const int* n_dev, int ndim, int batch_size,
                                const int32_t* spatial_shape, const int32_t* ksize, const int32_t* stride,
                                const int32_t* padding, const int32_t* dilation, int32_t* n_out_dev, int cap_out, int* overflow,
                                void* ws, size_t ws_bytes, cudaStream_t stream) {
    Geom g;
    int rc = make_geom(g, ndim, batch_size, spatial_shape, ksize, stride, padding, dilation);
    if (rc) return rc;
    VC_CHECK_ARG(n >= 0 && n_out_dev && ws && batch_size > 0, "bad arguments");
    ConvWs w = conv_ws_layout(ndim, batch_size, g.oshape, ws);
    VC_CHECK_ARG(w.n_words < (1LL << 31), "output grid too large");
    if (ws_bytes < w.bytes) {
        set_error("conv rulebook workspace %zu < %zu", ws_bytes, w.bytes);
        return VC_ERR_WORKSPACE;
    }
    VC_CUDA(cudaMemsetAsync(w.bitmap, 0, (size_t)w.n_words * 4, stream));
    if (n > 0) {
        conv_mark_kernel<<<dim3(cdiv(n, 256), g.K), 256, 0, stream>>>(indices, n, n_dev, g, w.bitmap);
        VC_LAUNCH_CHECK();
    }
    scan_local_kernel<<<w.n_blocks, 256, 0, stream>>>(w.bitmap, w.n_words, w.word_rank, w.block_sum);
    VC_LAUNCH_CHECK();
    scan_blocks_kernel<<<1, 1024, 0, stream>>>(w.block_sum, w.n_blocks, n_out_dev, cap_out, overflow);
    VC_LAUNCH_CHECK();
    return VC_OK;
}

extern "C" int vc_conv_rulebook_count(const int32_t* indices, int n, int ndim, int batch_size,
                                      const int32_t* spatial_shape, const int32_t* ksize, const int32_t* stride,
                                      const int32_t* padding, const int32_t* dilation, int32_t* n_out_dev, void* ws,
                                      size_t ws_bytes, vc_stream_t stream_) {
    return vc::conv_rulebook_count_dev(indices, n, nullptr, ndim, batch_size, spatial_shape, ksize, stride, padding, dilation,
                                       n_out_dev, 0, nullptr, ws, ws_bytes, (cudaStream_t)stream_);
}

extern "C" int vc_conv_rulebook_fill(const int32_t* indices, int n, int ndim, int batch_size,
                                     const int32_t* spatial_shape, const int32_t* ksize, const int32_t* stride,
                                     const int32_t* padding, const int32_t* dilation, int n_out, int32_t* out_indices,
                                     int32_t* nbr_fwd, int32_t* nbr_bwd, int32_t* pair_num, void* ws, size_t ws_bytes,
                                     vc_stream_t stream_) {
    return vc::conv_rulebook_fill_phases(indices, n, nullptr, ndim, batch_size, spatial_shape, ksize, stride, padding, dilation, n_out,
                                         out_indices, nbr_fwd, nbr_bwd, pair_num, ws, ws_bytes, (cudaStream_t)stream_, 3);
}

// phases: bit 0 = emit the output indices, bit 1 = build the neighbour tables.  The plan executor runs the emit phase of
// every strided conv first (the next conv's row count depends on it and the host waits for that), the tables later.
// (n_dev != NULL: n / n_out are the capacities of the input / output index sets, the real input count is *n_dev)
int vc::conv_rulebook_fill_phases(const int32_t* indices, int n, const int* n_dev, int ndim, int batch_size, const int32_t* spatial_shape,
                                  const int32_t* ksize, const int32_t* stride, const int32_t* padding,
                                  const int32_t* dilation, int n_out, int32_t* out_indices, int32_t* nbr_fwd,
                                  int32_t* nbr_bwd, int32_t* pair_num, void* ws, size_t ws_bytes, cudaStream_t stream,
                                  int phases) {
    Geom g;
    int rc = make_geom(g, ndim, batch_size, spatial_shape, ksize, stride, padding, dilation);
    if (rc) return rc;
    ConvWs w = conv_ws_layout(ndim, batch_size, g.oshape, ws);
    if (ws_bytes < w.bytes) {
        set_error("conv rulebook workspace %zu < %zu", ws_bytes, w.bytes);
        return VC_ERR_WORKSPACE;
    }
    if ((phases & 1) && n_out > 0) {
        VC_CHECK_ARG(out_indices, "null output pointer");
        conv_emit_indices_kernel<<<cdiv(w.n_words, 256), 256, 0, stream>>>(w.bitmap, w.word_rank, w.block_sum,
                                                                            w.n_words, g, out_indices, n_out);
        VC_LAUNCH_CHECK();
    }
    if (!(phases & 2)) return VC_OK;
    if (pair_num) VC_CUDA(cudaMemsetAsync(pair_num, 0, sizeof(int32_t) * g.K, stream));
    if (n_out > 0) {
        VC_CHECK_ARG(nbr_fwd, "null output pointer");
        VC_CUDA(cudaMemsetAsync(nbr_fwd, 0xFF, (size_t)g.K * n_out * 4, stream));
    }
    if (n > 0) {
        VC_CHECK_ARG(nbr_bwd, "null nbr_bwd");
        conv_tables_kernel<<<dim3(cdiv(n, 256), g.K), 256, 0, stream>>>(indices, n, n_dev, n_out, g, w.bitmap, w.word_rank, w.block_sum,
                                                             nbr_fwd, nbr_bwd, pair_num);
        VC_LAUNCH_CHECK();
    }
    return VC_OK;
}

extern "C" int vc_pairs_from_nbr(const int32_t* nbr, int K, int n, int32_t* pairs, int32_t* pair_num,
                                 vc_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    VC_CHECK_ARG(K > 0 && n >= 0 && pairs && pair_num, "bad arguments");
    if (n == 0) {
        VC_CUDA(cudaMemsetAsync(pair_num, 0, sizeof(int32_t) * K, stream));
        return VC_OK;
    }
    pairs_kernel<<<K, 1024, 0, stream>>>(nbr, K, n, pairs, pair_num);
    VC_LAUNCH_CHECK();
    return VC_OK;
}
