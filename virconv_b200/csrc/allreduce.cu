// One-shot gradient all-reduce over NVLink / NVSwitch peer memory (the only collective of the path).
//
// The backbone's gradients are ONE flat fp32 buffer of 1.7 MB: a latency-bound message (NCCL: 84 us at 2 GPUs, 176 us at 8,
// fully exposed after the backward — profiles/bench_r2_graph_bf16_n{2,8}.json).  Every rank keeps a SYMMETRIC buffer (mapped
// into all peers: torch.distributed._symmetric_memory) and a symmetric array of flags.  One kernel per step and rank:
//   1. block b copies its slice of the local gradients into the rank's symmetric buffer (parity half `epoch & 1`);
//   2. block b signals "slice b of epoch e is there" into every peer's flag array (system-scope release) and waits for the same
//      signal from every peer (acquire; time-bounded like the pipeline waits of the conv kernels);
//   3. block b sums slice b over all peers straight out of their buffers (16-byte loads over NVLink) and writes the scaled
//      result back into the local gradient buffer.
// No second barrier: the data buffer is double buffered by epoch parity, and a rank re-writes half p only two epochs later — after
// a barrier that every peer signs only once it has finished READING half p (block b reads exactly the slice block b waits for).
// Replaces `DistributedDataParallel`'s bucketed NCCL all-reduce (tools/train.py:140-141) for this one message.
#include "common.cuh"

namespace vc {
namespace {

constexpr int AR_BLOCKS = 64, AR_THREADS = 512, AR_MAX_WORLD = 16;

struct ArArgs {
    float* bufs[AR_MAX_WORLD];            // every rank's symmetric data buffer [2][n_pad] as mapped HERE
    unsigned* flags[AR_MAX_WORLD];        // every rank's symmetric flag array [world][AR_BLOCKS] as mapped HERE
    float* grads;                         // local flat gradients (in place)
    long long n, n_pad;
    int rank, world;
    unsigned epoch;
    float scale;
    int* err;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(AR_THREADS) allreduce_peer_kernel(const ArArgs a) {
    const int b = blockIdx.x, tid = threadIdx.x;
    const long long n4 = (a.n + 3) / 4;                       // float4 slots (the buffers are padded to a multiple of 4)
    const long long per = (n4 + AR_BLOCKS - 1) / AR_BLOCKS;
    const long long lo = (long long)b * per, hi = min(n4, lo + per);
    const size_t half = (size_t)(a.epoch & 1u) * (size_t)a.n_pad;
    // 1. publish the local slice
    {
        float4* mine = reinterpret_cast<float4*>(a.bufs[a.rank] + half);
        const float* g = a.grads;
        for (long long i = lo + tid; i < hi; i += AR_THREADS) {
            float4 v;
            if (4 * i + 3 < a.n) {
                v = *reinterpret_cast<const float4*>(g + 4 * i);
            } else {
                v.x = 4 * i + 0 < a.n ? g[4 * i + 0] : 0.f;
                v.y = 4 * i + 1 < a.n ? g[4 * i + 1] : 0.f;
                v.z = 4 * i + 2 < a.n ? g[4 * i + 2] : 0.f;
                v.w = 0.f;
            }
            mine[i] = v;
        }
    }
    __threadfence_system();
    __syncthreads();
    // 2. exchange "slice b of this epoch is in place"
    __shared__ int failed;
    if (tid == 0) failed = 0;
    __syncthreads();
    if (tid < a.world) {
        st_release_sys(a.flags[tid] + (size_t)a.rank * AR_BLOCKS + b, a.epoch);
        const unsigned* mine = a.flags[a.rank] + (size_t)tid * AR_BLOCKS + b;
        unsigned long long t0 = 0;
        for (unsigned spin = 0;; ++spin) {
            // (epochs only grow; the signed difference tolerates the 32-bit wrap)
            if ((int)(ld_acquire_sys(mine) - a.epoch) >= 0) break;
            if ((spin & 1023u) == 1023u) {
                unsigned long long t;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
                if (t0 == 0) t0 = t;
                else if (t - t0 > 4000000000ULL) {             // a peer never arrived: give up instead of hanging the GPU
                    if (a.err) atomicCAS(a.err, 0, 0x400 + tid);
                    failed = 1;
                    break;
                }
            }
        }
    }
    __syncthreads();
    if (failed) return;
    // 3. reduce slice b over all peers (own copy first, the others rotated so that the ranks do not all hit one peer at once)
    for (long long i = lo + tid; i < hi; i += AR_THREADS) {
        float4 acc = reinterpret_cast<const float4*>(a.bufs[a.rank] + half)[i];
        for (int k = 1; k < a.world; ++k) {
            const int r = (a.rank + k) % a.world;
            const float4 v = reinterpret_cast<const float4*>(a.bufs[r] + half)[i];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        acc.x *= a.scale; acc.y *= a.scale; acc.z *= a.scale; acc.w *= a.scale;
        if (4 * i + 3 < a.n) {
            *reinterpret_cast<float4*>(a.grads + 4 * i) = acc;
        } else {
            if (4 * i + 0 < a.n) a.grads[4 * i + 0] = acc.x;
            if (4 * i + 1 < a.n) a.grads[4 * i + 1] = acc.y;
            if (4 * i + 2 < a.n) a.grads[4 * i + 2] = acc.z;
        }
    }
}

}  // namespace
}  // namespace vc

extern "C" int vc_allreduce_peer_flag_words(int world) { return world * vc::AR_BLOCKS; }

extern "C" int vc_allreduce_peer_f32(const uint64_t* peer_bufs, const uint64_t* peer_flags, int rank, int world, float* grads,
                                     long long n, long long n_pad, unsigned epoch, float scale, int32_t* err_flag, vc_stream_t stream_) {
    using namespace vc;
    VC_CHECK_ARG(peer_bufs && peer_flags && grads, "null pointer");
    VC_CHECK_ARG(world >= 2 && world <= AR_MAX_WORLD && rank >= 0 && rank < world, "bad rank %d / world %d", rank, world);
    VC_CHECK_ARG(n > 0 && n_pad >= n && (n_pad & 3) == 0 && (reinterpret_cast<uintptr_t>(grads) & 15u) == 0,
                 "gradient buffer must be 16-byte aligned and the symmetric buffer padded to a multiple of 4 floats");
    ArArgs a;
    for (int r = 0; r < world; ++r) {
        a.bufs[r] = reinterpret_cast<float*>(peer_bufs[r]);
        a.flags[r] = reinterpret_cast<unsigned*>(peer_flags[r]);
        VC_CHECK_ARG(a.bufs[r] && a.flags[r], "peer %d is not mapped", r);
    }
    a.grads = grads; a.n = n; a.n_pad = n_pad; a.rank = rank; a.world = world; a.epoch = epoch; a.scale = scale; a.err = err_flag;
    allreduce_peer_kernel<<<AR_BLOCKS, AR_THREADS, 0, (cudaStream_t)stream_>>>(a);
    VC_LAUNCH_CHECK();
    return VC_OK;
}
