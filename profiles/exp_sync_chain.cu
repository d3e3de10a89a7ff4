// How fast can the mbarrier skeleton of a warp-specialised pipeline turn over, with NO data movement and NO MMA?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/exp_sync profiles/exp_sync_chain.cu && /tmp/exp_sync
// conv_tc2.cu without copies and without MMAs (build variants VC_DBG_NO_COPY + VC_DBG_NO_MMA) still needs ~480 ns per ring
// stage (profiles/microbench_conv2_r2_h_*.txt) — this isolates the barrier traffic:
//   8 producer warps:  wait empty[s] -> (work) -> arrive on full[s]      arrive style: every thread | lane 0 of every warp
//   1 consumer warp :  wait full[s]  -> release empty[s]                  release style: plain arrive | tcgen05.commit
//   E extra warps spinning on a barrier that completes once per 16 stages (the epilogue warps of the real kernel)
// ring of S stages, 1 CTA per SM, N stages per CTA; reports ns per stage.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar), done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}
// one lane polls, the rest of the warp waits at a warp barrier
__device__ __forceinline__ void mbar_wait_lane0(uint64_t* bar, uint32_t parity) {
    if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
    __syncwarp();
}
__device__ __forceinline__ void commit_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(smem_u32(bar))
        : "memory");
}

constexpr int S = 8, PW = 8;

// ARR: 0 every producer thread arrives, 1 lane 0 of every warp, 2 cp.async.mbarrier.arrive.noinc by every thread (no copies)
// REL: 0 plain arrive by consumer lane 0, 1 tcgen05.commit
// POLL: 0 all lanes poll, 1 lane 0 polls + __syncwarp
template <int ARR, int REL, int POLL>
__global__ void __launch_bounds__(32 * (1 + PW + 4), 1) chain_kernel(int n_stages, int extra_warps, int* sink) {
    __shared__ __align__(8) uint64_t full_bar[S], empty_bar[S], slow_bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(32u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full_bar[s], ARR == 1 ? PW : 32 * PW);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&slow_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) {
        // consumer
        int s = 0;
        uint32_t ph = 0;
        for (int i = 0; i < n_stages; ++i) {
            if (POLL) mbar_wait_lane0(&full_bar[s], ph); else mbar_wait(&full_bar[s], ph);
            if (REL == 1) commit_elect(&empty_bar[s]);
            else if (lane == 0) mbar_arrive(&empty_bar[s]);
            if ((i & 15) == 15 && lane == 0) mbar_arrive(&slow_bar);
            __syncwarp();
            if (++s == S) { s = 0; ph ^= 1u; }
        }
    } else if (warp <= PW) {
        int s = 0, wr = 0;
        for (int i = 0; i < n_stages; ++i) {
            if (wr > 0) {
                if (POLL) mbar_wait_lane0(&empty_bar[s], (uint32_t)((wr - 1) & 1)); else mbar_wait(&empty_bar[s], (uint32_t)((wr - 1) & 1));
            }
            if (ARR == 0) mbar_arrive(&full_bar[s]);
            else if (ARR == 2) asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full_bar[s])) : "memory");
            else {
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_bar[s]);
            }
            if (++s == S) { s = 0; ++wr; }
        }
    } else if (warp - PW - 1 < extra_warps) {
        // "epilogue" warps: wait for an event that happens once per 16 stages
        for (int i = 0; i < n_stages / 16; ++i) {
            if (POLL) mbar_wait_lane0(&slow_bar, (uint32_t)(i & 1)); else mbar_wait(&slow_bar, (uint32_t)(i & 1));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(32u));
    }
    if (n_stages < 0) sink[0] = 1;
}

template <int ARR, int REL, int POLL>
void run(int extra, int sms, int* sink) {
    const int n = 4096;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    chain_kernel<ARR, REL, POLL><<<sms, 32 * (1 + PW + 4)>>>(n, extra, sink);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    chain_kernel<ARR, REL, POLL><<<sms, 32 * (1 + PW + 4)>>>(n, extra, sink);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    const char* an[3] = {"256 thread arrives", "8 warp arrives   ", "256 noinc arrives "};
    printf("arrive: %s  release: %-14s  poll: %-9s  extra spinning warps %d : %7.1f ns / stage\n", an[ARR], REL ? "tcgen05.commit" : "plain arrive",
           POLL ? "lane 0" : "all lanes", extra, ms * 1e6 / n);
}

int main() {
    int dev = 0, sms = 148;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int* sink;
    CK(cudaMalloc(&sink, 4));
    for (int extra : {0, 4}) {
        run<0, 0, 0>(extra, sms, sink);
        run<1, 0, 0>(extra, sms, sink);
        run<2, 0, 0>(extra, sms, sink);
        run<0, 1, 0>(extra, sms, sink);
        run<1, 1, 0>(extra, sms, sink);
        run<2, 1, 0>(extra, sms, sink);
        run<0, 0, 1>(extra, sms, sink);
        run<1, 0, 1>(extra, sms, sink);
        run<2, 1, 1>(extra, sms, sink);
        run<1, 1, 1>(extra, sms, sink);
    }
    return 0;
}
