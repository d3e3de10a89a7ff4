// Row-gather global -> shared memory on sm_100a: which path moves scattered 32 / 64 / 128-byte rows fastest?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/exp_gp profiles/exp_gather_paths.cu && /tmp/exp_gp
// The persistent tcgen05 conv kernel (conv_tc2.cu) spends its time waiting for gathered rows: the trace
// (profiles/trace_tc2_r2_b.txt) shows one 16 KB ring stage landing every ~550 ns whatever the number of producer warps (8 / 16 /
// 24), whether missing neighbours are zero-filled by cp.async or by st.shared, i.e. ~30 SM cycles per cp.async.16 WARP
// INSTRUCTION.  This experiment isolates the copy paths (no MMA, no barriers besides what the path itself needs):
//   mode 0  cp.async.cg 16 B  (LDGSTS.128), D groups in flight per warp
//   mode 1  ld.global.nc.v4 -> registers -> st.shared.v4 (LDG.128 + STS.128), D loads in flight per thread
//   mode 2  like 0, but one lane copies a whole row with ROWB/16 consecutive cp.async (row-per-lane instead of chunk-per-lane)
//   mode 3  like 1, row-per-lane
// 1 CTA per SM, NW warps, every warp walks its share of a table of row indices (shared memory, like the kernel's neighbour table);
// rows come from a [N, ROWB] array (N = 140500, L2 resident), indices either sequential, or "local" (sorted random subset, what a
// rulebook column looks like), with a given fraction of missing neighbours (-1: slot skipped by predicate in every mode).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int TBL = 8192;         // row indices per CTA (shared), walked cyclically
constexpr int D = 8;              // copies in flight per warp / thread

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int ROWB, int MODE>
__global__ void __launch_bounds__(1024, 1) gather_kernel(const unsigned char* __restrict__ src, const int* __restrict__ idx_g, int iters,
                                                         float* sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    int* tbl = reinterpret_cast<int*>(smem);                         // [TBL]
    unsigned char* ring = smem + TBL * 4;                            // per warp: D x 512 B (modes 0/1) or D x 32 rows (modes 2/3)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
    for (int i = tid; i < TBL; i += blockDim.x) tbl[i] = idx_g[(size_t)blockIdx.x * TBL + i];
    __syncthreads();
    constexpr int CPR = ROWB / 16;
    float acc = 0.f;
    if (MODE == 0 || MODE == 1) {
        constexpr int CW = CPR < 4 ? CPR : 4;         // lanes per row (full sectors), as in conv_tc2.cu
        constexpr int RPI = 32 / CW;                  // rows per warp instruction
        constexpr int NCG = CPR / CW;                 // instructions per row group
        const int c_sub = lane % CW, r_sub = lane / CW;
        unsigned char* my = ring + (size_t)warp * D * 512;
        int pos = warp * RPI;                         // table cursor of this warp
        if (MODE == 0) {
            for (int it = 0; it < iters; ++it) {
                const int r = tbl[(pos + r_sub) & (TBL - 1)];
                pos += nw * RPI;
#pragma unroll
                for (int cg = 0; cg < NCG; ++cg) {
                    const int slot = (it * NCG + cg) % D;
                    if (r >= 0)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(my + slot * 512 + lane * 16)),
                                     "l"(src + (size_t)r * ROWB + (cg * CW + c_sub) * 16));
                    asm volatile("cp.async.commit_group;");
                    asm volatile("cp.async.wait_group %0;" ::"n"(D - 1));
                }
            }
            asm volatile("cp.async.wait_group 0;");
        } else {
            uint4 v[D];
#pragma unroll
            for (int d = 0; d < D; ++d) v[d] = make_uint4(0, 0, 0, 0);
            for (int it = 0; it < iters; it += D / NCG) {
                // D loads in flight, then D stores
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const int g = d / NCG, cg = d % NCG;
                    const int r = tbl[(pos + g * nw * RPI + r_sub) & (TBL - 1)];
                    if (r >= 0)
                        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[d].x), "=r"(v[d].y), "=r"(v[d].z), "=r"(v[d].w)
                                     : "l"(src + (size_t)r * ROWB + (cg * CW + c_sub) * 16));
                }
                pos += (D / NCG) * nw * RPI;
#pragma unroll
                for (int d = 0; d < D; ++d) *reinterpret_cast<uint4*>(my + d * 512 + lane * 16) = v[d];
            }
        }
        acc = reinterpret_cast<float*>(my)[lane];
    } else {
        // row per lane: 32 rows per warp "instruction group"
        unsigned char* my = ring + (size_t)warp * 2 * 32 * ROWB;   // 2 slots of 32 rows
        int pos = warp * 32;
        if (MODE == 2) {
            for (int it = 0; it < iters; ++it) {
                const int r = tbl[(pos + lane) & (TBL - 1)];
                pos += nw * 32;
                const int slot = it & 1;
                if (r >= 0) {
#pragma unroll
                    for (int c = 0; c < CPR; ++c)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(my + (slot * 32 + lane) * ROWB + ((c + lane) % CPR) * 16)),
                                     "l"(src + (size_t)r * ROWB + ((c + lane) % CPR) * 16));
                }
                asm volatile("cp.async.commit_group;");
                asm volatile("cp.async.wait_group 1;");
            }
            asm volatile("cp.async.wait_group 0;");
        } else {
            for (int it = 0; it < iters; ++it) {
                const int r = tbl[(pos + lane) & (TBL - 1)];
                pos += nw * 32;
                const int slot = it & 1;
                uint4 v[CPR];
#pragma unroll
                for (int c = 0; c < CPR; ++c) v[c] = make_uint4(0, 0, 0, 0);
                if (r >= 0) {
#pragma unroll
                    for (int c = 0; c < CPR; ++c)
                        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[c].x), "=r"(v[c].y), "=r"(v[c].z), "=r"(v[c].w)
                                     : "l"(src + (size_t)r * ROWB + c * 16));
                }
#pragma unroll
                for (int c = 0; c < CPR; ++c)
                    *reinterpret_cast<uint4*>(my + (slot * 32 + lane) * ROWB + ((c + lane) % CPR) * 16) = v[c];
            }
        }
        acc = reinterpret_cast<float*>(my)[lane];
    }
    if (acc == 123.456f) sink[0] = acc;
}

template <int ROWB, int MODE>
void run(const unsigned char* src, const int* idx, float* sink, int nw, int sms, const char* pat, double fill) {
    constexpr int CPR = ROWB / 16;
    constexpr int CW = CPR < 4 ? CPR : 4;
    const int rows_per_it = (MODE < 2) ? 32 / CW : 32;         // rows one warp covers per iteration
    const long long slots_per_cta = 1 << 18;                  // row slots per CTA
    int iters = (int)(slots_per_cta / ((long long)nw * rows_per_it));
    iters = iters / D * D;
    const size_t smem = TBL * 4 + (size_t)nw * ((MODE < 2) ? D * 512 : 2 * 32 * ROWB);
    auto k = gather_kernel<ROWB, MODE>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (smem > 200 * 1024) { printf("skip (smem)\n"); return; }
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    k<<<sms, nw * 32, smem>>>(src, idx, iters, sink);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(a));
        k<<<sms, nw * 32, smem>>>(src, idx, iters, sink);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        best = std::min(best, ms);
    }
    const double slots = (double)iters * nw * rows_per_it;           // per CTA
    const double ns = best * 1e6;
    const char* names[4] = {"cp.async chunk/lane", "LDG+STS  chunk/lane", "cp.async row/lane  ", "LDG+STS  row/lane  "};
    printf("row %3d B  %-7s fill %3.0f%%  %2d warps  %s : %8.1f us  %6.3f row-slots/ns/SM  %6.1f B/ns/SM (slots)  %6.1f B/ns/SM (valid)\n", ROWB,
           pat, fill * 100, nw, names[MODE], best * 1e3, slots / ns, slots * ROWB / ns, slots * ROWB * fill / ns);
}

int main() {
    int dev = 0, sms = 148;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int N = 140500;
    unsigned char* src;
    CK(cudaMalloc(&src, (size_t)N * 128));
    CK(cudaMemset(src, 1, (size_t)N * 128));
    float* sink;
    CK(cudaMalloc(&sink, 4));
    int* idx;
    CK(cudaMalloc(&idx, (size_t)sms * TBL * 4));
    std::vector<int> h((size_t)sms * TBL);
    for (int pat = 0; pat < 3; ++pat) {
        const double fill = pat == 0 ? 1.0 : 0.43;
        const char* pname = pat == 0 ? "seq" : (pat == 1 ? "seq" : "local");
        srand(1234);
        for (int c = 0; c < sms; ++c) {
            // CTA c covers a window of the array (tiles of consecutive output rows -> neighbours are near by)
            const int base = (int)((long long)c * (N - 9000) / sms);
            for (int i = 0; i < TBL; ++i) {
                int r;
                if (pat < 2) r = base + i % 8192;
                else r = base + (i % 128) + ((i / 128) * 131) % 8000 + (rand() % 3) * 40;       // per "offset" a shifted, slightly ragged run
                if (r >= N) r = N - 1;
                if ((rand() % 1000) >= fill * 1000) r = -1;
                h[(size_t)c * TBL + i] = r;
            }
        }
        CK(cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
        printf("---- pattern %s, fill %.0f%%\n", pname, fill * 100);
        for (int nw : {8, 16, 24}) {
            run<32, 0>(src, idx, sink, nw, sms, pname, fill);
            run<32, 1>(src, idx, sink, nw, sms, pname, fill);
            run<32, 2>(src, idx, sink, nw, sms, pname, fill);
            run<32, 3>(src, idx, sink, nw, sms, pname, fill);
            run<64, 0>(src, idx, sink, nw, sms, pname, fill);
            run<64, 1>(src, idx, sink, nw, sms, pname, fill);
            run<64, 3>(src, idx, sink, nw, sms, pname, fill);
            run<128, 0>(src, idx, sink, nw, sms, pname, fill);
            run<128, 1>(src, idx, sink, nw, sms, pname, fill);
            run<128, 3>(src, idx, sink, nw, sms, pname, fill);
        }
    }
    return 0;
}
