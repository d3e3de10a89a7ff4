// BatchNorm1d(eps, momentum) + ReLU over the active rows of a sparse tensor — sm_100a.
//
// Replaces the `norm_fn(out_channels)` / `nn.ReLU()` members of every spconv.SparseSequential
// (spconv_backbone.py:101-105, :160 eps=1e-3 momentum=0.01, :561-567).  HBM-bound elementwise work.
//
// Statistics never get their own pass: the conv epilogue adds each tile's channel sums (sum x, sum x^2) into a
// [2, C] float64 accumulator with atomics, and the apply kernel below derives mean / invstd / scale / shift from
// it per thread (block 0 also publishes them for backward and updates the running statistics).  Backward is two
// kernels: a reduce that accumulates (sum g, sum g*xhat) the same way, and an apply that derives its per-channel
// coefficients from those sums.  The float64 accumulators make the fp32-rounded results independent of the
// atomic arrival order for all practical purposes (each addend is an fp32 tile partial, < 2^12 of them).
#include <cooperative_groups.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace vc {

__device__ __forceinline__ uint2 pack_bf16x4(float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    return o;
}

// y = relu(x*scale + shift); train: statistics from `sums`, eval: from the running buffers.
// stats_out [4, c] = scale, shift, mean, invstd (block 0).  grid-stride over float4 with stride % (c/4) == 0.
__global__ void __launch_bounds__(256) bn_apply_relu_kernel(
    const float4* __restrict__ x, const double* __restrict__ sums, int n_rows, const int* __restrict__ n_dev, int c,
    const float* __restrict__ gamma, const float* __restrict__ beta, float* running_mean, float* running_var,
    long long* num_batches_tracked, float momentum, float eps, int training, float4* __restrict__ y, int y_ld4,
    uint2* __restrict__ y_bf16, int yb_ld4, float* __restrict__ stats_out, int relu, int tail_zero) {
    // n_dev (static mode): the row count lives in device memory, n_rows is the capacity of the buffers; with tail_zero the
    // rows [count, capacity) of the outputs are written as zeros (published tensors keep a defined tail).
    // y_ld4 / yb_ld4: row pitch of the outputs in float4 / uint2 units (>= c/4: the output may be a column slice of a
    // wider matrix — the concat of an NRConv block is written in place, spconv_backbone.py:227).
    pdl_wait();                 // the conv's output and channel sums
    pdl_launch_dependents();
    const int cap_rows = n_rows;
    if (n_dev != nullptr) n_rows = min(n_rows, __ldg(n_dev));
    const int c4 = c / 4;
    const size_t n4 = (size_t)n_rows * c4;
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const int cg = (int)(i0 % c4);
    __shared__ float sc_s[128], sh_s[128];
    if (threadIdx.x < c) {   // one thread per channel derives scale / shift (float64 once per block, not per thread)
        const int ch = threadIdx.x;
        float mean, invstd;
        double var = 0.0;
        if (training) {
            const double n = (double)(n_rows > 0 ? n_rows : 1);
            const double m = sums[ch] / n;
            var = sums[c + ch] / n - m * m;
            if (var < 0.0) var = 0.0;
            mean = (float)m;
            invstd = (float)(1.0 / sqrt(var + (double)eps));
        } else {
            mean = running_mean[ch];
            invstd = 1.f / sqrtf(running_var[ch] + eps);
        }
        const float scv = gamma[ch] * invstd;
        const float shv = beta[ch] - mean * scv;
        sc_s[ch] = scv;
        sh_s[ch] = shv;
        if (blockIdx.x == 0) {
            stats_out[ch] = scv;
            stats_out[c + ch] = shv;
            stats_out[2 * c + ch] = mean;
            stats_out[3 * c + ch] = invstd;
            if (training) {
                const double n = (double)n_rows;
                const double unbiased = n_rows > 1 ? var * n / (n - 1.0) : var;
                running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * mean;
                running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
            }
        }
    }
    if (training && blockIdx.x == 0 && threadIdx.x == 0 && num_batches_tracked != nullptr) *num_batches_tracked += 1;
    __syncthreads();
    float sc[4], sh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        sc[j] = sc_s[cg * 4 + j];
        sh[j] = sh_s[cg * 4 + j];
    }
    const bool dense = y_ld4 == c4 && yb_ld4 == c4;
    for (size_t i = i0; i < n4; i += stride) {
        float4 v = x[i];
        v.x = fmaf(v.x, sc[0], sh[0]); v.y = fmaf(v.y, sc[1], sh[1]);
        v.z = fmaf(v.z, sc[2], sh[2]); v.w = fmaf(v.w, sc[3], sh[3]);
        if (relu) {
            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        }
        const size_t r = i / c4;
        const size_t oy = dense ? i : r * y_ld4 + cg, ob = dense ? i : r * yb_ld4 + cg;
        if (y != nullptr) y[oy] = v;
        if (y_bf16 != nullptr) y_bf16[ob] = pack_bf16x4(v);   // shadow copy: the next conv's tensor-core operand
    }
    if (tail_zero && cap_rows > n_rows) {
        const size_t t4 = (size_t)cap_rows * c4;
        for (size_t i = n4 + i0; i < t4; i += stride) {
            const size_t r = i / c4;
            if (y != nullptr) y[r * y_ld4 + cg] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (y_bf16 != nullptr) y_bf16[r * yb_ld4 + cg] = make_uint2(0u, 0u);
        }
    }
}

// backward pass 1: sum g and sum g*xhat over rows, g = dy*(y>0); block partial -> float64 atomics on bsums [2, c]
// (the ReLU mask is recomputed from x with the forward's own scale / shift and fmaf: bit-identical to testing y > 0,
//  so the activation itself is not read again; dy may be a column slice of a wider matrix: row pitch dy_ld4 float4)
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float4* __restrict__ dy, int dy_ld4, const float4* __restrict__ x,
                                                            const float* __restrict__ stats, int n, const int* __restrict__ n_dev,
                                                            int c, double* __restrict__ bsums) {
    extern __shared__ float shf[];  // [rowlanes][2][c]
    pdl_wait();
    pdl_launch_dependents();
    if (n_dev != nullptr) n = min(n, __ldg(n_dev));
    const int c4 = c / 4;
    const int cg = threadIdx.x % c4, rl = threadIdx.x / c4, rowlanes = blockDim.x / c4;
    const float4 sc = reinterpret_cast<const float4*>(stats)[cg];
    const float4 sh = reinterpret_cast<const float4*>(stats + c)[cg];
    const float4 m = reinterpret_cast<const float4*>(stats + 2 * c)[cg];
    const float4 is = reinterpret_cast<const float4*>(stats + 3 * c)[cg];
    float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    for (int row = blockIdx.x * rowlanes + rl; row < n; row += gridDim.x * rowlanes) {
        size_t i = (size_t)row * c4 + cg;
        float4 g = dy[(size_t)row * dy_ld4 + cg], xx = x[i];
        g.x = fmaf(xx.x, sc.x, sh.x) > 0.f ? g.x : 0.f; g.y = fmaf(xx.y, sc.y, sh.y) > 0.f ? g.y : 0.f;
        g.z = fmaf(xx.z, sc.z, sh.z) > 0.f ? g.z : 0.f; g.w = fmaf(xx.w, sc.w, sh.w) > 0.f ? g.w : 0.f;
        s[0] += g.x; s[1] += g.y; s[2] += g.z; s[3] += g.w;
        q[0] = fmaf(g.x, (xx.x - m.x) * is.x, q[0]); q[1] = fmaf(g.y, (xx.y - m.y) * is.y, q[1]);
        q[2] = fmaf(g.z, (xx.z - m.z) * is.z, q[2]); q[3] = fmaf(g.w, (xx.w - m.w) * is.w, q[3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        shf[(rl * 2 + 0) * c + cg * 4 + j] = s[j];
        shf[(rl * 2 + 1) * c + cg * 4 + j] = q[j];
    }
    __syncthreads();
    if (threadIdx.x < 2 * c) {
        int which = threadIdx.x / c, ch = threadIdx.x % c;
        float v = 0.f;
        for (int l = 0; l < rowlanes; ++l) v += shf[(l * 2 + which) * c + ch];
        atomicAdd(bsums + which * c + ch, (double)v);
    }
}

// backward pass 2:  train: dx = gi*(g - S/n - xhat*Q/n);  eval: dx = gi*g;   gi = gamma*invstd
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float4* __restrict__ dy, int dy_ld4, const float4* __restrict__ x,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ stats,
                                                           const double* __restrict__ bsums, int n_rows,
                                                           const int* __restrict__ n_dev, int c, int training,
                                                           float4* __restrict__ dx, uint2* __restrict__ dx_bf16,
                                                           float* __restrict__ dgamma, float* __restrict__ dbeta, int tail_zero) {
    pdl_wait();                 // bsums from the reduce kernel
    pdl_launch_dependents();
    const int cap_rows = n_rows;
    if (n_dev != nullptr) n_rows = min(n_rows, __ldg(n_dev));
    const int c4 = c / 4;
    const size_t n4 = (size_t)n_rows * c4;
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const int cg = (int)(i0 % c4);
    __shared__ float a_s[128], b_s[128], d_s[128];
    if (threadIdx.x < c) {
        const int ch = threadIdx.x;
        const double S = bsums[ch], Q = bsums[c + ch];
        const float mean = stats[2 * c + ch], invstd = stats[3 * c + ch];
        const float gi = gamma[ch] * invstd;
        if (training) {
            const double nn = (double)(n_rows > 0 ? n_rows : 1);
            const float k1 = (float)(Q / nn) * invstd;
            a_s[ch] = gi;
            b_s[ch] = -gi * k1;
            d_s[ch] = gi * (k1 * mean - (float)(S / nn));
        } else {
            a_s[ch] = gi;
            b_s[ch] = 0.f;
            d_s[ch] = 0.f;
        }
        if (blockIdx.x == 0) {
            dbeta[ch] = (float)S;
            dgamma[ch] = (float)Q;
        }
    }
    __syncthreads();
    float a[4], b[4], d[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        a[j] = a_s[cg * 4 + j];
        b[j] = b_s[cg * 4 + j];
        d[j] = d_s[cg * 4 + j];
    }
    const float4 sc = reinterpret_cast<const float4*>(stats)[cg];
    const float4 sh = reinterpret_cast<const float4*>(stats + c)[cg];
    for (size_t i = i0; i < n4; i += stride) {
        float4 g = dy[dy_ld4 == c4 ? i : (i / c4) * dy_ld4 + cg], xx = x[i], r;
        g.x = fmaf(xx.x, sc.x, sh.x) > 0.f ? g.x : 0.f; g.y = fmaf(xx.y, sc.y, sh.y) > 0.f ? g.y : 0.f;
        g.z = fmaf(xx.z, sc.z, sh.z) > 0.f ? g.z : 0.f; g.w = fmaf(xx.w, sc.w, sh.w) > 0.f ? g.w : 0.f;
        r.x = fmaf(a[0], g.x, fmaf(b[0], xx.x, d[0])); r.y = fmaf(a[1], g.y, fmaf(b[1], xx.y, d[1]));
        r.z = fmaf(a[2], g.z, fmaf(b[2], xx.z, d[2])); r.w = fmaf(a[3], g.w, fmaf(b[3], xx.w, d[3]));
        if (dx != nullptr) dx[i] = r;
        if (dx_bf16 != nullptr) dx_bf16[i] = pack_bf16x4(r);
    }
    if (tail_zero && cap_rows > n_rows) {
        // static mode: rows between the count and the capacity are operands of the tensor-core wgrad tiles (multiplied by
        // zero-filled gathers): they must be finite
        const size_t t4 = (size_t)cap_rows * c4;
        for (size_t i = n4 + i0; i < t4; i += stride) {
            if (dx != nullptr) dx[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (dx_bf16 != nullptr) dx_bf16[i] = make_uint2(0u, 0u);
        }
    }
}

// backward, both passes in ONE cooperative launch (grid-wide barrier between the reduction and the apply): 20 launches fewer per
// step than reduce + apply, and the second read of dy / x comes out of L2.  Same arithmetic as the two kernels above.
__global__ void __launch_bounds__(256) bn_bwd_fused_kernel(const float4* __restrict__ dy, int dy_ld4, const float4* __restrict__ x,
                                                           const float* __restrict__ gamma, const float* __restrict__ stats,
                                                           double* __restrict__ bsums, int n_rows, const int* __restrict__ n_dev, int c,
                                                           int training, float4* __restrict__ dx, uint2* __restrict__ dx_bf16,
                                                           float* __restrict__ dgamma, float* __restrict__ dbeta, int tail_zero) {
    extern __shared__ float shf[];  // [rowlanes][2][c]
    const int cap_rows = n_rows;
    if (n_dev != nullptr) n_rows = min(n_rows, __ldg(n_dev));
    const int c4 = c / 4;
    {
        const int cg = threadIdx.x % c4, rl = threadIdx.x / c4, rowlanes = blockDim.x / c4;
        const float4 sc = reinterpret_cast<const float4*>(stats)[cg];
        const float4 sh = reinterpret_cast<const float4*>(stats + c)[cg];
        const float4 m = reinterpret_cast<const float4*>(stats + 2 * c)[cg];
        const float4 is = reinterpret_cast<const float4*>(stats + 3 * c)[cg];
        float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
        for (int row = blockIdx.x * rowlanes + rl; row < n_rows; row += gridDim.x * rowlanes) {
            size_t i = (size_t)row * c4 + cg;
            float4 g = dy[(size_t)row * dy_ld4 + cg], xx = x[i];
            g.x = fmaf(xx.x, sc.x, sh.x) > 0.f ? g.x : 0.f; g.y = fmaf(xx.y, sc.y, sh.y) > 0.f ? g.y : 0.f;
            g.z = fmaf(xx.z, sc.z, sh.z) > 0.f ? g.z : 0.f; g.w = fmaf(xx.w, sc.w, sh.w) > 0.f ? g.w : 0.f;
            s[0] += g.x; s[1] += g.y; s[2] += g.z; s[3] += g.w;
            q[0] = fmaf(g.x, (xx.x - m.x) * is.x, q[0]); q[1] = fmaf(g.y, (xx.y - m.y) * is.y, q[1]);
            q[2] = fmaf(g.z, (xx.z - m.z) * is.z, q[2]); q[3] = fmaf(g.w, (xx.w - m.w) * is.w, q[3]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            shf[(rl * 2 + 0) * c + cg * 4 + j] = s[j];
            shf[(rl * 2 + 1) * c + cg * 4 + j] = q[j];
        }
        __syncthreads();
        if (threadIdx.x < 2 * c) {
            int which = threadIdx.x / c, ch = threadIdx.x % c;
            float v = 0.f;
            for (int l = 0; l < rowlanes; ++l) v += shf[(l * 2 + which) * c + ch];
            atomicAdd(bsums + which * c + ch, (double)v);
        }
    }
    __threadfence();
    cooperative_groups::this_grid().sync();
    const size_t n4 = (size_t)n_rows * c4;
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const int cg = (int)(i0 % c4);
    float* a_s = shf;                 // (the reduction's staging is dead: reuse it)
    float* b_s = shf + 128;
    float* d_s = shf + 256;
    __syncthreads();
    if (threadIdx.x < c) {
        const int ch = threadIdx.x;
        const double S = __ldcg(bsums + ch), Q = __ldcg(bsums + c + ch);
        const float mean = stats[2 * c + ch], invstd = stats[3 * c + ch];
        const float gi = gamma[ch] * invstd;
        if (training) {
            const double nn = (double)(n_rows > 0 ? n_rows : 1);
            const float k1 = (float)(Q / nn) * invstd;
            a_s[ch] = gi;
            b_s[ch] = -gi * k1;
            d_s[ch] = gi * (k1 * mean - (float)(S / nn));
        } else {
            a_s[ch] = gi;
            b_s[ch] = 0.f;
            d_s[ch] = 0.f;
        }
        if (blockIdx.x == 0) {
            dbeta[ch] = (float)S;
            dgamma[ch] = (float)Q;
        }
    }
    __syncthreads();
    float a[4], b[4], d[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        a[j] = a_s[cg * 4 + j];
        b[j] = b_s[cg * 4 + j];
        d[j] = d_s[cg * 4 + j];
    }
    const float4 sc = reinterpret_cast<const float4*>(stats)[cg];
    const float4 sh = reinterpret_cast<const float4*>(stats + c)[cg];
    for (size_t i = i0; i < n4; i += stride) {
        float4 g = dy[dy_ld4 == c4 ? i : (i / c4) * dy_ld4 + cg], xx = x[i], r;
        g.x = fmaf(xx.x, sc.x, sh.x) > 0.f ? g.x : 0.f; g.y = fmaf(xx.y, sc.y, sh.y) > 0.f ? g.y : 0.f;
        g.z = fmaf(xx.z, sc.z, sh.z) > 0.f ? g.z : 0.f; g.w = fmaf(xx.w, sc.w, sh.w) > 0.f ? g.w : 0.f;
        r.x = fmaf(a[0], g.x, fmaf(b[0], xx.x, d[0])); r.y = fmaf(a[1], g.y, fmaf(b[1], xx.y, d[1]));
        r.z = fmaf(a[2], g.z, fmaf(b[2], xx.z, d[2])); r.w = fmaf(a[3], g.w, fmaf(b[3], xx.w, d[3]));
        if (dx != nullptr) dx[i] = r;
        if (dx_bf16 != nullptr) dx_bf16[i] = pack_bf16x4(r);
    }
    if (tail_zero && cap_rows > n_rows) {
        const size_t t4 = (size_t)cap_rows * c4;
        for (size_t i = n4 + i0; i < t4; i += stride) {
            if (dx != nullptr) dx[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (dx_bf16 != nullptr) dx_bf16[i] = make_uint2(0u, 0u);
        }
    }
}

int g_bn_fused = 1;     // 1: reduce + apply in one cooperative launch; 0: two launches (vc_set_bn_fused)

static bool c_ok(int c) { return c > 0 && c % 4 == 0 && c <= 128; }

static int ew_blocks(size_t n4) {
    size_t b = (n4 + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace vc

using namespace vc;

// internal entry points (plan executor): device row count, strided outputs / gradients, optional outputs
int vc::bn_apply_relu_dev(const float* x, const double* sums, int n_rows, const int* n_dev, int c, const float* gamma,
                          const float* beta, float* running_mean, float* running_var, long long* num_batches_tracked,
                          float momentum, float eps, int training, float* y, int y_ld, void* y_bf16, int yb_ld,
                          float* stats_out, int relu, int tail_zero, cudaStream_t stream) {
    VC_CHECK_ARG(c_ok(c) && n_rows >= 0, "bad bn args c=%d rows=%d", c, n_rows);
    VC_CHECK_ARG(gamma && beta && running_mean && running_var && stats_out, "null pointer");
    VC_CHECK_ARG(!training || sums, "training-mode BN needs the conv's channel sums");
    VC_CHECK_ARG(!training || n_rows > 0, "training-mode BN over zero rows");
    VC_CHECK_ARG(n_rows == 0 || (x && (y || y_bf16)), "null pointer");
    VC_CHECK_ARG(y_ld % 4 == 0 && yb_ld % 4 == 0 && y_ld >= c && yb_ld >= c, "bad output pitch");
    VC_LAUNCH_CHAIN(bn_apply_relu_kernel, dim3(ew_blocks((size_t)n_rows * c / 4)), dim3(256), 0, stream, (const float4*)x, sums,
                    n_rows, n_dev, c, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps, training,
                    (float4*)y, y_ld / 4, (uint2*)y_bf16, yb_ld / 4, stats_out, relu, tail_zero);
    return VC_OK;
}

int vc::bn_relu_bwd_dev(const float* dy, int dy_ld, const float* x, const float* gamma, const float* stats, float* dx,
                        void* dx_bf16, float* dgamma, float* dbeta, int n, const int* n_dev, int c, int training,
                        double* bsums, int tail_zero, cudaStream_t stream) {
    VC_CHECK_ARG(c_ok(c) && n >= 0, "bad args n=%d c=%d", n, c);
    VC_CHECK_ARG(dgamma && dbeta && bsums && stats && gamma, "null pointer");
    VC_CHECK_ARG(dy_ld % 4 == 0 && dy_ld >= c, "bad gradient pitch");
    if (n > 0 && g_bn_fused) {
        VC_CHECK_ARG(dy && x && (dx || dx_bf16), "null pointer");
        // the grid must be co-resident for the grid-wide barrier: <= 2 blocks per SM (the row loop and the apply loop stride)
        static int max_blocks = 0;
        if (max_blocks == 0) {
            int dev = 0, sms = 148, per_sm = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bn_bwd_fused_kernel, 256, 8192) != cudaSuccess || per_sm < 1) per_sm = 1;
            max_blocks = sms * (per_sm < 2 ? per_sm : 2);
        }
        const int c4 = c / 4, rowlanes = 256 / c4;
        int blocks = (n + rowlanes - 1) / rowlanes;
        if (blocks > max_blocks) blocks = max_blocks;
        const size_t smem = (size_t)rowlanes * 2 * c * sizeof(float) < 3 * 128 * sizeof(float) ? 3 * 128 * sizeof(float)
                                                                                               : (size_t)rowlanes * 2 * c * sizeof(float);
        const float4* dy4 = (const float4*)dy;
        const float4* x4 = (const float4*)x;
        int dy_ld4 = dy_ld / 4;
        float4* dx4 = (float4*)dx;
        uint2* dxb = (uint2*)dx_bf16;
        void* args[] = {&dy4, &dy_ld4, &x4, &gamma, &stats, &bsums, &n, &n_dev, &c, &training, &dx4, &dxb, &dgamma, &dbeta, &tail_zero};
        const cudaError_t ce = cudaLaunchCooperativeKernel((const void*)bn_bwd_fused_kernel, dim3(blocks), dim3(256), args, smem, stream);
        if (ce == cudaSuccess) {
            vc::count_launch();
            return VC_OK;
        }
        // (a device / driver configuration without cooperative launches: fall back to the two-kernel form for good)
        cudaGetLastError();
        g_bn_fused = 0;
    }
    if (n > 0) {
        VC_CHECK_ARG(dy && x && (dx || dx_bf16), "null pointer");
        int c4 = c / 4;
        int rowlanes = 256 / c4;
        int blocks = (n + rowlanes - 1) / rowlanes;
        if (blocks > 296) blocks = 296;
        size_t smem = (size_t)rowlanes * 2 * c * sizeof(float);
        VC_LAUNCH_CHAIN(bn_bwd_reduce_kernel, dim3(blocks), dim3(256), smem, stream, (const float4*)dy, dy_ld / 4, (const float4*)x,
                        stats, n, n_dev, c, bsums);
    }
    VC_LAUNCH_CHAIN(bn_bwd_apply_kernel, dim3(ew_blocks((size_t)n * c / 4)), dim3(256), 0, stream, (const float4*)dy, dy_ld / 4,
                    (const float4*)x, gamma, stats, bsums, n, n_dev, c, training, (float4*)dx, (uint2*)dx_bf16, dgamma, dbeta,
                    tail_zero);
    return VC_OK;
}

extern "C" int vc_bn_apply_relu_f32(const float* x, const double* sums, int n_rows, int c, const float* gamma,
                                    const float* beta, float* running_mean, float* running_var,
                                    long long* num_batches_tracked, float momentum, float eps, int training, float* y,
                                    void* y_bf16, float* stats_out, int relu, vc_stream_t stream_) {
    VC_CHECK_ARG(n_rows == 0 || y, "null pointer");
    return vc::bn_apply_relu_dev(x, sums, n_rows, nullptr, c, gamma, beta, running_mean, running_var, num_batches_tracked, momentum,
                                 eps, training, y, c, y_bf16, c, stats_out, relu, 0, (cudaStream_t)stream_);
}

/* (y is no longer read: the ReLU mask is recomputed from x and the saved scale / shift — kept in the signature) */
extern "C" int vc_bn_relu_bwd_f32(const float* dy, const float* x, const float* y, const float* gamma, const float* stats,
                                  float* dx, void* dx_bf16, float* dgamma, float* dbeta, int n, int c, int training,
                                  double* bsums, vc_stream_t stream_) {
    (void)y;
    VC_CHECK_ARG(n == 0 || dx, "null pointer");
    return vc::bn_relu_bwd_dev(dy, c, x, gamma, stats, dx, dx_bf16, dgamma, dbeta, n, nullptr, c, training, bsums, 0,
                               (cudaStream_t)stream_);
}

extern "C" int vc_set_bn_fused(int enable) {
    vc::g_bn_fused = enable ? 1 : 0;
    return VC_OK;
}
