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
    const float4* __restrict__ x, const double* __restrict__ sums, int n_rows, int c, const float* __restrict__ gamma,
    const float* __restrict__ beta, float* running_mean, float* running_var, long long* num_batches_tracked,
    float momentum, float eps, int training, float4* __restrict__ y, uint2* __restrict__ y_bf16,
    float* __restrict__ stats_out, int relu) {
    pdl_wait();                 // the conv's output and channel sums
    pdl_launch_dependents();
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
            const double n = (double)n_rows;
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
    for (size_t i = i0; i < n4; i += stride) {
        float4 v = x[i];
        v.x = fmaf(v.x, sc[0], sh[0]); v.y = fmaf(v.y, sc[1], sh[1]);
        v.z = fmaf(v.z, sc[2], sh[2]); v.w = fmaf(v.w, sc[3], sh[3]);
        if (relu) {
            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        }
        y[i] = v;
        if (y_bf16 != nullptr) y_bf16[i] = pack_bf16x4(v);   // shadow copy: the next conv's tensor-core operand
    }
}

// backward pass 1: sum g and sum g*xhat over rows, g = dy*(y>0); block partial -> float64 atomics on bsums [2, c]
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float4* __restrict__ dy, const float4* __restrict__ x,
                                                            const float4* __restrict__ y,
                                                            const float* __restrict__ stats, int n, int c,
                                                            double* __restrict__ bsums) {
    extern __shared__ float shf[];  // [rowlanes][2][c]
    pdl_wait();
    pdl_launch_dependents();
    const int c4 = c / 4;
    const int cg = threadIdx.x % c4, rl = threadIdx.x / c4, rowlanes = blockDim.x / c4;
    const float4 m = reinterpret_cast<const float4*>(stats + 2 * c)[cg];
    const float4 is = reinterpret_cast<const float4*>(stats + 3 * c)[cg];
    float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    for (int row = blockIdx.x * rowlanes + rl; row < n; row += gridDim.x * rowlanes) {
        size_t i = (size_t)row * c4 + cg;
        float4 g = dy[i], yy = y[i], xx = x[i];
        g.x = yy.x > 0.f ? g.x : 0.f; g.y = yy.y > 0.f ? g.y : 0.f;
        g.z = yy.z > 0.f ? g.z : 0.f; g.w = yy.w > 0.f ? g.w : 0.f;
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
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float4* __restrict__ dy, const float4* __restrict__ x,
                                                           const float4* __restrict__ y, const float* __restrict__ gamma,
                                                           const float* __restrict__ stats,
                                                           const double* __restrict__ bsums, int n_rows, int c, int training,
                                                           float4* __restrict__ dx, uint2* __restrict__ dx_bf16,
                                                           float* __restrict__ dgamma, float* __restrict__ dbeta) {
    pdl_wait();                 // bsums from the reduce kernel
    pdl_launch_dependents();
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
            const float k1 = (float)(Q / (double)n_rows) * invstd;
            a_s[ch] = gi;
            b_s[ch] = -gi * k1;
            d_s[ch] = gi * (k1 * mean - (float)(S / (double)n_rows));
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
    for (size_t i = i0; i < n4; i += stride) {
        float4 g = dy[i], yy = y[i], xx = x[i], r;
        g.x = yy.x > 0.f ? g.x : 0.f; g.y = yy.y > 0.f ? g.y : 0.f;
        g.z = yy.z > 0.f ? g.z : 0.f; g.w = yy.w > 0.f ? g.w : 0.f;
        r.x = fmaf(a[0], g.x, fmaf(b[0], xx.x, d[0])); r.y = fmaf(a[1], g.y, fmaf(b[1], xx.y, d[1]));
        r.z = fmaf(a[2], g.z, fmaf(b[2], xx.z, d[2])); r.w = fmaf(a[3], g.w, fmaf(b[3], xx.w, d[3]));
        dx[i] = r;
        if (dx_bf16 != nullptr) dx_bf16[i] = pack_bf16x4(r);
    }
}

static bool c_ok(int c) { return c > 0 && c % 4 == 0 && c <= 128; }

static int ew_blocks(size_t n4) {
    size_t b = (n4 + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace vc

using namespace vc;

extern "C" int vc_bn_apply_relu_f32(const float* x, const double* sums, int n_rows, int c, const float* gamma,
                                    const float* beta, float* running_mean, float* running_var,
                                    long long* num_batches_tracked, float momentum, float eps, int training, float* y,
                                    void* y_bf16, float* stats_out, int relu, vc_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    VC_CHECK_ARG(c_ok(c) && n_rows >= 0, "bad bn args c=%d rows=%d", c, n_rows);
    VC_CHECK_ARG(gamma && beta && running_mean && running_var && stats_out, "null pointer");
    VC_CHECK_ARG(!training || sums, "training-mode BN needs the conv's channel sums");
    VC_CHECK_ARG(!training || n_rows > 0, "training-mode BN over zero rows");
    VC_CHECK_ARG(n_rows == 0 || (x && y), "null pointer");
    VC_LAUNCH_CHAIN(bn_apply_relu_kernel, dim3(ew_blocks((size_t)n_rows * c / 4)), dim3(256), 0, stream, (const float4*)x, sums,
                    n_rows, c, gamma, beta, running_mean, running_var, num_batches_tracked, momentum, eps, training, (float4*)y,
                    (uint2*)y_bf16, stats_out, relu);
    return VC_OK;
}

extern "C" int vc_bn_relu_bwd_f32(const float* dy, const float* x, const float* y, const float* gamma, const float* stats,
                                  float* dx, void* dx_bf16, float* dgamma, float* dbeta, int n, int c, int training,
                                  double* bsums, vc_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    VC_CHECK_ARG(c_ok(c) && n >= 0, "bad args n=%d c=%d", n, c);
    VC_CHECK_ARG(dgamma && dbeta && bsums && stats && gamma, "null pointer");
    if (n > 0) {
        VC_CHECK_ARG(dy && x && y && dx, "null pointer");
        int c4 = c / 4;
        int rowlanes = 256 / c4;
        int blocks = (n + rowlanes - 1) / rowlanes;
        if (blocks > 296) blocks = 296;
        size_t smem = (size_t)rowlanes * 2 * c * sizeof(float);
        VC_LAUNCH_CHAIN(bn_bwd_reduce_kernel, dim3(blocks), dim3(256), smem, stream, (const float4*)dy, (const float4*)x,
                        (const float4*)y, stats, n, c, bsums);
    }
    VC_LAUNCH_CHAIN(bn_bwd_apply_kernel, dim3(ew_blocks((size_t)n * c / 4)), dim3(256), 0, stream, (const float4*)dy,
                    (const float4*)x, (const float4*)y, gamma, stats, bsums, n, c, training, (float4*)dx, (uint2*)dx_bf16, dgamma,
                    dbeta);
    return VC_OK;
}
