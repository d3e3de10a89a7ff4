// BatchNorm1d(eps, momentum) + ReLU over the active rows of a sparse tensor — sm_100a.
//
// Replaces the `norm_fn(out_channels)` / `nn.ReLU()` members of every spconv.SparseSequential
// (spconv_backbone.py:101-105, :160 eps=1e-3 momentum=0.01, :561-567).  HBM-bound elementwise work:
// forward apply reads N*C*4 and writes N*C*4 bytes; the channel statistics come for free from the conv
// epilogue (per-tile partial sums) and are reduced here in a fixed order (deterministic).
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

// one block; thread (j, ch): j strides over tiles.  double accumulation of the fp32 tile partials.
__global__ void __launch_bounds__(1024) bn_train_finalize_kernel(
    const float* __restrict__ partial, int n_tiles, int n_rows, int c, const float* __restrict__ gamma,
    const float* __restrict__ beta, float* running_mean, float* running_var, float momentum, float eps,
    float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ save_mean,
    float* __restrict__ save_invstd) {
    extern __shared__ double sh[];  // [lanes][2][c]
    int ch = threadIdx.x % c, j = threadIdx.x / c, lanes = blockDim.x / c;
    double s = 0.0, q = 0.0, s1 = 0.0, q1 = 0.0;
    int t = j;
    for (; t + lanes < n_tiles; t += 2 * lanes) {     // two independent chains: loads of both are in flight together
        s += (double)__ldg(partial + ((size_t)t * 2 + 0) * c + ch);
        q += (double)__ldg(partial + ((size_t)t * 2 + 1) * c + ch);
        s1 += (double)__ldg(partial + ((size_t)(t + lanes) * 2 + 0) * c + ch);
        q1 += (double)__ldg(partial + ((size_t)(t + lanes) * 2 + 1) * c + ch);
    }
    for (; t < n_tiles; t += lanes) {
        s += (double)__ldg(partial + ((size_t)t * 2 + 0) * c + ch);
        q += (double)__ldg(partial + ((size_t)t * 2 + 1) * c + ch);
    }
    s += s1;
    q += q1;
    sh[(j * 2 + 0) * c + ch] = s;
    sh[(j * 2 + 1) * c + ch] = q;
    __syncthreads();
    if (j == 0) {
        for (int l = 1; l < lanes; ++l) {
            s += sh[(l * 2 + 0) * c + ch];
            q += sh[(l * 2 + 1) * c + ch];
        }
        double n = (double)n_rows;
        double mean = s / n;
        double var = q / n - mean * mean;
        if (var < 0.0) var = 0.0;
        float invstd = (float)(1.0 / sqrt(var + (double)eps));
        float sc = gamma[ch] * invstd;
        scale[ch] = sc;
        shift[ch] = beta[ch] - (float)mean * sc;
        save_mean[ch] = (float)mean;
        save_invstd[ch] = invstd;
        double unbiased = n_rows > 1 ? var * n / (n - 1.0) : var;
        running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)mean;
        running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
    }
}

__global__ void bn_eval_affine_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ rm, const float* __restrict__ rv, float eps, int c,
                                      float* scale, float* shift, float* save_mean, float* save_invstd) {
    int ch = threadIdx.x;
    if (ch >= c) return;
    float invstd = 1.f / sqrtf(rv[ch] + eps);
    float sc = gamma[ch] * invstd;
    scale[ch] = sc;
    shift[ch] = beta[ch] - rm[ch] * sc;
    save_mean[ch] = rm[ch];
    save_invstd[ch] = invstd;
}

__global__ void __launch_bounds__(256) affine_relu_kernel(const float4* __restrict__ x, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, float4* __restrict__ y,
                                                          uint2* __restrict__ y_bf16, size_t n4, int c4, int relu) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) {
        int cg = (int)(i % c4);
        float4 v = x[i];
        float4 sc = reinterpret_cast<const float4*>(scale)[cg];
        float4 sh = reinterpret_cast<const float4*>(shift)[cg];
        v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
        v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
        if (relu) {
            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        }
        y[i] = v;
        if (y_bf16 != nullptr) y_bf16[i] = pack_bf16x4(v);   // shadow copy: the next conv's tensor-core operand
    }
}

// backward pass 1: per-block partial sums of g = dy*(y>0) and g*xhat      partial [blocks][2][c]
static constexpr int BWD_BLOCKS = 296;
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float4* __restrict__ dy, const float4* __restrict__ x,
                                                            const float4* __restrict__ y,
                                                            const float* __restrict__ mean,
                                                            const float* __restrict__ invstd, int n, int c,
                                                            float* __restrict__ partial) {
    extern __shared__ float shf[];  // [rowlanes][2][c]
    int c4 = c / 4;
    int cg = threadIdx.x % c4, rl = threadIdx.x / c4, rowlanes = blockDim.x / c4;
    float4 m = reinterpret_cast<const float4*>(mean)[cg];
    float4 is = reinterpret_cast<const float4*>(invstd)[cg];
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
        partial[((size_t)blockIdx.x * 2 + which) * c + ch] = v;
    }
}

// pass 2 (one block, lanes x c threads): dgamma, dbeta and the per-channel coefficients of
//   dx = coef[0]*g + coef[1]*x + coef[2]
__global__ void __launch_bounds__(1024) bn_bwd_finalize_kernel(const float* __restrict__ partial, int blocks, int n, int c,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ mean,
                                                                const float* __restrict__ invstd, int training,
                                                                float* dgamma, float* dbeta, float* __restrict__ coef) {
    extern __shared__ double shd[];  // [lanes][2][c]
    int ch = threadIdx.x % c, j = threadIdx.x / c, lanes = blockDim.x / c;
    double s = 0.0, q = 0.0;
    for (int b = j; b < blocks; b += lanes) {
        s += (double)partial[((size_t)b * 2 + 0) * c + ch];
        q += (double)partial[((size_t)b * 2 + 1) * c + ch];
    }
    shd[(j * 2 + 0) * c + ch] = s;
    shd[(j * 2 + 1) * c + ch] = q;
    __syncthreads();
    if (j != 0) return;
    for (int l = 1; l < lanes; ++l) {
        s += shd[(l * 2 + 0) * c + ch];
        q += shd[(l * 2 + 1) * c + ch];
    }
    dbeta[ch] = (float)s;
    dgamma[ch] = (float)q;
    float gi = gamma[ch] * invstd[ch];
    if (training) {
        // dx = gi*(g - s/n - xhat*q/n),  xhat = (x-mean)*invstd
        float k1 = (float)(q / (double)n) * invstd[ch];
        coef[ch] = gi;
        coef[c + ch] = -gi * k1;
        coef[2 * c + ch] = gi * (k1 * mean[ch] - (float)(s / (double)n));
    } else {
        coef[ch] = gi;
        coef[c + ch] = 0.f;
        coef[2 * c + ch] = 0.f;
    }
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float4* __restrict__ dy, const float4* __restrict__ x,
                                                           const float4* __restrict__ y, const float* __restrict__ coef,
                                                           float4* __restrict__ dx, uint2* __restrict__ dx_bf16, size_t n4,
                                                           int c) {
    int c4 = c / 4;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) {
        int cg = (int)(i % c4);
        float4 a = reinterpret_cast<const float4*>(coef)[cg];
        float4 b = reinterpret_cast<const float4*>(coef + c)[cg];
        float4 d = reinterpret_cast<const float4*>(coef + 2 * c)[cg];
        float4 g = dy[i], yy = y[i], xx = x[i], r;
        g.x = yy.x > 0.f ? g.x : 0.f; g.y = yy.y > 0.f ? g.y : 0.f;
        g.z = yy.z > 0.f ? g.z : 0.f; g.w = yy.w > 0.f ? g.w : 0.f;
        r.x = fmaf(a.x, g.x, fmaf(b.x, xx.x, d.x)); r.y = fmaf(a.y, g.y, fmaf(b.y, xx.y, d.y));
        r.z = fmaf(a.z, g.z, fmaf(b.z, xx.z, d.z)); r.w = fmaf(a.w, g.w, fmaf(b.w, xx.w, d.w));
        dx[i] = r;
        if (dx_bf16 != nullptr) dx_bf16[i] = pack_bf16x4(r);
    }
}

static bool c_ok(int c) { return c > 0 && c % 4 == 0 && c <= 256; }

}  // namespace vc

using namespace vc;

extern "C" int vc_bn_train_finalize(const float* bn_partial, int n_tiles, int n_rows, int c, const float* gamma,
                                    const float* beta, float* running_mean, float* running_var, float momentum,
                                    float eps, float* scale, float* shift, float* save_mean, float* save_invstd,
                                    vc_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    VC_CHECK_ARG(c_ok(c) && n_tiles >= 0 && n_rows > 0, "bad bn args c=%d tiles=%d rows=%d", c, n_tiles, n_rows);
    VC_CHECK_ARG(bn_partial && gamma && beta && running_mean && running_var && scale && shift && save_mean && save_invstd,
                 "null pointer");
    int lanes = 1024 / c;
    if (lanes > 32) lanes = 32;
    size_t smem = (size_t)lanes * 2 * c * sizeof(double);
    bn_train_finalize_kernel<<<1, lanes * c, smem, stream>>>(bn_partial, n_tiles, n_rows, c, gamma, beta, running_mean,
                                                              running_var, momentum, eps, scale, shift, save_mean,
                                                              save_invstd);
    VC_LAUNCH_CHECK();
    return VC_OK;
}

extern "C" int vc_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean,
                                 const float* running_var, float eps, int c, float* scale, float* shift,
                                 float* save_mean, float* save_invstd, vc_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    VC_CHECK_ARG(c_ok(c), "bad channel count %d", c);
    bn_eval_affine_kernel<<<1, 256, 0, stream>>>(gamma, beta, running_mean, running_var, eps, c, scale, shift,
                                                  save_mean, save_invstd);
    VC_LAUNCH_CHECK();
    return VC_OK;
}

extern "C" int vc_affine_relu_f32(const float* x, const float* scale, const float* shift, float* y, void* y_bf16, int n,
                                  int c, int relu, vc_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    VC_CHECK_ARG(c_ok(c) && n >= 0, "bad args n=%d c=%d", n, c);
    if (n == 0) return VC_OK;
    size_t n4 = (size_t)n * c / 4;
    int blocks = (int)((n4 + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    affine_relu_kernel<<<blocks, 256, 0, stream>>>((const float4*)x, scale, shift, (float4*)y, (uint2*)y_bf16, n4, c / 4,
                                                   relu);
    VC_LAUNCH_CHECK();
    return VC_OK;
}

extern "C" size_t vc_bn_bwd_ws_bytes(int n, int c) { return ((size_t)BWD_BLOCKS * 2 * c + 3 * c) * sizeof(float); }

extern "C" int vc_bn_relu_bwd_f32(const float* dy, const float* x, const float* y, const float* gamma,
                                  const float* save_mean, const float* save_invstd, float* dx, void* dx_bf16,
                                  float* dgamma, float* dbeta, int n, int c, int training, void* ws, size_t ws_bytes,
                                  vc_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    VC_CHECK_ARG(c_ok(c) && n >= 0 && c <= 64 * 4, "bad args n=%d c=%d", n, c);
    VC_CHECK_ARG(dgamma && dbeta && ws, "null pointer");
    if (ws_bytes < vc_bn_bwd_ws_bytes(n, c)) {
        set_error("bn bwd workspace %zu < %zu", ws_bytes, vc_bn_bwd_ws_bytes(n, c));
        return VC_ERR_WORKSPACE;
    }
    if (n == 0) {
        VC_CUDA(cudaMemsetAsync(dgamma, 0, c * 4, stream));
        VC_CUDA(cudaMemsetAsync(dbeta, 0, c * 4, stream));
        return VC_OK;
    }
    float* partial = (float*)ws;
    float* coef = partial + (size_t)BWD_BLOCKS * 2 * c;
    int c4 = c / 4;
    int rowlanes = 256 / c4;
    int blocks = (n + rowlanes - 1) / rowlanes;
    if (blocks > BWD_BLOCKS) blocks = BWD_BLOCKS;
    size_t smem = (size_t)rowlanes * 2 * c * sizeof(float);
    bn_bwd_reduce_kernel<<<blocks, 256, smem, stream>>>((const float4*)dy, (const float4*)x, (const float4*)y, save_mean,
                                                        save_invstd, n, c, partial);
    VC_LAUNCH_CHECK();
    int flanes = 1024 / c;
    if (flanes > 32) flanes = 32;
    bn_bwd_finalize_kernel<<<1, flanes * c, (size_t)flanes * 2 * c * sizeof(double), stream>>>(
        partial, blocks, n, c, gamma, save_mean, save_invstd, training, dgamma, dbeta, coef);
    VC_LAUNCH_CHECK();
    size_t n4 = (size_t)n * c4;
    int ablocks = (int)((n4 + 255) / 256);
    if (ablocks > 148 * 16) ablocks = 148 * 16;
    bn_bwd_apply_kernel<<<ablocks, 256, 0, stream>>>((const float4*)dy, (const float4*)x, (const float4*)y, coef,
                                                     (float4*)dx, (uint2*)dx_bf16, n4, c);
    VC_LAUNCH_CHECK();
    return VC_OK;
}
