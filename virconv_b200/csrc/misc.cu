// index2uv, dense(), row gather, error plumbing — sm_100a.
#include <cuda_bf16.h>
#include <stdarg.h>

#include "common.cuh"

namespace vc {

static thread_local char g_err[512] = "";
static long long g_launches = 0;
int g_pdl = 1;
void count_launch() { ++g_launches; }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

struct UVGrid {
    float vx, vy, vz, mx, my, mz;
};

// Voxel index -> pixel cell.  Every float op is an explicitly rounded fp32 multiply / add / divide in the
// order oracle/index2uv.py fixes (no FMA contraction), so the integer result is bit-identical to it.
// Reference arithmetic: spconv_backbone.py:8-24,54-83; X_transform.py:139-154; calibration_kitti.py:120-153.
__global__ void __launch_bounds__(256) index2uv_kernel(const int4* __restrict__ idx, int n, const int* __restrict__ n_dev,
                                                       int batch_size, const float* __restrict__ params, UVGrid g, int stride,
                                                       int u_max, int v_max, int32_t* __restrict__ uv) {
    if (n_dev != nullptr) n = min(n, __ldg(n_dev));
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int4 v = idx[i];  // (b, z, y, x)
    int b = v.x;
    float x = __fadd_rn(__fmul_rn((float)v.w, g.vx), g.mx);
    float y = __fadd_rn(__fmul_rn((float)v.z, g.vy), g.my);
    float z = __fadd_rn(__fmul_rn((float)v.y, g.vz), g.mz);
    int ui = 0, vi = 0;
    if (b >= 0 && b < batch_size) {
        const float* p = params + (size_t)b * 28;
        if (p[20] != 0.f) {
            float sc = p[21];
            x = __fdiv_rn(x, sc);
            y = __fdiv_rn(y, sc);
            z = __fdiv_rn(z, sc);
            if (p[22] != 0.f) y = -y;
            float c = p[23], s = p[24];
            float xr = __fadd_rn(__fmul_rn(x, c), __fmul_rn(y, -s));
            float yr = __fadd_rn(__fmul_rn(x, s), __fmul_rn(y, c));
            x = xr;
            y = yr;
        }
        float r[3];
#pragma unroll
        for (int j = 0; j < 3; ++j)
            r[j] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, p[0 * 3 + j]), __fmul_rn(y, p[1 * 3 + j])),
                                       __fmul_rn(z, p[2 * 3 + j])),
                             p[3 * 3 + j]);
        float h[2];
#pragma unroll
        for (int j = 0; j < 2; ++j)
            h[j] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r[0], p[12 + 0 * 2 + j]), __fmul_rn(r[1], p[12 + 1 * 2 + j])),
                                       __fmul_rn(r[2], p[12 + 2 * 2 + j])),
                             p[12 + 3 * 2 + j]);
        float uf = __fdiv_rn(h[0], r[2]);
        float vf = __fdiv_rn(h[1], r[2]);
        ui = __float2int_rz(uf);  // saturating, NaN -> 0: what `.int()` gives on the reference's CUDA path
        vi = __float2int_rz(vf);
    }
    ui = min(max(ui, 0), u_max - 1) / stride;
    vi = min(max(vi, 0), v_max - 1) / stride;
    int32_t* o = uv + (size_t)i * 3;
    o[0] = b;
    o[1] = ui;
    o[2] = vi;
}

__global__ void dense_kernel(const float* __restrict__ feat, const int32_t* __restrict__ idx, int n, int c, int ndim,
                             long long spatial, const int* __restrict__ shape_dev_unused, int s0, int s1, int s2,
                             float* __restrict__ out, int backward, float* __restrict__ dfeat) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n * c) return;
    int row = (int)(t / c), ch = (int)(t % c);
    const int32_t* p = idx + (size_t)row * (1 + ndim);
    long long cell = p[1];
    if (ndim >= 2) cell = cell * s1 + p[2];
    if (ndim >= 3) cell = cell * s2 + p[3];
    long long o = ((long long)p[0] * c + ch) * spatial + cell;
    if (!backward)
        out[o] = feat[t];
    else
        dfeat[t] = out[o];
}

__global__ void gather_rows_kernel(const uint4* __restrict__ in, const int32_t* __restrict__ rows, uint4* __restrict__ out,
                                   int n_rows, int chunks) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_rows * chunks) return;
    int r = (int)(t / chunks), c = (int)(t % chunks);
    out[t] = in[(size_t)rows[r] * chunks + c];
}
__global__ void gather_rows4_kernel(const uint32_t* __restrict__ in, const int32_t* __restrict__ rows,
                                    uint32_t* __restrict__ out, int n_rows, int words) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_rows * words) return;
    int r = (int)(t / words), c = (int)(t % words);
    out[t] = in[(size_t)rows[r] * words + c];
}

// out[:, :ca] = a, out[:, ca:] = b (fp32) plus an optional bf16 shadow of the same matrix; float4 granularity
// (n_dev: device row count, n = capacity; rows [count, capacity) are written as zeros — the concat is a published tensor)
__global__ void cat2_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ out,
                            uint2* __restrict__ out_bf16, int n, const int* __restrict__ n_dev, int ca4, int cb4) {
    pdl_wait();
    pdl_launch_dependents();
    const int cnt = n_dev != nullptr ? min(n, __ldg(n_dev)) : n;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int w = ca4 + cb4;
    if (t >= (long long)n * w) return;
    int r = (int)(t / w), c = (int)(t % w);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < cnt) v = c < ca4 ? a[(size_t)r * ca4 + c] : b[(size_t)r * cb4 + (c - ca4)];
    out[t] = v;
    if (out_bf16 != nullptr) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&lo);
        o.y = *reinterpret_cast<uint32_t*>(&hi);
        out_bf16[t] = o;
    }
}

}  // namespace vc

using namespace vc;

int vc::cat2_dev(const float* a, const float* b, float* out, void* out_bf16, int n, const int* n_dev, int ca, int cb,
                 cudaStream_t stream) {
    VC_CHECK_ARG(n >= 0 && ca > 0 && cb > 0 && ca % 4 == 0 && cb % 4 == 0, "bad cat2 arguments");
    if (n == 0) return VC_OK;
    VC_CHECK_ARG(a && b && out, "null pointer");
    long long total = (long long)n * (ca + cb) / 4;
    VC_LAUNCH_CHAIN(cat2_kernel, dim3(cdiv(total, 256)), dim3(256), 0, stream, (const float4*)a, (const float4*)b, (float4*)out,
                    (uint2*)out_bf16, n, n_dev, ca / 4, cb / 4);
    return VC_OK;
}

extern "C" int vc_cat2_f32(const float* a, const float* b, float* out, void* out_bf16, int n, int ca, int cb,
                           vc_stream_t stream_) {
    return vc::cat2_dev(a, b, out, out_bf16, n, nullptr, ca, cb, (cudaStream_t)stream_);
}

extern "C" int vc_version(void) { return 100; }
extern "C" const char* vc_last_error(void) { return g_err; }
extern "C" long long vc_launch_count(void) { return g_launches; }
extern "C" int vc_set_pdl(int enable) {
    vc::g_pdl = enable != 0;
    return VC_OK;
}

int vc::index2uv_dev(const int32_t* indices, int n, const int* n_dev, int batch_size, const float* params, const float* grid,
                     int stride, int u_max, int v_max, int32_t* uv_out, cudaStream_t stream) {
    VC_CHECK_ARG(n >= 0 && batch_size > 0 && stride > 0 && u_max > 0 && v_max > 0 && grid, "bad index2uv arguments");
    if (n == 0) return VC_OK;
    VC_CHECK_ARG(indices && params && uv_out, "null pointer");
    UVGrid g{grid[0], grid[1], grid[2], grid[3], grid[4], grid[5]};
    index2uv_kernel<<<cdiv(n, 256), 256, 0, stream>>>((const int4*)indices, n, n_dev, batch_size, params, g, stride, u_max, v_max,
                                                      uv_out);
    VC_LAUNCH_CHECK();
    return VC_OK;
}

extern "C" int vc_index2uv(const int32_t* indices, int n, int batch_size, const float* params, const float* grid,
                           int stride, int u_max, int v_max, int32_t* uv_out, vc_stream_t stream_) {
    return vc::index2uv_dev(indices, n, nullptr, batch_size, params, grid, stride, u_max, v_max, uv_out, (cudaStream_t)stream_);
}

static int dense_common(const float* features, const int32_t* indices, int n, int c, int ndim, int batch_size,
                        const int32_t* shape, float* dense, float* dfeat, int backward, cudaStream_t stream) {
    VC_CHECK_ARG(n >= 0 && c > 0 && ndim >= 1 && ndim <= VC_MAX_NDIM && batch_size > 0 && shape, "bad dense arguments");
    if (n == 0) return VC_OK;
    long long spatial = 1;
    for (int d = 0; d < ndim; ++d) spatial *= shape[d];
    long long total = (long long)n * c;
    dense_kernel<<<cdiv(total, 256), 256, 0, stream>>>(features, indices, n, c, ndim, spatial, nullptr, shape[0],
                                                       ndim > 1 ? shape[1] : 1, ndim > 2 ? shape[2] : 1, dense, backward,
                                                       dfeat);
    VC_LAUNCH_CHECK();
    return VC_OK;
}

extern "C" int vc_dense_f32(const float* features, const int32_t* indices, int n, int c, int ndim, int batch_size,
                            const int32_t* spatial_shape, float* out, vc_stream_t stream_) {
    return dense_common(features, indices, n, c, ndim, batch_size, spatial_shape, out, nullptr, 0, (cudaStream_t)stream_);
}

extern "C" int vc_dense_bwd_f32(const float* dout, const int32_t* indices, int n, int c, int ndim, int batch_size,
                                const int32_t* spatial_shape, float* dfeatures, vc_stream_t stream_) {
    return dense_common(nullptr, indices, n, c, ndim, batch_size, spatial_shape, const_cast<float*>(dout), dfeatures, 1,
                        (cudaStream_t)stream_);
}

extern "C" int vc_gather_rows(const void* in, const int32_t* rows, void* out, int n_rows, int row_bytes,
                              vc_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    VC_CHECK_ARG(n_rows >= 0 && row_bytes > 0 && row_bytes % 4 == 0, "bad gather arguments");
    if (n_rows == 0) return VC_OK;
    VC_CHECK_ARG(in && rows && out, "null pointer");
    if (row_bytes % 16 == 0 && ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0)) {
        int chunks = row_bytes / 16;
        gather_rows_kernel<<<cdiv((long long)n_rows * chunks, 256), 256, 0, stream>>>((const uint4*)in, rows, (uint4*)out,
                                                                                      n_rows, chunks);
    } else {
        int words = row_bytes / 4;
        gather_rows4_kernel<<<cdiv((long long)n_rows * words, 256), 256, 0, stream>>>((const uint32_t*)in, rows,
                                                                                      (uint32_t*)out, n_rows, words);
    }
    VC_LAUNCH_CHECK();
    return VC_OK;
}
