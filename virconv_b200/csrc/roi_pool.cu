// Voxel-RoI pooling primitives of the consumers of x_conv3 / x_conv4 — sm_100a.
//
// Replaces, with identical results, the stacked pointnet2 ops the RoI head calls on the backbone's outputs:
//   voxel_query   pcdet/ops/pointnet2/pointnet2_stack/src/voxel_query_gpu.cu:10-89  (`VoxelQuery.forward`,
//                 voxel_query_utils.py:10-44; called from `VoxelQueryAndGrouping.forward` :80 for every RoI grid point,
//                 ted_head.py:533-541)
//   group_points  .../src/group_points_gpu.cu:71-103 (+ grad :15-45)  (`GroupingOperation`, pointnet2_utils.py:48-102)
// The reference walks the (2r+1)^3 neighbourhood of a query point with ONE thread (up to 729 dependent-latency loads of
// the dense voxel->row map); here a warp owns the query point: 32 cells of the z-major scan order are probed at once and
// a ballot + prefix popcount keeps the reference's "first nsample hits in scan order" rule exactly.  Grouping moves
// whole feature rows (coalesced) and transposes through shared memory into the [M, C, nsample] layout instead of one
// 4-byte load per output element; its backward uses the same tiles with float atomics, like the reference.
// HBM/L2-bound integer + copy work: per query ~ (2r+1)^3 * 4 B of map probes (L2 resident: 24 MB at stride 4) and
// nsample rows of C floats.
#include "common.cuh"

namespace vc {
namespace {

constexpr int QWARPS = 8;   // query points per block

__global__ void __launch_bounds__(QWARPS * 32) voxel_query_kernel(int M, int R1, int R2, int R3, int nsample, float radius,
                                                                  int z_range, int y_range, int x_range,
                                                                  const float* __restrict__ new_xyz,
                                                                  const float* __restrict__ xyz,
                                                                  const int* __restrict__ new_coords,
                                                                  const int* __restrict__ point_indices,
                                                                  int* __restrict__ idx, unsigned char* __restrict__ empty) {
    const int lane = threadIdx.x & 31;
    const int pt = blockIdx.x * QWARPS + (threadIdx.x >> 5);
    if (pt >= M) return;
    const float radius2 = radius * radius;
    const float nx = new_xyz[pt * 3 + 0], ny = new_xyz[pt * 3 + 1], nz = new_xyz[pt * 3 + 2];
    const int b = new_coords[pt * 4 + 0], cz = new_coords[pt * 4 + 1], cy = new_coords[pt * 4 + 2], cx = new_coords[pt * 4 + 3];
    const int dz_n = 2 * z_range + 1, dy_n = 2 * y_range + 1, dx_n = 2 * x_range + 1;
    const int cells = dz_n * dy_n * dx_n;
    int* out = idx + (size_t)pt * nsample;
    int cnt = 0, first = -1;
    for (int c0 = 0; c0 < cells && cnt < nsample; c0 += 32) {
        const int c = c0 + lane;
        int nb = -1;
        if (c < cells) {
            const int dx = c % dx_n - x_range, dy = (c / dx_n) % dy_n - y_range, dz = c / (dx_n * dy_n) - z_range;
            const int z = cz + dz, y = cy + dy, x = cx + dx;
            if (z >= 0 && z < R1 && y >= 0 && y < R2 && x >= 0 && x < R3) {
                // same int arithmetic as the reference (voxel_query_gpu.cu:54-57)
                const int index = b * R1 * R2 * R3 + z * R2 * R3 + y * R3 + x;
                const int cand = __ldg(point_indices + index);
                if (cand >= 0) {
                    const float xp = __ldg(xyz + cand * 3 + 0), yp = __ldg(xyz + cand * 3 + 1), zp = __ldg(xyz + cand * 3 + 2);
                    // same expression as the reference (:65), compiled with the same default contraction
                    const float dist2 = (xp - nx) * (xp - nx) + (yp - ny) * (yp - ny) + (zp - nz) * (zp - nz);
                    if (!(dist2 > radius2)) nb = cand;
                }
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, nb >= 0);
        if (m) {
            if (first < 0) first = __shfl_sync(0xffffffffu, nb, __ffs(m) - 1);
            const int rank = cnt + __popc(m & ((1u << lane) - 1u));
            if (nb >= 0 && rank < nsample) out[rank] = nb;
            cnt += __popc(m);
        }
    }
    if (cnt > nsample) cnt = nsample;
    // the reference pre-fills all nsample slots with the first hit (:72-76); an empty ball is row 0.. = 0 after
    // `idx[empty_ball_mask] = 0` (voxel_query_utils.py:38-39)
    for (int l = cnt + lane; l < nsample; l += 32) out[l] = first < 0 ? 0 : first;
    if (lane == 0) empty[pt] = first < 0 ? 1 : 0;
}

// out[pt, c, s] = features[start[pt's batch] + idx[pt, s], c]; one block = 8 query points, smem tile [nsample][C+1]
constexpr int GP = 8;

__device__ __forceinline__ int batch_start(int pt, int B, const int* __restrict__ idx_batch_cnt,
                                           const int* __restrict__ features_batch_cnt) {
    // group_points_gpu.cu:87-95: which sample the query point belongs to, start row of that sample's features
    int bs = 0, pt_cnt = idx_batch_cnt[0];
    for (int k = 1; k < B; k++) {
        if (pt < pt_cnt) break;
        pt_cnt += idx_batch_cnt[k];
        bs = k;
    }
    int start = 0;
    for (int k = 0; k < bs; k++) start += features_batch_cnt[k];
    return start;
}

__global__ void __launch_bounds__(256) group_points_kernel(int B, int M, int C, int nsample,
                                                           const float* __restrict__ features,
                                                           const int* __restrict__ features_batch_cnt,
                                                           const int* __restrict__ idx,
                                                           const int* __restrict__ idx_batch_cnt, float* __restrict__ out) {
    extern __shared__ float tile[];   // [GP][nsample][C + 1]
    const int pt0 = blockIdx.x * GP;
    const int stride = C + 1;
    // gather: consecutive threads read consecutive channels of one (point, sample) row
    for (int q = threadIdx.x; q < GP * nsample * C; q += blockDim.x) {
        const int c = q % C, s = (q / C) % nsample, p = q / (C * nsample);
        const int pt = pt0 + p;
        if (pt < M) {
            const int start = batch_start(pt, B, idx_batch_cnt, features_batch_cnt);
            const int row = start + idx[(size_t)pt * nsample + s];
            tile[(p * nsample + s) * stride + c] = __ldg(features + (size_t)row * C + c);
        }
    }
    __syncthreads();
    // scatter: consecutive threads write consecutive samples of one (point, channel)
    for (int q = threadIdx.x; q < GP * C * nsample; q += blockDim.x) {
        const int s = q % nsample, c = (q / nsample) % C, p = q / (nsample * C);
        const int pt = pt0 + p;
        if (pt < M) out[((size_t)pt * C + c) * nsample + s] = tile[(p * nsample + s) * stride + c];
    }
}

__global__ void __launch_bounds__(256) group_points_grad_kernel(int B, int M, int C, int nsample,
                                                                const float* __restrict__ grad_out,
                                                                const int* __restrict__ idx,
                                                                const int* __restrict__ idx_batch_cnt,
                                                                const int* __restrict__ features_batch_cnt,
                                                                float* __restrict__ grad_features) {
    extern __shared__ float tile[];   // [GP][nsample][C + 1]
    const int pt0 = blockIdx.x * GP;
    const int stride = C + 1;
    for (int q = threadIdx.x; q < GP * C * nsample; q += blockDim.x) {
        const int s = q % nsample, c = (q / nsample) % C, p = q / (nsample * C);
        const int pt = pt0 + p;
        if (pt < M) tile[(p * nsample + s) * stride + c] = __ldg(grad_out + ((size_t)pt * C + c) * nsample + s);
    }
    __syncthreads();
    for (int q = threadIdx.x; q < GP * nsample * C; q += blockDim.x) {
        const int c = q % C, s = (q / C) % nsample, p = q / (C * nsample);
        const int pt = pt0 + p;
        if (pt < M) {
            const int start = batch_start(pt, B, idx_batch_cnt, features_batch_cnt);
            const int row = start + idx[(size_t)pt * nsample + s];
            atomicAdd(grad_features + (size_t)row * C + c, tile[(p * nsample + s) * stride + c]);
        }
    }
}

}  // namespace
}  // namespace vc

using namespace vc;

extern "C" int vc_voxel_query(int M, int R1, int R2, int R3, int nsample, float radius, int z_range, int y_range, int x_range,
                              const float* new_xyz, const float* xyz, const int32_t* new_coords,
                              const int32_t* point_indices, int32_t* idx, unsigned char* empty_mask, vc_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    VC_CHECK_ARG(M >= 0 && R1 > 0 && R2 > 0 && R3 > 0 && nsample > 0 && z_range >= 0 && y_range >= 0 && x_range >= 0,
                 "bad voxel_query arguments");
    if (M == 0) return VC_OK;
    VC_CHECK_ARG(new_xyz && xyz && new_coords && point_indices && idx && empty_mask, "null pointer");
    voxel_query_kernel<<<cdiv(M, QWARPS), QWARPS * 32, 0, stream>>>(M, R1, R2, R3, nsample, radius, z_range, y_range, x_range,
                                                                   new_xyz, xyz, new_coords, point_indices, idx, empty_mask);
    VC_LAUNCH_CHECK();
    return VC_OK;
}

static int group_smem(int C, int nsample, size_t* bytes) {
    *bytes = (size_t)GP * nsample * (C + 1) * sizeof(float);
    if (*bytes > 200 * 1024) {
        set_error("group_points: tile of %d samples x %d channels does not fit shared memory", nsample, C);
        return VC_ERR_UNSUPPORTED;
    }
    return VC_OK;
}

extern "C" int vc_group_points(int B, int M, int C, int nsample, const float* features, const int32_t* features_batch_cnt,
                               const int32_t* idx, const int32_t* idx_batch_cnt, float* out, vc_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    VC_CHECK_ARG(B > 0 && M >= 0 && C > 0 && nsample > 0, "bad group_points arguments");
    if (M == 0) return VC_OK;
    VC_CHECK_ARG(features && features_batch_cnt && idx && idx_batch_cnt && out, "null pointer");
    size_t smem;
    int rc = group_smem(C, nsample, &smem);
    if (rc) return rc;
    VC_CUDA(cudaFuncSetAttribute(group_points_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    group_points_kernel<<<cdiv(M, GP), 256, smem, stream>>>(B, M, C, nsample, features, features_batch_cnt, idx, idx_batch_cnt, out);
    VC_LAUNCH_CHECK();
    return VC_OK;
}

extern "C" int vc_group_points_grad(int B, int M, int C, int N, int nsample, const float* grad_out, const int32_t* idx,
                                    const int32_t* idx_batch_cnt, const int32_t* features_batch_cnt, float* grad_features,
                                    vc_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    VC_CHECK_ARG(B > 0 && M >= 0 && C > 0 && N >= 0 && nsample > 0, "bad group_points_grad arguments");
    if (M == 0) return VC_OK;
    VC_CHECK_ARG(grad_out && idx && idx_batch_cnt && features_batch_cnt && grad_features, "null pointer");
    size_t smem;
    int rc = group_smem(C, nsample, &smem);
    if (rc) return rc;
    VC_CUDA(cudaFuncSetAttribute(group_points_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    group_points_grad_kernel<<<cdiv(M, GP), 256, smem, stream>>>(B, M, C, nsample, grad_out, idx, idx_batch_cnt,
                                                               features_batch_cnt, grad_features);
    VC_LAUNCH_CHECK();
    return VC_OK;
}
