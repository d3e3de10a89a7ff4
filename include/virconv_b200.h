/*
 * virconv_b200 — C ABI of the B200 (sm_100a) sparse-convolution hot path of VirConv.
 *
 * The reference (hailanyi/VirConv) has no FFI of its own for this path: every sparse operator is a
 * call into the third-party Python package spconv 2.1 (`import spconv.pytorch as spconv`,
 * pcdet/utils/spconv_utils.py:33-36).  Each entry point below therefore cites the reference CALL
 * SITE (file:line under /root/reference) whose spconv work it replaces; the Python classes in
 * `virconv_b200/spconv_compat.py` re-create the spconv.pytorch surface on top of these calls
 * (binding shown in INTEGRATION.md).
 *
 * Contract:
 *   - plain pointers + sizes; all pointers are DEVICE pointers unless marked "host";
 *   - the caller (PyTorch) owns every input, output and workspace buffer; the library never
 *     allocates, frees or retains a pointer past the call;
 *   - all work is enqueued on the cudaStream_t passed as `stream`; no hidden device sync,
 *     re-entrant, one host thread per device;
 *   - return 0 on success, <0 on error (message: vc_last_error(), thread-local); no exceptions,
 *     no exit();
 *   - row-major everywhere: features [N, C]; indices [N, 1+ndim] int32 = (batch, z, y, x) or
 *     (batch, u, v); conv weight in spconv-2.x layout [C_out, K, C_in] with K = prod(kernel) and
 *     offsets numbered z-major (k = (kz*Ky+ky)*Kx+kx), see detector3d_template.py:358-370;
 *   - a rulebook is a NEIGHBOUR TABLE nbr[K, N_out] int32: input row feeding output row o through
 *     kernel offset k, or -1 (canonical form; oracle/rulebook.py is the definition).
 */
#ifndef VIRCONV_B200_H
#define VIRCONV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VC_OK 0
#define VC_ERR_INVALID (-1)      /* bad argument */
#define VC_ERR_CUDA (-2)         /* CUDA runtime error (launch / config) */
#define VC_ERR_WORKSPACE (-3)    /* workspace too small */
#define VC_ERR_UNSUPPORTED (-4)  /* channel count / ndim not compiled in */
#define VC_ERR_PIPELINE (-5)     /* a tensor-core kernel's mbarrier wait timed out earlier (err_flag set): results invalid */

#define VC_MAX_NDIM 3
#define VC_TILE_ROWS 128         /* output rows per CTA tile in the conv kernels (BN partial granularity) */

typedef void* vc_stream_t;       /* cudaStream_t */

int vc_version(void);
const char* vc_last_error(void);
/* number of kernels this library has launched so far in this process (bench.py's gpu_launches) */
long long vc_launch_count(void);
/* Programmatic dependent launch for the kernels of the conv / BN chain (default on): each starts while its predecessor
 * drains and blocks in `griddepcontrol.wait` before touching dependent data.  0 = plain stream-ordered launches. */
int vc_set_pdl(int enable);
/* BatchNorm backward as ONE cooperative launch (reduction, grid-wide barrier, apply; default on) or as two launches (0). */
int vc_set_bn_fused(int enable);
/* Tensor-core conv forward / gather-dgrad kernel: 1 (default) = persistent kernel with a deep cp.async operand ring
 * (csrc/conv_tc2.cu; C in {8,16,32,64}); 0 = the round-1 one-tile-per-CTA kernel (csrc/conv_tc.cu; C in {16,32,64}).
 * A/B measurements only — weight images built under one variant are not valid under the other. */
int vc_set_tc_variant(int variant);
/* A/B aid for the persistent tensor-core conv kernel (conv_tc2.cu): CTAs per SM, 0 = automatic (two when each still gets a
 * ring of >= 2 stages in half the shared memory), 1 or 2 = forced. */
int vc_conv_tc2_config(int ctas_per_sm);
/* The plan executor's weight-gradient kernel: variant 1 (default) = wgrad_tc3.cu (half-tile ring stages, <= 111 KB shared memory
 * and <= 256 TMEM columns per CTA: shares an SM with the dgrad kernels of the main stream), 0 = wgrad_tc2.cu (one CTA per SM,
 * 32 KB stages); max_ctas caps the CTAs of variant 1 (0 = half the SMs, the measured optimum). */
int vc_conv_wgrad_tc3_config(int variant, int max_ctas);
/* CTAs of the persistent tensor-core wgrad kernel (csrc/wgrad_tc2.cu): 0 (default) = one per SM.  Every CTA adds one
 * [K, C_in, C_out] partial with L2 vector reductions, so fewer CTAs trade main-loop parallelism for reduction traffic. */
int vc_conv_wgrad_tc2_config(int max_ctas);

/* ------------------------------------------------------------------------------------------------
 * Rulebooks.  Replaces spconv `ops.get_indice_pairs`, reached from every conv call site:
 * spconv_backbone.py:89 (SubMConv3d), :113 (SubMConv2d on image indices built at :217-222),
 * :92-93 and :563-564 (SparseConv3d).
 * ---------------------------------------------------------------------------------------------- */

/* Submanifold: output rows == input rows.  Hash table of linearised coordinates (lowest row wins on
 * duplicate coordinates), 1 thread per (row, offset) probe, per-offset pair counts through a
 * shared-memory histogram.  ws >= vc_subm_rulebook_ws_bytes(n). */
size_t vc_subm_rulebook_ws_bytes(int n);
int vc_subm_rulebook(const int32_t* indices, int n, int ndim, int batch_size,
                     const int32_t* spatial_shape /*host[ndim]*/, const int32_t* ksize /*host[ndim]*/,
                     const int32_t* dilation /*host[ndim]*/, int32_t* nbr /*[K,n]*/,
                     int32_t* pair_num /*[K], may be NULL*/, void* ws, size_t ws_bytes, vc_stream_t stream);

/* Regular (strided) sparse conv, two phases because the output row count is data dependent.
 *   count: marks every reachable output cell in a bitmap over the OUTPUT grid and ranks it (prefix sum
 *          of popcounts) -> *n_out_dev; also writes out_shape (host).  Output rows are therefore
 *          ordered by ascending linear index, batch most significant (voxel_query_utils.py:85-91 needs
 *          batch-contiguous rows).
 *   fill : after the caller has read n_out and allocated, emits out_indices, nbr_fwd[K,n_out],
 *          nbr_bwd[K,n] (output row per input row; the dgrad table) and pair counts.
 * The same `ws` (>= vc_conv_rulebook_ws_bytes) must be passed to both calls, untouched in between. */
size_t vc_conv_rulebook_ws_bytes(int ndim, int batch_size, const int32_t* out_shape /*host[ndim]*/);
int vc_conv_out_shape(int ndim, const int32_t* spatial_shape, const int32_t* ksize, const int32_t* stride,
                      const int32_t* padding, const int32_t* dilation, int32_t* out_shape /*host[ndim]*/);
int vc_conv_rulebook_count(const int32_t* indices, int n, int ndim, int batch_size,
                           const int32_t* spatial_shape, const int32_t* ksize, const int32_t* stride,
                           const int32_t* padding, const int32_t* dilation,
                           int32_t* n_out_dev /*device int32[1]*/, void* ws, size_t ws_bytes, vc_stream_t stream);
int vc_conv_rulebook_fill(const int32_t* indices, int n, int ndim, int batch_size,
                          const int32_t* spatial_shape, const int32_t* ksize, const int32_t* stride,
                          const int32_t* padding, const int32_t* dilation, int n_out,
                          int32_t* out_indices /*[n_out,1+ndim]*/, int32_t* nbr_fwd /*[K,n_out]*/,
                          int32_t* nbr_bwd /*[K,n]*/, int32_t* pair_num /*[K], may be NULL*/,
                          void* ws, size_t ws_bytes, vc_stream_t stream);

/* spconv-style `indice_pairs [2,K,n]` (-1 padded) + `indice_pair_num [K]` from a neighbour table, each
 * offset's pairs ordered by ascending output row.  Only for API parity / tests; the conv kernels
 * consume the neighbour table directly. */
int vc_pairs_from_nbr(const int32_t* nbr, int K, int n, int32_t* pairs /*[2,K,n]*/,
                      int32_t* pair_num /*[K]*/, vc_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Convolution (output-stationary implicit gather-GEMM).  Replaces spconv `ops.indice_conv` and its
 * autograd backward for the same call sites.  fp32 in / fp32 accumulate / fp32 out.
 * Supported channel counts: cin, cout in {8,16,32,64}.
 * fwd / dgrad take a workspace of vc_conv_ws_bytes(cin, cout, K) for the re-laid-out weights
 * ([K, C_in, C_out] for forward, [K, C_out, C_in] for dgrad), rebuilt on every call.
 * ---------------------------------------------------------------------------------------------- */
size_t vc_conv_ws_bytes(int cin, int cout, int K);

/* out[o, co] = sum_k sum_ci in[nbr[k,o], ci] * w[co, k, ci].
 * bn_sums (may be NULL): [2, cout] float64 accumulator the kernel ADDS the output's channel sums (sum x,
 * sum x^2) to, for the BatchNorm1d that follows every conv (spconv_backbone.py:101-105); caller zeroes it. */
int vc_conv_fwd_f32(const float* in, const float* w, const int32_t* nbr, float* out, int n_out, int cin,
                    int cout, int K, double* bn_sums, void* ws, size_t ws_bytes, vc_stream_t stream);

/* din[i, ci] = sum_k sum_co dout[nbr_t[k,i], co] * w[co, kk, ci],  kk = mirror ? K-1-k : k.
 * nbr_t is the table indexed by INPUT row: the rulebook's nbr_bwd for a regular conv, or the
 * submanifold table itself with mirror=1 (valid when coordinates are unique). */
int vc_conv_dgrad_f32(const float* dout, const float* w, const int32_t* nbr_t, float* din, int n_in, int cin,
                      int cout, int K, int mirror, void* ws, size_t ws_bytes, vc_stream_t stream);

/* din[nbr[k,o], :] += dout[o, :] @ w[:, k, :]  with float atomics; din must be zeroed by the caller.
 * Needed for the 2-D image branch, whose table is many-to-one (duplicate pixel coordinates). */
int vc_conv_dgrad_scatter_f32(const float* dout, const float* w, const int32_t* nbr, float* din, int n_out,
                              int cin, int cout, int K, void* ws, size_t ws_bytes, vc_stream_t stream);

/* dw[co, k, ci] = sum_o in[nbr[k,o], ci] * dout[o, co].  Deterministic two-pass reduction.
 * ws >= vc_conv_wgrad_ws_bytes(n_out, cin, cout, K). */
size_t vc_conv_wgrad_ws_bytes(int n_out, int cin, int cout, int K);
int vc_conv_wgrad_f32(const float* in, const float* dout, const int32_t* nbr, float* dw, int n_out, int cin,
                      int cout, int K, void* ws, size_t ws_bytes, vc_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * bf16 tensor-core variants of forward / dgrad (tcgen05.mma, fp32 accumulators in TMEM).  Same tables,
 * same tiling, fp32 weights (converted per call) and fp32 outputs; the GATHERED operand (input features for
 * forward, output gradients for dgrad) is bf16 [N, C] row major — vc_cast_f32_bf16 makes it.
 * cin, cout in {16,32,64}.  ws >= vc_conv_tc_ws_bytes(cin, cout, K).  err_flag (device int32, may be NULL)
 * is set to 1 if a pipeline wait ever times out (bounded spin instead of a hang); results are then invalid.
 * ---------------------------------------------------------------------------------------------- */
int vc_cast_f32_bf16(const float* in, void* out_bf16, long long n_elements /* multiple of 4 */, vc_stream_t stream);
size_t vc_conv_tc_ws_bytes(int cin, int cout, int K);
int vc_conv_fwd_tc(const void* in_bf16, const float* w, const int32_t* nbr, float* out, int n_out, int cin, int cout,
                   int K, double* bn_sums, void* ws, size_t ws_bytes, int32_t* err_flag, vc_stream_t stream);
/* din[nbr[k,o], :] += dout[o, :] @ w[:, k, :] on tensor cores, for many-to-one tables (image branch); din zeroed by the
 * caller; the dout tile is contiguous, the result is scattered with 16-byte vector reductions. */
int vc_conv_dgrad_scatter_tc(const void* dout_bf16, const float* w, const int32_t* nbr, float* din, int n_out, int cin,
                             int cout, int K, void* ws, size_t ws_bytes, int32_t* err_flag, vc_stream_t stream);
int vc_conv_dgrad_tc(const void* dout_bf16, const float* w, const int32_t* nbr_t, float* din, int n_in, int cin,
                     int cout, int K, int mirror, void* ws, size_t ws_bytes, int32_t* err_flag, vc_stream_t stream);
/* dw[co,k,ci] = sum_o in[nbr[k,o],ci] * dout[o,co] on tensor cores: both operands bf16 (MN-major UMMA), 128/cin
 * kernel offsets stacked along the MMA M dimension, accumulators resident in TMEM across all tiles of a
 * persistent CTA, fixed-order reduction of the per-CTA partials.  ws >= vc_conv_wgrad_tc_ws_bytes(...). */
size_t vc_conv_wgrad_tc_ws_bytes(int n_out, int cin, int cout, int K);
/* Process-wide launch shape of vc_conv_wgrad_tc: at most max_ctas persistent CTAs (default 148 = one per SM), each asking
 * for at least smem_floor_bytes of dynamic shared memory (default 0).  The plan executor's backward uses (fewer CTAs,
 * a floor no gather-GEMM CTA fits beside) to run wgrad on its own stream next to the dgrad chain.  Affects
 * vc_conv_wgrad_tc_ws_bytes. */
int vc_conv_wgrad_tc_config(int max_ctas, int smem_floor_bytes);
int vc_conv_wgrad_tc(const void* in_bf16, const void* dout_bf16, const int32_t* nbr, float* dw, int n_out, int cin,
                     int cout, int K, void* ws, size_t ws_bytes, int32_t* err_flag, vc_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * BatchNorm1d(eps, momentum) + ReLU over the active rows.  Replaces the `norm_fn(out_channels)`,
 * `nn.ReLU()` members of every `spconv.SparseSequential` (spconv_backbone.py:101-105, :160, :561-567).
 * ---------------------------------------------------------------------------------------------- */

/* y = relu(BN(x)) in one pass.  training != 0: batch statistics from `sums` (the [2,c] float64 accumulator the conv
 * filled): mean, biased var -> scale = gamma*invstd, shift = beta - mean*scale; running_mean / running_var updated in
 * place with the unbiased variance and *num_batches_tracked (int64, may be NULL) incremented — torch.nn.BatchNorm1d
 * semantics.  training == 0: statistics from the running buffers.  stats_out [4,c] = scale, shift, mean, invstd is
 * kept for backward.  y_bf16 (may be NULL): bf16 shadow of y = the next conv's tensor-core operand.  In place
 * (y == x) allowed. */
int vc_bn_apply_relu_f32(const float* x, const double* sums, int n_rows, int c, const float* gamma, const float* beta,
                         float* running_mean, float* running_var, long long* num_batches_tracked, float momentum,
                         float eps, int training, float* y, void* y_bf16, float* stats_out, int relu, vc_stream_t stream);
/* Backward of y = relu(gamma*(x-mean)*invstd + beta), stats = the forward's stats_out:
 *   g = dy * (y > 0); dbeta = sum g; dgamma = sum g*xhat;
 *   train: dx = gamma*invstd*(g - dbeta/n - xhat*dgamma/n);   eval: dx = gamma*invstd*g.
 * bsums: [2,c] float64 scratch accumulator, zeroed by the caller.  dx_bf16 (may be NULL): bf16 shadow of dx. */
int vc_bn_relu_bwd_f32(const float* dy, const float* x, const float* y, const float* gamma, const float* stats,
                       float* dx, void* dx_bf16, float* dgamma, float* dbeta, int n, int c, int training,
                       double* bsums, vc_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Voxel index -> image pixel index.  Replaces `index2uv` + `index2points` + the per-sample
 * `X_TRANS.backward_with_param` / `Calibration.lidar_to_rect_cuda` / `rect_to_img_cuda` loop
 * (spconv_backbone.py:8-24, :54-83; X_transform.py:139-154; calibration_kitti.py:120-153).
 * params: [batch_size, 28] float32 per sample:
 *   [0..11]  M = V2C^T @ R0^T  (4x3, row major)        [12..19] first two columns of P2^T (4x2)
 *   [20] has_transform  [21] scale  [22] flip  [23] cos(-rot)  [24] sin(-rot)  [25..27] unused
 * grid: 6 floats (vx, vy, vz, min_x, min_y, min_z) = voxel size * stride and range min + half voxel.
 * uv_out [n,3] int32 = (batch, clamp(u,0,u_max-1)/stride, clamp(v,0,v_max-1)/stride).
 * ---------------------------------------------------------------------------------------------- */
int vc_index2uv(const int32_t* indices /*[n,4] b,z,y,x*/, int n, int batch_size, const float* params,
                const float* grid /*host[6]*/, int stride, int u_max, int v_max, int32_t* uv_out,
                vc_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Hash-grid voxelisation + MeanVFE.  Replaces `VoxelGeneratorWrapper.generate` -> spconv
 * `Point2VoxelCPU3d.point_to_voxel` (pcdet/datasets/processor/data_processor.py:43-59, called :156,:173) and
 * `MeanVFE.forward` (pcdet/models/backbones_3d/vfe/mean_vfe.py:39-58) with identical results.
 * points [n, 1+c] float32 rows (batch, x, y, z, features...), samples contiguous in ascending batch order (what
 * `collate_batch` produces, dataset.py:349-353).  Per sample: voxels in order of first appearance, at most max_voxels,
 * at most max_points (<= 8) points per voxel in point order.  Outputs (capacity batch_size*max_voxels rows):
 * out_features [*, c] = per-voxel mean (last channel = max over the zero-padded slots if vfe_max_last),
 * out_coords [*, 4] = (b, z, y, x), out_num [*], optional out_voxels [*, max_points, c] (caller zeroes), *n_out_dev.
 * pc_range host[6] = (x0,y0,z0,x1,y1,z1), voxel_size host[3].  ws >= vc_voxelize_ws_bytes(...).
 * ---------------------------------------------------------------------------------------------- */
size_t vc_voxelize_ws_bytes(int n_points, int batch_size, int max_points);
int vc_voxelize_mean(const float* points, int n_points, int c, int batch_size, const float* pc_range,
                     const float* voxel_size, int max_points, int max_voxels, int vfe_max_last, float* out_features,
                     int32_t* out_coords, int32_t* out_num, float* out_voxels /*may be NULL*/, int32_t* n_out_dev,
                     void* ws, size_t ws_bytes, vc_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * StVD input point discard.  Replaces `DatasetTemplate.partition` + `DatasetTemplate.input_point_discard`
 * (pcdet/datasets/dataset.py:120-189; applied to the virtual points of every frame at :275-290): bin_num range bins
 * along x of width max_dis/bin_num (last bin open-ended, x < 0 / NaN in no bin), emitted far -> near, the nearest bins
 * randomly subsampled.  The random draws stay on the host so that they come from the same numpy generator, in the same
 * order, as the reference's (virconv_b200/preprocess.py); the device does the order-preserving partition and the gather.
 *   vc_stvd_partition: points [n, c] fp32 (x in column 0) -> totals_dev[16] = points per bin, and in ws the per-bin point
 *                      lists (original order inside a bin).
 *   vc_stvd_gather   : segs host [n_seg][4] = (bin, out_base, count, sel_base): output rows [out_base, out_base+count)
 *                      are the bin's points with in-bin ranks sel_dev[sel_base + r] (sel_base < 0: rank r itself).
 * ws >= vc_stvd_ws_bytes(n), the same buffer for both calls.
 * ---------------------------------------------------------------------------------------------- */
size_t vc_stvd_ws_bytes(int n_points);
int vc_stvd_partition(const float* points, int n, int c, int bin_num, double max_dis, int32_t* totals_dev /*[16]*/,
                      void* ws, size_t ws_bytes, vc_stream_t stream);
int vc_stvd_gather(const float* points, int n, int c, const int32_t* segs /*host*/, int n_seg, const int32_t* sel_dev,
                   float* out /*[n_out, c]*/, int n_out, void* ws, size_t ws_bytes, vc_stream_t stream);

/* `generate_voxel2pinds` (pcdet/utils/spconv_utils.py:13-21): dense [B, *spatial_shape] int32 map voxel -> row, -1 where
 * empty (the RoI head's look-up table, ted_head.py:527,625). */
int vc_voxel2pinds(const int32_t* indices, int n, int ndim, int batch_size, const int32_t* spatial_shape, int32_t* out,
                   vc_stream_t stream);

/* `SparseConvTensor.dense()` (height_compression.py:29): out [B, C, *shape] must be zeroed by the caller. */
int vc_dense_f32(const float* features, const int32_t* indices, int n, int c, int ndim, int batch_size,
                 const int32_t* spatial_shape, float* out, vc_stream_t stream);
/* Gradient of dense(): dfeatures[n, c] = dout[b, c, coords]. */
int vc_dense_bwd_f32(const float* dout, const int32_t* indices, int n, int c, int ndim, int batch_size,
                     const int32_t* spatial_shape, float* dfeatures, vc_stream_t stream);

/* NRConv channel concat (spconv_backbone.py:227): out[:, :ca] = a, out[:, ca:] = b, plus an optional bf16 shadow. */
int vc_cat2_f32(const float* a, const float* b, float* out, void* out_bf16 /*may be NULL*/, int n, int ca, int cb,
                vc_stream_t stream);

/* Row gather for StVD layer discard (spconv_backbone.py:134-147): out[r] = in[rows[r]]. */
int vc_gather_rows(const void* in, const int32_t* rows, void* out, int n_rows, int row_bytes, vc_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Voxel-RoI pooling primitives: the stacked pointnet2 ops the RoI head runs on x_conv3 / x_conv4 (ted_head.py:496-552).
 * Same arguments, layouts and results as the reference launchers they replace:
 *   vc_voxel_query       <- voxel_query_kernel_launcher_stack (pointnet2_stack/src/voxel_query_gpu.cu:92-113) plus the
 *                           `idx[empty_ball_mask] = 0` of VoxelQuery.forward (voxel_query_utils.py:38-39): idx [M, nsample]
 *                           = the first nsample rows found scanning the (2r+1)^3 cells around new_coords (b,z,y,x) in
 *                           z-major order whose centre xyz lies within `radius` of new_xyz, padded with the first hit;
 *                           empty_mask [M] uint8 = 1 and idx row = 0 where nothing was found.  point_indices: dense
 *                           [B, R1, R2, R3] voxel -> row map (vc_voxel2pinds).
 *   vc_group_points      <- group_points_kernel_launcher_stack (group_points_gpu.cu:106-125): out [M, C, nsample] =
 *                           features[start(batch of m) + idx[m, s], c], idx batch-local.
 *   vc_group_points_grad <- group_points_grad_kernel_launcher_stack (:47-68): grad_features (zeroed by the caller)
 *                           += grad_out, float atomics.
 * ---------------------------------------------------------------------------------------------- */
int vc_voxel_query(int M, int R1, int R2, int R3, int nsample, float radius, int z_range, int y_range, int x_range,
                   const float* new_xyz, const float* xyz, const int32_t* new_coords, const int32_t* point_indices,
                   int32_t* idx, unsigned char* empty_mask, vc_stream_t stream);
int vc_group_points(int B, int M, int C, int nsample, const float* features, const int32_t* features_batch_cnt,
                    const int32_t* idx, const int32_t* idx_batch_cnt, float* out, vc_stream_t stream);
int vc_group_points_grad(int B, int M, int C, int N, int nsample, const float* grad_out, const int32_t* idx,
                         const int32_t* idx_batch_cnt, const int32_t* features_batch_cnt, float* grad_features,
                         vc_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Plan executor: a whole backbone forward / backward in ONE call.  Replaces the Python layer loop of
 * `VirConvL8x.forward` (spconv_backbone.py:609-699), `NRConvBlock.forward` (:207-229) and the two streams of
 * `VirConv8x.forward` (:339-535) together with autograd's reverse walk over them: the host side of ~100 operator
 * calls per step becomes native.  The caller describes the layer graph once as a PLAN:
 *   ops_i [n_ops, 24] int32, ops_f [n_ops, 8] float32, one row per op, executed in order:
 *     [0] kind  1 SUBM_RB   index set a            -> rulebook c            ([21] = 1 if coordinates are unique)
 *               2 CONV_RB   index set a            -> index set b, rulebook c   (host reads the row count once)
 *               3 INDEX2UV  3-D index set a        -> 2-D index set b       ([6],[7] = image shape, [21] stride,
 *                                                                            [22] u_max, [23] v_max, ops_f[0..5] = grid)
 *               4 CBR       features a, rulebook c -> features b            conv + BatchNorm1d + ReLU of layer [20]
 *                                                                            ([18] cin, [19] cout, [21] = input needs grad)
 *               5 CAT       features a, features b -> features c            channel concat
 *     [1] stream (0 main: feature ops, 1 side: index ops)   [2..4] a, b, c   [5] ndim   [6..8] kernel size
 *     [9..11] stride   [12..14] padding   [15..17] dilation
 *   layer_ptrs [n_layers, 9] uint64: weight, gamma, beta, running_mean, running_var, num_batches_tracked, then the
 *   caller's gradient buffers d_weight, d_gamma, d_beta (written by vc_exec_backward); layer_f [n_layers, 2] = eps, momentum.
 *   Feature slot 0 / index set 0 are the network input (feats0 [n0, c0] fp32, idx0 [n0, 4] int32, shape0 host[3]).
 * Memory: every activation, rulebook and scratch buffer is carved out of `arena` (bump allocation, nothing reused
 * inside a step; VC_ERR_WORKSPACE if it is too small).  `state` (>= vc_exec_state_bytes(), host) receives the buffer
 * addresses and row counts that vc_exec_backward and vc_exec_query read; the library keeps nothing else between calls
 * (apart from a pool of CUDA events for cross-stream ordering).  `pinned_host`: >= 16 int32 of page-locked host memory
 * for the row-count read-backs.  precision 0 = fp32 kernels, 1 = bf16 tensor-core kernels where channels allow.
 * Index ops run on side_stream (may be NULL = single stream); on return every later use of main_stream is ordered
 * after all of the step's work.  side_waits_main != 0: the side stream first waits for everything already queued on
 * main_stream (always safe).  0: the caller guarantees idx0 / proj_params are valid in side_stream order and that the
 * arena may be written from side_stream right away (e.g. it was allocated there) — the index ops of this step then
 * overlap the previous step's backward.
 * Static mode (n0_dev != NULL): nothing in the call reads a device value on the host, so the whole forward (and the
 * backward that follows it) can be captured into a CUDA graph (SURVEY §8b "CUDA-graph-capturable").  n0 is then the
 * CAPACITY of feats0 / idx0, *n0_dev the number of valid rows; caps[i] (host, indexed by index-set id) is the row
 * capacity of the strided convs' output sets; every data-dependent row count stays in device memory
 * (vc_exec_query(state, 2, id)[6] = its address), published buffers have capacity rows with a zero / -1 tail, and a row
 * count that does not fit its capacity is clamped and reported as max() into *overflow_flag (device int32; zero it
 * before the step, re-run with larger capacities when it is non-zero afterwards).  Use side_waits_main = 1 under capture.
 * ---------------------------------------------------------------------------------------------- */
size_t vc_exec_state_bytes(void);
int vc_exec_forward(const int32_t* ops_i, const float* ops_f, int n_ops, const uint64_t* layer_ptrs, const float* layer_f,
                    int n_layers, const float* feats0, int c0, const int32_t* idx0, int n0, const int32_t* shape0 /*host[3]*/,
                    int batch_size, const float* proj_params /*device [B,28], may be NULL without INDEX2UV ops*/,
                    int training, int precision, int want_pair_num, void* arena, size_t arena_bytes,
                    int32_t* pinned_host, int32_t* err_flag, void* state, size_t state_bytes, vc_stream_t main_stream,
                    vc_stream_t side_stream, int side_waits_main, const int32_t* caps /*host, may be NULL*/,
                    const int32_t* n0_dev /*device, NULL = exact mode*/, int32_t* overflow_flag /*device*/,
                    vc_stream_t side2_stream /*may be NULL: ops planned for stream 2 (image branch) then share side_stream*/);
/* Reverse walk.  pub_slots [n_pub]: feature slots whose gradients come from outside (the published tensors), ext_grads
 * [n_pub]: device pointers to those gradients ([rows, c] fp32 contiguous, read only) or 0.  Writes d_weight / d_gamma /
 * d_beta of every layer (zeros where nothing flowed back).  Same arena as the forward (it continues allocating).
 * wgrad_stream (may be NULL): the weight-gradient kernels go there, concurrent with the BN-backward / dgrad chain on
 * `stream`; `stream` is joined with it before the call returns. */
int vc_exec_backward(const int32_t* ops_i, const float* ops_f, int n_ops, const uint64_t* layer_ptrs, const float* layer_f,
                     int n_layers, const int32_t* pub_slots, const uint64_t* ext_grads, int n_pub, void* arena,
                     size_t arena_bytes, int32_t* err_flag, void* state, vc_stream_t stream, vc_stream_t wgrad_stream);
/* Read a state blob.  what 0: out[0] = arena bytes in use;  1 (feature slot id): f32 ptr, bf16 ptr, rows, c;
 * 2 (index set): idx ptr, n, ndim, shape[3];  3 (rulebook): nbr, nbr_bwd, pair_num ptrs, K, n_in, n_out, subm, unique;
 * 4 (layer): x ptr, y ptr, stats ptr, use_tc, rulebook id, in slot, out slot.  out: >= 8 int64 (host). */
int vc_exec_query(const void* state, int what, int id, long long* out);
/* Optional per-launch timing of the conv kernels (bench.py's roofline pass): vc_exec_timing(1) clears and arms it,
 * vc_exec_timing_read synchronises and returns up to max_records (ms, [kind, layer]) records; kind 0 fwd f32, 1 fwd tc,
 * 2 dgrad f32, 3 dgrad tc, 4 dgrad scatter, 5 wgrad f32, 6 wgrad tc. */
int vc_exec_timing(int enable);
int vc_exec_timing_read(float* ms_out, int32_t* kind_layer_out, int max_records);

/* ---- gradient all-reduce over NVLink peer memory (replaces DistributedDataParallel's NCCL all-reduce of
 * tools/train.py:140-141 for the backbone's one flat 1.7 MB gradient buffer) ----
 * Every rank owns a SYMMETRIC fp32 buffer of 2 * n_pad floats and a symmetric array of vc_allreduce_peer_flag_words(world)
 * zero-initialised uint32 flags, both mapped into every peer (torch.distributed._symmetric_memory); peer_bufs / peer_flags are
 * HOST arrays of `world` device addresses: where rank r's buffer / flags are mapped in THIS process.  One call per step on every
 * rank with the same, strictly increasing `epoch` (first call: 1): grads[i] <- scale * sum over ranks of grads[i], in place.
 * A peer that does not show up within 4 s sets *err_flag (0x400 + peer) instead of hanging the GPU. */
int vc_allreduce_peer_flag_words(int world);
int vc_allreduce_peer_f32(const uint64_t* peer_bufs, const uint64_t* peer_flags, int rank, int world, float* grads, long long n,
                          long long n_pad, unsigned epoch, float scale, int32_t* err_flag, vc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VIRCONV_B200_H */
