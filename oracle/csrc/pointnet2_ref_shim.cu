// TEST INFRASTRUCTURE (see oracle/__init__.py).  C-linkage doors onto the REFERENCE's own CUDA launchers, which
// oracle/ref_build.py compiles from the sources where they lie under /root/reference (never copied) into
// oracle/_ref/libpointnet2_stack_ref.so.  Declarations restate pointnet2_stack/src/voxel_query_gpu.h:16-18 and
// group_points_gpu.h:20-21,28-29.  The launchers use the legacy default stream.
void voxel_query_kernel_launcher_stack(int M, int R1, int R2, int R3, int nsample, float radius, int z_range, int y_range,
                                       int x_range, const float* new_xyz, const float* xyz, const int* new_coords,
                                       const int* point_indices, int* idx);
void group_points_kernel_launcher_stack(int B, int M, int C, int nsample, const float* features, const int* features_batch_cnt,
                                        const int* idx, const int* idx_batch_cnt, float* out);
void group_points_grad_kernel_launcher_stack(int B, int M, int C, int N, int nsample, const float* grad_out, const int* idx,
                                             const int* idx_batch_cnt, const int* features_batch_cnt, float* grad_features);

extern "C" void ref_voxel_query(int M, int R1, int R2, int R3, int nsample, float radius, int z_range, int y_range, int x_range,
                                const float* new_xyz, const float* xyz, const int* new_coords, const int* point_indices,
                                int* idx) {
    voxel_query_kernel_launcher_stack(M, R1, R2, R3, nsample, radius, z_range, y_range, x_range, new_xyz, xyz, new_coords,
                                      point_indices, idx);
}
extern "C" void ref_group_points(int B, int M, int C, int nsample, const float* features, const int* features_batch_cnt,
                                 const int* idx, const int* idx_batch_cnt, float* out) {
    group_points_kernel_launcher_stack(B, M, C, nsample, features, features_batch_cnt, idx, idx_batch_cnt, out);
}
extern "C" void ref_group_points_grad(int B, int M, int C, int N, int nsample, const float* grad_out, const int* idx,
                                      const int* idx_batch_cnt, const int* features_batch_cnt, float* grad_features) {
    group_points_grad_kernel_launcher_stack(B, M, C, N, nsample, grad_out, idx, idx_batch_cnt, features_batch_cnt, grad_features);
}
