// legacy_shim.cu -- extern "C" doors onto the REFERENCE's own launchers.
//
// TEST / BASELINE INFRASTRUCTURE ONLY.  This file contains no algorithm: it declares the
// launcher prototypes that the reference's unmodified .cu files define
//   pointrcnn/pointnet2_lib/pointnet2/src/{sampling,ball_query,group_points,interpolate}_gpu.cu
//   pointrcnn/lib/utils/iou3d/src/iou3d_kernel.cu, pointrcnn/lib/utils/roipool3d/src/roipool3d_kernel.cu
// and forwards raw pointers to them.  oracle/build_ref.py compiles those sources where they
// lie under /root/reference together with this shim into oracle/_ref/libpn2_legacy.so
// (git-ignored, shipped to the GPU box).  The reference's .cpp wrappers cannot be built on
// torch 2.11 (THC removed), which is why the door is at the launcher level.
#include <cuda_runtime.h>

void furthest_point_sampling_kernel_launcher(int b, int n, int m, const float *dataset, float *temp, int *idxs, cudaStream_t stream);
void gather_points_kernel_launcher_fast(int b, int c, int n, int npoints, const float *points, const int *idx, float *out, cudaStream_t stream);
void gather_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints, const float *grad_out, const int *idx, float *grad_points, cudaStream_t stream);
void ball_query_kernel_launcher_fast(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz, int *idx, cudaStream_t stream);
void group_points_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx, float *out, cudaStream_t stream);
void group_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample, const float *grad_out, const int *idx, float *grad_points, cudaStream_t stream);
void three_nn_kernel_launcher_fast(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx, cudaStream_t stream);
void three_interpolate_kernel_launcher_fast(int b, int c, int m, int n, const float *points, const int *idx, const float *weight, float *out, cudaStream_t stream);
void three_interpolate_grad_kernel_launcher_fast(int b, int c, int n, int m, const float *grad_out, const int *idx, const float *weight, float *grad_points, cudaStream_t stream);
void boxesoverlapLauncher(const int num_a, const float *boxes_a, const int num_b, const float *boxes_b, float *ans_overlap);
void boxesioubevLauncher(const int num_a, const float *boxes_a, const int num_b, const float *boxes_b, float *ans_iou);
void nmsLauncher(const float *boxes, unsigned long long *mask, int boxes_num, float nms_overlap_thresh);
void nmsNormalLauncher(const float *boxes, unsigned long long *mask, int boxes_num, float nms_overlap_thresh);
void roipool3dLauncher(int batch_size, int pts_num, int boxes_num, int feature_in_len, int sampled_pts_num, const float *xyz, const float *boxes3d, const float *pts_feature, float *pooled_features, int *pooled_empty_flag);

extern "C" {
void legacy_fps(int b, int n, int m, const float *xyz, float *temp, int *idx, cudaStream_t s) { furthest_point_sampling_kernel_launcher(b, n, m, xyz, temp, idx, s); }
void legacy_gather(int b, int c, int n, int m, const float *p, const int *idx, float *out, cudaStream_t s) { gather_points_kernel_launcher_fast(b, c, n, m, p, idx, out, s); }
void legacy_gather_grad(int b, int c, int n, int m, const float *g, const int *idx, float *gp, cudaStream_t s) { gather_points_grad_kernel_launcher_fast(b, c, n, m, g, idx, gp, s); }
void legacy_ball_query(int b, int n, int m, float r, int ns, const float *new_xyz, const float *xyz, int *idx, cudaStream_t s) { ball_query_kernel_launcher_fast(b, n, m, r, ns, new_xyz, xyz, idx, s); }
void legacy_group(int b, int c, int n, int m, int ns, const float *p, const int *idx, float *out, cudaStream_t s) { group_points_kernel_launcher_fast(b, c, n, m, ns, p, idx, out, s); }
void legacy_group_grad(int b, int c, int n, int m, int ns, const float *g, const int *idx, float *gp, cudaStream_t s) { group_points_grad_kernel_launcher_fast(b, c, n, m, ns, g, idx, gp, s); }
void legacy_three_nn(int b, int n, int m, const float *u, const float *k, float *d2, int *idx, cudaStream_t s) { three_nn_kernel_launcher_fast(b, n, m, u, k, d2, idx, s); }
void legacy_three_interpolate(int b, int c, int m, int n, const float *p, const int *idx, const float *w, float *out, cudaStream_t s) { three_interpolate_kernel_launcher_fast(b, c, m, n, p, idx, w, out, s); }
void legacy_three_interpolate_grad(int b, int c, int n, int m, const float *g, const int *idx, const float *w, float *gp, cudaStream_t s) { three_interpolate_grad_kernel_launcher_fast(b, c, n, m, g, idx, w, gp, s); }
// the iou3d / roipool3d launchers use the legacy default stream and (roipool3d) cudaMalloc inside
void legacy_boxes_overlap_bev(int na, const float *a, int nb, const float *b, float *out) { boxesoverlapLauncher(na, a, nb, b, out); }
void legacy_boxes_iou_bev(int na, const float *a, int nb, const float *b, float *out) { boxesioubevLauncher(na, a, nb, b, out); }
void legacy_nms_mask(const float *boxes, unsigned long long *mask, int n, float thresh) { nmsLauncher(boxes, mask, n, thresh); }
void legacy_nms_normal_mask(const float *boxes, unsigned long long *mask, int n, float thresh) { nmsNormalLauncher(boxes, mask, n, thresh); }
void legacy_roipool3d(int b, int n, int m, int c, int s, const float *xyz, const float *boxes, const float *feat, float *pooled, int *empty) { roipool3dLauncher(b, n, m, c, s, xyz, boxes, feat, pooled, empty); }
int legacy_sync() { return (int)cudaDeviceSynchronize(); }
}
