/*
 * pn2_b200.h -- C ABI of libpn2_b200.so, the sm_100a implementation of the PointRCNN
 * inference hot path of cxy1997/3D_adapt_auto_driving.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless its name
 *     starts with h_; the caller owns every buffer including scratch (nothing is allocated,
 *     freed or synchronised inside, unlike the reference's roipool3d_kernel.cu:214 /
 *     iou3d.cpp:87 which cudaMalloc per call and block on cudaMemcpy);
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued on it and the call
 *     returns immediately;
 *   - the return value is a status: 0 ok, 1 invalid argument, 2 launch failure,
 *     3 unsupported size.  pn2_last_error() gives the message.  Nothing ever calls exit()
 *     (the reference does: e.g. sampling_gpu.cu:249-252);
 *   - the library is stateless and re-entrant apart from the thread-local error string (and the opt-in stopwatch /
 *     debug hooks of the profiling tools, pn2_sa_fused_tc_set_profile / _set_debug, never used by the product).
 *
 * Each declaration cites the reference interface it replaces (paths relative to the
 * reference repository root).  INTEGRATION.md shows the binding a maintainer adds.
 */
#ifndef PN2_B200_H
#define PN2_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PN2_OK 0
#define PN2_ERR_INVALID 1
#define PN2_ERR_LAUNCH 2
#define PN2_ERR_UNSUPPORTED 3

const char *pn2_last_error(void);
int pn2_abi_version(void);

/* ---- pointnet2_cuda (pointrcnn/pointnet2_lib/pointnet2/src/pointnet2_api.cpp:10-24) ---- */

/* furthest_point_sampling_wrapper(b,n,m,points,temp,idx)  sampling.cpp:36-46,
 * kernel sampling_gpu.cu:93-253.  xyz (B,N,3) f32, temp (B,N) f32 caller scratch pre-filled
 * with 1e10 (mutated like the reference; NULL = implicit 1e10, not written back),
 * idx (B,M) int32.  Bit-exact incl. the reference's tie order. */
int pn2_fps_f32(const float *xyz, float *temp, int32_t *idx, int b, int n, int m, void *stream);
/* cuda_utils.h:10-14 opt_n_threads: the reference block size that fixes the tie order. */
int pn2_fps_ref_block_size(int n);
/* pn2_fps_f32 with the CTAs-per-cloud thread-block-cluster size given explicitly (1, 2, 4, 8; 0 = heuristic): tests and
 * tuning.  Same result for every value; no process-global state. */
int pn2_fps_cluster_f32(const float *xyz, float *temp, int32_t *idx, int b, int n, int m, int cluster_size, void *stream);
/* pn2_fps_f32 through the pruned one-CTA kernel (csrc/fps_cells.cu: Hilbert-ordered cells of 128 points with bounding
 * boxes; a round only touches the cells the new centre can change -- an exact test, same indices bit for bit), which is
 * what pn2_fps_f32 picks by itself for 2048 < N <= 16384.  N <= 16384; warps = CTA size in warps: 0 (heuristic = 8), 4, 8, 16. */
int pn2_fps_cells_f32(const float *xyz, float *temp, int32_t *idx, int b, int n, int m, int warps, void *stream);
/* Exact parallel test "does furthest_point_sample(xyz, m) return 0, 1, ..., m-1?" (true for every SA level of the
 * backbone after the first: pointnet2_msg.py:131-137 feeds level l the FPS-ordered centres of level l-1, and FPS of a
 * prefix of an FPS ordering is that prefix unless two candidates tie at the maximum).  viol (B) int32 (zeroed by this
 * call, on the stream), dmin (B, m) f32 scratch; afterwards viol[c] == 0 iff at every round the due point is the STRICT arg-max of the
 * running min-distances, computed with the reference's own float expressions (sampling_gpu.cu:129-138), so the answer
 * does not depend on the reference's tie order.  N * m independent pair evaluations instead of m dependent rounds. */
int pn2_fps_prefix_check_f32(const float *xyz, float *dmin, int32_t *viol, int b, int n, int m, void *stream);
/* pn2_fps_f32 (temp = NULL) that writes idx = 0..m-1 for the clouds with viol[c] == 0 and runs the round loop for the
 * others: bit-exact in both cases. */
int pn2_fps_guarded_f32(const float *xyz, int32_t *idx, const int32_t *viol, int b, int n, int m, void *stream);
/* pn2_fps_f32 (temp = NULL) / pn2_fps_guarded_f32 that also write new_xyz (B, m, 3) = the coordinates of the picked points:
 * the gather the callers of furthest_point_sample run next (pointnet2_modules.py:29-33) costs the kernel three stores a
 * round instead of a cast, a gather and a copy launch. */
int pn2_fps_xyz_f32(const float *xyz, int32_t *idx, float *new_xyz, int b, int n, int m, void *stream);
int pn2_fps_guarded_xyz_f32(const float *xyz, int32_t *idx, float *new_xyz, const int32_t *viol, int b, int n, int m,
                            void *stream);

/* gather_points_wrapper(b,c,n,npoints,points,idx,out)  sampling.cpp:11-21, sampling_gpu.cu:8-44.
 * points (B,C,N), idx (B,M) int32 -> out (B,C,M). */
int pn2_gather_points_f32(const float *points, const int32_t *idx, float *out, int b, int c, int n, int m, void *stream);
/* gather_points_grad_wrapper  sampling.cpp:24-34, sampling_gpu.cu:46-84. grad_points pre-zeroed. */
int pn2_gather_points_grad_f32(const float *grad_out, const int32_t *idx, float *grad_points, int b, int c, int n,
                               int m, void *stream);

/* ball_query_wrapper(b,n,m,radius,nsample,new_xyz,xyz,idx)  ball_query.cpp:14-24,
 * ball_query_gpu.cu:9-67.  idx (B,M,nsample) int32 must be zero-initialised by the caller
 * (pointnet2_utils.py:218): rows without a neighbour are not written.  Bit-exact. */
int pn2_ball_query_f32(const float *new_xyz, const float *xyz, int32_t *idx, int b, int n, int m, float radius,
                       int nsample, void *stream);
/* Both MSG scales of one SA layer (pointnet2_modules.py:37-38 loops over groupers) in one scan. */
int pn2_ball_query_dual_f32(const float *new_xyz, const float *xyz, int32_t *idx0, int32_t *idx1, int b, int n, int m,
                            float radius0, int nsample0, float radius1, int nsample1, void *stream);
/* Same idx as the two entries above (nsample1 == 0: single radius) through a spatially culled scan:
 * centres put in Hilbert-curve order per cloud into `order` (caller scratch, b*m int32), candidates
 * compacted by warp ballot against each warp's grown bounding box, hits appended by ballot too.
 * order == NULL or n < 128 -> brute force. */
int pn2_ball_query_culled_f32(const float *new_xyz, const float *xyz, int32_t *idx0, int32_t *idx1, int32_t *order,
                              int b, int n, int m, float radius0, int nsample0, float radius1, int nsample1,
                              void *stream);
/* The same for index lists the caller has NOT zeroed (the reference allocates them zeroed, pointnet2_utils.py:177, and its
 * kernel leaves the list of a centre without neighbours untouched): such lists are written as zeros here, every other
 * list is complete anyway.  Same results as a zero fill followed by pn2_ball_query_culled_f32.
 * hits0 / hits1 (B, M) int32 or NULL: neighbours found per centre, capped at nsample (input of pn2_group_compact_lists_i32). */
int pn2_ball_query_culled_fill_f32(const float *new_xyz, const float *xyz, int32_t *idx0, int32_t *idx1, int32_t *order,
                                   int b, int n, int m, float radius0, int nsample0, float radius1, int nsample1,
                                   int32_t *hits0, int32_t *hits1, void *stream);

/* group_points_wrapper(b,c,n,npoints,nsample,points,idx,out)  group_points.cpp:24-35,
 * group_points_gpu.cu:47-86.  points (B,C,N), idx (B,M,ns) -> out (B,C,M,ns). */
int pn2_group_points_f32(const float *points, const int32_t *idx, float *out, int b, int c, int n, int m, int nsample,
                         void *stream);
/* group_points_grad_wrapper  group_points.cpp:11-21, group_points_gpu.cu:8-44. */
int pn2_group_points_grad_f32(const float *grad_out, const int32_t *idx, float *grad_points, int b, int c, int n, int m,
                              int nsample, void *stream);

/* three_nn_wrapper(b,n,m,unknown,known,dist2,idx)  interpolate.cpp:14-24,
 * interpolate_gpu.cu:9-74.  unknown (B,n,3), known (B,m,3) -> dist2 (B,n,3) f32 SQUARED,
 * idx (B,n,3) int32.  Bit-exact. */
int pn2_three_nn_f32(const float *unknown, const float *known, float *dist2, int32_t *idx, int b, int n, int m,
                     void *stream);
/* Same dist2 / idx through a spatially culled scan (unknown points Hilbert-ordered into `order`,
 * caller scratch of b*n int32; known points compacted by warp ballot against the warp's bounding box
 * grown by the current third-neighbour bound).  order == NULL or a small level -> brute force. */
int pn2_three_nn_culled_f32(const float *unknown, const float *known, float *dist2, int32_t *idx, int32_t *order, int b,
                            int n, int m, void *stream);
/* three_interpolate_wrapper(b,c,m,n,points,idx,weight,out)  interpolate.cpp:27-38,
 * interpolate_gpu.cu:77-117.  points (B,C,m) -> out (B,C,n).  Bit-exact. */
int pn2_three_interpolate_f32(const float *points, const int32_t *idx, const float *weight, float *out, int b, int c,
                              int m, int n, void *stream);
/* three_interpolate_grad_wrapper(b,c,n,m,...)  interpolate.cpp:40-53, interpolate_gpu.cu:120-160. */
int pn2_three_interpolate_grad_f32(const float *grad_out, const int32_t *idx, const float *weight, float *grad_points,
                                   int b, int c, int n, int m, void *stream);

/* ---- roipool3d_cuda (pointrcnn/lib/utils/roipool3d/src/roipool3d.cpp:198-203) ---- */

/* forward(xyz, boxes3d, pts_feature, pooled_features, pooled_empty_flag)  roipool3d.cpp:17-45,
 * kernels roipool3d_kernel.cu:97-232.  xyz (B,N,3), boxes3d (B,M,7) already enlarged
 * (roipool3d_utils.py:18), feat (B,N,C) -> pooled (B,M,sampled,3+C) f32 and empty (B,M) int32,
 * both pre-zeroed by the caller (roipool3d_utils.py:20-22).  Indices bit-exact. */
int pn2_roipool3d_f32(const float *xyz, const float *boxes3d, const float *feat, float *pooled, int32_t *empty, int b,
                      int n, int m, int c, int sampled, void *stream);
/* The same pooling (same sampled points, same order) with the per-point features given in two
 * pieces and a padded output row [x y z | feat (c) | 0.. | feat2 (c2) at column off2 | 0..] of
 * ld_out floats, so the rpn feature block is 16-byte aligned for the tensor-core MLP that reads
 * it (replaces the torch.cat of rcnn_net.py:137-139 plus roipool3d_utils.py:7-28 on the fused path). */
int pn2_roipool3d_split_f32(const float *xyz, const float *boxes3d, const float *feat, int c, const float *feat2, int c2,
                            float *pooled, int ld_out, int off2, int32_t *empty, int b, int n, int m, int sampled,
                            void *stream);

/* ---- iou3d_cuda (pointrcnn/lib/utils/iou3d/src/iou3d.cpp:174-179) ---- */

/* boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap)  iou3d.cpp:31-50, iou3d_kernel.cu:223-234.
 * a (na,5), b (nb,5) [x1,y1,x2,y2,ry] -> out (na,nb) intersection area. */
int pn2_boxes_overlap_bev_f32(const float *a, int na, const float *b, int nb, float *out, void *stream);
/* boxes_iou_bev_gpu  iou3d.cpp:52-71, iou3d_kernel.cu:236-248. */
int pn2_boxes_iou_bev_f32(const float *a, int na, const float *b, int nb, float *out, void *stream);
/* boxes_iou3d_gpu (lib/utils/iou3d/iou3d_utils.py:21-53) in one launch: a (na,7), b (nb,7) [x,y,z,h,w,l,ry] -> out (na,nb)
 * 3-D IoU = rotated BEV intersection x height overlap / union volume; float operations in the order of the reference's torch
 * statements. */
int pn2_boxes_iou3d_f32(const float *a, int na, const float *b, int nb, float *out, void *stream);

/* nms_gpu / nms_normal_gpu (boxes, keep, thresh) -> num  iou3d.cpp:73-169 (mask kernels
 * iou3d_kernel.cu:250-348 + host greedy pass), batched and device-resident: boxes
 * (problems, stride, 5) sorted by descending score, problem p uses its first counts[p] boxes
 * (counts NULL: n for all); keep (problems, max_keep) int64, num (problems) int32.
 * max_keep = n gives the reference's full keep list; keep indices are bit-exact. */
int pn2_nms_bev_f32(const float *boxes, int problems, int stride, int n, const int32_t *counts, float thresh,
                    int rotated, int max_keep, long long *keep, int32_t *num, void *stream);
/* The same for TWO sets of problems in one launch (the near and the far band of the proposal layer,
 * lib/rpn/proposal_layer.py:58-119, differ in candidate count and keep limit): same results as two calls. */
int pn2_nms_bev_pair_f32(const float *boxes0, int problems0, int stride0, int n0, const int32_t *counts0, int max_keep0,
                         long long *keep0, int32_t *num0, const float *boxes1, int problems1, int stride1, int n1,
                         const int32_t *counts1, int max_keep1, long long *keep1, int32_t *num1, float thresh, int rotated,
                         void *stream);

/* ---- shared-MLP layers (pointnet2_lib/pointnet2/pytorch_utils.py:5-101 SharedMLP/Conv1d/Conv2d,
 *      pointnet2_modules.py:37-48 group -> MLP -> max_pool2d), point-major activations ---- */

/* Y[r,0:cout] = act(X[r,0:cin] . W[c,0:cin]^T + bias[c] [+ R[r,c]]), optional max over `pool`
 * consecutive rows (pool = nsample: F.max_pool2d of pointnet2_modules.py:42).  Within 1e-4 rel
 * of the cuDNN fp32 result. */
int pn2_linear_f32(const float *x, int ldx, const float *w, int ldw, const float *bias, const float *res, int ldr,
                   float *y, int ldy, long long rows, int cin, int cout, int relu, int pool, void *stream);
/* QueryAndGroup.forward (pointnet2_utils.py:241-264) + SharedMLP layers 1 and 2 fused: the
 * grouped tensor is never materialised.  h (clouds*n, ldh) = per-point part of layer 1. */
int pn2_sa_group_linear_f32(const float *h, int ldh, const int32_t *idx, const float *xyz, const float *centres,
                            const float *wxyz, const float *w, int ldw, const float *bias, float *y, int ldy, int clouds,
                            int n, int m, int ns, int c1, int cout, int relu, int pool, void *stream);
/* three_interpolate on point-major features (internal layout of the fused FP module,
 * pointnet2_modules.py:139-149). */
int pn2_three_interpolate_pm_f32(const float *feats, int ldf, const int32_t *idx, const float *weight, float *out,
                                 int ldo, int b, int c, int m, int n, void *stream);
/* The same fed with the SQUARED distances of three_nn_wrapper: sqrt (pointnet2_utils.py:104), 1 / (dist + 1e-8), the sum of
 * the three and the division (pointnet2_modules.py:209-211) happen in the kernel, with the IEEE operations torch executes
 * for those statements in the same order (bit-identical weights, tests/test_pn2_ops_gpu.py).
 * sum_order = association of the three-term sum, 0: (r0 + r1) + r2, 1: (r0 + r2) + r1 (torch.sum's kernel), 2: r0 + (r1 + r2). */
int pn2_three_interpolate_pm_d2_f32(const float *feats, int ldf, const int32_t *idx, const float *dist2, float *out,
                                    int ldo, int b, int c, int m, int n, int sum_order, void *stream);

/* ---- the same shared-MLP layers on the tcgen05 tensor cores (BF16x3 split, fp32 accumulation in
 *      TMEM); wblob is the host-packed weight image described in csrc/linear_tc.cu.  When pool (nsample)
 *      is 64 or 128 the pooled output y must be ZERO-FILLED by the caller: pooling groups that span
 *      several warps are combined with atomicMax on the (non-negative) float bits. ---- */
int pn2_linear_tc_f32(const float *x, int ldx, const void *wblob, int ntile, int nchunks, int nkb, const float *bias,
                      const float *res, int ldr, float *y, int ldy, long long rows, int cin, int cout, int relu,
                      int pool, void *stream);
/* pn2_linear_tc_f32 on the never-materialised concatenation [x (c_a columns) | x2 (cin - c_a columns)]:
 * merge_down_layer on cat[xyz_feature, rpn_feature] (rcnn_net.py:174-176) without the cat. */
int pn2_linear_tc2_f32(const float *x, int ldx, int c_a, const float *x2, int ldx2, const void *wblob, int ntile,
                       int nchunks, int nkb, const float *bias, const float *res, int ldr, float *y, int ldy,
                       long long rows, int cin, int cout, int relu, int pool, void *stream);
/* Two layers in one launch when the first one is tiny (RCNN xyz_up_layer [5 -> 128 -> 128], rcnn_net.py:41-47):
 * y = act(relu(x[:, :cpre] . Wpre^T + bpre) . W^T + b); wpre (cpre + 1, c1) = the first layer's weight rows
 * (input-major) then its bias row.  cpre = 5. */
int pn2_linear_pre_tc_f32(const float *x, int ldx, int cpre, const float *wpre, const void *wblob, int ntile, int nchunks,
                          int nkb, const float *bias, float *y, int ldy, long long rows, int c1, int cout, int relu,
                          int pool, void *stream);
int pn2_sa_group_linear_tc_f32(const float *h, int ldh, const int32_t *idx, const float *xyz, const float *centres,
                               const float *wxyz, const void *wblob, int ntile, int nchunks, int nkb, const float *bias,
                               float *y, int ldy, int clouds, int n, int m, int ns, int c1, int cout, int relu, int pool,
                               void *stream);

/* One SA scale end to end on chip: QueryAndGroup (pointnet2_utils.py:241-264) + SharedMLP layers 1-3
 * (pytorch_utils.py:5-101) + F.max_pool2d over nsample (pointnet2_modules.py:42), given the per-point
 * half h of layer 1.  Returns PN2_ERR_UNSUPPORTED when the shape does not fit TMEM / shared memory
 * (callers then use the layer-by-layer entry points above).  csrc/sa_fused_tc.cu. */
int pn2_sa_fused_tc_f32(const float *h, int ldh, const int32_t *idx, const float *xyz, const float *centres,
                        const float *wxyz, const void *w2blob, int n2, int nkb1, const float *b2, const void *w3blob,
                        int n3, int nkb2, const float *b3, float *y, int ldy, int clouds, int n, int m, int ns, int c1,
                        int c2, int c3, void *stream);

/* ---- evaluate/rotate_iou.py:294-329 rotate_iou_gpu_eval (numba kernel rotate_iou_kernel_eval :261-291) ----
 * boxes (n,5), qboxes (k,5) rows [cx, cy, w, h, angle] f32 -> out (n,k) f32; criterion -1 IoU,
 * 0 inter/area(query), 1 inter/area(box), 2 inter.  Bit-exact with the numba build. */
int pn2_rotate_iou_eval_f32(const float *boxes, int n, const float *qboxes, int k, float *out, int criterion,
                            void *stream);

/* tuning hook: per-CTA stopwatch buffer (32 u64 per CTA, device memory) for the fused SA kernel, NULL = off */
/* pn2_sa_fused_tc_f32 with the last layer computed transposed (W3 resident in tensor memory, rows of the
 * tile as accumulator columns, so the max over nsample is an in-thread reduction): the shapes with 128 or
 * 256 output channels.  w3hi / w3lo: (c3, c2 / 2) uint32, bf16 hi / lo pairs of W3.  csrc/sa_fused_t_tc.cu. */
int pn2_sa_fused_t_tc_f32(const float *h, int ldh, const int32_t *idx, const float *xyz, const float *centres,
                          const float *wxyz, const void *w2blob, int n2, int nkb1, const float *b2, const void *w3hi,
                          const void *w3lo, const float *b3, float *y, int ldy, int clouds, int n, int m, int ns, int c1,
                          int c2, int c3, const int32_t *cmap, const int32_t *jmap, const long long *rows_dev, int cmap_align,
                          void *stream);
/* The unique rows of ball-query groups (a group with cnt < nsample hits is padded with copies of its first hit, which
 * cannot change the max-pool): cnt (G) from idx (G, ns); with the exclusive prefix sum offs (G) int64 of cnt, the compact
 * row list cmap[u] = group, jmap[u] = neighbour index (u < sum cnt).  Passed to pn2_sa_fused_t_tc_f32 (cmap, jmap and
 * the device-resident row count) the SA kernel processes only those rows.  align (1 .. 16, a power of two dividing ns)
 * tops every group up to a multiple of `align` rows with further copies of its row 0, so that an aligned run of `align` list
 * rows has ONE centre: pn2_sa_fused_t_tc_f32 pools such a list (cmap_align 8) eight columns at a time.  csrc/group_compact.cu. */
int pn2_group_unique_count_i32(const int32_t *idx, long long g, int ns, int align, int32_t *cnt, void *stream);
int pn2_group_compact_i32(const int32_t *idx, long long g, int ns, const int32_t *cnt, const long long *offs,
                          int32_t *cmap, int32_t *jmap, void *stream);
/* Both steps and the exclusive scan between them in two launches: cnt (G) and block_sum (ceil(G / 256)) int32 scratch,
 * *total (device int64) = number of list rows.  Same lists as count -> prefix sum -> pn2_group_compact_i32.  hits (G) or
 * NULL: the hit counts of pn2_ball_query_culled_fill_f32 for these lists; the count pass then does not read the lists. */
int pn2_group_compact_lists_i32(const int32_t *idx, long long g, int ns, int align, const int32_t *hits, int32_t *cnt,
                                int32_t *block_sum, int32_t *cmap, int32_t *jmap, long long *total, void *stream);
void pn2_sa_fused_tc_set_profile(void *buf);
void pn2_sa_fused_t_set_profile(void *buf);    /* tools/prof_sat.py: in-kernel stopwatch of pn2_sa_fused_t_tc_f32 */
void pn2_sa_fused_t_set_debug(int bits);       /* tools/prof_sat.py: what-if switches of the stopwatch build (garbage results) */
void pn2_rcnn_front_set_profile(void *buf);   /* tools/prof_front.py: in-kernel stopwatch of pn2_rcnn_front_tc_f32 */
void pn2_fps_cells_set_profile(void *buf);    /* tools/prof_fps_cells.py: in-kernel stopwatch of pn2_fps_cells_f32 (16384-point shapes) */
void pn2_rcnn_front_set_mode(int bits);       /* tools/prof_front.py: tuning variants of the same kernel (same results) */
/* profiling experiments only: bit0 / bit1 switch off the TMEM traffic of the pooling / conversion epilogue
 * (results become garbage); 0 restores the product behaviour. */
void pn2_sa_fused_tc_set_debug(int bits);

/* ---- GPU data path of KittiRCNNDataset.get_rpn_sample (csrc/scene_prepare.cu) ----
 * pn2_scene_filter_f32: per scene lidar -> rect (lib/utils/calibration.py:51-58), projection into the image
 * (:60-71), the image / PC_AREA_SCOPE filter (lib/datasets/kitti_rcnn_dataset.py:201-222) and the ordered
 * near / far index lists (:291-296), numpy's float32 arithmetic reproduced bit for bit.
 *   raw (total, 4) f32 = the .bin clouds of the batch concatenated, offsets (b + 1) int64, calib (b, 32) f32 =
 *   {np.dot(V2C.T, R0.T) (4,3), P2.T (4,3), image width, height, pad}; x0 .. z1 = PC_AREA_SCOPE;
 *   valid (b, cap, 4) f32 rect x, y, z, intensity of the kept points in input order; near_list / far_list (b, cap)
 *   int32 positions in `valid`; counts (b, 4) int32 = {valid, near, far, 0}.  One CTA per scene.
 * pn2_scene_gather_f32: pts (b, npoints, 3) [and feat (b, npoints) = intensity - 0.5, or NULL] of the selection
 *   sel (b, npoints) int32 drawn on the host from the counts (datasets/gpu_loader.py: draw_selection). */
int pn2_scene_filter_f32(const float *raw, const long long *offsets, const float *calib, double x0, double x1, double y0,
                         double y1, double z0, double z1, int reduce_by_range, float near_z, float *valid,
                         int32_t *near_list, int32_t *far_list, int32_t *counts, int b, long long cap, void *stream);
int pn2_scene_gather_f32(const float *valid, const int32_t *near_list, const int32_t *far_list, const int32_t *sel,
                         float *pts, float *feat, int b, int npoints, long long cap, void *stream);

/* HOST function: pn2_scene_filter_f32's per-point pipeline for ONE scene on the CPU (the DataLoader workers of the
 * unmodified eval_rcnn.py): raw (n, 4) f32, m / p = the (4, 3) row-major float32 matrices lidar -> rect and rect -> image,
 * scope {x0, x1, y0, y1, z0, z1} -> valid (n, 4) capacity [rect x, y, z, intensity] compacted in input order, near_list /
 * far_list (n) capacity, counts[3] = {n_valid, n_near, n_far}.  Same float32 arithmetic as numpy's (FMA chains). */
int pn2_scene_filter_host_f32(const float *raw, long long n, const float *m, const float *p, float width, float height,
                              int reduce_by_range, const double *scope, float near_z, float *valid, int32_t *near_list,
                              int32_t *far_list, long long *counts);

/* HOST functions (no device work): the np.random draws of KittiRCNNDataset._sample_indices
 * (lib/datasets/kitti_rcnn_dataset.py:291-320) on an explicit MT19937 state, bit for bit numpy's legacy
 * RandomState.choice / shuffle / randint.  key (624) + pos as in np.random.get_state().  csrc/mt_select.cu. */
void pn2_mt_seed(uint32_t seed, uint32_t *key, int32_t *pos);
int pn2_mt_draw_selection(uint32_t *key, int32_t *pos, int n_valid, int n_near, int n_far, int npoints,
                          int npoints_faraway, int with_replace, int32_t *sel, int32_t *scratch);

/* ---- KITTI AP evaluator around rotate_iou (evaluate/eval2.py; csrc/kitti_eval.cu) ----
 * pn2_d3_overlap_f64 (DEVICE): d3_box_overlap_kernel (eval2.py:136-162): camera boxes (., 7) float64 and the rotated
 *   BEV intersection areas rinc (n, k) float32 (pn2_rotate_iou_eval_f32, criterion 2) -> out (n, k) float64.
 * Host functions (float64 / int64 arrays, no device work):
 *   pn2_eval_image_box_overlap   image_box_overlap (:104-127)
 *   pn2_eval_collect_thresholds  compute_statistics_jit(thresh 0, compute_fp False) over the images of one dataset
 *                                part (:506-520): scores of the true positives
 *   pn2_eval_fused_statistics    fused_compute_statistics (:311-358): pr (n_thresholds, 4) += [tp, fp, fn, similarity]
 *   pn2_eval_image_statistics    compute_statistics_jit (:172-298) for one image, every mode: out4 = {tp, fp, fn,
 *                                similarity (-1 = undefined)}, thresholds_out (gt_size) / *n_thresholds = tp scores
 *   overlaps: (total_dt, total_gt) of the part; gt_datas (., 5), dt_datas (., 6), dontcares (., 4). */
int pn2_d3_overlap_f64(const double *boxes, long long n, const double *qboxes, long long k, const float *rinc,
                       int criterion, double *out, void *stream);
int pn2_eval_image_box_overlap(const double *boxes, long long n, const double *qboxes, long long k, int criterion,
                               double *out);
int pn2_eval_collect_thresholds(const double *overlaps, long long total_dt, long long total_gt, const long long *gt_nums,
                                const long long *dt_nums, const long long *dc_nums, long long n_img,
                                const double *gt_datas, const double *dt_datas, const double *dontcares,
                                const long long *ignored_gts, const long long *ignored_dets, int metric,
                                double min_overlap, double *thresholds_out, long long *n_out);
int pn2_eval_image_statistics(const double *overlaps, long long ldo, const double *gt_datas, long long gt_size,
                              const double *dt_datas, long long det_size, const long long *ignored_gt,
                              const long long *ignored_det, const double *dc_bboxes, long long n_dc, int metric,
                              double min_overlap, double thresh, int compute_fp, int compute_aos, double *out4,
                              double *thresholds_out, long long *n_thresholds);
int pn2_eval_fused_statistics(const double *overlaps, long long total_dt, long long total_gt, double *pr,
                              const long long *gt_nums, const long long *dt_nums, const long long *dc_nums,
                              long long n_img, const double *gt_datas, const double *dt_datas, const double *dontcares,
                              const long long *ignored_gts, const long long *ignored_dets, int metric, double min_overlap,
                              const double *thresholds, long long n_thresholds, int compute_aos);

/* ---- Statistical Normalization point rescale on the GPU (stat_norm/norm.py:186-244, 42-45;
 * csrc/stat_norm.cu), default options (avoid_conflict = align_front = False): float64 arithmetic in numpy's dgemm
 * order, output = the float32 rows of the rescaled .bin, bit-identical to rescale_ptc + format_lidar_data.
 * mats (b, 42) f64 = {V2C^T (4,3), R0 (3,3), inv(R0) (3,3), C2V^T (4,3)}; boxes (nboxes, 18) f64 =
 * {t[3], R[9], l/2, h, w/2, scale[3]} of the rescaled-class objects in label order, box_offsets (b + 1). */
int pn2_stat_rescale_f64(const float *raw, const long long *offsets, const double *mats, const double *boxes,
                         const int32_t *box_offsets, double *rect, unsigned char *untouched, float *out,
                         int32_t *out_counts, int32_t *box_counts, int b, long long cap, long long cap_out, void *stream);
/* The same with convert's options (norm.py:205-240): avoid_conflict = per box the largest ratio of np.arange(1, -0.1, -0.1)
 * whose scaled patch swallows fewer than ten foreign points, found on the device; align_front = shift of the patch that
 * keeps the face nearest to the sensor in place.  box_opts (nboxes, 78) f64 = {scale[11][3] = mapping(obj, ratio_r),
 * shift[11][4] = up to two (dx, dz) pairs per ratio, nshift}, all from the reference's numpy expressions on the host;
 * box_ratio (nboxes) int32 receives the index of the chosen ratio. */
int pn2_stat_rescale_opts_f64(const float *raw, const long long *offsets, const double *mats, const double *boxes,
                              const int32_t *box_offsets, const double *box_opts, int avoid_conflict, double *rect,
                              unsigned char *untouched, float *out, int32_t *out_counts, int32_t *box_counts,
                              int32_t *box_ratio, int b, long long cap, long long cap_out, void *stream);

/* ---- the small stages between the big kernels (csrc/glue.cu, csrc/roipool3d.cu): one launch each for what the
 * reference writes as 20-60 torch statements.  Same IEEE float operations in the same order (bit-identical outputs,
 * tests/test_glue_gpu.py); Python scalars arrive as doubles and are rounded where torch rounds them.  rot_mode: rounding
 * of the K = 2 batched matmul in rotate_pc_along_y_torch (2 = two rounded products and an add, what torch's kernel does
 * on B200; 0 / 1 = the two FMA orders). ----
 * pn2_decode_bbox_f32         decode_bbox_target (lib/utils/bbox_transform.py:24-121); roi (rows, 3 | 7), reg (rows, c),
 *                             h_anchor HOST float[3]; y_bottom: y += h / 2 (lib/rpn/proposal_layer.py:23)
 * pn2_proposal_select_f32     distance_based_proposal up to the NMS input (proposal_layer.py:58-100): order (B, N) int64
 *                             descending-score order, props (B, N, 7) -> candidate point ids cidx0 (B, pre0) / cidx1 (B, pre1),
 *                             their BEV boxes bev0 / bev1 (., 5) (kitti_utils.py:134-147), cnt (2, B) int32
 * pn2_proposal_assemble_f32   proposal_layer.py:107-119, :38-44: keep lists -> zero-padded rois (B, post0 + post1, 7), scores
 * pn2_rcnn_post_prepare_f32   tools/eval_rcnn.py:516-535, 611-620: decode, sigmoid(raw) > thresh, stable descending score
 *                             order, BEV boxes; m <= 256 rois per scene, single foreground class
 * pn2_rcnn_post_assemble_f32  eval_rcnn.py:621-627: keep (B, m) int64 / num (B) -> rec (B, m, 8) [box7, raw score], rest zero
 * pn2_roipool3d_canon_f32     lib/net/rcnn_net.py:126-154 (enlarge_box3d + roipool3d_gpu + canonical transform, extras =
 *                             [seg mask, depth / depth_norm - 0.5]); rois NOT enlarged; empty (B, M) zeroed by the caller */
int pn2_decode_bbox_f32(const float *roi, int roi_dim, const float *reg, int c, float *out, long long rows,
                        double loc_scope, double loc_bin_size, int num_head_bin, const float *h_anchor, int get_xz_fine,
                        int get_y_by_bin, double loc_y_scope, double loc_y_bin_size, int get_ry_fine, int y_bottom,
                        int rot_mode, void *stream);
/* torch.sort(scores, dim=1, descending=True)[1] of lib/rpn/proposal_layer.py:26 as one launch (one CTA per row, bitonic
 * network in shared memory): order (B, N) int64, descending score, ascending index among equal scores; N <= 16384. */
int pn2_argsort_desc_f32(const float *scores, long long *order, int b, int n, void *stream);
int pn2_proposal_select_f32(const long long *order, const float *props, int b, int n, int pre0, int pre1, int32_t *cidx0,
                            int32_t *cidx1, float *bev0, float *bev1, int32_t *cnt, void *stream);
int pn2_proposal_assemble_f32(const float *props, const float *scores, int b, int n, const int32_t *cidx0,
                              const int32_t *cidx1, int pre0, int pre1, const long long *keep0, const long long *keep1,
                              const int32_t *num0, const int32_t *num1, int post0, int post1, float *rois,
                              float *roi_scores, void *stream);
int pn2_rcnn_post_prepare_f32(const float *rois, const float *reg, int c, const float *cls, int b, int m, double loc_scope,
                              double loc_bin_size, int num_head_bin, const float *h_anchor, int get_y_by_bin,
                              double loc_y_scope, double loc_y_bin_size, double score_thresh, int rot_mode,
                              float *boxes_sorted, float *scores_sorted, float *bev, int32_t *counts, void *stream);
int pn2_rcnn_post_assemble_f32(const float *boxes_sorted, const float *scores_sorted, const long long *keep,
                               const int32_t *num, int b, int m, float *rec, void *stream);
int pn2_roipool3d_canon_f32(const float *xyz, const float *rois, double extra_width, const float *score,
                            double score_thresh, const float *depth, double depth_norm, const float *feat2, int c2,
                            float *pooled, int ld_out, int off2, int32_t *empty, int b, int n, int m, int sampled,
                            int rot_mode, void *stream);

/* The RCNN input chain in one tcgen05 launch (csrc/rcnn_front_tc.cu): xyz_up_layer [5 -> 128 -> 128] (lib/net/rcnn_net.py:41-47,
 * :168-171), merge_down_layer on cat[xyz feature, rpn feature] (:174-176) and the per-point half of SA1's first layer
 * (pointnet2_modules.py:38-44).  x (rows, ldx) pooled rows [5 extras | pad | 128 rpn features at column off_f] (16-byte
 * aligned rows); wpre (6, 128) f32 = the five input-major weight rows of the first layer + its bias; w_up2 / w_merge / w_sa:
 * fused.pack_tc images (bf16 hi/lo, 128B-swizzled K-major) of the (128x128), (128x256), (128x128) folded weights;
 * h (rows, ldh): PRE-activation W1f . merged + b1.  BF16x3, fp32 accumulate. */
int pn2_rcnn_front_tc_f32(const float *x, int ldx, int off_f, const float *wpre, const void *w_up2, const float *b_up2,
                          const void *w_merge, const float *b_merge, const void *w_sa, const float *b_sa, float *h, int ldh,
                          long long rows, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PN2_B200_H */
