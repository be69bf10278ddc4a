/* libsstb200 - C ABI of the B200-native SST/FSD hot path.
 *
 * Drop-in boundary.  The reference (tusen-ai/SST @ 7c95376) reaches its native code
 * through three extension modules; each entry point below names the reference interface it replaces:
 *
 *   pybind module `voxel_layer`          mmdet3d/ops/voxel/src/voxelization.cpp:6-11, voxelization.h:51-130
 *   torch_scatter.scatter / scatter_max   call sites mmdet3d/ops/sst/sst_ops.py:173-175
 *   TorchEx ingroup_indices.forward       call site  mmdet3d/ops/sst/sst_ops.py:246-264
 *
 * plus fused entry points for the parts of the path the reference runs as chains of ATen ops
 * (window partition / bucketing, the Sparse-Regional-Attention encoder layer, DynamicVFE, SIRLayer).
 *
 * Conventions
 *   - plain C types only: device pointers, sizes, a cudaStream_t passed as void*.  No torch types.
 *   - every function returns 0 on success, <0 on error; sstb200_last_error(ctx) describes it.
 *   - all work is enqueued on the context's stream (sstb200_set_stream); no function synchronises
 *     unless it has a `*_host` out-parameter that is non-NULL (documented per function).
 *   - outputs are caller-allocated.  Data-dependent row counts (number of voxels / windows) are
 *     returned in device memory (`*_dev`); capacity of such outputs is the worst case (= #inputs).
 *   - temporaries come from a per-context arena that grows on demand.
 *   - one context per (device, stream, thread).
 */
#ifndef SSTB200_H_
#define SSTB200_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct sstb200_ctx sstb200_ctx;

#define SSTB200_REDUCE_SUM 0
#define SSTB200_REDUCE_MEAN 1
#define SSTB200_REDUCE_MAX 2

int sstb200_version(void);
sstb200_ctx* sstb200_create(int device);
void sstb200_destroy(sstb200_ctx* ctx);
int sstb200_set_stream(sstb200_ctx* ctx, void* cuda_stream);
const char* sstb200_last_error(sstb200_ctx* ctx);
int sstb200_num_sms(sstb200_ctx* ctx);

/* V1  voxel_layer.dynamic_voxelize(points, coors, voxel_size, coors_range, NDim=3)
 *     (voxelization.h:72-88, kernel voxelization_cuda.cu:25-65).  points [P,F] fp32 row-major,
 *     coors [P,3] int32 (z,y,x), clamped to the grid like this fork does.  No device sync
 *     (the reference's cudaDeviceSynchronize at voxelization_cuda.cu:371 is not reproduced). */
int sstb200_dynamic_voxelize(sstb200_ctx* ctx, const float* points, int num_points, int num_features,
                             const float voxel_size[3], const float coors_range[6], int32_t* coors);

/* V2  voxel_layer.dynamic_point_to_voxel_forward(feats, coors, reduce_type)
 *     (voxelization.h:96-108 -> scatter_points_cuda.cu:183-234).
 *     feats [P,C] fp32, coors [P,3] int32.  coor_lo/hi: inclusive bounds of the non-negative
 *     coordinates per column (the reference needs none because it sorts; this build ranks through a
 *     bitmap over the bounding grid).  Outputs (capacity P rows): reduced [.,C], out_coors [.,3],
 *     coors_map [P] (-1 = dropped), reduce_count [.].  num_voxels_dev: device int32.
 *     If num_voxels_host != NULL the call synchronises the stream and stores the count there.
 *     Reproduces the reference's unconditional removal of the first sorted row (:207-210). */
int sstb200_dynamic_point_to_voxel_forward(sstb200_ctx* ctx, const float* feats, const int32_t* coors,
                                           int num_points, int num_feats, int reduce_type,
                                           const int32_t coor_lo[3], const int32_t coor_hi[3],
                                           float* reduced, int32_t* out_coors, int32_t* coors_map,
                                           int32_t* reduce_count, int32_t* num_voxels_dev,
                                           int32_t* num_voxels_host);

/* V3  voxel_layer.dynamic_point_to_voxel_backward (voxelization.h:110-130 ->
 *     scatter_points_cuda.cu:236-303).  grad_feats [P,C] is fully written (zero where no grad). */
int sstb200_dynamic_point_to_voxel_backward(sstb200_ctx* ctx, float* grad_feats, const float* grad_reduced,
                                            const float* feats, const float* reduced,
                                            const int32_t* coors_map, const int32_t* reduce_count,
                                            int num_points, int num_voxels, int num_feats, int reduce_type);

/* V5a torch.unique(coors, dim=0, return_inverse=True, return_counts=True) on int64 rows
 *     (sst_ops.py:158-160,170; voxel_encoder.py:561; sir.py:70).  coors [P,ndim] int64, ndim<=4,
 *     every value inside [coor_lo, coor_hi].  Outputs: new_coors [.,ndim] (lexicographically sorted),
 *     inverse [P] int64, counts [.] int32 (may be NULL). */
int sstb200_unique_rows_i64(sstb200_ctx* ctx, const int64_t* coors, int num_rows, int ndim,
                            const int64_t* coor_lo, const int64_t* coor_hi, int64_t* new_coors,
                            int64_t* inverse, int32_t* counts, int32_t* num_unique_dev,
                            int32_t* num_unique_host);

/* V5b torch_scatter.scatter(src, index, dim=0, reduce=sum|mean) / scatter_max (sst_ops.py:173-175).
 *     src [P,C] fp32, index [P] int64 in [0,num_segments), out [num_segments,C] (empty segments: 0).
 *     argmax (int64 [num_segments,C], may be NULL): lowest row attaining the max, P if empty. */
int sstb200_segment_reduce(sstb200_ctx* ctx, const float* src, const int64_t* index, int num_rows,
                           int num_feats, int num_segments, int reduce_type, float* out, int64_t* argmax);

/* B2  ingroup_indices.forward(group_inds, out_inds) (sst_ops.py:246-264): rank of each element among
 *     the elements carrying the same id, in stable (input) order == get_inner_win_inds_slow
 *     (sst_input_layer.py:200-208).  group ids in [0, max_group_id]. */
int sstb200_ingroup_indices(sstb200_ctx* ctx, const int64_t* group_inds, int num, int64_t max_group_id,
                            int64_t* out_inds);

#ifdef __cplusplus
}
#endif
#endif
