/*
 * rtg_splat_b200.h -- C ABI of librtg_splat_b200.so (sm_100a).
 *
 * Drop-in boundary for RTG-SLAM's data-parallel hot path. Every entry point names the
 * reference interface it replaces ("RAST/" = submodules/diff-gaussian-rasterizer-depth/ of
 * MisEty/RTG-SLAM). Plain pointers and sizes only; no torch types. All device pointers are
 * CUDA device memory on the current device, fp32 / int32, contiguous. Every call is
 * asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream) and
 * performs no host<->device synchronisation.
 *
 * Return value: 0 on success, negative RtgStatus on error; rtg_last_error() holds the message
 * (thread-local).
 */
#ifndef RTG_SPLAT_B200_H
#define RTG_SPLAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum RtgStatus {
    RTG_OK = 0,
    RTG_ERR_INVALID_ARGUMENT = -1,
    RTG_ERR_CUDA = -2,
    RTG_ERR_UNSUPPORTED = -3,
};

const char *rtg_last_error(void);
int rtg_version(void);

/* Per-view constants: replaces GaussianRasterizationSettings
 * (RAST/diff_gaussian_rasterization_depth/__init__.py:284-303) as marshalled by
 * RasterizeGaussiansCUDA (RAST/rasterize_points.cu:37-64). Matrices stay on the device, exactly as
 * the reference passes them (viewmatrix = W2C transposed, projmatrix = full projection,
 * scene/cameras.py:96-110). */
typedef struct RtgSplatView {
    int32_t image_height, image_width;
    float tanfovx, tanfovy;
    float cx, cy;
    float scale_modifier;
    float color_sigma;
    float opaque_threshold;
    float depth_threshold;
    float normal_threshold; /* cosine */
    float T_threshold;
    int32_t sh_degree;
    int32_t prefiltered;
    const float *viewmatrix; /* device, 16 floats */
    const float *projmatrix; /* device, 16 floats */
    const float *campos;     /* device, 3 floats  */
    const float *bg;         /* device, 3 floats  */
} RtgSplatView;

/* Device counters written by rtg_splat_forward (int32 each). */
enum { RTG_CNT_NUM_RENDERED = 0, RTG_CNT_NUM_TILES = 1, RTG_CNT_OVERFLOW = 2, RTG_CNT_MAX_TILE_LEN = 3,
       RTG_CNT_ENTRIES_NEEDED = 4, /* instances with every tile's bucket padded to 4 entries: what R_cap must hold */
       RTG_CNT_WORDS = 8 };

/* Sizes of the three state buffers kept between forward and backward. Replaces the
 * geomBuffer / imgBuffer / binningBuffer resize callbacks (RAST/rasterize_points.cu:27-35,
 * required<T>() in RAST/cuda_rasterizer/rasterizer_impl.h). `R_cap` is the capacity, in
 * (Gaussian,tile) instances, of the binning buffer: the reference sizes it after a blocking
 * read-back of num_rendered (rasterizer_impl.cu:304); here the caller provides a capacity and
 * the forward raises counters[RTG_CNT_OVERFLOW] (and renders nothing) if it was too small. */
int rtg_splat_workspace_bytes(int32_t P, int32_t H, int32_t W, int64_t R_cap, size_t *geom_bytes, size_t *img_bytes,
                              size_t *bin_bytes);

/* Replaces CudaRasterizer::Rasterizer::forward (RAST/cuda_rasterizer/rasterizer.h:28-66,
 * rasterizer_impl.cu:205-437) together with the output initialisation of
 * RasterizeGaussiansCUDA (rasterize_points.cu:79-87): every element of every output is written.
 * Exactly one of (shs | colors_precomp) and one of (scales+rotations | cov3D_precomp) must be
 * non-NULL, as in GaussianRasterizer.forward (__init__.py:335-347). `M` = SH coefficients per
 * Gaussian. `counters`: device, RTG_CNT_WORDS int32. `counters_host` (may be NULL): pinned, device-mapped
 * host memory that receives a copy of the counters straight from the scan kernel; `scan_done_event` (may be
 * NULL): a cudaEvent_t recorded right after that kernel, so a caller can learn num_rendered / overflow
 * while the rest of the forward is still executing (the reference blocks the whole device instead,
 * rasterizer_impl.cu:304,346). */
int rtg_splat_forward(const RtgSplatView *view, int32_t P, int32_t M, const float *means3D, const float *shs,
                      const float *colors_precomp, const float *opacities, const float *scales, const float *rotations,
                      const float *cov3D_precomp, const int32_t *tile_mask, void *geom_ws, void *img_ws, void *bin_ws,
                      int64_t R_cap, float *out_color, float *out_depth, int32_t *out_hit_color, int32_t *out_hit_depth,
                      float *out_hit_color_weight, float *out_hit_depth_weight, float *out_T, int32_t *radii,
                      int32_t *counters, int32_t *counters_host, void *scan_done_event, void *stream);

/* Replaces CudaRasterizer::Rasterizer::backward (rasterizer.h:68-105, rasterizer_impl.cu:441-560)
 * and the zero-initialisation of RasterizeGaussiansBackwardCUDA (rasterize_points.cu:195-203):
 * every element of every non-NULL gradient output is written. The three state buffers, `R_cap` and
 * `counters` are those of the matching forward call; `final_T` is the forward's out_T,
 * `hit_image` its out_hit_depth (__init__.py:172-235). `grad2d_scratch`: P*16 floats that must be
 * all-zero on entry and are left all-zero on exit. Optional outputs (may be NULL):
 * dL_dcolors_precomp, dL_dcov3D, dL_dmeans2D. Internally the zero-fill of the culled rows runs on a library-owned
 * side stream that forks from and joins back into `stream` within this call (event fork/join, graph-capturable);
 * to the caller all work is ordered on `stream`. */
int rtg_splat_backward(const RtgSplatView *view, int32_t P, int32_t M, const float *means3D, const float *shs,
                       const float *colors_precomp, const float *scales, const float *rotations,
                       const float *cov3D_precomp, const int32_t *radii, const void *geom_ws, const void *img_ws,
                       const void *bin_ws, int64_t R_cap, const int32_t *counters, const float *final_T, const int32_t *hit_image, const float *dL_dcolor,
                       const float *dL_ddepth, float *grad2d_scratch, float *dL_dmeans3D, float *dL_dsh,
                       float *dL_dcolors_precomp, float *dL_dopacity, float *dL_dscales, float *dL_drotations,
                       float *dL_dcov3D, float *dL_dmeans2D, void *stream);

/* rtg_splat_backward without the zero fill: only the gradient rows of Gaussians with radii > 0 are written, the
 * others keep whatever the buffers held. For consumers that take `radii` themselves (rtg_map_adam_step): saves the
 * 105 MB of zero stores per 1 M Gaussians that a dense consumer needs. Same arguments. */
int rtg_splat_backward_visible(const RtgSplatView *view, int32_t P, int32_t M, const float *means3D, const float *shs,
                               const float *colors_precomp, const float *scales, const float *rotations,
                               const float *cov3D_precomp, const int32_t *radii, const void *geom_ws, const void *img_ws,
                               const void *bin_ws, int64_t R_cap, const int32_t *counters, const float *final_T,
                               const int32_t *hit_image, const float *dL_dcolor, const float *dL_ddepth, float *grad2d_scratch,
                               float *dL_dmeans3D, float *dL_dsh, float *dL_dcolors_precomp, float *dL_dopacity,
                               float *dL_dscales, float *dL_drotations, float *dL_dcov3D, float *dL_dmeans2D, void *stream);

/* The same backward in two calls with the same argument list, for callers that put an exchange step between the
 * compositing backward and the per-Gaussian backward (tile-sharded multi-GPU rendering):
 *   rtg_splat_backward_render : zero-fill of the culled rows + backward of renderCUDA (backward.cu:808-1066); on return
 *                               `grad2d_scratch` holds, per Gaussian, the 16-float record of partial sums
 *                               {colour 3, six moments of opacity*G*dL/dalpha (the 2-D gradients are linear in them),
 *                               depth-path dL/dmean 3, depth-path dL/d(surfel normal) 3, unused 1}
 *                               over the tiles this forward rendered -- sum it across ranks (e.g. ncclAllReduce);
 *   rtg_splat_backward_finish : computeCov2DCUDA + preprocessCUDA backward (backward.cu:273-548) from the records,
 *                               which it clears again. Visibility (radii) does not depend on tile_mask, so ranks
 *                               that render different tiles of one frame agree on which records exist.
 * rtg_splat_backward == rtg_splat_backward_render followed by rtg_splat_backward_finish. */
int rtg_splat_backward_render(const RtgSplatView *view, int32_t P, int32_t M, const float *means3D, const float *shs,
                              const float *colors_precomp, const float *scales, const float *rotations,
                              const float *cov3D_precomp, const int32_t *radii, const void *geom_ws, const void *img_ws,
                              const void *bin_ws, int64_t R_cap, const int32_t *counters, const float *final_T,
                              const int32_t *hit_image, const float *dL_dcolor, const float *dL_ddepth, float *grad2d_scratch,
                              float *dL_dmeans3D, float *dL_dsh, float *dL_dcolors_precomp, float *dL_dopacity,
                              float *dL_dscales, float *dL_drotations, float *dL_dcov3D, float *dL_dmeans2D, void *stream);
int rtg_splat_backward_finish(const RtgSplatView *view, int32_t P, int32_t M, const float *means3D, const float *shs,
                              const float *colors_precomp, const float *scales, const float *rotations,
                              const float *cov3D_precomp, const int32_t *radii, const void *geom_ws, const void *img_ws,
                              const void *bin_ws, int64_t R_cap, const int32_t *counters, const float *final_T,
                              const int32_t *hit_image, const float *dL_dcolor, const float *dL_ddepth, float *grad2d_scratch,
                              float *dL_dmeans3D, float *dL_dsh, float *dL_dcolors_precomp, float *dL_dopacity,
                              float *dL_dscales, float *dL_drotations, float *dL_dcov3D, float *dL_dmeans2D, void *stream);

/* ---- Gaussian-sharded rendering (BASELINE configs[4]) ----------------------------------
 * Rank r owns the Gaussians [p_begin, p_end) of the P-Gaussian map (parameters, gradients, optimizer state: arrays of
 * p_end - p_begin rows) and a subset of the 16x16 tiles. The forward of rtg_splat_forward is split around the exchange
 * of the per-Gaussian records, the backward around the exchange of the gradient records:
 *   rtg_splat_forward_preprocess  clears the binning counters and runs preprocessCUDA (forward.cu:238-354) on the owned
 *                                 Gaussians: writes rows [p_begin, p_end) of the geometry workspace (sized for P) and of
 *                                 `radii` (P entries) -- then all-gather those rows across ranks (their byte offsets inside
 *                                 the workspace: rtg_splat_geom_layout; splat 32 B, colour 16 B, surfel 32 B per Gaussian);
 *   rtg_splat_forward_render      tile histogram of ALL P records over the tiles of `tile_mask` (this rank's tiles), scan,
 *                                 scatter, per-tile sort, compositing: outputs as rtg_splat_forward, valid on those tiles.
 *                                 [tile_row_begin, tile_row_end): tile rows outside which tile_mask is all zero (a rank that
 *                                 owns a band of rows passes it so that the binning passes clip every rectangle to the band
 *                                 before expanding it: per-rank work / N instead of constant); 0, 0 = no clipping;
 *   rtg_splat_backward_render_shard   compositing backward over this rank's tiles into `grad2d_scratch` (P records, zero on
 *                                 entry) + zero-fill of the owned gradient rows -- then reduce-scatter the records so that
 *                                 the owner of [p_begin, p_end) holds their sums;
 *   rtg_splat_backward_finish_shard   per-Gaussian backward of the owned Gaussians; here `grad2d_scratch` must be indexable
 *                                 at records [p_begin, p_end) (pass the reduce-scatter output minus p_begin records).
 * In the *_shard calls means3D / shs / ... / dL_d* point to the OWNED rows (row 0 = Gaussian p_begin); radii, the
 * workspaces and the image-sized arrays are global. */
int rtg_splat_geom_layout(int32_t P, size_t *splat_offset, size_t *rgb_offset, size_t *hit_offset, size_t *vis_list_offset);
int rtg_splat_forward_preprocess(const RtgSplatView *view, int32_t P, int32_t p_begin, int32_t p_end, int32_t M, const float *means3D,
                                 const float *shs, const float *colors_precomp, const float *opacities, const float *scales,
                                 const float *rotations, const float *cov3D_precomp, void *geom_ws, void *bin_ws, int64_t R_cap,
                                 int32_t *radii, void *stream);
int rtg_splat_forward_render(const RtgSplatView *view, int32_t P, const int32_t *tile_mask, void *geom_ws, void *img_ws, void *bin_ws,
                             int64_t R_cap, float *out_color, float *out_depth, int32_t *out_hit_color, int32_t *out_hit_depth,
                             float *out_hit_color_weight, float *out_hit_depth_weight, float *out_T, const int32_t *radii,
                             int32_t *counters, int32_t *counters_host, void *scan_done_event, int32_t tile_row_begin,
                             int32_t tile_row_end, void *stream);
int rtg_splat_backward_render_shard(int32_t p_begin, int32_t p_end, const RtgSplatView *view, int32_t P, int32_t M,
                                    const float *means3D, const float *shs, const float *colors_precomp, const float *scales,
                                    const float *rotations, const float *cov3D_precomp, const int32_t *radii, const void *geom_ws,
                                    const void *img_ws, const void *bin_ws, int64_t R_cap, const int32_t *counters,
                                    const float *final_T, const int32_t *hit_image, const float *dL_dcolor, const float *dL_ddepth,
                                    float *grad2d_scratch, float *dL_dmeans3D, float *dL_dsh, float *dL_dcolors_precomp,
                                    float *dL_dopacity, float *dL_dscales, float *dL_drotations, float *dL_dcov3D,
                                    float *dL_dmeans2D, void *stream);
int rtg_splat_backward_finish_shard(int32_t p_begin, int32_t p_end, const RtgSplatView *view, int32_t P, int32_t M,
                                    const float *means3D, const float *shs, const float *colors_precomp, const float *scales,
                                    const float *rotations, const float *cov3D_precomp, const int32_t *radii, const void *geom_ws,
                                    const void *img_ws, const void *bin_ws, int64_t R_cap, const int32_t *counters,
                                    const float *final_T, const int32_t *hit_image, const float *dL_dcolor, const float *dL_ddepth,
                                    float *grad2d_scratch, float *dL_dmeans3D, float *dL_dsh, float *dL_dcolors_precomp,
                                    float *dL_dopacity, float *dL_dscales, float *dL_drotations, float *dL_dcov3D,
                                    float *dL_dmeans2D, void *stream);

/* Replaces CudaRasterizer::Rasterizer::markVisible (rasterizer.h:21-26, rasterizer_impl.cu:145-157). */
int rtg_splat_mark_visible(int32_t P, const float *means3D, const float *viewmatrix, const float *projmatrix,
                           uint8_t *present, void *stream);

/* ---- optimizer step ------------------------------------------------------------------------
 * Replaces torch.optim.Adam(l, lr=0.0, eps=1e-15).step() over the parameter groups built by
 * GaussianPointCloud.parametrize (SLAM/gaussian_pointcloud.py:245-284; SLAM/multiprocess/
 * mapper.py:156,452): one launch for all groups, bias-corrected Adam, no weight decay, no amsgrad. */
#define RTG_ADAM_MAX_GROUPS 8
typedef struct RtgAdamGroup {
    float *param;
    const float *grad;
    float *exp_avg;
    float *exp_avg_sq;
    int64_t numel;
    float lr;
    float _pad;
} RtgAdamGroup;
int rtg_adam_step(const RtgAdamGroup *groups, int32_t n_groups, float beta1, float beta2, float eps, int32_t step,
                  void *stream);

/* ---- fused map-parameter step ---------------------------------------------------------------
 * One pass over the Gaussian map that replaces, per optimisation iteration of Mapping.loss_update
 * (SLAM/multiprocess/mapper.py:376-468):
 *   - the activation backward of get_scaling / get_rotation / get_opacity (torch.exp, F.normalize, torch.sigmoid;
 *     SLAM/gaussian_pointcloud.py:16-25,511-523,574-581) and the slice backward of get_features' torch.cat,
 *   - the gradient of the "attach" regulariser 1000 * (l2(_scaling) + l2(_xyz) + l2(_rotation)) on the rows with
 *     attach_mask (mapper.py:384-401; l2_loss = mean of squares, utils/loss_utils.py:34-36),
 *   - torch.optim.Adam(l, lr=0.0, eps=1e-15).step() over the six groups of parametrize (gaussian_pointcloud.py:245-284),
 *   - `_confidence[(f_dc.grad.abs() != 0).any(-1)] += 1` (mapper.py:455-456),
 *   - the activation forward (and get_normal, gaussian_pointcloud.py:539-550) for the next iteration's render.
 * Gradients are the rasterizer's, i.e. with respect to the ACTIVATED tensors. With `radii` (the forward's output)
 * rows with radii <= 0 are treated as zero gradient and never read, so the backward may be
 * rtg_splat_backward_visible, which skips the zero fill of those rows. `sh` is ONE (P,16,3) block whose [:,0:1] /
 * [:,1:] slices are _features_dc / _features_rest (no torch.cat). All pointers device memory; optional ones may be
 * NULL: radii, attach_mask (+ xyz0, scaling0, rotation0), confidence, normal_out. */
typedef struct RtgMapStep {
    int32_t P;
    int32_t step;                 /* Adam step number, >= 1 */
    float *xyz, *sh, *opacity_raw, *scaling_raw, *rotation_raw;                 /* raw parameters, updated in place */
    float *m_xyz, *m_sh, *m_opacity, *m_scaling, *m_rotation;                   /* exp_avg */
    float *v_xyz, *v_sh, *v_opacity, *v_scaling, *v_rotation;                   /* exp_avg_sq */
    const float *g_means3D, *g_sh, *g_opacity, *g_scales, *g_rotations;         /* rasterizer gradients */
    const int32_t *radii;
    const uint8_t *attach_mask;
    const float *xyz0, *scaling0, *rotation0;                                   /* init_stat: _xyz, _scaling, _rotation (raw) */
    float attach_weight;          /* 1000 in the reference */
    int32_t attach_count;         /* number of rows with attach_mask (the means' denominators) */
    float lr_xyz, lr_f_dc, lr_f_rest, lr_opacity, lr_scaling, lr_rotation;
    float beta1, beta2, eps;
    float *scales_out, *rotations_out, *opacities_out;                          /* activated values after the update */
    float *normal_out;            /* (P,3) get_normal after the update */
    float *confidence;            /* (P) float */
} RtgMapStep;
int rtg_map_adam_step(const RtgMapStep *step, void *stream);
/* Activation forward only (initialisation): scales_out = exp(scaling_raw), rotations_out = normalize(rotation_raw),
 * opacities_out = sigmoid(opacity_raw), normal_out = get_normal (may be NULL). */
int rtg_map_activate(int32_t P, const float *scaling_raw, const float *rotation_raw, const float *opacity_raw,
                     float *scales_out, float *rotations_out, float *opacities_out, float *normal_out, void *stream);

/* Mapping.history_merge (SLAM/multiprocess/mapper.py:212-250) with slerp of SLAM/utils.py:593-652: after an optimisation
 * call, pull the raw parameters back towards their pre-optimisation values (`history_stat`, mapper.py:146-155):
 *   w = max_weight * hist_confidence / (confidence + 1e-6)                                  per Gaussian,
 *   _xyz          <- hist_xyz * w + (1 - w) * _xyz,
 *   _features_dc, _features_rest, _scaling  <- hist * w[0] + (1 - w[0]) * current    (the reference indexes
 *                    `history_weight[0]`: the FIRST Gaussian's weight for every row -- kept),
 *   _rotation     <- slerp(hist_rotation, normalize(_rotation), 1 - w)   (linear where |dot| > 0.9995 or NaN).
 * One launch instead of ~45 eager kernels; all updates in place. hist_rotation is get_rotation before the optimisation
 * (normalised); rotation_raw is `_rotation` (16-byte aligned). features_dc / features_rest rows start every
 * `*_stride` floats (3 / 45 for separate contiguous tensors, 48 for the two slices of one (P,16,3) block);
 * hist arrays are contiguous. max_weight <= 0 or P == 0: nothing is done (mapper.py:213-214). */
typedef struct RtgHistoryMerge {
    int32_t P;
    float max_weight;             /* history_merge_max_weight, configs/base.yaml:54 */
    const float *hist_confidence, *confidence;          /* (P) */
    const float *hist_xyz;      float *xyz;             /* (P,3) */
    const float *hist_features_dc;   float *features_dc;    /* rows of 3 */
    const float *hist_features_rest; float *features_rest;  /* rows of features_rest_width (45 for SH degree 3) */
    const float *hist_scaling;  float *scaling;         /* (P,3) */
    const float *hist_rotation; float *rotation_raw;    /* (P,4) */
    int32_t features_dc_stride, features_rest_stride, features_rest_width;
    int32_t _pad;
} RtgHistoryMerge;
int rtg_map_history_merge(const RtgHistoryMerge *merge, void *stream);

/* ---- projective point-to-plane ICP ---------------------------------------------------------
 * Pyramid level: replaces nn.MaxPool2d(pool) (SLAM/icp.py:343-345,374) + compute_vertex_map
 * (SLAM/utils.py:65-75) + compute_normal_map (SLAM/utils.py:100-122). `depth`: (H,W) full
 * resolution; outputs (H/pool, W/pool, 3) channels-last. fx..cy are the level intrinsics
 * (K*downscale, SLAM/utils.py:517-519). `ws`: rtg_icp_workspace_bytes() bytes. */
size_t rtg_icp_workspace_bytes(int32_t H, int32_t W);
int rtg_icp_build_level(const float *depth, int32_t H, int32_t W, int32_t pool, float fx, float fy, float cx, float cy,
                        float *vertex_out, float *normal_out, void *ws, void *stream);

/* `iters` Gauss-Newton iterations on one level: replaces the loop of ICP.icp (SLAM/icp.py:33-48):
 * compute_residuals_jacobian (:52-104), compute_jtj/jtr (:107-119), lev_mar_H (:248),
 * least_square_solve/invH (:313-333), exp_se3 (:271), forward_update_pose (:259). "0" is the
 * current frame, "1" the previous/model frame, as inside icp(). `pose`: device, 16 floats
 * row-major, updated in place. `valid_ratio`: device float (may be NULL). */
int rtg_icp_solve_level(const float *vertex0, const float *normal0, const float *vertex1, const float *normal1,
                        int32_t H, int32_t W, float fx, float fy, float cx, float cy, float distance_threshold,
                        float normal_cos_threshold, float damping, int32_t iters, float *pose, float *valid_ratio,
                        void *ws, void *stream);

/* point2plane_loss(p_t0, p_t1 @ R^T + t, n_t0, "mean") (SLAM/icp.py:7-13,444-447); `loss`: device float. */
int rtg_icp_point2plane_loss(const float *vertex_t0, const float *vertex_t1, const float *normal_t0, int32_t H, int32_t W,
                             const float *pose, float *loss, void *ws, void *stream);

/* IcpTracker.update_last_status depth filling (SLAM/icp.py:397-415): in place on render_depth. */
int rtg_icp_fill_model_depth(float *render_depth, const float *frame_depth, const float *render_normal,
                             const float *frame_normal, int32_t H, int32_t W, float distance_threshold,
                             float normal_threshold, void *stream);

/* Whole pyramid in one launch: build_vertex_pyramid + build_normal_pyramid (SLAM/utils.py:511-527) as IcpTracker.
 * update_curr_status / predict_pose call them (SLAM/icp.py:385-395,423-425). Level l max-pools `depth` (H,W) by pools[l]
 * and back-projects with the level intrinsics {fx,fy,cx,cy}[l] (K*downscale, SLAM/utils.py:517-519); vertex_out[l] /
 * normal_out[l] are device (H/pools[l], W/pools[l], 3) buffers. 1 <= n_levels <= 4. One cooperative launch (max-pool +
 * vertices of all levels, grid barrier for the per-level depth extremes, normals). */
int rtg_icp_build_pyramid(const float *depth, int32_t H, int32_t W, int32_t n_levels, const int32_t *pools, const float *fx,
                          const float *fy, const float *cx, const float *cy, float *const *vertex_out, float *const *normal_out,
                          void *ws, void *stream);

/* IcpTracker.predict_pose (SLAM/icp.py:417-452) in one cooperative launch: for every level (coarse to fine) `iters`
 * Gauss-Newton iterations as in rtg_icp_solve_level, the pose carried from level to level starting at pose_init (device,
 * 16 floats; NULL = identity), then point2plane_loss of the final pose on (p2p_vertex_t0, p2p_vertex_t1, p2p_normal_t0),
 * each (p2p_H, p2p_W, 3). In every level "0" is the CURRENT frame and "1" the previous / model frame (the argument swap of
 * SLAM/icp.py:438-441). out: 18 device floats = pose (row-major 4x4), loss, valid ratio of the last iteration; out_host
 * (may be NULL): the same 18 floats written straight into mapped pinned host memory by the kernel, valid once work
 * enqueued after this call on `stream` has been reached (record an event and wait for it) -- the only read-back of a
 * predict_pose. */
typedef struct RtgIcpLevel {
    const float *vertex0, *normal0, *vertex1, *normal1; /* (H,W,3) channels-last, device */
    int32_t H, W;
    float fx, fy, cx, cy;
    int32_t iters;
    int32_t _pad;
} RtgIcpLevel;
int rtg_icp_predict_pose(const RtgIcpLevel *levels, int32_t n_levels, float distance_threshold, float normal_cos_threshold,
                         float damping, const float *pose_init, const float *p2p_vertex_t0, const float *p2p_vertex_t1,
                         const float *p2p_normal_t0, int32_t p2p_H, int32_t p2p_W, float *out, float *out_host, void *ws,
                         void *stream);

/* ---- image-space glue between the rasterizer forward and backward -------------
 * Fused masked L1 colour + depth loss of Mapping.loss_update (SLAM/multiprocess/mapper.py:402-431,444-451, with
 * l1_loss of utils/loss_utils.py:27-31) and its gradients w.r.t. the rendered colour and depth:
 *   colour = mean |render - gt_color| over the render_mask pixels (NULL = all) x 3 channels,
 *   depth  = mean |depth - gt_depth| over pixels with depth_index != -1, gt_depth > 0, depth - gt_depth <
 *            depth_error_max and the render mask,  loss = color_weight*colour + depth_weight*depth.
 * render (3,H,W), depth (1,H,W), depth_index (1,H,W) are rasterizer outputs; gt_color is (H,W,3) if
 * gt_channels_last else (3,H,W); gt_depth (H,W); render_mask (H,W) bytes. loss_out: 4 device floats
 * {loss, colour, depth, n_depth}. An empty selection contributes 0 (torch's mean of an empty tensor is NaN).
 * ws: rtg_loss_workspace_bytes() bytes. */
size_t rtg_loss_workspace_bytes(void);
int rtg_loss_l1(const float *render, const float *depth, const int32_t *depth_index, const float *gt_color, const float *gt_depth,
                const uint8_t *render_mask, int32_t H, int32_t W, int32_t gt_channels_last, float color_weight, float depth_weight,
                float depth_error_max, float *dL_dcolor, float *dL_ddepth, float *loss_out, void *ws, void *stream);

/* The same loss with the cosine normal term of Mapping.loss_update (mapper.py:433-442):
 *   normal = mean of 1 - cosine_similarity(render_normal, gt_normal) over pixels with render_mask, depth_index != -1 and
 *            gt_normal != 0,  loss = color_weight*colour + depth_weight*depth + normal_weight*normal.
 * render_normal (3,H,W) is Renderer.render's "normal"; gt_normal (H,W,3); dL_dnormal (3,H,W). With normal_weight == 0 the
 * three normal pointers may be NULL. loss_out: 8 device floats {loss, colour, depth, n_depth, normal, n_normal, 0, 0}:
 * everything Mapping.loss_update reports (mapper.py:459-466) in one buffer, one read-back instead of six .item() calls. */
int rtg_loss_mapping(const float *render, const float *depth, const float *render_normal, const int32_t *depth_index,
                     const float *gt_color, const float *gt_depth, const float *gt_normal, const uint8_t *render_mask, int32_t H, int32_t W,
                     int32_t gt_channels_last, float color_weight, float depth_weight, float normal_weight, float depth_error_max,
                     float *dL_dcolor, float *dL_ddepth, float *dL_dnormal, float *loss_out, void *ws, void *stream);

/* SSIM term of Mapping.loss_update (mapper.py:411-415: ssim_loss = 1 - ssim(render, gt), evaluated only without a render
 * mask) with ssim of utils/loss_utils.py:40-100: 11x11 Gaussian window (sigma 1.5, one per channel), zero padding,
 * C1 = 0.01^2, C2 = 0.03^2, mean over all C*H*W entries of the ssim map. img1, img2: (C,H,W).
 * loss_out: 2 device floats {1 - mean(ssim_map), mean(ssim_map)}; dL_dimg1 (C,H,W) = d loss_out[0] / d img1, or NULL for the
 * value only. Replaces five grouped 11x11 convolutions, ~15 elementwise kernels and their autograd twins by three launches;
 * the mean is reduced in a fixed order (bit-reproducible). ws: rtg_ssim_workspace_bytes(C, H, W) bytes. */
size_t rtg_ssim_workspace_bytes(int32_t C, int32_t H, int32_t W);
int rtg_ssim_loss(const float *img1, const float *img2, int32_t C, int32_t H, int32_t W, float *dL_dimg1, float *loss_out, void *ws,
                  void *stream);

/* render_normal of Renderer.render (SLAM/render.py:130-133): out (3,H,W) = normal[depth_index] where the index is
 * > -1, zeros elsewhere; normal is (P,3). */
int rtg_normal_map(const float *normal, const int32_t *depth_index, int32_t H, int32_t W, float *out, void *stream);

/* ---- tracker-side preprocessing of an incoming depth frame --------------------------
 * Tracker.map_preprocess (SLAM/multiprocess/tracker.py:97-132): optional bilateral depth filter (bilateralFilter_torch,
 * SLAM/utils.py:550-589; taps with i^2+j^2 <= radius^2, zero padding, zero-depth neighbours ignored), valid-range mask
 * (min_depth, max_depth) exclusive, vertex map and Sobel normal map (compute_vertex_map / compute_normal_map,
 * SLAM/utils.py:65-122), confidence = |cos(normal, pixel ray)| (compute_confidence_map, SLAM/utils.py:125-138), and the
 * invalid-confidence masking: where the normal is zero or the confidence is below the threshold, depth, vertex, normal
 * and confidence are set to 0 and invalid_mask_out (may be NULL) to 1. depth_in / depth_out (H,W); vertex, normal
 * (H,W,3) channels-last; confidence (H,W). depth_in == depth_out is allowed only without the filter. With vertex_out,
 * normal_out, confidence_out and invalid_mask_out all NULL only the filter and the range mask run.
 * ws: rtg_icp_workspace_bytes() bytes. */
int rtg_frame_preprocess(const float *depth_in, int32_t H, int32_t W, int32_t depth_filter, int32_t radius, float sigma_color,
                         float sigma_space, float min_depth, float max_depth, float fx, float fy, float cx, float cy,
                         float invalid_confidence_thresh, float *depth_out, float *vertex_out, float *normal_out,
                         float *confidence_out, uint8_t *invalid_mask_out, void *ws, void *stream);

/* ---- per-frame consumers of the rasterizer's index / transmittance maps ---------
 * accumulate_gaussian_error of submodules/cuda_utils (cuda_utils.cu:17-60, map_process.cu:33-245), called by
 * Mapping.error_gaussians_remove (SLAM/multiprocess/mapper.py:546-559): for every pixel, the colour error goes to the
 * Gaussian named by color_index, the depth and normal errors to the one named by depth_index (indices outside [0,P)
 * are skipped); per Gaussian the maximum (check_max) or the mean of its pixels; rescale_counter counts the pixels
 * whose colour / depth / normal error exceeds its threshold (as a float, like the reference). All inputs (H*W), all
 * outputs (P); outputs need not be initialised. counters: 2*P int32 of scratch, only used when !check_max. */
int rtg_accumulate_gaussian_error(int32_t H, int32_t W, int32_t P, const float *screen_color_error, const float *screen_depth_error,
                                  const float *screen_normal_error, const int32_t *screen_color_index,
                                  const int32_t *screen_depth_index, float color_threshold, float depth_threshold,
                                  float normal_threshold, int32_t check_max, float *gs_color_error, float *gs_depth_error,
                                  float *gs_normal_error, float *gs_rescale_counter, int32_t *counters, void *stream);

/* 16x16 average pooling of an (H,W) float image with zero padding, as F.avg_pool2d is used by transmission2tilemask /
 * colorerror2tilemask (SLAM/utils.py:695-734): tile_mean (tiles_y, tiles_x) and / or tile_mask = (mean > ratio), either
 * may be NULL. */
int rtg_tile_mean(int32_t H, int32_t W, const float *pixels, float ratio, float *tile_mean, int32_t *tile_mask, void *stream);

/* render_mask = (T_map != 1) and tile_mask = transmission2tilemask(render_mask, 16, ratio) in one pass
 * (Mapping.evaluate_render_range, mapper.py:503-505). render_mask (H,W) bytes, may be NULL. */
int rtg_transmission_tile_mask(int32_t H, int32_t W, const float *T_map, float ratio, uint8_t *render_mask, int32_t *tile_mask,
                               void *stream);

/* colour error image of mapper.py:481-487: sum over channels of |render - gt|, 0 where the rendered pixel is black;
 * render, gt (3,H,W), out (H,W). */
int rtg_color_error(int32_t H, int32_t W, const float *render, const float *gt, float *out, void *stream);

/* ---- map surgery on the Gaussian SoA and nearest neighbours ----------------------------
 * rtg_soa_compact: GaussianPointCloud.delete / remove (SLAM/gaussian_pointcloud.py:195-235) -- keep the rows whose mask
 * byte is non-zero (invert = 0) or zero (invert = 1; `delete(mask)` keeps ~mask) of n_arrays attribute arrays that share the
 * row index; array a has words_per_row[a] 4-byte words per row (xyz 3, features_rest 45, rotation 4, counters 1, ...).
 * One scan of the mask, one gather launch for all arrays; rows keep their order. out[a] must hold P rows; n_kept (device
 * uint32) and n_kept_host (mapped pinned uint32, may be NULL) receive the number of rows kept. n_arrays <= 16.
 * ws: rtg_soa_compact_workspace_bytes(P) bytes. in / out: host arrays of device pointers. */
#define RTG_SOA_MAX_ARRAYS 16
size_t rtg_soa_compact_workspace_bytes(int64_t P);
int rtg_soa_compact(const uint8_t *mask, int32_t invert, int64_t P, int32_t n_arrays, const void *const *in, void *const *out,
                    const int32_t *words_per_row, uint32_t *n_kept, uint32_t *n_kept_host, void *ws, void *stream);

/* rtg_knn: exact K nearest reference points of every query point, 1 <= K <= 8, squared Euclidean distances ascending.
 * Replaces distCUDA2 of submodules/simple-knn (simple_knn.cu:169-251; spatial.cu) as GaussianPointCloud.update_geometry
 * uses it (gaussian_pointcloud.py:376: query == ref, K = 3, skip_self = 1, out_mean = (d0+d1+d2)/3 and out_idx) and
 * pytorch3d.ops.knn_points as Mapping.temp_points_filter / gaussians_isolated call it (mapper.py:812-819,903-910:
 * skip_self = 0). query (n_query,3), ref (n_ref,3) fp32; out_d2 (n_query,K), out_idx (n_query,K) int32, out_mean (n_query)
 * -- each may be NULL. With fewer than K candidates the missing slots hold FLT_MAX / INT_MAX (simple-knn's sentinels).
 * Uniform-grid search sized on the device from the bounding box of `ref`; no host synchronisation.
 * ws: rtg_knn_workspace_bytes(n_ref) bytes. */
size_t rtg_knn_workspace_bytes(int64_t n_ref);
int rtg_knn(const float *query, int64_t n_query, const float *ref, int64_t n_ref, int32_t K, int32_t skip_self, float *out_d2,
            int32_t *out_idx, float *out_mean, void *ws, void *stream);

/* ---- measurement hook (no reference counterpart) ---------------------------------------------
 * When enabled, every kernel launch of this library is bracketed by CUDA events on its launching stream.
 * rtg_profile_read synchronises the device and returns, per kernel id, the summed duration (ms) and the number
 * of launches since the last reset. bench.py uses it for the roofline line and the launch count. */
int rtg_profile_enable(int32_t on);
int rtg_profile_kernel_count(void);
const char *rtg_profile_kernel_name(int32_t id);
int rtg_profile_read(double *total_ms, int64_t *launches, int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* RTG_SPLAT_B200_H */
