// Host-side launch entry points of the kernels (internal; the public surface is include/rtg_splat_b200.h).
#pragma once
#include "common.cuh"

namespace rtg {

void launch_preprocess_fwd(const ViewParams &vp, int P, int M, const float *means, const float *scales, const float *rots,
                           const float *opac, const float *shs, const float *colors_precomp, const float *cov3D_precomp,
                           const int *tile_mask, const GeomState &g, int *radii, uint32_t *tile_count, uint32_t *tile_touched,
                           uint32_t *vis_count, cudaStream_t s);
void launch_mark_visible(int P, const float *means, const float *view, const float *proj, uint8_t *present, cudaStream_t s);
void launch_bwd_zero(int P, int M, bool has_sh, bool has_sr, const int *radii, float *dL_dmeans, float *dL_dsh, float *dL_dcolors,
                     float *dL_dopacity, float *dL_dscales, float *dL_drot, float *dL_dcov3D, float *dL_dmeans2D, cudaStream_t s);
void launch_preprocess_bwd(const ViewParams &vp, int P, int M, const float *means, const float *scales, const float *rots,
                           const float *shs, const float *cov3D_precomp, const GeomState &g, const uint32_t *vis_count, float *rec,
                           float *dL_dmeans, float *dL_dsh, float *dL_dcolors, float *dL_dopacity, float *dL_dscales, float *dL_drot,
                           float *dL_dcov3D, float *dL_dmeans2D, cudaStream_t s);

// binning.cu
void launch_tile_scan(const BinState &b, int T, int64_t R_cap, int32_t *counters, int32_t *counters_host, cudaStream_t s);
void launch_scatter(const ViewParams &vp, int P, const GeomState &g, const int *radii, const int *tile_mask, const BinState &b,
                    int64_t R_cap, const int32_t *counters, cudaStream_t s);
void launch_tile_sort(const BinState &b, int T, const int32_t *counters, cudaStream_t s);

// render.cu
void launch_render_fwd(const ViewParams &vp, const GeomState &g, const BinState &b, const ImgState &img, const int32_t *counters,
                       float *out_color, float *out_depth, int *out_hit_color, int *out_hit_depth, float *out_hcw,
                       float *out_hdw, float *out_T, cudaStream_t s);
void launch_render_bwd(const ViewParams &vp, const GeomState &g, const BinState &b, const ImgState &img, const int32_t *counters,
                       const float *means, const float *scales, const float *rots, const float *final_T, const int *hit_image,
                       const float *dL_dcolor, const float *dL_ddepth, float *rec, cudaStream_t s);

// mapstats.cu
void launch_gs_error(int H, int W, int P, const float *color_err, const float *depth_err, const float *normal_err,
                     const int *color_index, const int *depth_index, float thr_c, float thr_d, float thr_n, bool check_max,
                     float *gs_color, float *gs_depth, float *gs_normal, float *rescale, int *counters, cudaStream_t s);
void launch_tile_pool(int H, int W, const float *pixels, int mode, float ratio, uint8_t *pixel_mask, float *tile_mean, int *tile_mask,
                      cudaStream_t s);
void launch_color_error(int H, int W, const float *render, const float *gt, float *err, cudaStream_t s);

// frameprep.cu
void launch_frame_preprocess(const float *depth_in, int H, int W, int filter, int radius, float sigma_color, float sigma_space,
                             float min_depth, float max_depth, float fx, float fy, float cx, float cy, float conf_thresh,
                             float *depth_out, float *vertex, float *normal, float *confidence, uint8_t *invalid, void *ws,
                             cudaStream_t s);

// adam.cu / icp.cu declared in their own sections of capi.cu
}  // namespace rtg
