// Host-side launch entry points of the kernels (internal; the public surface is include/rtg_splat_b200.h).
#pragma once
#include "common.cuh"

namespace rtg {

void launch_preprocess_fwd(const ViewParams &vp, int P, int M, const float *means, const float *scales, const float *rots,
                           const float *opac, const float *shs, const float *colors_precomp, const float *cov3D_precomp,
                           const int *tile_mask, const GeomState &g, int *radii, uint32_t *tile_count, uint32_t *tile_touched,
                           uint32_t *vis_count, int p_begin, cudaStream_t s);
void launch_mark_visible(int P, const float *means, const float *view, const float *proj, uint8_t *present, cudaStream_t s);
void launch_bwd_zero(int p_begin, int P, int M, bool has_sh, bool has_sr, const int *radii, float *dL_dmeans, float *dL_dsh,
                     float *dL_dcolors, float *dL_dopacity, float *dL_dscales, float *dL_drot, float *dL_dcov3D, float *dL_dmeans2D,
                     cudaStream_t s);
void launch_preprocess_bwd(const ViewParams &vp, int P, int M, const float *means, const float *scales, const float *rots,
                           const float *shs, const float *cov3D_precomp, const GeomState &g, const uint32_t *vis_count, float *rec,
                           float *dL_dmeans, float *dL_dsh, float *dL_dcolors, float *dL_dopacity, float *dL_dscales, float *dL_drot,
                           float *dL_dcov3D, float *dL_dmeans2D, cudaStream_t s);

// binning.cu
void launch_tile_histogram(const ViewParams &vp, int P, const GeomState &g, const int *radii, const int *tile_mask, const BinState &b,
                           cudaStream_t s);
void launch_tile_scan(const BinState &b, int T, int64_t R_cap, int32_t *counters, int32_t *counters_host, cudaStream_t s);
void launch_scatter(const ViewParams &vp, int P, const GeomState &g, const int *radii, const int *tile_mask, const BinState &b,
                    int64_t R_cap, const int32_t *counters, cudaStream_t s);
void launch_tile_sort(const BinState &b, int T, const int32_t *counters, cudaStream_t s);

// render.cu
void launch_render_fwd(const ViewParams &vp, const GeomState &g, const BinState &b, const ImgState &img, const int32_t *counters,
                       float *out_color, float *out_depth, int *out_hit_color, int *out_hit_depth, float *out_hcw,
                       float *out_hdw, float *out_T, cudaStream_t s);
void launch_render_bwd(const ViewParams &vp, const GeomState &g, const BinState &b, const ImgState &img, const int32_t *counters,
                       const float *final_T, const int *hit_image, const float *dL_dcolor, const float *dL_ddepth, float *rec,
                       cudaStream_t s);

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

// icp.cu: persistent (cooperative) kernels
#define RTG_ICP_MAX_LEVELS 4
struct IcpLevel {
    const float *v0, *n0, *v1, *n1;  // "0" = current frame, "1" = previous / model frame
    int H, W;
    float fx, fy, cx, cy;
    int iters;
};
struct IcpPredict {
    IcpLevel lv[RTG_ICP_MAX_LEVELS];
    int n_levels;
    float dist_thr, cos_thr, damping;
    const float *pose_in;
    const float *p2p_v_t0, *p2p_v_t1, *p2p_n_t0;
    int p2p_HW;
    float *out, *out_host;
};
struct IcpPyramid {
    const float *depth;
    int H, W, n_levels;
    int pool[RTG_ICP_MAX_LEVELS];
    float fx[RTG_ICP_MAX_LEVELS], fy[RTG_ICP_MAX_LEVELS], cx[RTG_ICP_MAX_LEVELS], cy[RTG_ICP_MAX_LEVELS];
    float *vertex[RTG_ICP_MAX_LEVELS], *normal[RTG_ICP_MAX_LEVELS];
};
cudaError_t launch_icp_predict(const IcpPredict &prm, void *ws, cudaStream_t s);
cudaError_t launch_icp_pyramid(const IcpPyramid &prm, void *ws, cudaStream_t s);

// mapsurgery.cu
#ifndef RTG_SOA_MAX_ARRAYS
#define RTG_SOA_MAX_ARRAYS 16
#endif
size_t soa_compact_ws_bytes(int64_t P);
void launch_soa_compact(const uint8_t *mask, int invert, int64_t P, int n_arrays, const void *const *in, void *const *out,
                        const int32_t *words_per_row, uint32_t *n_kept, uint32_t *n_kept_host, void *ws, cudaStream_t s);
size_t knn_ws_bytes(int64_t n_ref);
int launch_knn(const float *query, int64_t n_query, const float *ref, int64_t n_ref, int K, int skip_self, float *out_d2, int32_t *out_idx,
               float *out_mean, void *ws, cudaStream_t s);

// adam.cu and the per-level icp.cu entry points are declared in their own sections of capi.cu
}  // namespace rtg
