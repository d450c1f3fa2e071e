// C ABI of librtg_splat_b200.so -- see include/rtg_splat_b200.h for the contract of every entry point.
#include <cstdio>
#include <cstring>
#include <cmath>
#include <string>

#include "../../include/rtg_splat_b200.h"
#include "common.cuh"
#include "kernels.h"
#include "prof.h"
#include <vector>
#include <mutex>
#include <utility>

namespace rtg {
int launch_adam(const RtgAdamGroup *groups, int n_groups, float beta1, float beta2, float eps, int step, cudaStream_t s);
void launch_map_adam_step(const RtgMapStep &st, cudaStream_t s);
void launch_history_merge(const RtgHistoryMerge &m, cudaStream_t s);
void launch_map_activate(int P, const float *scaling_raw, const float *rotation_raw, const float *opacity_raw, float *scales_out,
                         float *rotations_out, float *opacities_out, float *normal_out, cudaStream_t s);
size_t icp_ws_bytes();
void launch_icp_build_level(const float *depth, int H, int W, int pool, float fx, float fy, float cx, float cy, float *vertex,
                            float *normal, void *ws, cudaStream_t s);
void launch_icp_solve_level(const float *v0, const float *n0, const float *v1, const float *n1, int H, int W, float fx, float fy,
                            float cx, float cy, float dist_thr, float cos_thr, float damping, int iters, float *pose,
                            float *valid_ratio, void *ws, cudaStream_t s);
void launch_icp_p2p(const float *v_t0, const float *v_t1, const float *n_t0, int H, int W, const float *pose, float *loss,
                    void *ws, cudaStream_t s);
size_t loss_ws_bytes();
size_t ssim_ws_bytes(int C, int H, int W);
void launch_ssim_loss(const float *img1, const float *img2, int C, int H, int W, float *dL_dimg1, float *loss_out, void *ws,
                      cudaStream_t s);
void launch_loss_l1(const float *render, const float *depth, const int *depth_index, const float *gt_color, const float *gt_depth,
                    const uint8_t *mask, int H, int W, int channels_last, float color_weight, float depth_weight,
                    float depth_error_max, float *dL_dcolor, float *dL_ddepth, float *loss_out, void *ws, cudaStream_t s);
void launch_loss_mapping(const float *render, const float *depth, const float *render_normal, const int *depth_index,
                         const float *gt_color, const float *gt_depth, const float *gt_normal, const uint8_t *mask, int H, int W,
                         int channels_last, float color_weight, float depth_weight, float normal_weight, float depth_error_max,
                         float *dL_dcolor, float *dL_ddepth, float *dL_dnormal, float *loss_out, void *ws, cudaStream_t s);
void launch_normal_map(const float *normal, const int *depth_index, int H, int W, float *out, cudaStream_t s);
void launch_icp_fill(float *render_depth, const float *frame_depth, const float *rn, const float *fn, int H, int W, float dthr,
                     float nthr, cudaStream_t s);
}  // namespace rtg


// ---------------------------------------------------------------- event profiler
namespace rtg {
struct ProfRec { int id; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::mutex g_prof_mu;
static std::vector<cudaEvent_t> g_prof_pool;
static std::vector<ProfRec> g_prof_recs;
static thread_local cudaEvent_t t_open[K_COUNT];

static cudaEvent_t prof_get_event() {
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
void prof_begin(int id, cudaStream_t s) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEvent_t e = prof_get_event();
    cudaEventRecord(e, s);
    t_open[id] = e;
}
void prof_end(int id, cudaStream_t s) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEvent_t e = prof_get_event();
    cudaEventRecord(e, s);
    g_prof_recs.push_back({id, t_open[id], e});
}
static const char *kKernelNames[K_COUNT] = {"preprocess_fwd", "tile_scan", "scatter", "tile_sort", "render_fwd", "render_bwd",
                                            "preprocess_bwd", "adam", "icp_build_level", "icp_iter", "image_glue", "bwd_zero"};
}  // namespace rtg

static thread_local std::string g_err;

// one non-blocking side stream (+ fork / join events) per device, created on first use
struct SideStream {
    cudaStream_t stream;
    cudaEvent_t fork, join;
    std::mutex use;  // held across record(fork) .. wait(join): two host threads running a backward on the same device must not
                     // re-record the shared events between another thread's record and wait
};
static SideStream *side_stream() {
    static std::mutex mu;
    static SideStream table[64];
    static bool made[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    if (!made[dev]) {
        if (cudaStreamCreateWithFlags(&table[dev].stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        cudaEventCreateWithFlags(&table[dev].fork, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&table[dev].join, cudaEventDisableTiming);
        made[dev] = true;
    }
    return &table[dev];
}

static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

static int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(RTG_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return RTG_OK;
}

static rtg::ViewParams make_view(const RtgSplatView *v) {
    rtg::ViewParams p;
    p.H = v->image_height;
    p.W = v->image_width;
    p.tiles_x = (p.W + RTG_TILE - 1) / RTG_TILE;
    p.tiles_y = (p.H + RTG_TILE - 1) / RTG_TILE;
    p.tanfovx = v->tanfovx;
    p.tanfovy = v->tanfovy;
    // rasterizer_impl.cu:244-245
    p.focal_y = p.H / (2.0f * v->tanfovy);
    p.focal_x = p.W / (2.0f * v->tanfovx);
    p.cx = v->cx;
    p.cy = v->cy;
    p.scale_modifier = v->scale_modifier;
    p.color_sigma = v->color_sigma;
    p.opaque_thr = v->opaque_threshold;
    p.depth_thr = v->depth_threshold;
    p.normal_thr = v->normal_threshold;
    p.T_thr = v->T_threshold;
    p.sh_degree = v->sh_degree;
    p.prefiltered = v->prefiltered;
    p.view = v->viewmatrix;
    p.proj = v->projmatrix;
    p.campos = v->campos;
    p.bg = v->bg;
    p.row_begin = 0;
    p.row_end = p.tiles_y;
    return p;
}

// rows of a LOCAL parameter array addressed with the global Gaussian id: base - p_begin rows (integer arithmetic: the
// shifted value is only ever dereferenced at ids >= p_begin)
template <typename T>
static const T *shift_rows(const T *base, int64_t p_begin, size_t elems_per_row) {
    return base ? reinterpret_cast<const T *>(reinterpret_cast<uintptr_t>(base) - (uintptr_t)p_begin * elems_per_row * sizeof(T)) : nullptr;
}
template <typename T>
static T *shift_rows(T *base, int64_t p_begin, size_t elems_per_row) {
    return base ? reinterpret_cast<T *>(reinterpret_cast<uintptr_t>(base) - (uintptr_t)p_begin * elems_per_row * sizeof(T)) : nullptr;
}

extern "C" {

const char *rtg_last_error(void) { return g_err.c_str(); }
int rtg_version(void) { return 100; }

int rtg_splat_workspace_bytes(int32_t P, int32_t H, int32_t W, int64_t R_cap, size_t *geom_bytes, size_t *img_bytes,
                              size_t *bin_bytes) {
    if (P < 0 || H <= 0 || W <= 0 || R_cap < 0) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_splat_workspace_bytes: bad sizes");
    const size_t T = (size_t)((W + RTG_TILE - 1) / RTG_TILE) * ((H + RTG_TILE - 1) / RTG_TILE);
    size_t g = 0, i = 0, b = 0;
    rtg::geom_from(nullptr, (size_t)P, &g);
    rtg::img_from(nullptr, (size_t)H * W, &i);
    rtg::bin_from(nullptr, T, (size_t)R_cap, &b);
    if (geom_bytes) *geom_bytes = g + 256;
    if (img_bytes) *img_bytes = i + 256;
    if (bin_bytes) *bin_bytes = b + 256;
    return RTG_OK;
}

static int validate_view(const RtgSplatView *v, const char *who) {
    if (!v) return fail(RTG_ERR_INVALID_ARGUMENT, std::string(who) + ": view is NULL");
    if (v->image_height <= 0 || v->image_width <= 0) return fail(RTG_ERR_INVALID_ARGUMENT, std::string(who) + ": bad image size");
    if (!v->viewmatrix || !v->projmatrix || !v->campos || !v->bg)
        return fail(RTG_ERR_INVALID_ARGUMENT, std::string(who) + ": view matrices / campos / bg must be device pointers");
    if (v->sh_degree < 0 || v->sh_degree > 3) return fail(RTG_ERR_INVALID_ARGUMENT, std::string(who) + ": sh_degree must be 0..3");
    return RTG_OK;
}

static int validate_inputs(const char *who, const RtgSplatView *view, int32_t n, int32_t M, const float *means3D, const float *shs,
                           const float *colors_precomp, const float *opacities, const float *scales, const float *rotations,
                           const float *cov3D_precomp, const int32_t *radii) {
    if (n <= 0) return RTG_OK;
    if (!means3D || !opacities || !radii) return fail(RTG_ERR_INVALID_ARGUMENT, std::string(who) + ": NULL means3D / opacities / radii");
    // same exactly-one-of rules as GaussianRasterizer.forward (__init__.py:335-347)
    if ((shs == nullptr) == (colors_precomp == nullptr))
        return fail(RTG_ERR_INVALID_ARGUMENT, "Please provide excatly one of either SHs or precomputed colors!");
    if (((scales == nullptr || rotations == nullptr) && cov3D_precomp == nullptr) ||
        ((scales != nullptr || rotations != nullptr) && cov3D_precomp != nullptr))
        return fail(RTG_ERR_INVALID_ARGUMENT, "Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
    if (shs && (M < (view->sh_degree + 1) * (view->sh_degree + 1) || M > 16))
        return fail(RTG_ERR_INVALID_ARGUMENT, std::string(who) + ": M must hold (sh_degree+1)^2 coefficients and be <= 16");
    if ((((uintptr_t)rotations) & 15) || (M == 16 && (((uintptr_t)shs) & 15)))
        return fail(RTG_ERR_INVALID_ARGUMENT, std::string(who) + ": rotations / shs must be 16-byte aligned");
    return RTG_OK;
}

// clear of the binning counters + forward preprocess of the Gaussians [p_begin, p_end); `fused_histogram`: the tile
// histogram of these Gaussians is built in the same pass (single-GPU forward)
static int forward_preprocess(const rtg::ViewParams &vp, int32_t p_begin, int32_t p_end, int32_t M, const float *means3D,
                              const float *shs, const float *colors_precomp, const float *opacities, const float *scales,
                              const float *rotations, const float *cov3D_precomp, const int32_t *tile_mask, const rtg::GeomState &g,
                              const rtg::BinState &b, int32_t *radii, bool fused_histogram, cudaStream_t s) {
    // tile_count, tile_fill, tile_touched and vis_count are adjacent (bin_from): one clear
    cudaError_t e = cudaMemsetAsync(b.tile_count, 0, (size_t)((char *)b.tile_offset - (char *)b.tile_count), s);
    if (e != cudaSuccess) return fail(RTG_ERR_CUDA, std::string("rtg_splat_forward memset: ") + cudaGetErrorString(e));
    rtg::launch_preprocess_fwd(vp, p_end, M, means3D, scales, rotations, opacities, shs, colors_precomp, cov3D_precomp, tile_mask, g,
                               radii, fused_histogram ? b.tile_count : nullptr, b.tile_touched, b.vis_count, p_begin, s);
    return RTG_OK;
}

// scan -> scatter -> sort -> compositing over the records of all P Gaussians
static int forward_bin_render(const rtg::ViewParams &vp, int32_t P, const int32_t *tile_mask, const rtg::GeomState &g,
                              const rtg::ImgState &img, const rtg::BinState &b, int64_t R_cap, float *out_color, float *out_depth,
                              int32_t *out_hit_color, int32_t *out_hit_depth, float *out_hit_color_weight, float *out_hit_depth_weight,
                              float *out_T, const int32_t *radii, int32_t *counters, int32_t *counters_host, void *scan_done_event,
                              cudaStream_t s) {
    const int T = vp.tiles_x * vp.tiles_y;
    rtg::launch_tile_scan(b, T, R_cap, counters, counters_host, s);
    if (scan_done_event) {
        cudaError_t e = cudaEventRecord(reinterpret_cast<cudaEvent_t>(scan_done_event), s);
        if (e != cudaSuccess) return fail(RTG_ERR_CUDA, std::string("rtg_splat_forward event: ") + cudaGetErrorString(e));
    }
    rtg::launch_scatter(vp, P, g, radii, tile_mask, b, R_cap, counters, s);
    rtg::launch_tile_sort(b, T, counters, s);
    rtg::launch_render_fwd(vp, g, b, img, counters, out_color, out_depth, out_hit_color, out_hit_depth, out_hit_color_weight,
                           out_hit_depth_weight, out_T, s);
    return RTG_OK;
}

int rtg_splat_forward(const RtgSplatView *view, int32_t P, int32_t M, const float *means3D, const float *shs,
                      const float *colors_precomp, const float *opacities, const float *scales, const float *rotations,
                      const float *cov3D_precomp, const int32_t *tile_mask, void *geom_ws, void *img_ws, void *bin_ws,
                      int64_t R_cap, float *out_color, float *out_depth, int32_t *out_hit_color, int32_t *out_hit_depth,
                      float *out_hit_color_weight, float *out_hit_depth_weight, float *out_T, int32_t *radii, int32_t *counters,
                      int32_t *counters_host, void *scan_done_event, void *stream) {
    int rc = validate_view(view, "rtg_splat_forward");
    if (rc) return rc;
    if (P < 0) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_splat_forward: P < 0");
    if (!out_color || !out_depth || !out_hit_color || !out_hit_depth || !out_hit_color_weight || !out_hit_depth_weight || !out_T ||
        !counters || !geom_ws || !img_ws || !bin_ws || !tile_mask)
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_splat_forward: NULL output / workspace / tile_mask pointer");
    rc = validate_inputs("rtg_splat_forward", view, P, M, means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp, radii);
    if (rc) return rc;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const rtg::ViewParams vp = make_view(view);
    const int T = vp.tiles_x * vp.tiles_y;
    rtg::GeomState g = rtg::geom_from(geom_ws, (size_t)P);
    rtg::ImgState img = rtg::img_from(img_ws, (size_t)vp.H * vp.W);
    rtg::BinState b = rtg::bin_from(bin_ws, (size_t)T, (size_t)R_cap);
    rc = forward_preprocess(vp, 0, P, M, means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp, tile_mask, g, b, radii,
                            true, s);
    if (rc) return rc;
    rc = forward_bin_render(vp, P, tile_mask, g, img, b, R_cap, out_color, out_depth, out_hit_color, out_hit_depth, out_hit_color_weight,
                            out_hit_depth_weight, out_T, radii, counters, counters_host, scan_done_event, s);
    if (rc) return rc;
    return check_launch("rtg_splat_forward");
}

int rtg_splat_geom_layout(int32_t P, size_t *splat_offset, size_t *rgb_offset, size_t *hit_offset, size_t *vis_list_offset) {
    if (P < 0) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_splat_geom_layout: P < 0");
    rtg::GeomState g = rtg::geom_from(nullptr, (size_t)P);
    if (splat_offset) *splat_offset = (size_t)reinterpret_cast<char *>(g.splat);
    if (rgb_offset) *rgb_offset = (size_t)reinterpret_cast<char *>(g.rgb_flags);
    if (hit_offset) *hit_offset = (size_t)reinterpret_cast<char *>(g.hit);
    if (vis_list_offset) *vis_list_offset = (size_t)reinterpret_cast<char *>(g.vis_list);
    return RTG_OK;
}

int rtg_splat_forward_preprocess(const RtgSplatView *view, int32_t P, int32_t p_begin, int32_t p_end, int32_t M, const float *means3D,
                                 const float *shs, const float *colors_precomp, const float *opacities, const float *scales,
                                 const float *rotations, const float *cov3D_precomp, void *geom_ws, void *bin_ws, int64_t R_cap,
                                 int32_t *radii, void *stream) {
    int rc = validate_view(view, "rtg_splat_forward_preprocess");
    if (rc) return rc;
    if (P < 0 || p_begin < 0 || p_end < p_begin || p_end > P) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_splat_forward_preprocess: bad range");
    if (!geom_ws || !bin_ws || (P > 0 && !radii)) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_splat_forward_preprocess: NULL workspace / radii");
    rc = validate_inputs("rtg_splat_forward_preprocess", view, p_end - p_begin, M, means3D, shs, colors_precomp, opacities, scales,
                         rotations, cov3D_precomp, radii);
    if (rc) return rc;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const rtg::ViewParams vp = make_view(view);
    const int T = vp.tiles_x * vp.tiles_y;
    rtg::GeomState g = rtg::geom_from(geom_ws, (size_t)P);
    rtg::BinState b = rtg::bin_from(bin_ws, (size_t)T, (size_t)R_cap);
    rc = forward_preprocess(vp, p_begin, p_end, M, shift_rows(means3D, p_begin, 3), shift_rows(shs, p_begin, (size_t)3 * M),
                            shift_rows(colors_precomp, p_begin, 3), shift_rows(opacities, p_begin, 1), shift_rows(scales, p_begin, 3),
                            shift_rows(rotations, p_begin, 4), shift_rows(cov3D_precomp, p_begin, 6), nullptr, g, b, radii, false, s);
    if (rc) return rc;
    return check_launch("rtg_splat_forward_preprocess");
}

int rtg_splat_forward_render(const RtgSplatView *view, int32_t P, const int32_t *tile_mask, void *geom_ws, void *img_ws, void *bin_ws,
                             int64_t R_cap, float *out_color, float *out_depth, int32_t *out_hit_color, int32_t *out_hit_depth,
                             float *out_hit_color_weight, float *out_hit_depth_weight, float *out_T, const int32_t *radii,
                             int32_t *counters, int32_t *counters_host, void *scan_done_event, int32_t tile_row_begin,
                             int32_t tile_row_end, void *stream) {
    int rc = validate_view(view, "rtg_splat_forward_render");
    if (rc) return rc;
    if (P < 0) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_splat_forward_render: P < 0");
    if (!out_color || !out_depth || !out_hit_color || !out_hit_depth || !out_hit_color_weight || !out_hit_depth_weight || !out_T ||
        !counters || !geom_ws || !img_ws || !bin_ws || !tile_mask || (P > 0 && !radii))
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_splat_forward_render: NULL output / workspace / tile_mask / radii pointer");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    rtg::ViewParams vp = make_view(view);
    if (tile_row_end > tile_row_begin) {  // the rows that hold this rank's tiles: rectangles are clipped to them before expansion
        vp.row_begin = tile_row_begin < 0 ? 0 : tile_row_begin;
        vp.row_end = tile_row_end > vp.tiles_y ? vp.tiles_y : tile_row_end;
    }
    const int T = vp.tiles_x * vp.tiles_y;
    rtg::GeomState g = rtg::geom_from(geom_ws, (size_t)P);
    rtg::ImgState img = rtg::img_from(img_ws, (size_t)vp.H * vp.W);
    rtg::BinState b = rtg::bin_from(bin_ws, (size_t)T, (size_t)R_cap);
    // a re-run after a capacity overflow must start from a clean histogram / cursors (vis_count stays)
    cudaError_t e = cudaMemsetAsync(b.tile_count, 0, (size_t)((char *)b.vis_count - (char *)b.tile_count), s);
    if (e != cudaSuccess) return fail(RTG_ERR_CUDA, std::string("rtg_splat_forward_render memset: ") + cudaGetErrorString(e));
    rtg::launch_tile_histogram(vp, P, g, radii, tile_mask, b, s);
    rc = forward_bin_render(vp, P, tile_mask, g, img, b, R_cap, out_color, out_depth, out_hit_color, out_hit_depth, out_hit_color_weight,
                            out_hit_depth_weight, out_T, radii, counters, counters_host, scan_done_event, s);
    if (rc) return rc;
    return check_launch("rtg_splat_forward_render");
}

// phase: 0 = whole backward, 1 = zero-fill + compositing backward (fills the gradient records), 2 = per-Gaussian backward
static int splat_backward_impl(int phase, int32_t p_begin, int32_t p_end, const RtgSplatView *view, int32_t P, int32_t M, const float *means3D, const float *shs,
                       const float *colors_precomp, const float *scales, const float *rotations, const float *cov3D_precomp,
                       const int32_t *radii, const void *geom_ws, const void *img_ws, const void *bin_ws, int64_t R_cap,
                       const int32_t *counters, const float *final_T, const int32_t *hit_image, const float *dL_dcolor, const float *dL_ddepth, float *grad2d_scratch,
                       float *dL_dmeans3D, float *dL_dsh, float *dL_dcolors_precomp, float *dL_dopacity, float *dL_dscales,
                       float *dL_drotations, float *dL_dcov3D, float *dL_dmeans2D, void *stream) {
    int rc = validate_view(view, "rtg_splat_backward");
    if (rc) return rc;
    if (P < 0 || p_begin < 0 || p_end < p_begin || p_end > P) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_splat_backward: bad P / range");
    if (P == 0) return RTG_OK;
    if (!means3D || !radii || !geom_ws || !img_ws || !bin_ws || !counters || !final_T || !hit_image || !dL_dcolor || !dL_ddepth ||
        !grad2d_scratch || !dL_dmeans3D || !dL_dopacity)
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_splat_backward: NULL pointer");
    if ((shs == nullptr) == (colors_precomp == nullptr))
        return fail(RTG_ERR_INVALID_ARGUMENT, "Please provide excatly one of either SHs or precomputed colors!");
    if (shs && !dL_dsh) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_splat_backward: dL_dsh is NULL");
    if (colors_precomp && !dL_dcolors_precomp) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_splat_backward: dL_dcolors_precomp is NULL");
    if (cov3D_precomp == nullptr && (!scales || !rotations || !dL_dscales || !dL_drotations))
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_splat_backward: scales / rotations and their gradients are required");
    if (cov3D_precomp != nullptr && !dL_dcov3D) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_splat_backward: dL_dcov3D is NULL");
    if ((((uintptr_t)grad2d_scratch) & 15) || (((uintptr_t)dL_drotations) & 15) || (M == 16 && (((uintptr_t)dL_dsh) & 15)))
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_splat_backward: scratch / dL_drotations / dL_dsh must be 16-byte aligned");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const rtg::ViewParams vp = make_view(view);
    const int T = vp.tiles_x * vp.tiles_y;
    rtg::GeomState g = rtg::geom_from(const_cast<void *>(geom_ws), (size_t)P);
    rtg::ImgState img = rtg::img_from(const_cast<void *>(img_ws), (size_t)vp.H * vp.W);
    rtg::BinState b = rtg::bin_from(const_cast<void *>(bin_ws), (size_t)T, (size_t)R_cap);
    // counters live in the scan output: recompute the overflow word location is not needed -- the
    // backward reads the same device counters the forward wrote.
    // fork: the zero-fill of the culled rows only depends on the forward; it streams to HBM on a side stream while
    // the compute-bound render backward runs, and joins before the call's work on `s` ends
    if (phase == 3) {  // visible rows only: no zero fill
        rtg::launch_render_bwd(vp, g, b, img, counters, final_T, hit_image, dL_dcolor, dL_ddepth, grad2d_scratch, s);
    } else if (phase != 2) {
        SideStream *ss = side_stream();
        std::unique_lock<std::mutex> side_lock;
        if (ss) side_lock = std::unique_lock<std::mutex>(ss->use);
        cudaError_t e = ss ? cudaEventRecord(ss->fork, s) : cudaErrorUnknown;
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ss->stream, ss->fork, 0);
        const bool forked = (e == cudaSuccess);
        rtg::launch_bwd_zero(p_begin, p_end, M, shs != nullptr, cov3D_precomp == nullptr, radii, shift_rows(dL_dmeans3D, p_begin, 3),
                             shift_rows(dL_dsh, p_begin, (size_t)3 * M), shift_rows(dL_dcolors_precomp, p_begin, 3),
                             shift_rows(dL_dopacity, p_begin, 1), shift_rows(dL_dscales, p_begin, 3), shift_rows(dL_drotations, p_begin, 4),
                             shift_rows(dL_dcov3D, p_begin, 6), shift_rows(dL_dmeans2D, p_begin, 3), forked ? ss->stream : s);
        if (forked) cudaEventRecord(ss->join, ss->stream);
        rtg::launch_render_bwd(vp, g, b, img, counters, final_T, hit_image, dL_dcolor, dL_ddepth, grad2d_scratch, s);
        if (forked) cudaStreamWaitEvent(s, ss->join, 0);  // enqueued after both: the zero-fill still overlaps render_bwd
    }
    if (phase != 1)  // walks the compact visible list, which holds only Gaussians of [p_begin, p_end)
        rtg::launch_preprocess_bwd(vp, p_end - p_begin, M, shift_rows(means3D, p_begin, 3), shift_rows(scales, p_begin, 3),
                                   shift_rows(rotations, p_begin, 4), shift_rows(shs, p_begin, (size_t)3 * M),
                                   shift_rows(cov3D_precomp, p_begin, 6), g, b.vis_count, grad2d_scratch,
                                   shift_rows(dL_dmeans3D, p_begin, 3), shift_rows(dL_dsh, p_begin, (size_t)3 * M),
                                   shift_rows(dL_dcolors_precomp, p_begin, 3), shift_rows(dL_dopacity, p_begin, 1),
                                   shift_rows(dL_dscales, p_begin, 3), shift_rows(dL_drotations, p_begin, 4),
                                   shift_rows(dL_dcov3D, p_begin, 6), shift_rows(dL_dmeans2D, p_begin, 3), s);
    return check_launch("rtg_splat_backward");
}

#define RTG_BWD_PARAMS                                                                                                             \
    const RtgSplatView *view, int32_t P, int32_t M, const float *means3D, const float *shs, const float *colors_precomp,          \
        const float *scales, const float *rotations, const float *cov3D_precomp, const int32_t *radii, const void *geom_ws,        \
        const void *img_ws, const void *bin_ws, int64_t R_cap, const int32_t *counters, const float *final_T,                     \
        const int32_t *hit_image, const float *dL_dcolor, const float *dL_ddepth, float *grad2d_scratch, float *dL_dmeans3D,       \
        float *dL_dsh, float *dL_dcolors_precomp, float *dL_dopacity, float *dL_dscales, float *dL_drotations, float *dL_dcov3D,  \
        float *dL_dmeans2D, void *stream
#define RTG_BWD_ARGS                                                                                                               \
    view, P, M, means3D, shs, colors_precomp, scales, rotations, cov3D_precomp, radii, geom_ws, img_ws, bin_ws, R_cap, counters,  \
        final_T, hit_image, dL_dcolor, dL_ddepth, grad2d_scratch, dL_dmeans3D, dL_dsh, dL_dcolors_precomp, dL_dopacity, dL_dscales, \
        dL_drotations, dL_dcov3D, dL_dmeans2D, stream
int rtg_splat_backward(RTG_BWD_PARAMS) { return splat_backward_impl(0, 0, P, RTG_BWD_ARGS); }
int rtg_splat_backward_visible(RTG_BWD_PARAMS) { return splat_backward_impl(3, 0, P, RTG_BWD_ARGS); }
int rtg_splat_backward_render(RTG_BWD_PARAMS) { return splat_backward_impl(1, 0, P, RTG_BWD_ARGS); }
int rtg_splat_backward_finish(RTG_BWD_PARAMS) { return splat_backward_impl(2, 0, P, RTG_BWD_ARGS); }
int rtg_splat_backward_render_shard(int32_t p_begin, int32_t p_end, RTG_BWD_PARAMS) {
    return splat_backward_impl(1, p_begin, p_end, RTG_BWD_ARGS);
}
int rtg_splat_backward_finish_shard(int32_t p_begin, int32_t p_end, RTG_BWD_PARAMS) {
    return splat_backward_impl(2, p_begin, p_end, RTG_BWD_ARGS);
}

int rtg_splat_mark_visible(int32_t P, const float *means3D, const float *viewmatrix, const float *projmatrix, uint8_t *present,
                           void *stream) {
    if (P < 0) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_splat_mark_visible: P < 0");
    if (P == 0) return RTG_OK;
    if (!means3D || !viewmatrix || !projmatrix || !present) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_splat_mark_visible: NULL pointer");
    rtg::launch_mark_visible(P, means3D, viewmatrix, projmatrix, present, reinterpret_cast<cudaStream_t>(stream));
    return check_launch("rtg_splat_mark_visible");
}

int rtg_adam_step(const RtgAdamGroup *groups, int32_t n_groups, float beta1, float beta2, float eps, int32_t step, void *stream) {
    if (!groups || n_groups <= 0 || n_groups > RTG_ADAM_MAX_GROUPS) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_adam_step: bad groups");
    if (step < 1) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_adam_step: step must be >= 1");
    for (int i = 0; i < n_groups; i++)
        if (groups[i].grad && (!groups[i].param || !groups[i].exp_avg || !groups[i].exp_avg_sq))
            return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_adam_step: NULL param / state pointer");
    rtg::launch_adam(groups, n_groups, beta1, beta2, eps, step, reinterpret_cast<cudaStream_t>(stream));
    return check_launch("rtg_adam_step");
}

int rtg_map_adam_step(const RtgMapStep *st, void *stream) {
    if (!st) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_map_adam_step: NULL step");
    if (st->P < 0 || st->step < 1) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_map_adam_step: P must be >= 0 and step >= 1");
    if (st->P == 0) return RTG_OK;
    const void *need[] = {st->xyz, st->sh, st->opacity_raw, st->scaling_raw, st->rotation_raw, st->m_xyz, st->m_sh, st->m_opacity,
                          st->m_scaling, st->m_rotation, st->v_xyz, st->v_sh, st->v_opacity, st->v_scaling, st->v_rotation,
                          st->g_means3D, st->g_sh, st->g_opacity, st->g_scales, st->g_rotations, st->scales_out, st->rotations_out,
                          st->opacities_out};
    for (const void *q : need)
        if (!q) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_map_adam_step: NULL parameter / state / gradient / output pointer");
    const void *al16[] = {st->sh, st->m_sh, st->v_sh, st->g_sh, st->rotation_raw, st->m_rotation, st->v_rotation, st->g_rotations,
                          st->rotations_out, st->rotation0};
    for (const void *q : al16)
        if (((uintptr_t)q) & 15) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_map_adam_step: sh / rotation tensors must be 16-byte aligned");
    if (st->attach_mask && (!st->xyz0 || !st->scaling0 || !st->rotation0))
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_map_adam_step: attach_mask needs xyz0, scaling0 and rotation0");
    if (st->attach_mask && st->attach_count < 0) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_map_adam_step: attach_count < 0");
    rtg::launch_map_adam_step(*st, reinterpret_cast<cudaStream_t>(stream));
    return check_launch("rtg_map_adam_step");
}

int rtg_map_activate(int32_t P, const float *scaling_raw, const float *rotation_raw, const float *opacity_raw, float *scales_out,
                     float *rotations_out, float *opacities_out, float *normal_out, void *stream) {
    if (P < 0) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_map_activate: P < 0");
    if (P == 0) return RTG_OK;
    if (!scaling_raw || !rotation_raw || !opacity_raw || !scales_out || !rotations_out || !opacities_out)
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_map_activate: NULL pointer");
    if ((((uintptr_t)rotation_raw) | ((uintptr_t)rotations_out)) & 15)
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_map_activate: rotation tensors must be 16-byte aligned");
    rtg::launch_map_activate(P, scaling_raw, rotation_raw, opacity_raw, scales_out, rotations_out, opacities_out, normal_out,
                             reinterpret_cast<cudaStream_t>(stream));
    return check_launch("rtg_map_activate");
}

int rtg_map_history_merge(const RtgHistoryMerge *m, void *stream) {
    if (!m) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_map_history_merge: merge is NULL");
    if (m->P < 0) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_map_history_merge: P < 0");
    if (m->P == 0 || !(m->max_weight > 0.f)) return RTG_OK;
    if (!m->hist_confidence || !m->confidence || !m->hist_xyz || !m->xyz || !m->hist_features_dc || !m->features_dc ||
        !m->hist_features_rest || !m->features_rest || !m->hist_scaling || !m->scaling || !m->hist_rotation || !m->rotation_raw)
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_map_history_merge: NULL pointer");
    if (m->features_rest_width < 0 || m->features_dc_stride < 3 || m->features_rest_stride < m->features_rest_width)
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_map_history_merge: bad feature strides");
    if ((((uintptr_t)m->hist_rotation) | ((uintptr_t)m->rotation_raw)) & 15)
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_map_history_merge: rotation tensors must be 16-byte aligned");
    rtg::launch_history_merge(*m, reinterpret_cast<cudaStream_t>(stream));
    return check_launch("rtg_map_history_merge");
}

size_t rtg_icp_workspace_bytes(int32_t H, int32_t W) {
    (void)H; (void)W;
    return rtg::icp_ws_bytes();
}

int rtg_icp_build_level(const float *depth, int32_t H, int32_t W, int32_t pool, float fx, float fy, float cx, float cy,
                        float *vertex_out, float *normal_out, void *ws, void *stream) {
    if (!depth || !vertex_out || !normal_out || !ws || H <= 0 || W <= 0 || pool < 1 || H / pool < 1 || W / pool < 1)
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_icp_build_level: bad arguments");
    rtg::launch_icp_build_level(depth, H, W, pool, fx, fy, cx, cy, vertex_out, normal_out, ws, reinterpret_cast<cudaStream_t>(stream));
    return check_launch("rtg_icp_build_level");
}

int rtg_icp_solve_level(const float *vertex0, const float *normal0, const float *vertex1, const float *normal1, int32_t H, int32_t W,
                        float fx, float fy, float cx, float cy, float distance_threshold, float normal_cos_threshold, float damping,
                        int32_t iters, float *pose, float *valid_ratio, void *ws, void *stream) {
    if (!vertex0 || !normal0 || !vertex1 || !normal1 || !pose || !ws || H <= 0 || W <= 0 || iters < 0)
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_icp_solve_level: bad arguments");
    rtg::launch_icp_solve_level(vertex0, normal0, vertex1, normal1, H, W, fx, fy, cx, cy, distance_threshold, normal_cos_threshold,
                                damping, iters, pose, valid_ratio, ws, reinterpret_cast<cudaStream_t>(stream));
    return check_launch("rtg_icp_solve_level");
}

int rtg_icp_build_pyramid(const float *depth, int32_t H, int32_t W, int32_t n_levels, const int32_t *pools, const float *fx,
                          const float *fy, const float *cx, const float *cy, float *const *vertex_out, float *const *normal_out,
                          void *ws, void *stream) {
    if (!depth || !pools || !fx || !fy || !cx || !cy || !vertex_out || !normal_out || !ws || H <= 0 || W <= 0 || n_levels < 1 ||
        n_levels > RTG_ICP_MAX_LEVELS)
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_icp_build_pyramid: bad arguments");
    rtg::IcpPyramid p;
    p.depth = depth; p.H = H; p.W = W; p.n_levels = n_levels;
    for (int l = 0; l < n_levels; l++) {
        if (pools[l] < 1 || H / pools[l] < 1 || W / pools[l] < 1 || !vertex_out[l] || !normal_out[l])
            return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_icp_build_pyramid: bad level");
        p.pool[l] = pools[l]; p.fx[l] = fx[l]; p.fy[l] = fy[l]; p.cx[l] = cx[l]; p.cy[l] = cy[l];
        p.vertex[l] = vertex_out[l]; p.normal[l] = normal_out[l];
    }
    // the grid is sized for the finest level: order the levels so that the last one has the smallest pool
    int finest = 0;
    for (int l = 1; l < n_levels; l++)
        if (p.pool[l] < p.pool[finest]) finest = l;
    if (finest != n_levels - 1) {
        std::swap(p.pool[finest], p.pool[n_levels - 1]); std::swap(p.fx[finest], p.fx[n_levels - 1]);
        std::swap(p.fy[finest], p.fy[n_levels - 1]); std::swap(p.cx[finest], p.cx[n_levels - 1]);
        std::swap(p.cy[finest], p.cy[n_levels - 1]); std::swap(p.vertex[finest], p.vertex[n_levels - 1]);
        std::swap(p.normal[finest], p.normal[n_levels - 1]);
    }
    cudaError_t e = rtg::launch_icp_pyramid(p, ws, reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(RTG_ERR_CUDA, std::string("rtg_icp_build_pyramid: ") + cudaGetErrorString(e));
    return check_launch("rtg_icp_build_pyramid");
}

int rtg_icp_predict_pose(const RtgIcpLevel *levels, int32_t n_levels, float distance_threshold, float normal_cos_threshold,
                         float damping, const float *pose_init, const float *p2p_vertex_t0, const float *p2p_vertex_t1,
                         const float *p2p_normal_t0, int32_t p2p_H, int32_t p2p_W, float *out, float *out_host, void *ws,
                         void *stream) {
    if (!levels || n_levels < 1 || n_levels > RTG_ICP_MAX_LEVELS || !p2p_vertex_t0 || !p2p_vertex_t1 || !p2p_normal_t0 || p2p_H <= 0 ||
        p2p_W <= 0 || !out || !ws)
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_icp_predict_pose: bad arguments");
    rtg::IcpPredict p;
    p.n_levels = n_levels;
    for (int l = 0; l < n_levels; l++) {
        const RtgIcpLevel &L = levels[l];
        if (!L.vertex0 || !L.normal0 || !L.vertex1 || !L.normal1 || L.H <= 0 || L.W <= 0 || L.iters < 0)
            return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_icp_predict_pose: bad level");
        p.lv[l] = rtg::IcpLevel{L.vertex0, L.normal0, L.vertex1, L.normal1, L.H, L.W, L.fx, L.fy, L.cx, L.cy, L.iters};
    }
    p.dist_thr = distance_threshold; p.cos_thr = normal_cos_threshold; p.damping = damping;
    p.pose_in = pose_init;
    p.p2p_v_t0 = p2p_vertex_t0; p.p2p_v_t1 = p2p_vertex_t1; p.p2p_n_t0 = p2p_normal_t0; p.p2p_HW = p2p_H * p2p_W;
    p.out = out; p.out_host = out_host;
    cudaError_t e = rtg::launch_icp_predict(p, ws, reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(RTG_ERR_CUDA, std::string("rtg_icp_predict_pose: ") + cudaGetErrorString(e));
    return check_launch("rtg_icp_predict_pose");
}

int rtg_icp_point2plane_loss(const float *vertex_t0, const float *vertex_t1, const float *normal_t0, int32_t H, int32_t W,
                             const float *pose, float *loss, void *ws, void *stream) {
    if (!vertex_t0 || !vertex_t1 || !normal_t0 || !pose || !loss || !ws || H <= 0 || W <= 0)
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_icp_point2plane_loss: bad arguments");
    rtg::launch_icp_p2p(vertex_t0, vertex_t1, normal_t0, H, W, pose, loss, ws, reinterpret_cast<cudaStream_t>(stream));
    return check_launch("rtg_icp_point2plane_loss");
}

int rtg_icp_fill_model_depth(float *render_depth, const float *frame_depth, const float *render_normal, const float *frame_normal,
                             int32_t H, int32_t W, float distance_threshold, float normal_threshold, void *stream) {
    if (!render_depth || !frame_depth || !render_normal || !frame_normal || H <= 0 || W <= 0)
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_icp_fill_model_depth: bad arguments");
    rtg::launch_icp_fill(render_depth, frame_depth, render_normal, frame_normal, H, W, distance_threshold, normal_threshold,
                         reinterpret_cast<cudaStream_t>(stream));
    return check_launch("rtg_icp_fill_model_depth");
}

size_t rtg_loss_workspace_bytes(void) { return rtg::loss_ws_bytes(); }

int rtg_loss_l1(const float *render, const float *depth, const int32_t *depth_index, const float *gt_color, const float *gt_depth,
                const uint8_t *render_mask, int32_t H, int32_t W, int32_t gt_channels_last, float color_weight, float depth_weight,
                float depth_error_max, float *dL_dcolor, float *dL_ddepth, float *loss_out, void *ws, void *stream) {
    if (!render || !depth || !depth_index || !gt_color || !gt_depth || !dL_dcolor || !dL_ddepth || !loss_out || !ws || H <= 0 || W <= 0)
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_loss_l1: bad arguments");
    rtg::launch_loss_l1(render, depth, depth_index, gt_color, gt_depth, render_mask, H, W, gt_channels_last, color_weight, depth_weight,
                        depth_error_max, dL_dcolor, dL_ddepth, loss_out, ws, reinterpret_cast<cudaStream_t>(stream));
    return check_launch("rtg_loss_l1");
}

int rtg_loss_mapping(const float *render, const float *depth, const float *render_normal, const int32_t *depth_index,
                     const float *gt_color, const float *gt_depth, const float *gt_normal, const uint8_t *render_mask, int32_t H, int32_t W,
                     int32_t gt_channels_last, float color_weight, float depth_weight, float normal_weight, float depth_error_max,
                     float *dL_dcolor, float *dL_ddepth, float *dL_dnormal, float *loss_out, void *ws, void *stream) {
    if (!render || !depth || !depth_index || !gt_color || !gt_depth || !dL_dcolor || !dL_ddepth || !loss_out || !ws || H <= 0 || W <= 0)
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_loss_mapping: bad arguments");
    if (normal_weight > 0.f && (!render_normal || !gt_normal || !dL_dnormal))
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_loss_mapping: normal_weight > 0 needs render_normal, gt_normal and dL_dnormal");
    rtg::launch_loss_mapping(render, depth, render_normal, depth_index, gt_color, gt_depth, gt_normal, render_mask, H, W, gt_channels_last,
                             color_weight, depth_weight, normal_weight, depth_error_max, dL_dcolor, dL_ddepth, dL_dnormal, loss_out, ws,
                             reinterpret_cast<cudaStream_t>(stream));
    return check_launch("rtg_loss_mapping");
}

size_t rtg_ssim_workspace_bytes(int32_t C, int32_t H, int32_t W) {
    if (C <= 0 || H <= 0 || W <= 0) return 0;
    return rtg::ssim_ws_bytes(C, H, W);
}

int rtg_ssim_loss(const float *img1, const float *img2, int32_t C, int32_t H, int32_t W, float *dL_dimg1, float *loss_out, void *ws,
                  void *stream) {
    if (!img1 || !img2 || !loss_out || !ws || C <= 0 || C > 65535 || H <= 0 || W <= 0 || (H + 15) / 16 > 65535)
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_ssim_loss: bad arguments");
    rtg::launch_ssim_loss(img1, img2, C, H, W, dL_dimg1, loss_out, ws, reinterpret_cast<cudaStream_t>(stream));
    return check_launch("rtg_ssim_loss");
}

int rtg_normal_map(const float *normal, const int32_t *depth_index, int32_t H, int32_t W, float *out, void *stream) {
    if (!normal || !depth_index || !out || H <= 0 || W <= 0) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_normal_map: bad arguments");
    rtg::launch_normal_map(normal, depth_index, H, W, out, reinterpret_cast<cudaStream_t>(stream));
    return check_launch("rtg_normal_map");
}

int rtg_frame_preprocess(const float *depth_in, int32_t H, int32_t W, int32_t depth_filter, int32_t radius, float sigma_color,
                         float sigma_space, float min_depth, float max_depth, float fx, float fy, float cx, float cy,
                         float invalid_confidence_thresh, float *depth_out, float *vertex_out, float *normal_out,
                         float *confidence_out, uint8_t *invalid_mask_out, void *ws, void *stream) {
    const bool maps = vertex_out || normal_out || confidence_out || invalid_mask_out;
    if (!depth_in || !depth_out || H <= 0 || W <= 0 || (maps && (!vertex_out || !normal_out || !confidence_out || !ws)))
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_frame_preprocess: bad arguments");
    if (depth_filter && (radius < 0 || radius > 15 || !(sigma_color > 0.f) || !(sigma_space > 0.f)))
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_frame_preprocess: bad filter parameters");
    if (depth_in == depth_out && depth_filter) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_frame_preprocess: the filter is not in-place");
    rtg::launch_frame_preprocess(depth_in, H, W, depth_filter, radius, sigma_color, sigma_space, min_depth, max_depth, fx, fy, cx, cy,
                                 invalid_confidence_thresh, depth_out, vertex_out, normal_out, confidence_out, invalid_mask_out, ws,
                                 reinterpret_cast<cudaStream_t>(stream));
    return check_launch("rtg_frame_preprocess");
}

int rtg_accumulate_gaussian_error(int32_t H, int32_t W, int32_t P, const float *screen_color_error, const float *screen_depth_error,
                                  const float *screen_normal_error, const int32_t *screen_color_index,
                                  const int32_t *screen_depth_index, float color_threshold, float depth_threshold,
                                  float normal_threshold, int32_t check_max, float *gs_color_error, float *gs_depth_error,
                                  float *gs_normal_error, float *gs_rescale_counter, int32_t *counters, void *stream) {
    if (H < 0 || W < 0 || P < 0) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_accumulate_gaussian_error: negative size");
    if (P == 0) return RTG_OK;
    if (!gs_color_error || !gs_depth_error || !gs_normal_error || !gs_rescale_counter || (!check_max && !counters))
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_accumulate_gaussian_error: NULL output");
    if ((size_t)H * W > 0 && (!screen_color_error || !screen_depth_error || !screen_normal_error || !screen_color_index ||
                              !screen_depth_index))
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_accumulate_gaussian_error: NULL input");
    rtg::launch_gs_error(H, W, P, screen_color_error, screen_depth_error, screen_normal_error, screen_color_index, screen_depth_index,
                         color_threshold, depth_threshold, normal_threshold, check_max != 0, gs_color_error, gs_depth_error,
                         gs_normal_error, gs_rescale_counter, counters, reinterpret_cast<cudaStream_t>(stream));
    return check_launch("rtg_accumulate_gaussian_error");
}

int rtg_tile_mean(int32_t H, int32_t W, const float *pixels, float ratio, float *tile_mean, int32_t *tile_mask, void *stream) {
    if (!pixels || H <= 0 || W <= 0 || (!tile_mean && !tile_mask)) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_tile_mean: bad arguments");
    rtg::launch_tile_pool(H, W, pixels, 0, ratio, nullptr, tile_mean, tile_mask, reinterpret_cast<cudaStream_t>(stream));
    return check_launch("rtg_tile_mean");
}

int rtg_transmission_tile_mask(int32_t H, int32_t W, const float *T_map, float ratio, uint8_t *render_mask, int32_t *tile_mask,
                               void *stream) {
    if (!T_map || !tile_mask || H <= 0 || W <= 0) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_transmission_tile_mask: bad arguments");
    rtg::launch_tile_pool(H, W, T_map, 1, ratio, render_mask, nullptr, tile_mask, reinterpret_cast<cudaStream_t>(stream));
    return check_launch("rtg_transmission_tile_mask");
}

int rtg_color_error(int32_t H, int32_t W, const float *render, const float *gt, float *out, void *stream) {
    if (!render || !gt || !out || H <= 0 || W <= 0) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_color_error: bad arguments");
    rtg::launch_color_error(H, W, render, gt, out, reinterpret_cast<cudaStream_t>(stream));
    return check_launch("rtg_color_error");
}

size_t rtg_soa_compact_workspace_bytes(int64_t P) { return rtg::soa_compact_ws_bytes(P < 0 ? 0 : P); }

int rtg_soa_compact(const uint8_t *mask, int32_t invert, int64_t P, int32_t n_arrays, const void *const *in, void *const *out,
                    const int32_t *words_per_row, uint32_t *n_kept, uint32_t *n_kept_host, void *ws, void *stream) {
    if (P < 0 || P > 0x7fffffffLL || n_arrays < 0 || n_arrays > RTG_SOA_MAX_ARRAYS || !n_kept)
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_soa_compact: bad sizes");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (P == 0) {
        cudaMemsetAsync(n_kept, 0, 4, s);
        if (n_kept_host) *n_kept_host = 0;
        return check_launch("rtg_soa_compact");
    }
    if (!mask || !ws || (n_arrays > 0 && (!in || !out || !words_per_row))) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_soa_compact: NULL pointer");
    for (int i = 0; i < n_arrays; i++)
        if (!in[i] || !out[i] || words_per_row[i] < 1 || in[i] == out[i])
            return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_soa_compact: bad array (NULL, in place, or words_per_row < 1)");
    rtg::launch_soa_compact(mask, invert, P, n_arrays, in, out, words_per_row, n_kept, n_kept_host, ws, s);
    return check_launch("rtg_soa_compact");
}

size_t rtg_knn_workspace_bytes(int64_t n_ref) { return rtg::knn_ws_bytes(n_ref < 0 ? 0 : n_ref); }

int rtg_knn(const float *query, int64_t n_query, const float *ref, int64_t n_ref, int32_t K, int32_t skip_self, float *out_d2,
            int32_t *out_idx, float *out_mean, void *ws, void *stream) {
    if (n_query < 0 || n_ref < 0 || n_query > 0x7fffffffLL || n_ref > 0x7fffffffLL || K < 1 || K > 8)
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_knn: bad sizes (1 <= K <= 8)");
    if (n_query == 0) return RTG_OK;
    if (!query || !ws || (n_ref > 0 && !ref)) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_knn: NULL pointer");
    if (skip_self && (n_query != n_ref)) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_knn: skip_self needs query == ref");
    if (rtg::launch_knn(query, n_query, ref, n_ref, K, skip_self, out_d2, out_idx, out_mean, ws, reinterpret_cast<cudaStream_t>(stream)))
        return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_knn: unsupported K");
    return check_launch("rtg_knn");
}

int rtg_profile_enable(int32_t on) {
    std::lock_guard<std::mutex> lk(rtg::g_prof_mu);
    rtg::g_prof_on = on != 0;
    return RTG_OK;
}

int rtg_profile_kernel_count(void) { return rtg::K_COUNT; }

const char *rtg_profile_kernel_name(int32_t id) { return (id >= 0 && id < rtg::K_COUNT) ? rtg::kKernelNames[id] : ""; }

int rtg_profile_read(double *total_ms, int64_t *launches, int32_t reset) {
    if (!total_ms || !launches) return fail(RTG_ERR_INVALID_ARGUMENT, "rtg_profile_read: NULL output");
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return fail(RTG_ERR_CUDA, std::string("rtg_profile_read: ") + cudaGetErrorString(e));
    std::lock_guard<std::mutex> lk(rtg::g_prof_mu);
    for (int i = 0; i < rtg::K_COUNT; i++) { total_ms[i] = 0.0; launches[i] = 0; }
    for (auto &r : rtg::g_prof_recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { total_ms[r.id] += ms; launches[r.id] += 1; }
    }
    if (reset) {
        for (auto &r : rtg::g_prof_recs) { rtg::g_prof_pool.push_back(r.a); rtg::g_prof_pool.push_back(r.b); }
        rtg::g_prof_recs.clear();
    }
    return RTG_OK;
}

}  // extern "C"
