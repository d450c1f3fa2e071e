// Per-frame consumers of the rasterizer's index / transmittance maps:
//  * accumulate_gaussian_error -- per-Gaussian max (or mean) of the colour / depth / normal error of the pixels whose
//    colour-index / depth-index map names the Gaussian, plus the outlier counter
//    (submodules/cuda_utils/map_process.cu:33-245, wrapper cuda_utils.cu:17-60);
//  * the tile-mask builders of SLAM/utils.py:681-734 (16x16 average pooling of a pixel mask or of the colour error).
// What differs by design: outputs are cleared by the same call (the reference allocates + fills 7 tensors with
// torch::full first), the float max is a single integer atomic (errors are >= 0 by construction; negative values take
// the unsigned-min path, so the result is the exact maximum for any input), the mean pass is fused behind the
// accumulation on the same stream, and the transmission mask is produced from T_map directly (render_mask = T != 1,
// mapper.py:503-505) together with the tile mask in one pass.
#include "common.cuh"
#include "kernels.h"
#include "prof.h"

namespace rtg {

__device__ __forceinline__ void atomic_max_float(float *addr, float v) {
    // IEEE order == signed-int order for v >= 0 and reversed unsigned order for v < 0 (initial value 0.0f)
    if (v >= 0.f) atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

__global__ void __launch_bounds__(256) gs_error_kernel(const int N, const int P, const float *__restrict__ color_err,
                                                       const float *__restrict__ depth_err, const float *__restrict__ normal_err,
                                                       const int *__restrict__ color_index, const int *__restrict__ depth_index,
                                                       const float thr_c, const float thr_d, const float thr_n, const bool check_max,
                                                       float *__restrict__ gs_color, float *__restrict__ gs_depth,
                                                       float *__restrict__ gs_normal, int *__restrict__ cnt_color,
                                                       int *__restrict__ cnt_depth, float *__restrict__ rescale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int ci = color_index[i], di = depth_index[i];
    if (ci >= 0 && ci < P) {
        const float e = color_err[i];
        if (check_max) atomic_max_float(gs_color + ci, e);
        else { atomicAdd(gs_color + ci, e); atomicAdd(cnt_color + ci, 1); }
        if (e > thr_c) atomicAdd(rescale + ci, 1.0f);
    }
    if (di >= 0 && di < P) {
        const float ed = depth_err[i], en = normal_err[i];
        if (check_max) { atomic_max_float(gs_depth + di, ed); atomic_max_float(gs_normal + di, en); }
        else { atomicAdd(gs_depth + di, ed); atomicAdd(gs_normal + di, en); atomicAdd(cnt_depth + di, 1); }
        // one counter, two increments of 1.0f: the sum of small integers is exact in fp32 and order-independent
        const float k = (ed > thr_d ? 1.0f : 0.f) + (en > thr_n ? 1.0f : 0.f);
        if (k > 0.f) atomicAdd(rescale + di, k);
    }
}

__global__ void __launch_bounds__(256) gs_error_mean_kernel(const int P, const int *__restrict__ cnt_color,
                                                            const int *__restrict__ cnt_depth, float *__restrict__ gs_color,
                                                            float *__restrict__ gs_depth, float *__restrict__ gs_normal) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int cc = cnt_color[i], cd = cnt_depth[i];  // the reference's depth and normal counters are always equal
    if (cc > 0) gs_color[i] = gs_color[i] / cc;
    if (cd > 0) { gs_depth[i] = gs_depth[i] / cd; gs_normal[i] = gs_normal[i] / cd; }
}

void launch_gs_error(int H, int W, int P, const float *color_err, const float *depth_err, const float *normal_err,
                     const int *color_index, const int *depth_index, float thr_c, float thr_d, float thr_n, bool check_max,
                     float *gs_color, float *gs_depth, float *gs_normal, float *rescale, int *counters, cudaStream_t s) {
    ProfScope ps(K_ICP_MISC, s);
    cudaMemsetAsync(gs_color, 0, sizeof(float) * (size_t)P, s);
    cudaMemsetAsync(gs_depth, 0, sizeof(float) * (size_t)P, s);
    cudaMemsetAsync(gs_normal, 0, sizeof(float) * (size_t)P, s);
    cudaMemsetAsync(rescale, 0, sizeof(float) * (size_t)P, s);
    if (!check_max) cudaMemsetAsync(counters, 0, sizeof(int) * 2 * (size_t)P, s);
    const int N = H * W;
    gs_error_kernel<<<(N + 255) / 256, 256, 0, s>>>(N, P, color_err, depth_err, normal_err, color_index, depth_index, thr_c, thr_d,
                                                    thr_n, check_max, gs_color, gs_depth, gs_normal, counters, counters + P, rescale);
    if (!check_max)
        gs_error_mean_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, counters, counters + P, gs_color, gs_depth, gs_normal);
}

// ------------------------------------------------------------------ tile pooling
// One CTA (256 threads) per 16x16 tile. `mode`: 0 = pixels is a float image, mean over the stride^2 cells of the
// zero-padded tile (F.avg_pool2d of SLAM/utils.py:697-704 / 714-723); 1 = pixels is T_map: the pooled quantity is
// (T != 1), which is also written to `pixel_mask` (mapper.py:503).
__global__ void __launch_bounds__(256) tile_pool_kernel(const int H, const int W, const float *__restrict__ pixels, const int mode,
                                                        const float ratio, uint8_t *__restrict__ pixel_mask,
                                                        float *__restrict__ tile_mean, int *__restrict__ tile_mask) {
    __shared__ float s_part[8];
    const int tx = blockIdx.x, ty = blockIdx.y;
    const int px = tx * RTG_TILE + (threadIdx.x & 15), py = ty * RTG_TILE + (threadIdx.x >> 4);
    float v = 0.f;
    if (px < W && py < H) {
        const float t = pixels[(size_t)py * W + px];
        if (mode == 1) {
            const bool m = (t != 1.0f);
            v = m ? 1.f : 0.f;
            if (pixel_mask) pixel_mask[(size_t)py * W + px] = m ? 1 : 0;
        } else {
            v = t;
        }
    }
    // fixed reduction order (row-major pairs -> warp tree -> 8 partials in order): deterministic
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++) sum += s_part[k];
        const float mean = sum / (float)(RTG_TILE * RTG_TILE);
        const int t = ty * gridDim.x + tx;
        if (tile_mean) tile_mean[t] = mean;
        if (tile_mask) tile_mask[t] = (mean > ratio) ? 1 : 0;
    }
}

void launch_tile_pool(int H, int W, const float *pixels, int mode, float ratio, uint8_t *pixel_mask, float *tile_mean, int *tile_mask,
                      cudaStream_t s) {
    ProfScope ps(K_ICP_MISC, s);
    dim3 grid((W + RTG_TILE - 1) / RTG_TILE, (H + RTG_TILE - 1) / RTG_TILE);
    tile_pool_kernel<<<grid, 256, 0, s>>>(H, W, pixels, mode, ratio, pixel_mask, tile_mean, tile_mask);
}

// colour error of Mapping.evaluate_render_range (mapper.py:481-487): sum_c |render - gt|, zero where the rendered pixel is
// black (render.sum == 0). render, gt: (3,H,W).
__global__ void __launch_bounds__(256) color_error_kernel(const int N, const float *__restrict__ render, const float *__restrict__ gt,
                                                          float *__restrict__ err) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float r0 = render[i], r1 = render[N + i], r2 = render[2 * N + i];
    const float e = (fabsf(r0 - gt[i]) + fabsf(r1 - gt[N + i])) + fabsf(r2 - gt[2 * N + i]);
    err[i] = ((r0 + r1) + r2 == 0.f) ? 0.f : e;
}

void launch_color_error(int H, int W, const float *render, const float *gt, float *err, cudaStream_t s) {
    ProfScope ps(K_ICP_MISC, s);
    const int N = H * W;
    color_error_kernel<<<(N + 255) / 256, 256, 0, s>>>(N, render, gt, err);
}

}  // namespace rtg
