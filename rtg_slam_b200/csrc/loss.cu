// Fused image-space kernels that sit between the rasterizer forward and backward in the mapping loop
// (SURVEY.md section 8(f) #1): the masked L1 colour + depth loss with its gradients, and the per-pixel normal map.
//
// Reference arithmetic (eager PyTorch, ~25 kernels incl. autograd and two boolean-mask compactions):
//   SLAM/multiprocess/mapper.py:402-431,444-451  loss_update (colour L1 on render_mask, depth L1 on the valid mask)
//   utils/loss_utils.py:27-31                     l1_loss = abs(a - b).mean()
//   SLAM/render.py:130-133                        render_normal[:, idx > -1] = normal[idx[idx > -1]]
#include "common.cuh"
#include "prof.h"

namespace rtg {

struct LossStats {  // device scratch, 4 doubles: sum|dC|, n_colour_pixels, sum|dd|, n_depth_pixels
    double v[4];
};

__device__ __forceinline__ float sgn(float x) { return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f); }

__global__ void loss_init_kernel(LossStats *st) {
    if (threadIdx.x < 4) st->v[threadIdx.x] = 0.0;
}

// gt_color: (H,W,3) if channels_last else (3,H,W); gt_depth: (H,W); mask: (H,W) bytes or NULL
__global__ void __launch_bounds__(256) loss_reduce_kernel(const float *__restrict__ render, const float *__restrict__ depth,
                                                          const int *__restrict__ depth_index, const float *__restrict__ gt_color,
                                                          const float *__restrict__ gt_depth, const uint8_t *__restrict__ mask,
                                                          int N, int channels_last, float depth_error_max, LossStats *st) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < N; p += gridDim.x * blockDim.x) {
        const bool m = mask ? (mask[p] != 0) : true;
        if (m) {
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float g = channels_last ? gt_color[3 * (size_t)p + c] : gt_color[(size_t)c * N + p];
                s += fabsf(render[(size_t)c * N + p] - g);
            }
            a0 += (double)s;
            a1 += 1.0;
            const float gd = gt_depth[p];
            const float e = depth[p] - gd;
            if (depth_index[p] != -1 && gd > 0.f && e < depth_error_max) { a2 += (double)fabsf(e); a3 += 1.0; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o); a3 += __shfl_xor_sync(0xffffffffu, a3, o);
    }
    __shared__ double s_p[8][4];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { s_p[w][0] = a0; s_p[w][1] = a1; s_p[w][2] = a2; s_p[w][3] = a3; }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int k = 0; k < 8; k++) t += s_p[k][threadIdx.x];
        atomicAdd(&st->v[threadIdx.x], t);
    }
}

__global__ void __launch_bounds__(256) loss_grad_kernel(const float *__restrict__ render, const float *__restrict__ depth,
                                                        const int *__restrict__ depth_index, const float *__restrict__ gt_color,
                                                        const float *__restrict__ gt_depth, const uint8_t *__restrict__ mask,
                                                        int N, int channels_last, float depth_error_max, float color_weight,
                                                        float depth_weight, const LossStats *st, float *__restrict__ dL_dcolor,
                                                        float *__restrict__ dL_ddepth, float *__restrict__ loss_out) {
    const double nc = st->v[1], nd = st->v[3];
    const float gcs = nc > 0.0 ? (float)((double)color_weight / (3.0 * nc)) : 0.f;  // mean over 3 * n_colour elements
    const float gds = nd > 0.0 ? (float)((double)depth_weight / nd) : 0.f;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const double cl = nc > 0.0 ? st->v[0] / (3.0 * nc) : 0.0, dl = nd > 0.0 ? st->v[2] / nd : 0.0;
        loss_out[0] = (float)((double)color_weight * cl + (double)depth_weight * dl);
        loss_out[1] = (float)cl;
        loss_out[2] = (float)dl;
        loss_out[3] = (float)nd;
    }
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < N; p += gridDim.x * blockDim.x) {
        const bool m = mask ? (mask[p] != 0) : true;
        float g0 = 0.f, g1 = 0.f, g2 = 0.f, gd_ = 0.f;
        if (m) {
            const float t0 = channels_last ? gt_color[3 * (size_t)p + 0] : gt_color[p];
            const float t1 = channels_last ? gt_color[3 * (size_t)p + 1] : gt_color[(size_t)N + p];
            const float t2 = channels_last ? gt_color[3 * (size_t)p + 2] : gt_color[2 * (size_t)N + p];
            g0 = gcs * sgn(render[p] - t0);
            g1 = gcs * sgn(render[(size_t)N + p] - t1);
            g2 = gcs * sgn(render[2 * (size_t)N + p] - t2);
            const float gd = gt_depth[p];
            const float e = depth[p] - gd;
            if (depth_index[p] != -1 && gd > 0.f && e < depth_error_max) gd_ = gds * sgn(e);
        }
        dL_dcolor[p] = g0; dL_dcolor[(size_t)N + p] = g1; dL_dcolor[2 * (size_t)N + p] = g2;
        dL_ddepth[p] = gd_;
    }
}

// out (3,H,W) = normal[idx] where idx > -1, zeros elsewhere
__global__ void __launch_bounds__(256) normal_map_kernel(const float *__restrict__ normal, const int *__restrict__ depth_index, int N,
                                                         float *__restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const int id = depth_index[p];
    float x = 0.f, y = 0.f, z = 0.f;
    if (id > -1) { x = normal[3 * (size_t)id]; y = normal[3 * (size_t)id + 1]; z = normal[3 * (size_t)id + 2]; }
    out[p] = x; out[(size_t)N + p] = y; out[2 * (size_t)N + p] = z;
}

size_t loss_ws_bytes() { return sizeof(LossStats); }

void launch_loss_l1(const float *render, const float *depth, const int *depth_index, const float *gt_color, const float *gt_depth,
                    const uint8_t *mask, int H, int W, int channels_last, float color_weight, float depth_weight,
                    float depth_error_max, float *dL_dcolor, float *dL_ddepth, float *loss_out, void *ws, cudaStream_t s) {
    ProfScope ps(K_ICP_MISC, s);
    LossStats *st = reinterpret_cast<LossStats *>(ws);
    const int N = H * W;
    int nb = (N + 255) / 256;
    if (nb > 148 * 8) nb = 148 * 8;
    loss_init_kernel<<<1, 32, 0, s>>>(st);
    loss_reduce_kernel<<<nb, 256, 0, s>>>(render, depth, depth_index, gt_color, gt_depth, mask, N, channels_last, depth_error_max, st);
    loss_grad_kernel<<<nb, 256, 0, s>>>(render, depth, depth_index, gt_color, gt_depth, mask, N, channels_last, depth_error_max,
                                        color_weight, depth_weight, st, dL_dcolor, dL_ddepth, loss_out);
}

void launch_normal_map(const float *normal, const int *depth_index, int H, int W, float *out, cudaStream_t s) {
    ProfScope ps(K_ICP_MISC, s);
    const int N = H * W;
    normal_map_kernel<<<(N + 255) / 256, 256, 0, s>>>(normal, depth_index, N, out);
}

}  // namespace rtg
