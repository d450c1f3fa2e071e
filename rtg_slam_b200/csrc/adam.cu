// One-launch Adam over all parameter groups of the Gaussian map.
//
// Replaces torch.optim.Adam(l, lr=0.0, eps=1e-15).step() over the six groups of
// GaussianPointCloud.parametrize (SLAM/gaussian_pointcloud.py:245-284, SLAM/multiprocess/
// mapper.py:156,452): bias-corrected, no weight decay, no amsgrad. Pure streaming work:
// 28 bytes per parameter (read p,g,m,v; write p,m,v), 128-bit accesses.
#include "common.cuh"
#include "prof.h"
#include "../../include/rtg_splat_b200.h"

namespace rtg {

struct AdamArgs {
    RtgAdamGroup g[RTG_ADAM_MAX_GROUPS];
    int n_groups;
    float beta1, beta2, eps, bc1, bc2_sqrt;
};

__global__ void __launch_bounds__(256) adam_kernel(const AdamArgs a) {
    const RtgAdamGroup grp = a.g[blockIdx.y];
    if (grp.grad == nullptr || grp.numel <= 0) return;
    const float lr_over_bc1 = grp.lr / a.bc1;
    const long long n = grp.numel;
    const bool vec = ((((uintptr_t)grp.param | (uintptr_t)grp.grad | (uintptr_t)grp.exp_avg | (uintptr_t)grp.exp_avg_sq) & 15) == 0);
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        const long long n4 = n >> 2;
        float4 *P4 = reinterpret_cast<float4 *>(grp.param);
        const float4 *G4 = reinterpret_cast<const float4 *>(grp.grad);
        float4 *M4 = reinterpret_cast<float4 *>(grp.exp_avg);
        float4 *V4 = reinterpret_cast<float4 *>(grp.exp_avg_sq);
        for (long long i = tid; i < n4; i += stride) {
            float4 p = P4[i], m = M4[i], v = V4[i];
            const float4 g = __ldg(G4 + i);
            adam_one(p.x, g.x, m.x, v.x, lr_over_bc1, a.beta1, a.beta2, a.eps, a.bc2_sqrt);
            adam_one(p.y, g.y, m.y, v.y, lr_over_bc1, a.beta1, a.beta2, a.eps, a.bc2_sqrt);
            adam_one(p.z, g.z, m.z, v.z, lr_over_bc1, a.beta1, a.beta2, a.eps, a.bc2_sqrt);
            adam_one(p.w, g.w, m.w, v.w, lr_over_bc1, a.beta1, a.beta2, a.eps, a.bc2_sqrt);
            P4[i] = p; M4[i] = m; V4[i] = v;
        }
        for (long long i = (n4 << 2) + tid; i < n; i += stride) {
            float p = grp.param[i], m = grp.exp_avg[i], v = grp.exp_avg_sq[i];
            adam_one(p, grp.grad[i], m, v, lr_over_bc1, a.beta1, a.beta2, a.eps, a.bc2_sqrt);
            grp.param[i] = p; grp.exp_avg[i] = m; grp.exp_avg_sq[i] = v;
        }
    } else {
        for (long long i = tid; i < n; i += stride) {
            float p = grp.param[i], m = grp.exp_avg[i], v = grp.exp_avg_sq[i];
            adam_one(p, grp.grad[i], m, v, lr_over_bc1, a.beta1, a.beta2, a.eps, a.bc2_sqrt);
            grp.param[i] = p; grp.exp_avg[i] = m; grp.exp_avg_sq[i] = v;
        }
    }
}

int launch_adam(const RtgAdamGroup *groups, int n_groups, float beta1, float beta2, float eps, int step, cudaStream_t s) {
    AdamArgs a;
    long long max_n = 0;
    for (int i = 0; i < n_groups; i++) {
        a.g[i] = groups[i];
        if (groups[i].grad != nullptr && groups[i].numel > max_n) max_n = groups[i].numel;
    }
    a.n_groups = n_groups;
    a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
    // torch computes the bias corrections in python doubles (torch/optim/adam.py, _single_tensor_adam)
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    a.bc1 = (float)bc1;
    a.bc2_sqrt = (float)sqrt(bc2);
    if (max_n == 0) return 0;
    long long blocks = (max_n / 4 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 16) blocks = 148 * 16;
    dim3 grid((unsigned)blocks, (unsigned)n_groups);
    ProfScope ps(K_ADAM, s);
    adam_kernel<<<grid, 256, 0, s>>>(a);
    return 0;
}

}  // namespace rtg
