// Fused map-parameter step: activation backward + attach regulariser + Adam + activation forward in one pass.
//
// Per optimisation iteration the reference runs (SLAM/multiprocess/mapper.py:376-468, SLAM/gaussian_pointcloud.py:
// 245-284,511-523,574-581) torch.exp / F.normalize / torch.sigmoid / torch.cat forward, their autograd nodes backward,
// three masked l2 losses forward and backward, one multi-tensor Adam chain per parameter group and the confidence
// update: ~60 launches that stream the 59 floats of every Gaussian a dozen times. Here a warp owns 32 consecutive rows
// and touches every array exactly once with coalesced accesses: p, m, v read + written, the rasterizer gradient read
// only where radii > 0 (a culled row's gradient is zero, so the backward need not even write it).
//
// HBM bytes per Gaussian: 59 * 24 (p, m, v both ways) + 59 * 4 * visible fraction (g) + 32 (activated outputs)
// + 12 (normal) + 4 (radii) + 41 (attach: mask + xyz0 + scaling0 + rotation0), i.e. ~1.55 kB -- the reference's
// optimizer.step() alone moves 1.65 kB.
#include "common.cuh"
#include "prof.h"
#include "../../include/rtg_splat_b200.h"

namespace rtg {

#define FULL 0xffffffffu

// -DMAPSTEP_EXACT_ADAM: torch's operation order with IEEE sqrt and divisions (adam_one) instead of adam_one_fast
#ifdef MAPSTEP_EXACT_ADAM
#define ADAM1 adam_one
#define ADAM_BC2 a.bc2_sqrt
#else
#define ADAM1 adam_one_fast
#define ADAM_BC2 a.inv_bc2_sqrt
#endif

struct MapStepArgs {
    RtgMapStep s;
    float k_xyz, k_dc, k_rest, k_opacity, k_scaling, k_rotation;  // lr / (1 - beta1^step)
    float bc2_sqrt, inv_bc2_sqrt;
    float attach3, attach4;  // d/dp of weight * mean((p - p0)^2) over count*3 resp. count*4 elements: 2 * weight / (count * dim)
};

__device__ __forceinline__ float sigmoidf(const float x) { return 1.0f / (1.0f + expf(-x)); }

// get_normal (gaussian_pointcloud.py:539-550): column argmin(scale) of R(q / |q|) (build_rotation, utils/general_utils.py:
// 108-131; q is already normalised here), divided by (its norm + 1e-8)
__device__ __forceinline__ void surfel_normal(const float4 q, const float sx, const float sy, const float sz, float n[3]) {
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    const int axis = arg_min3(sx, sy, sz);
    if (axis == 0) { n[0] = 1.f - 2.f * (y * y + z * z); n[1] = 2.f * (x * y + r * z); n[2] = 2.f * (x * z - r * y); }
    else if (axis == 1) { n[0] = 2.f * (x * y - r * z); n[1] = 1.f - 2.f * (x * x + z * z); n[2] = 2.f * (y * z + r * x); }
    else { n[0] = 2.f * (x * z + r * y); n[1] = 2.f * (y * z - r * x); n[2] = 1.f - 2.f * (x * x + y * y); }
    const float inv = 1.0f / (sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]) + 1e-8f);
    n[0] *= inv; n[1] *= inv; n[2] *= inv;
}

// F.normalize(q, dim=-1): q / max(|q|, 1e-12)
__device__ __forceinline__ float4 normalize4(const float4 q, float &nn) {
    nn = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
    return make_float4(q.x / nn, q.y / nn, q.z / nn, q.w / nn);
}

#define MAPSTEP_THREADS 256
#ifndef MAPSTEP_MIN_BLOCKS
#define MAPSTEP_MIN_BLOCKS 3  // 80 registers; measured 0.276 ms (2 -> 0.290, 4 -> 0.286 with spills)
#endif
__global__ void __launch_bounds__(MAPSTEP_THREADS, MAPSTEP_MIN_BLOCKS) map_adam_kernel(const MapStepArgs a) {
    const RtgMapStep &s = a.s;
    const int lane = threadIdx.x & 31;
    const int P = s.P;
    const int n_blk = (P + 31) >> 5;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int blk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; blk < n_blk; blk += warps) {
        const int row0 = blk << 5;
        const int row = row0 + lane;
        const bool in = row < P;
        const uint32_t vm = __ballot_sync(FULL, in && (s.radii == nullptr || __ldg(s.radii + row) > 0));   // rows with a gradient
        const uint32_t am = __ballot_sync(FULL, in && s.attach_mask != nullptr && __ldg(s.attach_mask + row) != 0);
        const int rows = min(32, P - row0);

        // Every phase first issues ALL of its loads into registers, then computes, then stores: the arrays may alias as
        // far as the compiler knows, so a load placed after a store is never hoisted above it -- written iteration by
        // iteration each of the ~18 small steps would cost its own round trip to HBM.

        // ---- xyz, then scaling: 32 rows x 3 floats each, lane + 32 k
        {
            float pp[3], mm[3], vv[3], gg[3], p0[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const int e = lane + 32 * k, r = e / 3;
                pp[k] = mm[k] = vv[k] = gg[k] = p0[k] = 0.f;
                if (r >= rows) continue;
                const size_t i = (size_t)row0 * 3 + e;
                pp[k] = s.xyz[i]; mm[k] = s.m_xyz[i]; vv[k] = s.v_xyz[i];
                if ((vm >> r) & 1u) gg[k] = __ldg(s.g_means3D + i);
                if ((am >> r) & 1u) p0[k] = __ldg(s.xyz0 + i);
            }
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const int e = lane + 32 * k, r = e / 3;
                if (r >= rows) continue;
                const size_t i = (size_t)row0 * 3 + e;
                float g = gg[k];
                if ((am >> r) & 1u) g += a.attach3 * (pp[k] - p0[k]);
                ADAM1(pp[k], g, mm[k], vv[k], a.k_xyz, s.beta1, s.beta2, s.eps, ADAM_BC2);
                s.xyz[i] = pp[k]; s.m_xyz[i] = mm[k]; s.v_xyz[i] = vv[k];
            }
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const int e = lane + 32 * k, r = e / 3;
                pp[k] = mm[k] = vv[k] = gg[k] = p0[k] = 0.f;
                if (r >= rows) continue;
                const size_t i = (size_t)row0 * 3 + e;
                pp[k] = s.scaling_raw[i]; mm[k] = s.m_scaling[i]; vv[k] = s.v_scaling[i];
                if ((vm >> r) & 1u) gg[k] = __ldg(s.g_scales + i);
                if ((am >> r) & 1u) p0[k] = __ldg(s.scaling0 + i);
            }
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const int e = lane + 32 * k, r = e / 3;
                if (r >= rows) continue;
                const size_t i = (size_t)row0 * 3 + e;
                float g = gg[k] * expf(pp[k]);  // d exp(p) / dp = exp(p); gg = 0 on rows without a gradient
                if ((am >> r) & 1u) g += a.attach3 * (pp[k] - p0[k]);
                ADAM1(pp[k], g, mm[k], vv[k], a.k_scaling, s.beta1, s.beta2, s.eps, ADAM_BC2);
                s.scaling_raw[i] = pp[k]; s.m_scaling[i] = mm[k]; s.v_scaling[i] = vv[k];
                s.scales_out[i] = expf(pp[k]);
            }
        }

        // ---- spherical harmonics: 32 rows x 48 floats = 384 float4, lane + 32 k, in batches of MAPSTEP_SH_BATCH; coefficient 0 is
        // _features_dc
        {
            float4 *P4 = reinterpret_cast<float4 *>(s.sh) + (size_t)row0 * 12;
            float4 *M4 = reinterpret_cast<float4 *>(s.m_sh) + (size_t)row0 * 12;
            float4 *V4 = reinterpret_cast<float4 *>(s.v_sh) + (size_t)row0 * 12;
            const float4 *G4 = reinterpret_cast<const float4 *>(s.g_sh) + (size_t)row0 * 12;
#ifndef MAPSTEP_SH_BATCH
#define MAPSTEP_SH_BATCH 2  // float4 per array in flight per lane; 4 needs 122 registers and is no faster
#endif
#pragma unroll 1
            for (int kb = 0; kb < 12; kb += MAPSTEP_SH_BATCH) {
                float4 p[MAPSTEP_SH_BATCH], m[MAPSTEP_SH_BATCH], v[MAPSTEP_SH_BATCH], g[MAPSTEP_SH_BATCH];
#pragma unroll
                for (int j = 0; j < MAPSTEP_SH_BATCH; j++) {
                    const int f = lane + 32 * (kb + j), r = f / 12;
                    g[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < rows) {
                        p[j] = P4[f]; m[j] = M4[f]; v[j] = V4[f];
                        if ((vm >> r) & 1u) g[j] = __ldg(G4 + f);
                    }
                }
#pragma unroll
                for (int j = 0; j < MAPSTEP_SH_BATCH; j++) {
                    const int f = lane + 32 * (kb + j), r = f / 12, c4 = f - r * 12;
                    if (r >= rows) continue;
                    const bool first = (c4 == 0);  // floats 0..2 of the row are the degree-0 coefficient
                    const float k3 = first ? a.k_dc : a.k_rest;
                    ADAM1(p[j].x, g[j].x, m[j].x, v[j].x, k3, s.beta1, s.beta2, s.eps, ADAM_BC2);
                    ADAM1(p[j].y, g[j].y, m[j].y, v[j].y, k3, s.beta1, s.beta2, s.eps, ADAM_BC2);
                    ADAM1(p[j].z, g[j].z, m[j].z, v[j].z, k3, s.beta1, s.beta2, s.eps, ADAM_BC2);
                    ADAM1(p[j].w, g[j].w, m[j].w, v[j].w, a.k_rest, s.beta1, s.beta2, s.eps, ADAM_BC2);
                    P4[f] = p[j]; M4[f] = m[j]; V4[f] = v[j];
                    // mapper.py:455-456: confidence += 1 where any _features_dc gradient is non-zero
                    if (first && s.confidence != nullptr && (g[j].x != 0.f || g[j].y != 0.f || g[j].z != 0.f))
                        s.confidence[row0 + r] += 1.0f;
                }
            }
        }

        float sc_new[3] = {0.f, 0.f, 0.f};  // this lane's row of new scales, for the normal
        if (s.normal_out != nullptr) {  // they were stored by other lanes of this warp (long ago: the SH phase lies between)
            __syncwarp();
            if (in) {
                sc_new[0] = s.scales_out[(size_t)row * 3];
                sc_new[1] = s.scales_out[(size_t)row * 3 + 1];
                sc_new[2] = s.scales_out[(size_t)row * 3 + 2];
            }
        }

        if (in) {
            // ---- rotation (one float4 per row) and opacity (one float), loads first
            float4 *Q4 = reinterpret_cast<float4 *>(s.rotation_raw), *M4 = reinterpret_cast<float4 *>(s.m_rotation),
                   *V4 = reinterpret_cast<float4 *>(s.v_rotation);
            const bool vis = (vm >> lane) & 1u, att = (am >> lane) & 1u;
            float4 q = Q4[row], m = M4[row], v = V4[row];
            float po = s.opacity_raw[row], mo = s.m_opacity[row], vo = s.v_opacity[row];
            float4 ga = make_float4(0.f, 0.f, 0.f, 0.f), q0 = ga;
            float go = 0.f;
            if (vis) { ga = __ldg(reinterpret_cast<const float4 *>(s.g_rotations) + row); go = __ldg(s.g_opacity + row); }
            if (att) q0 = __ldg(reinterpret_cast<const float4 *>(s.rotation0) + row);
            // y = q / max(|q|, eps): dL/dq = (g - y (y . g)) / max(|q|, eps)
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (vis) {
                float nn;
                const float4 y = normalize4(q, nn);
                const float d = y.x * ga.x + y.y * ga.y + y.z * ga.z + y.w * ga.w;
                g = make_float4((ga.x - y.x * d) / nn, (ga.y - y.y * d) / nn, (ga.z - y.z * d) / nn, (ga.w - y.w * d) / nn);
            }
            if (att) {
                g.x += a.attach4 * (q.x - q0.x); g.y += a.attach4 * (q.y - q0.y);
                g.z += a.attach4 * (q.z - q0.z); g.w += a.attach4 * (q.w - q0.w);
            }
            ADAM1(q.x, g.x, m.x, v.x, a.k_rotation, s.beta1, s.beta2, s.eps, ADAM_BC2);
            ADAM1(q.y, g.y, m.y, v.y, a.k_rotation, s.beta1, s.beta2, s.eps, ADAM_BC2);
            ADAM1(q.z, g.z, m.z, v.z, a.k_rotation, s.beta1, s.beta2, s.eps, ADAM_BC2);
            ADAM1(q.w, g.w, m.w, v.w, a.k_rotation, s.beta1, s.beta2, s.eps, ADAM_BC2);
            Q4[row] = q; M4[row] = m; V4[row] = v;
            float nn;
            const float4 y = normalize4(q, nn);
            reinterpret_cast<float4 *>(s.rotations_out)[row] = y;
            if (s.normal_out != nullptr) {
                float n[3];
                surfel_normal(y, sc_new[0], sc_new[1], sc_new[2], n);
                s.normal_out[(size_t)row * 3] = n[0]; s.normal_out[(size_t)row * 3 + 1] = n[1]; s.normal_out[(size_t)row * 3 + 2] = n[2];
            }
            // opacity: o = sigmoid(p), dL/dp = g o (1 - o)
            float gop = 0.f;
            if (vis) {
                const float o = sigmoidf(po);
                gop = go * ((1.f - o) * o);
            }
            ADAM1(po, gop, mo, vo, a.k_opacity, s.beta1, s.beta2, s.eps, ADAM_BC2);
            s.opacity_raw[row] = po; s.m_opacity[row] = mo; s.v_opacity[row] = vo;
            s.opacities_out[row] = sigmoidf(po);
        }
    }
}

__global__ void __launch_bounds__(256) map_activate_kernel(const int P, const float *__restrict__ scaling_raw,
                                                           const float *__restrict__ rotation_raw, const float *__restrict__ opacity_raw,
                                                           float *__restrict__ scales_out, float *__restrict__ rotations_out,
                                                           float *__restrict__ opacities_out, float *__restrict__ normal_out) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= P) return;
    const size_t i3 = (size_t)row * 3;
    const float sx = expf(scaling_raw[i3]), sy = expf(scaling_raw[i3 + 1]), sz = expf(scaling_raw[i3 + 2]);
    scales_out[i3] = sx; scales_out[i3 + 1] = sy; scales_out[i3 + 2] = sz;
    float nn;
    const float4 y = normalize4(reinterpret_cast<const float4 *>(rotation_raw)[row], nn);
    reinterpret_cast<float4 *>(rotations_out)[row] = y;
    opacities_out[row] = sigmoidf(opacity_raw[row]);
    if (normal_out != nullptr) {
        float n[3];
        surfel_normal(y, sx, sy, sz, n);
        normal_out[i3] = n[0]; normal_out[i3 + 1] = n[1]; normal_out[i3 + 2] = n[2];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Mapping.history_merge (SLAM/multiprocess/mapper.py:212-250): after an optimisation call the raw parameters are pulled
// back towards their pre-optimisation values with the weight  w = max_weight * confidence_before / (confidence_now + 1e-6).
// The reference evaluates it with ~45 eager kernels (slerp alone is ~30, SLAM/utils.py:593-652); here one launch: grid row
// y = 0 merges _xyz (per-row weight) and the rotation (slerp with t = 1 - w), rows y = 1..3 the three arrays that the
// reference merges with `history_weight[0]`, the weight of the FIRST Gaussian (mapper.py:229,234,239 -- kept as is).
// Products and sums are the reference's separate fp32 operations (no contraction into FMAs).
struct HistLerpSeg {
    const float *hist;  // (P, width) contiguous
    float *cur;         // row r at cur + r * stride
    int width, stride;
};

struct HistMergeArgs {
    RtgHistoryMerge m;
    HistLerpSeg seg[3];
};

__device__ __forceinline__ float hist_weight(const HistMergeArgs &a, const int row) {
    return __fdiv_rn(__fmul_rn(a.m.max_weight, __ldg(a.m.hist_confidence + row)), __fadd_rn(__ldg(a.m.confidence + row), 1e-6f));
}

// hist * w + (1 - w) * cur
__device__ __forceinline__ float hist_mix(const float h, const float c, const float w) {
    return __fadd_rn(__fmul_rn(h, w), __fmul_rn(__fsub_rn(1.0f, w), c));
}

// torch.lerp(start, end, weight) (ATen/native/Lerp.h)
__device__ __forceinline__ float torch_lerp(const float start, const float end, const float w) {
    const float diff = __fsub_rn(end, start);
    return (fabsf(w) < 0.5f) ? fmaf(w, diff, start) : fmaf(-diff, __fsub_rn(1.0f, w), end);
}

__global__ void __launch_bounds__(256) history_merge_kernel(const HistMergeArgs a) {
    const int P = a.m.P;
    if (blockIdx.y == 0) {
        const int row = blockIdx.x * blockDim.x + threadIdx.x;
        if (row >= P) return;
        const float w = hist_weight(a, row);
        float *x = a.m.xyz + 3 * (size_t)row;
        const float *hx = a.m.hist_xyz + 3 * (size_t)row;
        const float x0 = hist_mix(__ldg(hx), x[0], w), x1 = hist_mix(__ldg(hx + 1), x[1], w), x2 = hist_mix(__ldg(hx + 2), x[2], w);
        x[0] = x0; x[1] = x1; x[2] = x2;
        // slerp(history rotation, normalize(_rotation), 1 - w)  (SLAM/utils.py:593-652; get_rotation = F.normalize)
        const float4 v0 = __ldg(reinterpret_cast<const float4 *>(a.m.hist_rotation) + row);
        float4 *qp = reinterpret_cast<float4 *>(a.m.rotation_raw) + row;
        float nn;
        const float4 v1 = normalize4(*qp, nn);
        const float t = __fsub_rn(1.0f, w);
        const float n0 = sqrtf(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v0.x, v0.x), __fmul_rn(v0.y, v0.y)), __fmul_rn(v0.z, v0.z)), __fmul_rn(v0.w, v0.w)));
        const float n1 = sqrtf(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v1.x, v1.x), __fmul_rn(v1.y, v1.y)), __fmul_rn(v1.z, v1.z)), __fmul_rn(v1.w, v1.w)));
        const float dot = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(__fdiv_rn(v0.x, n0), __fdiv_rn(v1.x, n1)), __fmul_rn(__fdiv_rn(v0.y, n0), __fdiv_rn(v1.y, n1))),
                                              __fmul_rn(__fdiv_rn(v0.z, n0), __fdiv_rn(v1.z, n1))), __fmul_rn(__fdiv_rn(v0.w, n0), __fdiv_rn(v1.w, n1)));
        float4 o;
        if (isnan(dot) || fabsf(dot) > 0.9995f) {  // (nearly) collinear or a zero quaternion: linear
            o = make_float4(torch_lerp(v0.x, v1.x, t), torch_lerp(v0.y, v1.y, t), torch_lerp(v0.z, v1.z, t), torch_lerp(v0.w, v1.w, t));
        } else {
            const float theta0 = acosf(dot), sin0 = sinf(theta0), theta_t = __fmul_rn(theta0, t);
            const float s0 = __fdiv_rn(sinf(__fsub_rn(theta0, theta_t)), sin0), s1 = __fdiv_rn(sinf(theta_t), sin0);
            o = make_float4(__fadd_rn(__fmul_rn(s0, v0.x), __fmul_rn(s1, v1.x)), __fadd_rn(__fmul_rn(s0, v0.y), __fmul_rn(s1, v1.y)),
                            __fadd_rn(__fmul_rn(s0, v0.z), __fmul_rn(s1, v1.z)), __fadd_rn(__fmul_rn(s0, v0.w), __fmul_rn(s1, v1.w)));
        }
        *qp = o;
        return;
    }
    const HistLerpSeg &g = a.seg[blockIdx.y - 1];
    const float w0 = hist_weight(a, 0);  // history_weight[0]
    const size_t n = (size_t)P * (size_t)g.width;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / (size_t)g.width;
        float *c = g.cur + r * (size_t)g.stride + (i - r * (size_t)g.width);
        *c = hist_mix(__ldg(g.hist + i), *c, w0);
    }
}

void launch_history_merge(const RtgHistoryMerge &m, cudaStream_t s) {
    HistMergeArgs a;
    a.m = m;
    a.seg[0] = {m.hist_features_dc, m.features_dc, 3, m.features_dc_stride};
    a.seg[1] = {m.hist_features_rest, m.features_rest, m.features_rest_width, m.features_rest_stride};
    a.seg[2] = {m.hist_scaling, m.scaling, 3, 3};
    ProfScope ps(K_ADAM, s);
    history_merge_kernel<<<dim3((m.P + 255) / 256, 4), 256, 0, s>>>(a);
}

void launch_map_adam_step(const RtgMapStep &st, cudaStream_t s) {
    MapStepArgs a;
    a.s = st;
    // torch computes the bias corrections in python doubles (torch/optim/adam.py, _single_tensor_adam)
    const double bc1 = 1.0 - pow((double)st.beta1, (double)st.step);
    const double bc2 = 1.0 - pow((double)st.beta2, (double)st.step);
    const float fbc1 = (float)bc1;
    a.k_xyz = st.lr_xyz / fbc1; a.k_dc = st.lr_f_dc / fbc1; a.k_rest = st.lr_f_rest / fbc1;
    a.k_opacity = st.lr_opacity / fbc1; a.k_scaling = st.lr_scaling / fbc1; a.k_rotation = st.lr_rotation / fbc1;
    a.bc2_sqrt = (float)sqrt(bc2);
    a.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
    a.attach3 = a.attach4 = 0.f;
    if (st.attach_mask != nullptr && st.attach_count > 0) {
        a.attach3 = (float)(2.0 * (double)st.attach_weight / ((double)st.attach_count * 3.0));
        a.attach4 = (float)(2.0 * (double)st.attach_weight / ((double)st.attach_count * 4.0));
    } else {
        a.s.attach_mask = nullptr;  // mapper.py:387: no attach loss when the mask is empty
    }
    const int n_blk = (st.P + 31) / 32;
    int blocks = (n_blk + (MAPSTEP_THREADS / 32) - 1) / (MAPSTEP_THREADS / 32);
    static int resident = 0;  // one wave of resident CTAs, grid-stride over the 32-row blocks
    if (resident == 0) {
        int dev = 0, sms = 148, per_sm = MAPSTEP_MIN_BLOCKS;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, map_adam_kernel, MAPSTEP_THREADS, 0) != cudaSuccess || per_sm < 1)
            per_sm = MAPSTEP_MIN_BLOCKS;
        resident = sms * per_sm;
    }
    if (blocks > resident) blocks = resident;
    ProfScope ps(K_ADAM, s);
    map_adam_kernel<<<blocks, MAPSTEP_THREADS, 0, s>>>(a);
}

void launch_map_activate(int P, const float *scaling_raw, const float *rotation_raw, const float *opacity_raw, float *scales_out,
                         float *rotations_out, float *opacities_out, float *normal_out, cudaStream_t s) {
    map_activate_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, scaling_raw, rotation_raw, opacity_raw, scales_out, rotations_out,
                                                        opacities_out, normal_out);
}

}  // namespace rtg
