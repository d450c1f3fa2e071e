// Micro-benchmark: scalar FFMA vs packed FFMA2 issue throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float *out, int iters, float s) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
    const float2 m = make_float2(s, s * 0.5f), c = make_float2(0.25f, 0.125f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
            else a[i] = __ffma2_rn(a[i], m, c);
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
    float *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(d, iters, 0.999f); else k<1><<<148 * 8, 256>>>(d, iters, 0.999f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fma = 148.0 * 8 * 256 * iters * 16.0;
            if (rep) printf("mode %d (%s): %.3f ms, %.1f TFMA/s (%.1f TFLOP/s)\n", mode, mode ? "FFMA2" : "FFMA", ms, fma / ms / 1e9, 2 * fma / ms / 1e9);
        }
    }
    return 0;
}
