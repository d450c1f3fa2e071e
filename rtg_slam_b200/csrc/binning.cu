// Per-tile bucketing of (Gaussian, tile) instances without any host round trip.
//
// The reference emits (tile<<32 | depth) keys after a device-wide prefix sum over Gaussians, reads
// the instance count back to the host, runs a device-wide 64-bit radix sort, and compacts the
// non-empty tiles on the CPU (RAST/cuda_rasterizer/rasterizer_impl.cu:300-362). Here the tile
// histogram comes out of the preprocess pass, one small scan turns it into tile offsets (and into
// the ascending list of non-empty tiles), instances are scattered straight into their tile's
// bucket, and each bucket is depth-sorted on chip. Sorting by (depth bits, Gaussian id) reproduces
// the order of the reference's stable sort, whose ties are broken by emission (= Gaussian) order.
#include "common.cuh"
#include "kernels.h"
#include "prof.h"

namespace rtg {

// ---------------------------------------------------------------- tile scan (single CTA)
__global__ void __launch_bounds__(1024) tile_scan_kernel(BinState b, int T, long long R_cap, int *__restrict__ counters,
                                                         int *__restrict__ counters_host) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry, s_max, s_nact, s_sum;
    __shared__ uint32_t s_cls[66];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) { s_carry = 0; s_max = 0; s_nact = 0; s_sum = 0; }
    if (tid < 66) s_cls[tid] = 0;
    __syncthreads();
    uint32_t local_max = 0, local_act = 0, local_sum = 0;
    // (1) exclusive scan of the histogram -> tile offsets; (2) histogram of work classes for the launch order
    for (int base = 0; base < T; base += 1024) {
        const int i = base + tid;
        const uint32_t c = (i < T) ? b.tile_count[(size_t)i * RTG_CNT_STRIDE] : 0u;
        const uint32_t cpad = (c + (RTG_LIST_ALIGN - 1)) & ~(uint32_t)(RTG_LIST_ALIGN - 1);  // bucket capacity
        local_max = max(local_max, c);
        local_act += (c > 0);
        local_sum += c;
        uint32_t v = cpad;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += n;
        }
        if (lane == 31) s_warp[wid] = v;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t n = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += n;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const uint32_t incl = v + (wid > 0 ? s_warp[wid - 1] : 0u) + s_carry;
        if (i < T) {
            b.tile_offset[i] = incl - cpad;
            // class 0 = longest lists (>= 4032 entries) ... class 63 = 1..63 entries, class 64 = empty tiles
            const int cls = (c == 0) ? 64 : 63 - (int)min(63u, c >> 6);
            atomicAdd(&s_cls[cls], 1u);
        }
        __syncthreads();
        if (tid == 1023) s_carry = incl;
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
        local_act += __shfl_xor_sync(0xffffffffu, local_act, o);
        local_sum += __shfl_xor_sync(0xffffffffu, local_sum, o);
    }
    if (lane == 0) { atomicMax(&s_max, local_max); atomicAdd(&s_nact, local_act); atomicAdd(&s_sum, local_sum); }
    __syncthreads();
    // Launch order of the tiles: longest lists first (the block scheduler hands out CTAs in index order, so this
    // is longest-processing-time-first scheduling), empty tiles last. Order inside a class is irrelevant.
    if (tid == 0) {
        uint32_t run = 0;
        for (int k = 0; k < 65; k++) { const uint32_t n = s_cls[k]; s_cls[k] = run; run += n; }
    }
    __syncthreads();
    for (int i = tid; i < T; i += 1024) {
        const uint32_t c = b.tile_count[(size_t)i * RTG_CNT_STRIDE];
        const int cls = (c == 0) ? 64 : 63 - (int)min(63u, c >> 6);
        b.active[atomicAdd(&s_cls[cls], 1u)] = (uint32_t)i;
    }
    if (tid == 0) {
        const uint32_t Rpad = s_carry, R = s_sum;  // buffer entries needed (padded buckets) / instances
        b.tile_offset[T] = Rpad;
        const int ov = ((long long)Rpad > R_cap) ? 1 : 0;
        counters[0] = (int)R; counters[1] = (int)s_nact; counters[2] = ov; counters[3] = (int)s_max; counters[4] = (int)Rpad;
        if (counters_host) {
            counters_host[0] = (int)R; counters_host[1] = (int)s_nact; counters_host[2] = ov; counters_host[3] = (int)s_max;
            counters_host[4] = (int)Rpad;
        }
    }
}

// ---------------------------------------------------------------- scatter into tile buckets
__global__ void __launch_bounds__(256) scatter_kernel(const ViewParams vp, int P, const GeomState g, const int *__restrict__ radii,
                                                      const int *__restrict__ tile_mask, BinState b, long long R_cap,
                                                      const int *__restrict__ counters) {
    __shared__ uint32_t s_excl[8][32], s_rect[8][32];
    __shared__ float4 s_ga[8][32], s_gb[8][32];
    __shared__ uint32_t s_depth[8][32];
    if (counters[2]) return;  // capacity overflow: render nothing, the caller retries with a larger buffer
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int r = idx < P ? radii[idx] : 0;
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
    if (r > 0) {
        s0 = g.splat[2 * (size_t)idx];
        s1 = g.splat[2 * (size_t)idx + 1];
        tile_rect(make_float2(s0.x, s0.y), r, vp.tiles_x, vp.tiles_y, x0, y0, x1, y1);
        y0 = max(y0, vp.row_begin); y1 = max(y0, min(y1, vp.row_end));  // rows outside the rank's band hold no tile of its mask
    }
    const int npairs = (x1 - x0) * (y1 - y0);
    const int incl = warp_incl_scan(npairs, lane);
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;  // warp-uniform
    s_excl[w][lane] = (uint32_t)(incl - npairs);
    s_rect[w][lane] = pack_rect(x0, y0, max(x1 - x0, 1));
    const float2 nb = cut_slopes(s1.x, s1.y, s1.z);
    s_ga[w][lane] = make_float4(s0.x, s0.y, s0.z, nb.x);  // x, y, q_cut, -b/c
    s_gb[w][lane] = make_float4(s1.x, s1.y, s1.z, nb.y);  // conic, -b/a
    s_depth[w][lane] = (r > 0) ? __float_as_uint(g.hit[2 * (size_t)idx + 1].z) : 0u;  // view depth (forward.cu:336, rasterizer_impl.cu:96)
    __syncwarp();
    const int base = idx - lane;
    for (int k = lane; k < total; k += 32) {
        const int o = pair_owner(s_excl[w], (uint32_t)k);
        const uint32_t local = (uint32_t)k - s_excl[w][o], rc = s_rect[w][o];
        const uint32_t rw = rc >> 20;
        int cx, cy;
        rect_cell(local, rw, cx, cy);
        const int x = (int)(rc & 1023u) + cx, y = (int)((rc >> 10) & 1023u) + cy;
        const int t = y * vp.tiles_x + x;
        if (__ldg(tile_mask + t)) {
            const float4 ga = s_ga[w][o], gb = s_gb[w][o];
            const float fx0 = (float)(x * RTG_TILE), fy0 = (float)(y * RTG_TILE);
            // same (bit-identical) decision as the histogram pass in preprocess_fwd_kernel
            if (rect_below_cutoff(ga.x, ga.y, gb.x, gb.y, gb.z, ga.z, ga.w, gb.w, fx0, fx0 + (RTG_TILE - 1), fy0, fy0 + (RTG_TILE - 1))) continue;
            const uint32_t slot = atomicAdd(b.tile_fill + (size_t)t * RTG_CNT_STRIDE, 1u);
            const uint32_t begin = b.tile_offset[t];
            if (begin + slot < b.tile_offset[t + 1])  // never write outside the (padded) bucket
                b.keys[begin + slot] = ((uint64_t)s_depth[w][o] << 32) | (uint32_t)(base + o);
        }
    }
}

// ---------------------------------------------------------------- tile histogram from the records (Gaussian-sharded forward)
// The single-GPU forward builds the histogram inside the forward preprocess. When the Gaussians are sharded over ranks,
// every rank needs the histogram of ALL visible Gaussians over ITS tiles, which exists only after the records have been
// exchanged: this pass recomputes it from the records with the same pair expansion and the bit-identical cut-off test.
__global__ void __launch_bounds__(256) tile_histogram_kernel(const ViewParams vp, int P, const GeomState g, const int *__restrict__ radii,
                                                             const int *__restrict__ tile_mask, BinState b) {
    __shared__ uint32_t s_excl[8][32], s_rect[8][32];
    __shared__ float4 s_ga[8][32], s_gb[8][32];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int r = idx < P ? radii[idx] : 0;
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
    if (r > 0) {
        s0 = g.splat[2 * (size_t)idx];
        s1 = g.splat[2 * (size_t)idx + 1];
        tile_rect(make_float2(s0.x, s0.y), r, vp.tiles_x, vp.tiles_y, x0, y0, x1, y1);
        y0 = max(y0, vp.row_begin); y1 = max(y0, min(y1, vp.row_end));  // rows outside the rank's band hold no tile of its mask
    }
    const int npairs = (x1 - x0) * (y1 - y0);
    const int incl = warp_incl_scan(npairs, lane);
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;  // warp-uniform
    s_excl[w][lane] = (uint32_t)(incl - npairs);
    s_rect[w][lane] = pack_rect(x0, y0, max(x1 - x0, 1));
    const float2 nb = cut_slopes(s1.x, s1.y, s1.z);
    s_ga[w][lane] = make_float4(s0.x, s0.y, s0.z, nb.x);
    s_gb[w][lane] = make_float4(s1.x, s1.y, s1.z, nb.y);
    __syncwarp();
    for (int k = lane; k < total; k += 32) {
        const int o = pair_owner(s_excl[w], (uint32_t)k);
        const uint32_t local = (uint32_t)k - s_excl[w][o], rc = s_rect[w][o];
        const uint32_t rw = rc >> 20;
        int cx, cy;
        rect_cell(local, rw, cx, cy);
        const int x = (int)(rc & 1023u) + cx, y = (int)((rc >> 10) & 1023u) + cy;
        const int t = y * vp.tiles_x + x;
        if (__ldg(tile_mask + t)) {
            const float4 ga = s_ga[w][o], gb = s_gb[w][o];
            const float fx0 = (float)(x * RTG_TILE), fy0 = (float)(y * RTG_TILE);
            if (rect_below_cutoff(ga.x, ga.y, gb.x, gb.y, gb.z, ga.z, ga.w, gb.w, fx0, fx0 + (RTG_TILE - 1), fy0, fy0 + (RTG_TILE - 1)))
                b.tile_touched[t] = 1u;
            else
                atomicAdd(b.tile_count + (size_t)t * RTG_CNT_STRIDE, 1u);
        }
    }
}

void launch_tile_histogram(const ViewParams &vp, int P, const GeomState &g, const int *radii, const int *tile_mask, const BinState &b,
                           cudaStream_t s) {
    if (P <= 0) return;
    ProfScope ps(K_PREPROCESS_FWD, s);
    tile_histogram_kernel<<<(P + 255) / 256, 256, 0, s>>>(vp, P, g, radii, tile_mask, b);
}

// ---------------------------------------------------------------- per-tile depth sort
#define SORT_THREADS 256
#define SORT_SMEM_KEYS 4096

__device__ __forceinline__ int next_pow2(int n) {
    int N = 1;
    while (N < n) N <<= 1;
    return N;
}

// Register-resident bitonic sort of 256*K keys by one CTA of 256 threads: thread t holds the keys of positions
// t*K .. t*K+K-1. A compare-exchange sub-stage with partner distance j is
//   j <  K      : both keys in the same thread  -> register compare-swap with compile-time indices,
//   j <  32*K   : partner in the same warp       -> one 64-bit shuffle,
//   j >= 32*K   : partner in another warp        -> one pass through shared memory (striped: conflict-free),
// so only 6 of the 55 (N = 1024) .. 78 (N = 4096) sub-stages touch shared memory at all. `Nn` (a power of two,
// >= the number of real keys) bounds the merge size: blocks beyond it hold only +inf padding.
template <int K>
__device__ __forceinline__ void sort_in_registers(uint64_t (&key)[K], const int Nn, uint64_t *s_buf) {
    const int t = threadIdx.x;
    const int base = t * K;
    for (int size = 2; size <= Nn; size <<= 1) {
        const bool up_t = ((base & size) == 0);  // direction of this thread's keys when size >= K
        int j = size >> 1;
        for (; j >= 32 * K; j >>= 1) {
            const int tj = j / K;                  // partner thread = t ^ tj
            __syncthreads();                       // the previous pass has been read by everybody
#pragma unroll
            for (int r = 0; r < K; r++) s_buf[r * SORT_THREADS + t] = key[r];
            __syncthreads();
            const bool keep_min = (((t & tj) == 0) == up_t);
#pragma unroll
            for (int r = 0; r < K; r++) {
                const uint64_t o = s_buf[r * SORT_THREADS + (t ^ tj)];
                const bool o_less = o < key[r];
                key[r] = (o_less == keep_min) ? o : key[r];
            }
        }
        for (; j >= K; j >>= 1) {
            const int lj = j / K;                  // partner lane = lane ^ lj
            const bool keep_min = (((t & lj) == 0) == up_t);
#pragma unroll
            for (int r = 0; r < K; r++) {
                const uint64_t o = __shfl_xor_sync(0xffffffffu, key[r], lj);
                const bool o_less = o < key[r];
                key[r] = (o_less == keep_min) ? o : key[r];
            }
        }
#pragma unroll
        for (int jj = K / 2; jj > 0; jj >>= 1) {
            if (jj < size) {
#pragma unroll
                for (int r = 0; r < K; r++) {
                    if ((r & jj) == 0) {
                        const bool up = (((base + r) & size) == 0);
                        const uint64_t x = key[r], y = key[r | jj];
                        const bool sw = (x > y) == up;
                        key[r] = sw ? y : x;
                        key[r | jj] = sw ? x : y;
                    }
                }
            }
        }
    }
}

// Load (any order: the input is unsorted), sort, write front to back either the Gaussian ids of the n real keys
// (IDS) or the sorted keys themselves, in place (chunks of an oversized list). K must be the smallest power of two
// with n <= 256*K, so that next_pow2(n) covers every position that holds a real key.
template <int K, bool IDS>
__device__ __forceinline__ void sort_tile_in_registers(uint64_t *__restrict__ gk, uint32_t *__restrict__ out, const int n,
                                                       uint64_t *s_buf) {
    uint64_t key[K];
#pragma unroll
    for (int r = 0; r < K; r++) {
        const int i = r * SORT_THREADS + (int)threadIdx.x;  // coalesced
        key[r] = (i < n) ? gk[i] : 0xffffffffffffffffull;
    }
    if (!IDS) __syncthreads();  // in place: everybody has read its keys before sorted keys are written back
    sort_in_registers<K>(key, K == 1 ? next_pow2(n) : SORT_THREADS * K, s_buf);
#pragma unroll
    for (int r = 0; r < K; r++) {
        const int i = (int)threadIdx.x * K + r;
        if (i < n) {
            if (IDS) out[i] = (uint32_t)key[r];
            else gk[i] = key[r];
        }
    }
}

template <bool IDS>
__device__ __forceinline__ void sort_up_to_4096(uint64_t *__restrict__ gk, uint32_t *__restrict__ out, const int n, uint64_t *s_buf) {
    if (n <= 256) sort_tile_in_registers<1, IDS>(gk, out, n, s_buf);
    else if (n <= 512) sort_tile_in_registers<2, IDS>(gk, out, n, s_buf);
    else if (n <= 1024) sort_tile_in_registers<4, IDS>(gk, out, n, s_buf);
    else if (n <= 2048) sort_tile_in_registers<8, IDS>(gk, out, n, s_buf);
    else sort_tile_in_registers<16, IDS>(gk, out, n, s_buf);
    __syncthreads();
}

// One CTA per non-empty tile. Up to SORT_SMEM_KEYS keys are sorted in registers (above). Longer lists (rare: a few
// tiles in front of a dense surface) are sorted chunk by chunk in shared memory and merged by rank: with chunk c
// resident (sorted) in shared memory, every key of the tile adds the number of chunk-c keys below it (binary
// search on chip; keys are unique), which is its final position once all chunks have been visited.
__global__ void __launch_bounds__(SORT_THREADS) tile_sort_kernel(BinState b, const int *__restrict__ counters) {
    __shared__ uint64_t s_keys[SORT_SMEM_KEYS];
    if (counters[2]) return;
    const int nact = counters[1];
    for (int ai = blockIdx.x; ai < nact; ai += gridDim.x) {
        const uint32_t tile = b.active[ai];
        const uint32_t start = b.tile_offset[tile];
        const int n = (int)b.tile_count[(size_t)tile * RTG_CNT_STRIDE];
        uint64_t *gk = b.keys + start;
        uint32_t *out = b.point_list + start;
        if (n <= SORT_SMEM_KEYS) {
            sort_up_to_4096<true>(gk, out, n, s_keys);
            continue;
        }
        const int nchunks = (n + SORT_SMEM_KEYS - 1) / SORT_SMEM_KEYS;
        for (int c = 0; c < nchunks; c++) {  // sort every chunk in place
            const int c0 = c * SORT_SMEM_KEYS, cn = min(SORT_SMEM_KEYS, n - c0);
            sort_up_to_4096<false>(gk + c0, nullptr, cn, s_keys);
        }
        for (int c = 0; c < nchunks; c++) {  // accumulate ranks in `out` (used as scratch until the final permutation)
            const int c0 = c * SORT_SMEM_KEYS, cn = min(SORT_SMEM_KEYS, n - c0);
            for (int i = threadIdx.x; i < cn; i += SORT_THREADS) s_keys[i] = gk[c0 + i];
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += SORT_THREADS) {
                uint32_t r;
                if (i >= c0 && i < c0 + cn) {
                    r = (uint32_t)(i - c0);
                } else {
                    const uint64_t key = gk[i];
                    int lo = 0, hi = cn;  // first index with s_keys[idx] >= key
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (s_keys[mid] < key) lo = mid + 1; else hi = mid;
                    }
                    r = (uint32_t)lo;
                }
                out[i] = (c == 0) ? r : out[i] + r;
            }
            __syncthreads();
        }
        for (int i = threadIdx.x; i < n; i += SORT_THREADS) gk[i] = ((uint64_t)out[i] << 32) | (uint32_t)gk[i];  // (rank, id)
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += SORT_THREADS) {
            const uint64_t v = gk[i];
            out[(uint32_t)(v >> 32)] = (uint32_t)v;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- launchers
void launch_tile_scan(const BinState &b, int T, int64_t R_cap, int32_t *counters, int32_t *counters_host, cudaStream_t s) {
    ProfScope ps(K_TILE_SCAN, s);
    tile_scan_kernel<<<1, 1024, 0, s>>>(b, T, (long long)R_cap, counters, counters_host);
}

void launch_scatter(const ViewParams &vp, int P, const GeomState &g, const int *radii, const int *tile_mask, const BinState &b,
                    int64_t R_cap, const int32_t *counters, cudaStream_t s) {
    if (P <= 0) return;
    ProfScope ps(K_SCATTER, s);
    scatter_kernel<<<(P + 255) / 256, 256, 0, s>>>(vp, P, g, radii, tile_mask, b, (long long)R_cap, counters);
}

void launch_tile_sort(const BinState &b, int T, const int32_t *counters, cudaStream_t s) {
    if (T <= 0) return;
    ProfScope ps(K_TILE_SORT, s);
    tile_sort_kernel<<<T, SORT_THREADS, 0, s>>>(b, counters);
}

}  // namespace rtg
