// filter.cuh -- the filtered brute-force scan shared by the batched kNN (k_knn_filter) and
// radius (k_near_filtered) kernels of nodes.cu: centred fp32 shadow, conservative per-query
// threshold, and the TMA-staged packed-FMA scan over groups of 8 nodes.
#pragma once
#include <math.h>

#include "common.cuh"

// --------------------------------------- K4a': seeded filter + exact refine ---
// When every query already has an a-priori bound (seed2, the grid seeds below), the top-K
// state is not needed during the scan at all: the scan only has to FIND the few nodes inside
// the bound.  The filter runs on a per-call fp32 shadow of the nodes, centred on the middle c
// of their bounding box, with the squared norm precomputed (k_shadow_centered):
//     |n - q|^2 = |n^|^2 - 2 n^.q^ + |q^|^2,      n^ = fl32(n - c), q^ = fl32(q - c)
// so a pair costs D fused multiply-adds instead of D subtractions + D multiply-adds:
//     t = fma(n^_D, -2q^_D, ... fma(n^_1, -2q^_1, fl32(|n^|^2)))   vs   thr' = bound - |q^|^2.
// k_knn_filter streams the shadow through shared memory (TMA bulk copies, double buffered,
// one mbarrier per stage).  A thread owns FQ = 4 queries and evaluates 8 nodes per step for
// each of them with the packed fp32x2 pipe (FFMA2) and three-input minima (FMNMX3): every
// broadcast LDS.128 feeds 16 (query, node) pairs -- the shared-memory pipe moves only 8 bytes
// per wavefront on a broadcast, which bounded the one-query-per-thread version -- and the
// kernel is bound by the FMA pipe (1 packed FFMA2 = 2 pipe cycles per pair for d = 2).
//
// CONSERVATIVE threshold.  With u = 2^-24, M >= |n_j - c_j|, Mq >= |q_j - c_j|:
//   * |n^ - q^| <= |n - q| + a,  a = 4 u sqrt(D) max(M, Mq)   (roundings of n^ and q^);
//   * the D+1 roundings of the fma chain change t by at most
//     E = u D (D+1) M (M + Mq) (1 + 1e-5) + 1e-15 Mq^2    (partial sums are <= D M^2 + 2 j M Mq);
// so every pair with exact fp64 d^2 <= bound has t <= thr' := round_up((sqrt(bound) + a)^2 + E
// - |q^|^2).  The filter can only add false positives, never lose a neighbour.  A query
// whose band E exceeds (sqrt(bound) + a)^2 (huge extent / tiny neighbour distance, outliers),
// or that has no finite bound, is not filtered: it is flagged and recomputed by k_knn_batched.
//
// A group of 8 nodes whose minimum t passes is appended to the (query, chunk) candidate list;
// k_knn_refine then decides the candidates exactly in fp64 with the same TopK::offer as the
// one-kernel scan, visiting them in node-index order.
#define FILTER_STAGES 2
#define FILTER_Q 4          // queries per thread
// nodes per stage: 2 stages x (D + 1) x tile x 4 B of static shared memory (<= 40 KB)
__host__ __device__ constexpr int filter_tile(int D) { return D <= 4 ? 1024 : 512; }

__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float f_min3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t phase) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
    return ok != 0;
}

#define BOUNDS_BLOCKS 128
// Bounding box of the first DD coordinates: per-block partials [blocks][DD][2] = (min, max),
// then a one-warp fold (NaN coordinates are ignored by fmin / fmax).
template <int DD>
__global__ void __launch_bounds__(256)
k_bounds_partial(const double *__restrict__ soa, int64_t capacity, int64_t n, double *__restrict__ partial,
                 const int32_t *__restrict__ pending = nullptr) {
    if (pending && *pending == 0) return;          // every query already answered (k_knn_cells)
    __shared__ double sh[2 * DD][8];
    double mn[DD], mx[DD];
#pragma unroll
    for (int j = 0; j < DD; ++j) { mn[j] = INFINITY; mx[j] = -INFINITY; }
    for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < n; t += (int64_t)gridDim.x * 256) {
#pragma unroll
        for (int j = 0; j < DD; ++j) {
            const double v = soa[(int64_t)j * capacity + t];
            mn[j] = fmin(mn[j], v); mx[j] = fmax(mx[j], v);
        }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < DD; ++j) {
        for (int o = 16; o > 0; o >>= 1) {
            mn[j] = fmin(mn[j], __shfl_down_sync(0xffffffffu, mn[j], o));
            mx[j] = fmax(mx[j], __shfl_down_sync(0xffffffffu, mx[j], o));
        }
        if (lane == 0) { sh[2 * j][w] = mn[j]; sh[2 * j + 1][w] = mx[j]; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * DD) {
        double v = sh[threadIdx.x][0];
        for (int i = 1; i < 8; ++i) v = (threadIdx.x & 1) ? fmax(v, sh[threadIdx.x][i]) : fmin(v, sh[threadIdx.x][i]);
        partial[(int64_t)blockIdx.x * 2 * DD + threadIdx.x] = v;
    }
}

template <int DD>
__global__ void k_bounds_fold(const double *__restrict__ partial, int nb, double *__restrict__ out,
                              const int32_t *__restrict__ pending = nullptr) {
    if (pending && *pending == 0) return;
    const int c = threadIdx.x;                     // one thread per (coordinate, min | max)
    if (c >= 2 * DD) return;
    double v = partial[c];
    for (int i = 1; i < nb; ++i) v = (c & 1) ? fmax(v, partial[(int64_t)i * 2 * DD + c]) : fmin(v, partial[(int64_t)i * 2 * DD + c]);
    out[c] = v;
}

// Centred fp32 shadow [D + 1][cap32]: coordinates relative to the bounding-box centre and
// the squared norm of the ROUNDED coordinates (exact products, one rounding).  Slots in
// [n, cap32) hold the origin with norm 3e38 so that a tile tail never passes a threshold.
// info[0] = M (max |centred coordinate|, +inf when a bound is not finite), info[1+j] = c_j,
// info[1+D+j] = extent of coordinate j (max - min).
template <int D>
__global__ void k_shadow_centered(const double *__restrict__ soa, int64_t capacity, int64_t n,
                                  const double *__restrict__ bounds /*[D][2] = min, max*/,
                                  int64_t cap32, float *__restrict__ sh, double *__restrict__ info,
                                  const int32_t *__restrict__ pending = nullptr) {
    if (pending && *pending == 0) return;          // every query already answered (k_knn_cells)
    double c[D], M = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        const double mn = bounds[2 * j], mx = bounds[2 * j + 1];
        c[j] = 0.5 * mn + 0.5 * mx;
        const double h = fmax(mx - c[j], c[j] - mn);
        M = (h != h) ? INFINITY : fmax(M, h);
        if (!(mn <= mx)) M = INFINITY;                  // empty / NaN bounds
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        info[0] = M * (1.0 + 1e-9);
#pragma unroll
        for (int j = 0; j < D; ++j) {
            info[1 + j] = c[j];
            info[1 + D + j] = bounds[2 * j + 1] - bounds[2 * j];
        }
    }
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < cap32;
         t += (int64_t)gridDim.x * blockDim.x) {
        if (t < n) {
            double nn = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const float xc = (float)(soa[(int64_t)j * capacity + t] - c[j]);
                sh[(int64_t)j * cap32 + t] = xc;
                nn += (double)xc * (double)xc;
            }
            sh[(int64_t)D * cap32 + t] = (float)nn;
        } else {
            // padding: t = fma(0, ., 3e38) = 3e38 never passes a (finite) threshold
#pragma unroll
            for (int j = 0; j < D; ++j) sh[(int64_t)j * cap32 + t] = 0.0f;
            sh[(int64_t)D * cap32 + t] = 3.0e38f;
        }
    }
}

// Threshold of one query for the filter (see the derivation above).  `bound2` is the exact
// fp64 d^2 bound of the query (seed, or the radius threshold); xq its coordinates (stride of
// the caller), DD the number of leading coordinates the distance uses.  Returns false -- and
// thr = -inf, so that nothing passes -- when the query cannot be filtered conservatively.
// `max_expected` > 0 additionally refuses a query whose bound is too loose to be worth filtering:
// the expected number of nodes inside the bounding cube of its ball, n * prod_j min(1, 2 r /
// extent_j), exceeds it (seeds from an x-y grid say little about the other coordinates of
// isotropic clouds; such queries go to the one-kernel scan instead of overflowing their lists).
template <int DD>
__device__ __forceinline__ bool filter_query_setup(const double *__restrict__ info, const double *xq,
                                                   bool live, double bound2, float (&nq2)[DD],
                                                   float &thr, double n_nodes = 0.0,
                                                   double max_expected = 0.0) {
    const double M = info[0];
    double Mq = 0.0, qq = 0.0;
#pragma unroll
    for (int j = 0; j < DD; ++j) {
        const double v = live ? xq[j] - info[1 + j] : 0.0;
        const float qf = (float)v;
        nq2[j] = -2.0f * qf;
        qq += (double)qf * (double)qf;
        const double av = v != v ? INFINITY : fabs(v);
        Mq = av > Mq ? av : Mq;
    }
    const double mm = M > Mq ? M : Mq;
    const double a = 4.0 * 5.9604644775390625e-8 * sqrt((double)DD) * mm * (1.0 + 1e-6);
    const double E = 5.9604644775390625e-8 * (double)(DD * (DD + 1)) * M * (M + Mq) * (1.0 + 1e-5) +
                     1e-15 * Mq * Mq;
    const double ra = sqrt(bound2) + a;                              // NaN for bound2 < 0 or NaN
    const double pos = (ra * ra + E) * (1.0 + 1e-9);
    bool usable = live && mm < 1e15 && bound2 < INFINITY && E <= ra * ra;
    if (usable && max_expected > 0.0) {
        double expect = n_nodes;
#pragma unroll
        for (int j = 0; j < DD; ++j) {
            const double ext = info[1 + DD + j];
            expect *= (ext > 2.0 * ra) ? 2.0 * ra / ext : 1.0;
        }
        usable = expect <= max_expected;
    }
    thr = usable ? __double2float_ru(pos - qq) : -INFINITY;          // t (or NaN) never passes -inf
    return usable;
}

// The scan itself: streams node chunk [c0, c1) of the centred shadow through shared memory
// and calls on_group(r, g) for every (query slot r, group of 8 nodes g = node index / 8) whose
// minimum t passes thr[r].  All threads of the CTA must call it (barriers inside).
template <int DD, int NW, int FQ, class OnGroup>
__device__ __forceinline__ void filter_scan(const float *__restrict__ sh, int64_t cap32, int64_t c0,
                                            int64_t c1, const float (&nq2)[FQ][DD],
                                            const float (&thr)[FQ], OnGroup on_group) {
    constexpr int TILE = filter_tile(DD), STAGES = FILTER_STAGES;
    __shared__ __align__(128) float tile[STAGES][DD + 1][TILE];
    __shared__ __align__(8) uint64_t full[STAGES];
    const int T = c1 > c0 ? (int)((c1 - c0 + TILE - 1) / TILE) : 0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < STAGES; ++i) mbar_init(&full[i], 1);
        fence_barrier_init();
    }
    __syncthreads();
    // one thread feeds the stages: DD + 1 bulk copies of <= 4 KB per tile (cap32 is a multiple
    // of the tile, so every copy is a whole number of 16-byte units inside the allocation)
    auto issue = [&](int it) {
        const int st = it % STAGES;
        const int64_t t0 = c0 + (int64_t)it * TILE;
        const int cnt = (int)(c1 - t0 < TILE ? c1 - t0 : TILE);
        const uint32_t bytes = (uint32_t)((cnt + 7) & ~7) * 4u;
        mbar_expect_tx(&full[st], bytes * (DD + 1));
#pragma unroll
        for (int j = 0; j <= DD; ++j)
            bulk_g2s(&tile[st][j][0], sh + (int64_t)j * cap32 + t0, bytes, &full[st]);
    };
    if (threadIdx.x == 0)
        for (int it = 0; it < T && it < STAGES; ++it) issue(it);
    for (int it = 0; it < T; ++it) {
        const int st = it % STAGES;
        const uint32_t phase = (uint32_t)(it / STAGES) & 1u;
        while (!mbar_try_wait(&full[st], phase)) {}
        const int64_t t0 = c0 + (int64_t)it * TILE;
        const int cnt = (int)(c1 - t0 < TILE ? c1 - t0 : TILE);
        const int groups = (cnt + 7) >> 3;
        const int32_t g0 = (int32_t)(t0 >> 3);
        // Two groups of 8 nodes per step: the minima of both groups are combined before the
        // compare, so the hot path pays one compare per query and one branch per 16 nodes; the
        // (rare) candidate path tells the two groups apart again (and ignores the stale second
        // half of an odd tail).
        const int pairs = (groups + 1) >> 1;
#pragma unroll 1
        for (int gp = 0; gp < pairs; ++gp) {
            float mn[2][FQ];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int g = 2 * gp + h;
                uint64_t nd[DD + 1][4];      // 8 nodes: packed pairs of every component and the norm
#pragma unroll
                for (int j = 0; j <= DD; ++j) {
                    const float4 a = *reinterpret_cast<const float4 *>(&tile[st][j][g * 8]);
                    const float4 b = *reinterpret_cast<const float4 *>(&tile[st][j][g * 8 + 4]);
                    nd[j][0] = f2_pack(a.x, a.y); nd[j][1] = f2_pack(a.z, a.w);
                    nd[j][2] = f2_pack(b.x, b.y); nd[j][3] = f2_pack(b.z, b.w);
                }
#pragma unroll
                for (int r = 0; r < FQ; ++r) {
                    uint64_t acc[4];
#pragma unroll
                    for (int p = 0; p < 4; ++p) acc[p] = nd[DD][p];
#pragma unroll
                    for (int j = 0; j < DD; ++j) {
                        const uint64_t qj = f2_pack(nq2[r][j], nq2[r][j]);
#pragma unroll
                        for (int p = 0; p < 4; ++p) acc[p] = f2_fma(nd[j][p], qj, acc[p]);
                    }
                    float v[8];
#pragma unroll
                    for (int p = 0; p < 4; ++p) f2_unpack(acc[p], v[2 * p], v[2 * p + 1]);
                    // depth-2 tree (same four instructions as a chain, half the dependent latency)
                    mn[h][r] = f_min3(f_min3(v[0], v[1], v[2]), f_min3(v[3], v[4], v[5]), fminf(v[6], v[7]));
                }
            }
            bool any = false;
#pragma unroll
            for (int r = 0; r < FQ; ++r) any |= fminf(mn[0][r], mn[1][r]) <= thr[r];
            if (any) {
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int r = 0; r < FQ; ++r)
                        if (mn[h][r] <= thr[r] && 2 * gp + h < groups) on_group(r, g0 + 2 * gp + h);
            }
        }
        __syncthreads();                                         // every warp is done with stage st
        if (threadIdx.x == 0 && it + STAGES < T) issue(it + STAGES);
    }
}
