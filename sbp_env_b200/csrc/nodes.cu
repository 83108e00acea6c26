// nodes.cu -- device-resident append-only node store; K4 brute-force kNN with
// per-thread top-k and shuffle merge; K5 radius query with ordered compaction.
//
// Reference behaviour restated (file:line under /root/reference/sbp_env/):
//   planners' `poses` arrays and appends      planners/rrtPlanner.py:24-26, :106-113
//   RRTPlanner.find_nearest_neighbour_idx     planners/rrtPlanner.py:268-279
//   np.argsort(distances)[:20] candidates     planners/rrtPlanner.py:151-152, :226-227
//   prmPlanner.nearest_neighbours             planners/prmPlanner.py:28-44
//   radius gate on engine.dist                planners/rrtPlanner.py:177-181, engine.py:13-24
//
// Distance semantics: d = sqrt_rn(((dx*dx + dy*dy) + dz*dz) + ...),
// every operation rounded separately (no FMA), which is bit-identical to
// np.linalg.norm(poses - q, axis=1) for d <= 8.  Results are ordered by
// (d, index) with d the ROUNDED square root: distinct d^2 may collapse to one d,
// in which case the lower index wins, exactly like np.argmin / a stable argsort.
// The hot loop compares d^2 against round-up(d_k * d_k): sqrt_rn is monotone, so
// a candidate with d^2 >= that bound cannot have sqrt < d_k; only candidates
// below it take the slow path that computes the square root.
#include <math.h>

#include "common.cuh"
#include "device_fns.cuh"

#define KNN_TPB 128        // queries per CTA in the batched radius kernel
#define KNN_TILE 1024      // nodes staged in shared memory per step
#define WIDE_TPB 256

// ------------------------------------------------------------ node store ----
extern "C" int32_t sbp_nodes_create(sbp_ctx *ctx, int32_t d, int64_t capacity, sbp_nodes **out) {
    SBP_REQUIRE(ctx && out, "sbp_nodes_create: NULL argument");
    SBP_REQUIRE(d >= 1 && d <= 8, "sbp_nodes_create: 1 <= d <= 8");
    SBP_REQUIRE(capacity >= 0 && capacity < ((int64_t)1 << 31), "sbp_nodes_create: bad capacity");
    SBP_CUDA(cudaSetDevice(ctx->device));
    sbp_nodes *s = new sbp_nodes();
    s->ctx = ctx;
    s->d = d;
    s->capacity = capacity < 1024 ? 1024 : capacity;
    cudaError_t e = cudaMalloc(&s->d_soa, (size_t)s->capacity * d * sizeof(double));
    if (e != cudaSuccess) {
        delete s;
        sbp_set_error("sbp_nodes_create: cudaMalloc failed: %s", cudaGetErrorString(e));
        return SBP_ERR_CUDA;
    }
    *out = s;
    return SBP_OK;
}

extern "C" int32_t sbp_nodes_destroy(sbp_nodes *s) {
    if (!s) return SBP_OK;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    if (s->d_soa) cudaFree(s->d_soa);
    delete s;
    return SBP_OK;
}

__global__ void k_aos_to_soa(const double *__restrict__ X, int64_t m, int d, double *__restrict__ soa,
                             int64_t capacity, int64_t offset) {
    int64_t total = m * d;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        int64_t row = t / d;
        int j = (int)(t - row * d);
        soa[(int64_t)j * capacity + offset + row] = X[t];
    }
}

__global__ void k_soa_to_aos(const double *__restrict__ soa, int64_t capacity, int64_t first,
                             int64_t count, int d, double *__restrict__ X) {
    int64_t total = count * d;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        int64_t row = t / d;
        int j = (int)(t - row * d);
        X[t] = soa[(int64_t)j * capacity + first + row];
    }
}

static int32_t nodes_grow(sbp_nodes *s, int64_t need) {
    if (need <= s->capacity) return SBP_OK;
    int64_t cap = s->capacity;
    while (cap < need) cap *= 2;
    SBP_REQUIRE(cap < ((int64_t)1 << 31), "node store limited to 2^31-1 nodes");
    double *nw = nullptr;
    SBP_CUDA(cudaMalloc(&nw, (size_t)cap * s->d * sizeof(double)));
    for (int j = 0; j < s->d; ++j)
        SBP_CUDA(cudaMemcpyAsync(nw + (int64_t)j * cap, s->d_soa + (int64_t)j * s->capacity,
                                 (size_t)s->n * sizeof(double), cudaMemcpyDeviceToDevice,
                                 s->ctx->stream));
    SBP_CUDA(cudaStreamSynchronize(s->ctx->stream));
    SBP_CUDA(cudaFree(s->d_soa));
    s->d_soa = nw;
    s->capacity = cap;
    return SBP_OK;
}

extern "C" int32_t sbp_nodes_append(sbp_nodes *s, const double *X, int64_t m, int32_t on_device) {
    SBP_REQUIRE(s && (m == 0 || X), "sbp_nodes_append: NULL argument");
    SBP_REQUIRE(m >= 0, "sbp_nodes_append: m < 0");
    if (m == 0) return SBP_OK;
    sbp_ctx *ctx = s->ctx;
    SBP_CUDA(cudaSetDevice(ctx->device));
    int32_t rc = nodes_grow(s, s->n + m);
    if (rc) return rc;
    const double *src = X;
    if (!on_device) {
        void *stage = nullptr;
        rc = sbp_ctx_scratch(ctx, 7, (size_t)m * s->d * sizeof(double), &stage);
        if (rc) return rc;
        SBP_CUDA(cudaMemcpyAsync(stage, X, (size_t)m * s->d * sizeof(double), cudaMemcpyHostToDevice,
                                 ctx->stream));
        src = (const double *)stage;
    }
    int64_t total = m * s->d;
    int blocks = (int)((total + 255) / 256);
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    k_aos_to_soa<<<blocks, 256, 0, ctx->stream>>>(src, m, s->d, s->d_soa, s->capacity, s->n);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    // a pageable host source may be reused by the caller right after return
    if (!on_device) SBP_CUDA(cudaStreamSynchronize(ctx->stream));
    s->n += m;
    return SBP_OK;
}

extern "C" int32_t sbp_nodes_size(sbp_nodes *s, int64_t *n) {
    SBP_REQUIRE(s && n, "sbp_nodes_size: NULL argument");
    *n = s->n;
    return SBP_OK;
}

extern "C" int32_t sbp_nodes_truncate(sbp_nodes *s, int64_t n) {
    SBP_REQUIRE(s, "sbp_nodes_truncate: NULL argument");
    SBP_REQUIRE(n >= 0 && n <= s->n, "sbp_nodes_truncate: n out of range");
    s->n = n;
    return SBP_OK;
}

extern "C" int32_t sbp_nodes_read(sbp_nodes *s, int64_t first, int64_t count, double *out_host) {
    SBP_REQUIRE(s && (count == 0 || out_host), "sbp_nodes_read: NULL argument");
    SBP_REQUIRE(first >= 0 && count >= 0 && first + count <= s->n, "sbp_nodes_read: range");
    if (count == 0) return SBP_OK;
    sbp_ctx *ctx = s->ctx;
    SBP_CUDA(cudaSetDevice(ctx->device));
    void *stage = nullptr;
    int32_t rc = sbp_ctx_scratch(ctx, 7, (size_t)count * s->d * sizeof(double), &stage);
    if (rc) return rc;
    int64_t total = count * s->d;
    int blocks = (int)((total + 255) / 256);
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    k_soa_to_aos<<<blocks, 256, 0, ctx->stream>>>(s->d_soa, s->capacity, first, count, s->d,
                                                  (double *)stage);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    SBP_CUDA(cudaMemcpyAsync(out_host, stage, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost,
                             ctx->stream));
    SBP_CUDA(cudaStreamSynchronize(ctx->stream));
    return SBP_OK;
}

extern "C" int32_t sbp_nodes_rows_device(sbp_nodes *s, int64_t first, int64_t count,
                                         double *out_device) {
    SBP_REQUIRE(s && (count == 0 || out_device), "sbp_nodes_rows_device: NULL argument");
    SBP_REQUIRE(first >= 0 && count >= 0 && first + count <= s->n, "sbp_nodes_rows_device: range");
    if (count == 0) return SBP_OK;
    sbp_ctx *ctx = s->ctx;
    SBP_CUDA(cudaSetDevice(ctx->device));
    int64_t total = count * s->d;
    int blocks = (int)((total + 255) / 256);
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    k_soa_to_aos<<<blocks, 256, 0, ctx->stream>>>(s->d_soa, s->capacity, first, count, s->d,
                                                  out_device);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

// --------------------------------------------------------- distance core ----
// Per-thread sorted list of the K best (d, idx), d ascending, ties keep arrival
// order (callers feed indices in ascending order, so ties keep the lower index).
template <int K>
struct TopK {
    double d[K];
    int32_t i[K];
    double bound2;   // round-up(d[K-1]^2): candidates with d^2 >= bound2 are rejected
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int t = 0; t < K; ++t) { d[t] = INFINITY; i[t] = -1; }
        bound2 = INFINITY;
    }
    __device__ __forceinline__ void offer(double d2, int32_t idx) {
        if (d2 < bound2) {
            double dd = __dsqrt_rn(d2);
            if (dd < d[K - 1]) {
                bool placed = false;
#pragma unroll
                for (int t = K - 1; t >= 0; --t) {
                    if (!placed) {
                        if (t > 0 && dd < d[t - 1]) { d[t] = d[t - 1]; i[t] = i[t - 1]; }
                        else { d[t] = dd; i[t] = idx; placed = true; }
                    }
                }
                bound2 = __dmul_ru(d[K - 1], d[K - 1]);
            }
        }
    }
    // insert an already square-rooted distance (merging partial lists)
    __device__ __forceinline__ void offer_sorted(double dd, int32_t idx) {
        if (dd < d[K - 1]) {
            bool placed = false;
#pragma unroll
            for (int t = K - 1; t >= 0; --t) {
                if (!placed) {
                    if (t > 0 && dd < d[t - 1]) { d[t] = d[t - 1]; i[t] = i[t - 1]; }
                    else { d[t] = dd; i[t] = idx; placed = true; }
                }
            }
        }
    }
    // remove the head (used by the block merge)
    __device__ __forceinline__ void pop() {
#pragma unroll
        for (int t = 0; t + 1 < K; ++t) { d[t] = d[t + 1]; i[t] = i[t + 1]; }
        d[K - 1] = INFINITY; i[K - 1] = -1;
    }
};

// ---------------------------------------------- K4a: batched, thread/query ---
// grid = (ceil(m / KNN_TPB), S).  CTA (bx, s) scans node chunk s for its 128
// queries; node tiles are staged in shared memory (SoA) and read by broadcast.
template <int D, int K, int TPB>
__global__ void __launch_bounds__(TPB)
k_knn_batched(const double *__restrict__ soa, int64_t capacity, int64_t n, int S,
              const double *__restrict__ X, int64_t m, double *__restrict__ pd,
              int32_t *__restrict__ pi) {
    constexpr int TILE = TPB * 8;     // nodes staged in shared memory per step
    __shared__ double tile[D][TILE];
    const int64_t q = (int64_t)blockIdx.x * TPB + threadIdx.x;
    const int s = blockIdx.y;
    const int64_t chunk = (n + S - 1) / S;
    const int64_t c0 = (int64_t)s * chunk;
    const int64_t c1 = c0 + chunk < n ? c0 + chunk : n;
    double qv[D];
#pragma unroll
    for (int j = 0; j < D; ++j) qv[j] = q < m ? X[q * D + j] : 0.0;
    TopK<K> top;
    top.init();
    for (int64_t t0 = c0; t0 < c1; t0 += TILE) {
        int cnt = (int)(c1 - t0 < TILE ? c1 - t0 : TILE);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < D; ++j)
            for (int t = threadIdx.x; t < cnt; t += TPB)
                tile[j][t] = soa[(int64_t)j * capacity + t0 + t];
        __syncthreads();
        // Two-phase scan per block of 32 nodes.  Phase 1 is branch-free: six fp64
        // instructions per pair plus one predicated bit-set, with the top-k state
        // read-only (a fused scan+insert loop made ptxas copy the whole list between
        // unrolled iterations: ~45 instructions per pair).  Phase 2 revisits only the
        // flagged nodes; the bound can only have tightened meanwhile, and offer()
        // re-checks it, so stale flags are harmless.
        for (int b0 = 0; b0 < cnt; b0 += 32) {
            const int bn = cnt - b0 < 32 ? cnt - b0 : 32;
            const double bound2 = top.bound2;
            uint32_t mask = 0;
            if (bn == 32) {
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    double pv[D];
#pragma unroll
                    for (int j = 0; j < D; ++j) pv[j] = tile[j][b0 + t];
                    mask |= (dist2_full<D>(pv, qv) < bound2) ? (1u << t) : 0u;
                }
            } else {
                for (int t = 0; t < bn; ++t) {
                    double pv[D];
#pragma unroll
                    for (int j = 0; j < D; ++j) pv[j] = tile[j][b0 + t];
                    mask |= (dist2_full<D>(pv, qv) < bound2) ? (1u << t) : 0u;
                }
            }
            // lock-step over the warp: round r inserts every lane's r-th flagged node,
            // so insertions of different queries share one pass through offer()
            while (__any_sync(0xffffffffu, mask != 0u)) {
                double d2 = INFINITY;
                int32_t id = -1;
                if (mask) {
                    const int t = __ffs(mask) - 1;
                    mask &= mask - 1;
                    double pv[D];
#pragma unroll
                    for (int j = 0; j < D; ++j) pv[j] = tile[j][b0 + t];
                    d2 = dist2_full<D>(pv, qv);
                    id = (int32_t)(t0 + b0 + t);
                }
                top.offer(d2, id);
            }
        }
    }
    if (q < m) {
        int64_t base = (q * S + s) * K;
#pragma unroll
        for (int t = 0; t < K; ++t) { pd[base + t] = top.d[t]; pi[base + t] = top.i[t]; }
    }
}

// ------------------------------------------------- K4b: wide, CTA/(q,chunk) ---
// For few queries over many nodes (RRT nearest / argsort[:20]): threads stride
// the chunk with coalesced SoA loads, then the CTA merges its per-thread lists by
// repeatedly extracting the (d, idx)-smallest head with warp shuffles.

// K rounds of block-wide argmin over the list heads, key (d, idx).  Round r's winner
// goes to the partial buffers (pd/pi) or, when idx64 is given, to the final row.
template <int K>
__device__ __forceinline__ void block_merge_topk(TopK<K> &top, double *pd, int32_t *pi,
                                                 int64_t *idx64, double *dist, int k_out) {
    __shared__ double wd[WIDE_TPB / 32];
    __shared__ int32_t wi[WIDE_TPB / 32];
    __shared__ int wl[WIDE_TPB / 32];
    __shared__ int win_thread;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rounds = idx64 ? k_out : K;
    for (int r = 0; r < rounds; ++r) {
        double bd = top.d[0];
        int32_t bi = top.i[0];
        int bl = threadIdx.x;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double od = __shfl_xor_sync(FULL, bd, o);
            int32_t oi = __shfl_xor_sync(FULL, bi, o);
            int ol = __shfl_xor_sync(FULL, bl, o);
            // empty heads (idx -1, d inf) lose against anything real
            bool take = (oi >= 0) && (bi < 0 || od < bd || (od == bd && oi < bi));
            if (take) { bd = od; bi = oi; bl = ol; }
        }
        if (lane == 0) { wd[warp] = bd; wi[warp] = bi; wl[warp] = bl; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double fd = wd[0]; int32_t fi = wi[0]; int fl = wl[0];
            for (int w = 1; w < WIDE_TPB / 32; ++w) {
                bool take = (wi[w] >= 0) && (fi < 0 || wd[w] < fd || (wd[w] == fd && wi[w] < fi));
                if (take) { fd = wd[w]; fi = wi[w]; fl = wl[w]; }
            }
            if (idx64) {
                idx64[r] = fi;
                if (dist) dist[r] = fi >= 0 ? fd : INFINITY;
            } else {
                pd[r] = fi >= 0 ? fd : INFINITY;
                pi[r] = fi;
            }
            win_thread = fi >= 0 ? fl : -1;
        }
        __syncthreads();
        if ((int)threadIdx.x == win_thread) top.pop();
        __syncthreads();
    }
}

template <int D, int K>
__global__ void __launch_bounds__(WIDE_TPB)
k_knn_wide(const double *__restrict__ soa, int64_t capacity, int64_t n, int S,
           const double *__restrict__ X, int64_t m, double *__restrict__ pd,
           int32_t *__restrict__ pi, int64_t *__restrict__ idx, double *__restrict__ dist,
           int k_out) {
    const int64_t q = blockIdx.x;
    const int s = blockIdx.y;
    const int64_t chunk = (n + S - 1) / S;
    const int64_t c0 = (int64_t)s * chunk;
    const int64_t c1 = c0 + chunk < n ? c0 + chunk : n;
    double qv[D];
#pragma unroll
    for (int j = 0; j < D; ++j) qv[j] = X[q * D + j];
    TopK<K> top;
    top.init();
    // blocks of UNR nodes per thread: all loads issued first (memory-level
    // parallelism), flags gathered branch-free, flagged nodes inserted afterwards
    constexpr int UNR = 8;
    for (int64_t t0 = c0 + threadIdx.x; t0 < c1; t0 += (int64_t)WIDE_TPB * UNR) {
        double pv[UNR][D];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            int64_t t = t0 + (int64_t)u * WIDE_TPB;
#pragma unroll
            for (int j = 0; j < D; ++j) pv[u][j] = t < c1 ? soa[(int64_t)j * capacity + t] : INFINITY;
        }
        double d2[UNR];
        uint32_t mask = 0;
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            d2[u] = dist2_full<D>(pv[u], qv);      // inf for the padding slots: never flagged
            mask |= (d2[u] < top.bound2) ? (1u << u) : 0u;
        }
        if (mask) {
#pragma unroll
            for (int u = 0; u < UNR; ++u)
                if (mask & (1u << u)) top.offer(d2[u], (int32_t)(t0 + (int64_t)u * WIDE_TPB));
        }
    }
    if (S == 1) block_merge_topk<K>(top, nullptr, nullptr, idx + q * k_out, dist ? dist + q * k_out : nullptr, k_out);
    else block_merge_topk<K>(top, pd + (q * S + s) * K, pi + (q * S + s) * K, nullptr, nullptr, 0);
}

// Second level of the wide path: one CTA per query merges its S partial lists
// (S up to 1024; thread t takes lists t, t+256, ...; ascending chunk order keeps
// the lower index first among equal distances).
template <int K>
__global__ void __launch_bounds__(WIDE_TPB)
k_knn_merge_wide(const double *__restrict__ pd, const int32_t *__restrict__ pi, int S, int k_out,
                 int64_t *__restrict__ idx, double *__restrict__ dist) {
    const int64_t q = blockIdx.x;
    TopK<K> top;
    top.init();
    for (int s = threadIdx.x; s < S; s += WIDE_TPB) {
        const double *d0 = pd + (q * S + s) * K;
        const int32_t *i0 = pi + (q * S + s) * K;
#pragma unroll
        for (int t = 0; t < K; ++t)
            if (i0[t] >= 0) top.offer_sorted(d0[t], i0[t]);
    }
    block_merge_topk<K>(top, nullptr, nullptr, idx + q * k_out, dist ? dist + q * k_out : nullptr, k_out);
}

// ----------------------------------------------------- K4c: chunk merge ------
// One thread per query merges S sorted partial lists (chunk s holds indices below
// chunk s+1, so on equal d the lower chunk wins) into the final [m][k] rows.
__global__ void k_knn_merge(const double *__restrict__ pd, const int32_t *__restrict__ pi, int64_t m,
                            int S, int K, int k, int64_t *__restrict__ idx,
                            double *__restrict__ dist) {
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    const double *d0 = pd + q * S * K;
    const int32_t *i0 = pi + q * S * K;
    if (S == 1) {
        for (int t = 0; t < k; ++t) {
            idx[q * k + t] = i0[t];
            if (dist) dist[q * k + t] = i0[t] >= 0 ? d0[t] : INFINITY;
        }
        return;
    }
    unsigned char head[64];
    for (int s = 0; s < S; ++s) head[s] = 0;
    for (int t = 0; t < k; ++t) {
        int best = -1;
        double bd = INFINITY;
        for (int s = 0; s < S; ++s) {
            int h = head[s];
            if (h >= K) continue;
            int32_t ci = i0[s * K + h];
            if (ci < 0) continue;
            double cd = d0[s * K + h];
            if (best < 0 || cd < bd) { best = s; bd = cd; }
        }
        if (best < 0) {
            idx[q * k + t] = -1;
            if (dist) dist[q * k + t] = INFINITY;
        } else {
            idx[q * k + t] = i0[best * K + head[best]];
            if (dist) dist[q * k + t] = bd;
            head[best]++;
        }
    }
}

int g_knn_chunks = 0;   // tuning knob: node chunks of the batched kNN kernel (0 = auto)

template <int D, int K>
static int32_t knn_launch(sbp_nodes *s, const double *X, int64_t m, int32_t k, int64_t *idx,
                          double *dist) {
    sbp_ctx *ctx = s->ctx;
    const int64_t n = s->n;
    bool wide = m < 256;
    int S, tpb = 128;
    if (wide) {
        int64_t want = (4 * (int64_t)ctx->sm_count + m - 1) / m;
        int64_t maxs = (n + 2047) / 2048;
        S = (int)(want < maxs ? want : maxs);
    } else {
        // queries per CTA: 128, or 32 when single-warp CTAs balance the SMs better
        // (work per SM is ceil(tiles / SMs) * tpb query rows)
        int64_t sm = ctx->sm_count;
        int64_t w128 = ((m + 127) / 128 + sm - 1) / sm * 128;
        int64_t w32 = ((m + 31) / 32 + sm - 1) / sm * 32;
        tpb = (w32 * 100 < w128 * 95) ? 32 : 128;
        int64_t tiles = (m + tpb - 1) / tpb;
        int64_t want = (2 * sm + tiles - 1) / tiles;
        int64_t maxs = (n + 4095) / 4096;
        S = (int)(want < maxs ? want : maxs);
    }
    if (!wide && g_knn_chunks > 0) S = g_knn_chunks;
    if (S < 1) S = 1;
    if (S > (wide ? 1024 : 64)) S = wide ? 1024 : 64;
    void *pd = nullptr, *pi = nullptr;
    int32_t rc = sbp_ctx_scratch(ctx, 0, (size_t)m * S * K * sizeof(double), &pd);
    if (rc) return rc;
    rc = sbp_ctx_scratch(ctx, 1, (size_t)m * S * K * sizeof(int32_t), &pi);
    if (rc) return rc;
    if (wide) {
        dim3 grid((unsigned)m, (unsigned)S);
        k_knn_wide<D, K><<<grid, WIDE_TPB, 0, ctx->stream>>>(s->d_soa, s->capacity, n, S, X, m,
                                                              (double *)pd, (int32_t *)pi, idx, dist, k);
        ctx->launches++;
        SBP_CUDA(cudaGetLastError());
        if (S > 1) {
            k_knn_merge_wide<K><<<(unsigned)m, WIDE_TPB, 0, ctx->stream>>>(
                (const double *)pd, (const int32_t *)pi, S, k, idx, dist);
            ctx->launches++;
            SBP_CUDA(cudaGetLastError());
        }
        return SBP_OK;
    }
    if (tpb == 32) {
        dim3 grid((unsigned)((m + 31) / 32), (unsigned)S);
        k_knn_batched<D, K, 32><<<grid, 32, 0, ctx->stream>>>(s->d_soa, s->capacity, n, S, X, m,
                                                              (double *)pd, (int32_t *)pi);
    } else {
        dim3 grid((unsigned)((m + 127) / 128), (unsigned)S);
        k_knn_batched<D, K, 128><<<grid, 128, 0, ctx->stream>>>(s->d_soa, s->capacity, n, S, X, m,
                                                                (double *)pd, (int32_t *)pi);
    }
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    k_knn_merge<<<(unsigned)((m + 127) / 128), 128, 0, ctx->stream>>>(
        (const double *)pd, (const int32_t *)pi, m, S, K, k, idx, dist);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

template <int D>
static int32_t knn_dispatch_k(sbp_nodes *s, const double *X, int64_t m, int32_t k, int64_t *idx,
                              double *dist) {
    // list length K >= k; 11 and 21 are the PRM / RRT* cases (k = 10 or 20 plus self)
    if (k <= 1) return knn_launch<D, 1>(s, X, m, k, idx, dist);
    if (k <= 2) return knn_launch<D, 2>(s, X, m, k, idx, dist);
    if (k <= 4) return knn_launch<D, 4>(s, X, m, k, idx, dist);
    if (k <= 8) return knn_launch<D, 8>(s, X, m, k, idx, dist);
    if (k <= 11) return knn_launch<D, 11>(s, X, m, k, idx, dist);
    if (k <= 16) return knn_launch<D, 16>(s, X, m, k, idx, dist);
    if (k <= 21) return knn_launch<D, 21>(s, X, m, k, idx, dist);
    return knn_launch<D, 32>(s, X, m, k, idx, dist);
}

__global__ void k_fill_empty_knn(int64_t total, int64_t *idx, double *dist) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        idx[t] = -1;
        if (dist) dist[t] = INFINITY;
    }
}

extern "C" int32_t sbp_nodes_knn(sbp_nodes *s, const double *X, int64_t m, int32_t k,
                                 int32_t precision, int64_t *idx, double *dist) {
    SBP_REQUIRE(s && (m == 0 || (X && idx)), "sbp_nodes_knn: NULL argument");
    SBP_REQUIRE(m >= 0, "sbp_nodes_knn: m < 0");
    SBP_REQUIRE(k >= 1 && k <= 32, "sbp_nodes_knn: 1 <= k <= 32");
    SBP_REQUIRE(precision == 64, "sbp_nodes_knn: only the bit-exact fp64 path exists (precision=64)");
    if (m == 0) return SBP_OK;
    SBP_CUDA(cudaSetDevice(s->ctx->device));
    if (s->n == 0) {
        k_fill_empty_knn<<<64, 256, 0, s->ctx->stream>>>(m * k, idx, dist);
        s->ctx->launches++;
        SBP_CUDA(cudaGetLastError());
        return SBP_OK;
    }
    switch (s->d) {
        case 2: return knn_dispatch_k<2>(s, X, m, k, idx, dist);
        case 3: return knn_dispatch_k<3>(s, X, m, k, idx, dist);
        case 4: return knn_dispatch_k<4>(s, X, m, k, idx, dist);
        case 6: return knn_dispatch_k<6>(s, X, m, k, idx, dist);
        default:
            sbp_set_error("sbp_nodes_knn: dimension %d not instantiated (2, 3, 4, 6)", s->d);
            return SBP_ERR_ARG;
    }
}

// --------------------------------------------------------- K5: radius -------
// Membership is decided on d^2 against a host-computed threshold (see
// near_threshold below), so no square root is evaluated per pair.
// wide: CTA per (query, chunk).  PASS 0 counts, PASS 1 writes indices in order.
template <int D, bool XY, int PASS>
__global__ void __launch_bounds__(WIDE_TPB)
k_near_wide(const double *__restrict__ soa, int64_t capacity, int64_t n, int S,
            const double *__restrict__ X, int64_t m, double thr, int closed,
            int64_t *__restrict__ counts, const int64_t *__restrict__ offsets,
            int64_t *__restrict__ idx, int64_t idx_capacity, const int64_t *__restrict__ needed) {
    const int64_t q = blockIdx.x;
    const int s = blockIdx.y;
    if (PASS == 1 && *needed > idx_capacity) return;
    const int64_t chunk = (n + S - 1) / S;
    const int64_t c0 = (int64_t)s * chunk;
    const int64_t c1 = c0 + chunk < n ? c0 + chunk : n;
    constexpr int DD = XY ? 2 : D;
    double qv[D];
#pragma unroll
    for (int j = 0; j < D; ++j) qv[j] = j < DD ? X[q * D + j] : 0.0;
    __shared__ int wsum[WIDE_TPB / 32];
    __shared__ int64_t base_s;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t mycount = 0;
    if (PASS == 1 && threadIdx.x == 0) base_s = offsets[q * S + s];
    if (PASS == 1) __syncthreads();
    for (int64_t t0 = c0; t0 < c1; t0 += WIDE_TPB) {
        int64_t t = t0 + threadIdx.x;
        bool in = false;
        if (t < c1) {
            double pv[D];
#pragma unroll
            for (int j = 0; j < D; ++j) pv[j] = j < DD ? soa[(int64_t)j * capacity + t] : 0.0;
            in = near_in(near_d2<D, XY>(pv, qv), thr, closed);
        }
        if (PASS == 0) {
            mycount += in ? 1 : 0;
        } else {
            unsigned bal = __ballot_sync(FULL, in);
            int pre = __popc(bal & ((1u << lane) - 1u));
            if (lane == 0) wsum[warp] = __popc(bal);
            __syncthreads();
            int woff = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < WIDE_TPB / 32; ++w) {
                int v = wsum[w];
                if (w < warp) woff += v;
                tot += v;
            }
            int64_t b = base_s;
            if (in) idx[b + woff + pre] = t;
            __syncthreads();
            if (threadIdx.x == 0) base_s = b + tot;
            __syncthreads();
        }
    }
    if (PASS == 0) {
        for (int o = 16; o > 0; o >>= 1) mycount += __shfl_down_sync(FULL, mycount, o);
        __shared__ int64_t wc[WIDE_TPB / 32];
        if (lane == 0) wc[warp] = mycount;
        __syncthreads();
        if (threadIdx.x == 0) {
            int64_t tot = 0;
            for (int w = 0; w < WIDE_TPB / 32; ++w) tot += wc[w];
            counts[q * S + s] = tot;
        }
    }
}

// batched: thread per query over shared-memory node tiles; same branch-free 32-node
// blocks as the kNN scan (flags gathered in a mask, then counted or written in order).
template <int D, bool XY, int PASS, int TPB>
__global__ void __launch_bounds__(TPB)
k_near_batched(const double *__restrict__ soa, int64_t capacity, int64_t n, int S,
               const double *__restrict__ X, int64_t m, double thr, int closed,
               int64_t *__restrict__ counts, const int64_t *__restrict__ offsets,
               int64_t *__restrict__ idx, int64_t idx_capacity,
               const int64_t *__restrict__ needed) {
    constexpr int DD = XY ? 2 : D;
    constexpr int TILE = TPB * 8;
    __shared__ double tile[DD][TILE];
    if (PASS == 1 && *needed > idx_capacity) return;
    const int64_t q = (int64_t)blockIdx.x * TPB + threadIdx.x;
    const int s = blockIdx.y;
    const int64_t chunk = (n + S - 1) / S;
    const int64_t c0 = (int64_t)s * chunk;
    const int64_t c1 = c0 + chunk < n ? c0 + chunk : n;
    double qv[D];
#pragma unroll
    for (int j = 0; j < D; ++j) qv[j] = (q < m && j < DD) ? X[q * D + j] : 0.0;
    int64_t cnt_or_pos = 0;
    if (PASS == 1 && q < m) cnt_or_pos = offsets[q * S + s];
    for (int64_t t0 = c0; t0 < c1; t0 += TILE) {
        int cnt = (int)(c1 - t0 < TILE ? c1 - t0 : TILE);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < DD; ++j)
            for (int t = threadIdx.x; t < cnt; t += TPB)
                tile[j][t] = soa[(int64_t)j * capacity + t0 + t];
        __syncthreads();
        for (int b0 = 0; b0 < cnt; b0 += 32) {
            const int bn = cnt - b0 < 32 ? cnt - b0 : 32;
            uint32_t mask = 0;
            if (bn == 32) {
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    double pv[D];
#pragma unroll
                    for (int j = 0; j < D; ++j) pv[j] = j < DD ? tile[j < DD ? j : 0][b0 + t] : 0.0;
                    mask |= near_in(near_d2<D, XY>(pv, qv), thr, closed) ? (1u << t) : 0u;
                }
            } else {
                for (int t = 0; t < bn; ++t) {
                    double pv[D];
#pragma unroll
                    for (int j = 0; j < D; ++j) pv[j] = j < DD ? tile[j < DD ? j : 0][b0 + t] : 0.0;
                    mask |= near_in(near_d2<D, XY>(pv, qv), thr, closed) ? (1u << t) : 0u;
                }
            }
            if (PASS == 0) {
                cnt_or_pos += __popc(mask);
            } else if (q < m) {
                while (mask) {                                   // ascending node index
                    const int t = __ffs(mask) - 1;
                    mask &= mask - 1;
                    idx[cnt_or_pos++] = t0 + b0 + t;
                }
            }
        }
    }
    if (PASS == 0 && q < m) counts[q * S + s] = cnt_or_pos;
}

// Exclusive scan of counts[L] (single CTA; L = m*S) -> offsets[L]; row_ptr[q] =
// offsets[q*S], row_ptr[m] = total, *needed = total.
__global__ void __launch_bounds__(1024)
k_scan_counts(const int64_t *__restrict__ counts, int64_t L, int S, int64_t m,
              int64_t *__restrict__ offsets, int64_t *__restrict__ row_ptr,
              int64_t *__restrict__ needed) {
    __shared__ int64_t part[1024];
    const int tid = threadIdx.x;
    const int64_t seg = (L + 1023) / 1024;
    const int64_t a = (int64_t)tid * seg, b = a + seg < L ? a + seg : L;
    int64_t sum = 0;
    for (int64_t t = a; t < b; ++t) sum += counts[t];
    part[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        int64_t v = tid >= o ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    int64_t run = part[tid] - sum;   // exclusive prefix of my segment
    for (int64_t t = a; t < b; ++t) {
        offsets[t] = run;
        if (t % S == 0) row_ptr[t / S] = run;
        run += counts[t];
    }
    if (tid == 1023) {
        row_ptr[m] = part[1023];
        *needed = part[1023];
    }
}

template <int D, bool XY>
static int32_t near_launch(sbp_nodes *s, const double *X, int64_t m, double thr, int closed,
                           int64_t *row_ptr, int64_t *idx, int64_t idx_capacity, int64_t *needed) {
    sbp_ctx *ctx = s->ctx;
    const int64_t n = s->n;
    bool wide = m < 256;
    int S, tpb = 128;
    if (wide) {
        int64_t want = (4 * (int64_t)ctx->sm_count + m - 1) / m;
        int64_t maxs = (n + 2047) / 2048;
        S = (int)(want < maxs ? want : maxs);
    } else {
        int64_t sm = ctx->sm_count;
        int64_t w128 = ((m + 127) / 128 + sm - 1) / sm * 128;
        int64_t w32 = ((m + 31) / 32 + sm - 1) / sm * 32;
        tpb = (w32 * 100 < w128 * 95) ? 32 : 128;
        int64_t tiles = (m + tpb - 1) / tpb;
        // no per-chunk restart cost here (unlike the kNN lists): aim for >= 24 warps per SM
        int64_t want = (24 * sm * 32 / tpb + tiles - 1) / tiles;
        int64_t maxs = (n + 4095) / 4096;
        S = (int)(want < maxs ? want : maxs);
    }
    if (S < 1) S = 1;
    if (S > 1024) S = 1024;
    const int64_t L = m * S;
    void *ws = nullptr;
    int32_t rc = sbp_ctx_scratch(ctx, 0, (size_t)L * 2 * sizeof(int64_t), &ws);
    if (rc) return rc;
    int64_t *counts = (int64_t *)ws, *offsets = counts + L;
    dim3 gw((unsigned)m, (unsigned)S), gb((unsigned)((m + tpb - 1) / tpb), (unsigned)S);
#define NEAR_BATCHED(PASSV)                                                                        \
    do {                                                                                            \
        if (tpb == 32)                                                                              \
            k_near_batched<D, XY, PASSV, 32><<<gb, 32, 0, ctx->stream>>>(                           \
                s->d_soa, s->capacity, n, S, X, m, thr, closed, counts, offsets, idx, idx_capacity, \
                needed);                                                                            \
        else                                                                                        \
            k_near_batched<D, XY, PASSV, 128><<<gb, 128, 0, ctx->stream>>>(                         \
                s->d_soa, s->capacity, n, S, X, m, thr, closed, counts, offsets, idx, idx_capacity, \
                needed);                                                                            \
    } while (0)
    if (wide)
        k_near_wide<D, XY, 0><<<gw, WIDE_TPB, 0, ctx->stream>>>(s->d_soa, s->capacity, n, S, X, m, thr,
                                                                closed, counts, offsets, idx,
                                                                idx_capacity, needed);
    else
        NEAR_BATCHED(0);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    k_scan_counts<<<1, 1024, 0, ctx->stream>>>(counts, L, S, m, offsets, row_ptr, needed);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    if (idx && idx_capacity > 0) {
        if (wide)
            k_near_wide<D, XY, 1><<<gw, WIDE_TPB, 0, ctx->stream>>>(s->d_soa, s->capacity, n, S, X, m,
                                                                    thr, closed, counts, offsets, idx,
                                                                    idx_capacity, needed);
        else
            NEAR_BATCHED(1);
        ctx->launches++;
        SBP_CUDA(cudaGetLastError());
    }
#undef NEAR_BATCHED
    return SBP_OK;
}

__global__ void k_zero_rowptr(int64_t m, int64_t *row_ptr, int64_t *needed) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t <= m;
         t += (int64_t)gridDim.x * blockDim.x)
        row_ptr[t] = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) *needed = 0;
}

extern "C" int32_t sbp_nodes_near(sbp_nodes *s, const double *X, int64_t m, double r,
                                  int32_t metric, int32_t closed, int64_t *row_ptr, int64_t *idx,
                                  int64_t idx_capacity, int64_t *needed) {
    SBP_REQUIRE(s && row_ptr && needed && (m == 0 || X), "sbp_nodes_near: NULL argument");
    SBP_REQUIRE(m >= 0 && idx_capacity >= 0, "sbp_nodes_near: negative size");
    SBP_REQUIRE(metric == SBP_METRIC_FULL || metric == SBP_METRIC_XY, "sbp_nodes_near: bad metric");
    SBP_REQUIRE(metric == SBP_METRIC_FULL || s->d >= 2, "sbp_nodes_near: XY metric needs d >= 2");
    SBP_CUDA(cudaSetDevice(s->ctx->device));
    double thr;
    int cl;
    near_threshold(r, closed ? 1 : 0, &thr, &cl);
    if (m == 0 || s->n == 0) {
        k_zero_rowptr<<<64, 256, 0, s->ctx->stream>>>(m, row_ptr, needed);
        s->ctx->launches++;
        SBP_CUDA(cudaGetLastError());
        return SBP_OK;
    }
#define SBP_NEAR_CASE(DV)                                                                        \
    case DV:                                                                                      \
        return metric == SBP_METRIC_XY                                                            \
                   ? near_launch<DV, true>(s, X, m, thr, cl, row_ptr, idx, idx_capacity, needed)  \
                   : near_launch<DV, false>(s, X, m, thr, cl, row_ptr, idx, idx_capacity, needed);
    switch (s->d) {
        SBP_NEAR_CASE(2)
        SBP_NEAR_CASE(3)
        SBP_NEAR_CASE(4)
        SBP_NEAR_CASE(6)
        default:
            sbp_set_error("sbp_nodes_near: dimension %d not instantiated (2, 3, 4, 6)", s->d);
            return SBP_ERR_ARG;
    }
#undef SBP_NEAR_CASE
}

// -------------------------------------------- batched planner primitives -----
// Endpoints of candidate edges (row i -> nbr[i][j]) for a following visible call:
// the batched form of PRMPlanner.build_graph's inner loop (prmPlanner.py:91-101),
// which calls visible(m_g.pos, v.pos): P = neighbour, Q = row node.
__global__ void k_gather_edges(const double *__restrict__ soa, int64_t capacity, int64_t n, int d,
                               const int64_t *__restrict__ nbr, int64_t m, int k, int64_t first_row,
                               double *__restrict__ P, double *__restrict__ Q,
                               uint8_t *__restrict__ valid) {
    int64_t total = m * k;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        int64_t row = first_row + t / k;
        int64_t nb = nbr[t];
        bool ok = nb >= 0 && nb < n && nb != row && row < n;
        int64_t src = ok ? nb : (row < n ? row : 0);
        int64_t dst = row < n ? row : 0;
        for (int j = 0; j < d; ++j) {
            P[t * d + j] = soa[(int64_t)j * capacity + src];
            Q[t * d + j] = soa[(int64_t)j * capacity + dst];
        }
        if (valid) valid[t] = ok ? 1 : 0;
    }
}

extern "C" int32_t sbp_nodes_gather_edges(sbp_nodes *s, const int64_t *nbr, int64_t m, int32_t k,
                                          int64_t first_row, double *P, double *Q,
                                          uint8_t *valid) {
    SBP_REQUIRE(s && (m == 0 || (nbr && P && Q)), "sbp_nodes_gather_edges: NULL argument");
    SBP_REQUIRE(m >= 0 && k >= 1 && first_row >= 0, "sbp_nodes_gather_edges: bad sizes");
    if (m == 0) return SBP_OK;
    sbp_ctx *ctx = s->ctx;
    SBP_CUDA(cudaSetDevice(ctx->device));
    int64_t total = m * k;
    int blocks = (int)((total + 255) / 256);
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    k_gather_edges<<<blocks, 256, 0, ctx->stream>>>(s->d_soa, s->capacity, s->n, s->d, nbr, m, k,
                                                    first_row, P, Q, valid);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

// kNN rows -> candidate edges, self entry dropped (see sbp_gpu.h).  One thread per
// (row, output column): the position of the self column is found by a short scan of
// the row's k_in entries.
__global__ void k_knn_edges(const double *__restrict__ soa, int64_t capacity, int64_t n, int d,
                            const int64_t *__restrict__ nbr_in, int64_t m, int k_in, int k_out,
                            int64_t first_row, const double *__restrict__ X,
                            int64_t *__restrict__ nbr_out, double *__restrict__ P,
                            double *__restrict__ Q, uint8_t *__restrict__ valid) {
    int64_t total = m * k_out;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = t / k_out;
        int j = (int)(t - i * k_out);
        const int64_t *row = nbr_in + i * k_in;
        int src_col = j;
        int64_t self = -1;
        if (!X) {
            self = first_row + i;
            int drop = k_in - 1;
            for (int c = 0; c < k_in; ++c)
                if (row[c] == self) { drop = c; break; }
            src_col = j < drop ? j : j + 1;
        }
        int64_t nb = row[src_col];
        bool ok = nb >= 0 && nb < n;
        nbr_out[t] = nb;
        for (int c = 0; c < d; ++c) {
            double qv = X ? X[i * d + c] : soa[(int64_t)c * capacity + self];
            P[t * d + c] = ok ? soa[(int64_t)c * capacity + nb] : qv;
            Q[t * d + c] = qv;
        }
        valid[t] = ok ? 1 : 0;
    }
}

extern "C" int32_t sbp_nodes_knn_edges(sbp_nodes *s, const int64_t *nbr_in, int64_t m,
                                       int32_t k_in, int64_t first_row, const double *X,
                                       int64_t *nbr_out, double *P, double *Q, uint8_t *valid) {
    SBP_REQUIRE(s && (m == 0 || (nbr_in && nbr_out && P && Q && valid)),
                "sbp_nodes_knn_edges: NULL argument");
    int k_out = X ? k_in : k_in - 1;
    SBP_REQUIRE(m >= 0 && k_out >= 1, "sbp_nodes_knn_edges: bad sizes");
    SBP_REQUIRE(X || (first_row >= 0 && first_row + m <= s->n), "sbp_nodes_knn_edges: row range");
    if (m == 0) return SBP_OK;
    sbp_ctx *ctx = s->ctx;
    SBP_CUDA(cudaSetDevice(ctx->device));
    int64_t total = m * k_out;
    int blocks = (int)((total + 255) / 256);
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    k_knn_edges<<<blocks, 256, 0, ctx->stream>>>(s->d_soa, s->capacity, s->n, s->d, nbr_in, m, k_in,
                                                 k_out, first_row, X, nbr_out, P, Q, valid);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}
