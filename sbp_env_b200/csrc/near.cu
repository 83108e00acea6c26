// near.cu -- K5: radius queries on the node store with ordered compaction (count -> scan ->
// fill, rows index-ascending): filtered scan for batches (filter.cuh), CTA-per-query and
// thread-per-query kernels otherwise.
//
// Reference behaviour restated (file:line under /root/reference/sbp_env/):
//   prmPlanner.nearest_neighbours             planners/prmPlanner.py:28-44   (all dims, `<`)
//   radius gate on engine.dist                planners/rrtPlanner.py:177-181, engine.py:13-24
//                                             (first two dims, `<=`)
#include <math.h>

#include "common.cuh"
#include "device_fns.cuh"
#include "filter.cuh"

#define KNN_TPB 128        // queries per CTA in the batched radius kernel
#define KNN_TILE 1024      // nodes staged in shared memory per step
#define WIDE_TPB 256

// --------------------------------------------------------- K5: radius -------
// Radius query on the same scan (K5'): a group that passes the conservative fp32 test has
// its 8 nodes decided exactly in fp64 (near_d2 / near_in, the comparison of the unfiltered
// kernels); PASS 0 counts the matches of (query, chunk), PASS 1 writes them at the scanned
// offsets, ascending.  A query that cannot be filtered gets thr = +inf, i.e. every group is
// decided exactly: slower, same result.  DD = coordinates the metric uses (2 for XY).
template <int D, bool XY, int PASS, int NW>
__global__ void __launch_bounds__(NW * 32)
k_near_filtered(const float *__restrict__ sh, const double *__restrict__ info, int64_t cap32,
                const double *__restrict__ soa, int64_t capacity, int64_t n, int S, int64_t chunk,
                const double *__restrict__ X, int64_t m, double thr2, int closed,
                int64_t *__restrict__ counts, const int64_t *__restrict__ offsets,
                int64_t *__restrict__ idx, int64_t idx_capacity, const int64_t *__restrict__ needed) {
    constexpr int DD = XY ? 2 : D;
    constexpr int FQ = FILTER_Q, TPB = NW * 32;
    if (PASS == 1 && *needed > idx_capacity) return;
    const int64_t qb = (int64_t)blockIdx.x * (TPB * FQ) + threadIdx.x;
    const int s = blockIdx.y;
    const int64_t c0 = (int64_t)s * chunk;
    const int64_t c1 = c0 + chunk < n ? c0 + chunk : n;
    float nq2[FQ][DD];
    float thr[FQ];
    int64_t pos[FQ];         // PASS 0: matches so far; PASS 1: next output slot
#pragma unroll
    for (int r = 0; r < FQ; ++r) {
        const int64_t q = qb + (int64_t)r * TPB;
        const bool live = q < m;
        const bool usable = filter_query_setup<DD>(info, X + (live ? q : 0) * D, live, thr2, nq2[r], thr[r]);
        if (live && !usable) {                                       // decide every node exactly:
            thr[r] = INFINITY;                                       // t = |n^|^2 <= inf always passes
#pragma unroll
            for (int j = 0; j < DD; ++j) nq2[r][j] = 0.0f;
        }
        pos[r] = (PASS == 1 && live) ? offsets[q * S + s] : 0;
    }
    filter_scan<DD, NW, FQ>(sh, cap32, c0, c1, nq2, thr, [&](int r, int32_t g) {
        const int64_t q = qb + (int64_t)r * TPB;
        double qv[D];
#pragma unroll
        for (int j = 0; j < D; ++j) qv[j] = j < DD ? X[q * D + j] : 0.0;
        const int64_t b = (int64_t)g * 8;
#pragma unroll 1
        for (int u = 0; u < 8; ++u) {
            const int64_t id = b + u;
            if (id >= c1) break;
            double pv[D];
#pragma unroll
            for (int j = 0; j < D; ++j) pv[j] = j < DD ? soa[(int64_t)j * capacity + id] : 0.0;
            if (near_in(near_d2<D, XY>(pv, qv), thr2, closed)) {
                if (PASS == 1) idx[pos[r]] = id;
                ++pos[r];
            }
        }
    });
    if (PASS == 0) {
#pragma unroll
        for (int r = 0; r < FQ; ++r) {
            const int64_t q = qb + (int64_t)r * TPB;
            if (q < m) counts[q * S + s] = pos[r];
        }
    }
}

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
    if (!wide && n >= 4096 && ctx->tune.near_filter) {
        // filtered path: centred fp32 shadow + packed scan, exact fp64 decision of the groups
        // that pass (k_near_filtered); same counts -> scan -> fill structure as below
        constexpr int DD = XY ? 2 : D;
        constexpr int NW = 2, FT = filter_tile(DD);
        const int QPB = NW * 32 * FILTER_Q;
        const int64_t qblocks = (m + QPB - 1) / QPB;
        const int64_t tiles = (n + FT - 1) / FT;
        int64_t FS = (16 * (int64_t)ctx->sm_count + qblocks - 1) / qblocks;
        if (FS > tiles) FS = tiles;
        if (FS > 256) FS = 256;
        if (FS < 1) FS = 1;
        const int64_t chunk = (tiles + FS - 1) / FS * FT;
        FS = (n + chunk - 1) / chunk;
        const int64_t cap32 = tiles * FT, L = m * FS;
        void *ws = nullptr, *shbuf = nullptr;
        int32_t rc = sbp_ctx_scratch(ctx, 0, (size_t)L * 2 * sizeof(int64_t), &ws);
        if (rc) return rc;
        const size_t hdr = 128 + 128 + (size_t)BOUNDS_BLOCKS * 2 * DD * sizeof(double);   // info | bounds | partials
        rc = sbp_ctx_scratch(ctx, 10, hdr + (size_t)(DD + 1) * cap32 * sizeof(float), &shbuf);
        if (rc) return rc;
        double *info = (double *)shbuf, *bnd = (double *)((char *)shbuf + 128),
               *partial = (double *)((char *)shbuf + 128 + 128);
        float *shadow = (float *)((char *)shbuf + hdr);
        int64_t *counts = (int64_t *)ws, *offsets = counts + L;
        k_bounds_partial<DD><<<BOUNDS_BLOCKS, 256, 0, ctx->stream>>>(s->d_soa, s->capacity, n, partial);
        k_bounds_fold<DD><<<1, 32, 0, ctx->stream>>>(partial, BOUNDS_BLOCKS, bnd);
        int sb = (int)((cap32 + 255) / 256);
        if (sb > ctx->sm_count * 8) sb = ctx->sm_count * 8;
        k_shadow_centered<DD><<<sb, 256, 0, ctx->stream>>>(s->d_soa, s->capacity, n, bnd, cap32, shadow, info);
        dim3 fgrid((unsigned)qblocks, (unsigned)FS);
        k_near_filtered<D, XY, 0, NW><<<fgrid, NW * 32, 0, ctx->stream>>>(
            shadow, info, cap32, s->d_soa, s->capacity, n, (int)FS, chunk, X, m, thr, closed, counts,
            offsets, idx, idx_capacity, needed);
        k_scan_counts<<<1, 1024, 0, ctx->stream>>>(counts, L, (int)FS, m, offsets, row_ptr, needed);
        ctx->launches += 5;
        SBP_CUDA(cudaGetLastError());
        if (idx && idx_capacity > 0) {
            k_near_filtered<D, XY, 1, NW><<<fgrid, NW * 32, 0, ctx->stream>>>(
                shadow, info, cap32, s->d_soa, s->capacity, n, (int)FS, chunk, X, m, thr, closed, counts,
                offsets, idx, idx_capacity, needed);
            ctx->launches++;
            SBP_CUDA(cudaGetLastError());
        }
        return SBP_OK;
    }
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
    SBP_NVTX("sbp_nodes_near");
    SBP_REQUIRE(s && row_ptr && needed && (m == 0 || X), "sbp_nodes_near: NULL argument");
    SBP_REQUIRE(m >= 0 && idx_capacity >= 0, "sbp_nodes_near: negative size");
    SBP_REQUIRE(metric == SBP_METRIC_FULL || metric == SBP_METRIC_XY, "sbp_nodes_near: bad metric");
    SBP_REQUIRE(metric == SBP_METRIC_FULL || s->d >= 2, "sbp_nodes_near: XY metric needs d >= 2");
    SBP_DEVICE(s->ctx);
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

