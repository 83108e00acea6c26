// fused.cu -- one-launch "radius query + edge visibility" for the sequential planners.
//
// Reference loops replaced (file:line under /root/reference/sbp_env/planners/):
//   choose_least_cost_parent   rrtPlanner.py:173-196   for p in nodes: engine.dist(new,p) <= radius
//                                                       and cc.visible(new.pos, p.pos)
//   rewire (linear variant)    rrtPlanner.py:250-266   ... and cc.visible(n.pos, new.pos)
//   get_nearest_free           prmPlanner.py:103-126   cc.visible(node.pos, n.pos) over a radius set
//   Bi-RRT join                birrtPlanner.py:82-87
// In the reference this is an O(N) Python loop per accepted node (91 % of RRT* wall time,
// SURVEY.md 3.2).  Here one kernel scans the node store, and the thread that finds a node
// inside the radius immediately walks the edge (2-D Bresenham or the arm's interpolated
// link sweeps) and appends (index, visible) to a small per-query buffer.  The host sorts the
// handful of hits by index (the reference's loop order) -- one launch, one D2H copy.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "device_fns.cuh"

#define NV_TPB 256

template <int D, bool XY, bool ARM>
__global__ void __launch_bounds__(NV_TPB)
k_near_visible(const double *__restrict__ soa, int64_t capacity, int64_t n, int S, MapView mv,
               double L0, double L1, const double *__restrict__ X, double thr, int closed,
               int direction, int64_t *__restrict__ out_idx, uint8_t *__restrict__ out_vis, int cap,
               int32_t *__restrict__ count, int32_t *__restrict__ flags) {
    const int64_t q = blockIdx.y;
    const int s = blockIdx.x;
    const int64_t chunk = (n + S - 1) / S;
    const int64_t c0 = (int64_t)s * chunk;
    const int64_t c1 = c0 + chunk < n ? c0 + chunk : n;
    double qv[D];
#pragma unroll
    for (int j = 0; j < D; ++j) qv[j] = X[q * D + j];
    int myflag = 0;
    for (int64_t t = c0 + threadIdx.x; t < c1; t += NV_TPB) {
        double pv[D];
#pragma unroll
        for (int j = 0; j < D; ++j) pv[j] = soa[(int64_t)j * capacity + t];
        if (!near_in(near_d2<D, XY>(pv, qv), thr, closed)) continue;
        int vis;
        if (ARM) {
            double a[4] = {qv[0], qv[1], D > 2 ? qv[2] : 0.0, D > 3 ? qv[3] : 0.0};
            double b[4] = {pv[0], pv[1], D > 2 ? pv[2] : 0.0, D > 3 ? pv[3] : 0.0};
            vis = direction == 0 ? arm_edge_seq<false>(mv.words, mv, L0, L1, a, b, myflag)
                                 : arm_edge_seq<false>(mv.words, mv, L0, L1, b, a, myflag);
        } else {
            vis = direction == 0
                      ? edge_visible_seq<false>(mv.words, mv, qv[0], qv[1], pv[0], pv[1], myflag)
                      : edge_visible_seq<false>(mv.words, mv, pv[0], pv[1], qv[0], qv[1], myflag);
        }
        int pos = atomicAdd(count + q, 1);
        if (pos < cap) {
            out_idx[q * cap + pos] = t;
            out_vis[q * cap + pos] = (uint8_t)vis;
        }
    }
    if (myflag && flags) atomicOr(flags, myflag);
}

template <int D>
static int32_t nv_launch(sbp_nodes *s, sbp_map *map, int arm, double L0, double L1, const double *X,
                         int64_t m, double thr, int cl, int metric, int direction, int64_t *out_idx,
                         uint8_t *out_vis, int cap, int32_t *count, int32_t *flags) {
    sbp_ctx *ctx = s->ctx;
    const int64_t n = s->n;
    int64_t want = (4 * (int64_t)ctx->sm_count + m - 1) / m;
    int64_t maxs = (n + 1023) / 1024;
    int S = (int)(want < maxs ? want : maxs);
    if (S < 1) S = 1;
    dim3 grid((unsigned)S, (unsigned)m);
#define NV_GO(XYV, ARMV)                                                                          \
    k_near_visible<D, XYV, ARMV><<<grid, NV_TPB, 0, ctx->stream>>>(                                \
        s->d_soa, s->capacity, n, S, map->view, L0, L1, X, thr, cl, direction, out_idx, out_vis,   \
        cap, count, flags)
    if (metric == SBP_METRIC_XY) { if (arm) NV_GO(true, true); else NV_GO(true, false); }
    else                         { if (arm) NV_GO(false, true); else NV_GO(false, false); }
#undef NV_GO
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

extern "C" int32_t sbp_nodes_near_visible(sbp_nodes *s, sbp_map *map, int32_t arm, double L0,
                                          double L1, const double *X, int64_t m, double r,
                                          int32_t metric, int32_t closed, int32_t direction,
                                          int64_t *out_idx, uint8_t *out_vis, int32_t cap,
                                          int32_t *count, int32_t *flags) {
    SBP_NVTX("sbp_nodes_near_visible");
    SBP_REQUIRE(s && map && X && out_idx && out_vis && count, "sbp_nodes_near_visible: NULL argument");
    SBP_REQUIRE(m >= 1 && m <= 65535 && cap >= 1, "sbp_nodes_near_visible: 1 <= m <= 65535, cap >= 1");
    SBP_REQUIRE(s->d >= 2 && (!arm || s->d == 4), "sbp_nodes_near_visible: d >= 2 (arm: d == 4)");
    SBP_REQUIRE(s->ctx == map->ctx, "sbp_nodes_near_visible: store and map live in different contexts");
    sbp_ctx *ctx = s->ctx;
    SBP_DEVICE(ctx);
    SBP_CUDA(cudaMemsetAsync(count, 0, (size_t)m * sizeof(int32_t), ctx->stream));
    if (s->n == 0) return SBP_OK;
    double thr;
    int cl;
    near_threshold(r, closed ? 1 : 0, &thr, &cl);
    switch (s->d) {
        case 2: return nv_launch<2>(s, map, arm, L0, L1, X, m, thr, cl, metric, direction, out_idx, out_vis, cap, count, flags);
        case 3: return nv_launch<3>(s, map, arm, L0, L1, X, m, thr, cl, metric, direction, out_idx, out_vis, cap, count, flags);
        case 4: return nv_launch<4>(s, map, arm, L0, L1, X, m, thr, cl, metric, direction, out_idx, out_vis, cap, count, flags);
        case 6: return nv_launch<6>(s, map, arm, L0, L1, X, m, thr, cl, metric, direction, out_idx, out_vis, cap, count, flags);
        default:
            sbp_set_error("sbp_nodes_near_visible: dimension %d not instantiated (2, 3, 4, 6)", s->d);
            return SBP_ERR_ARG;
    }
}

// Host form for ONE query: results sorted by node index (the reference's loop order).
// Returns the number of hits in *n_hits; when it exceeds cap the call must be repeated
// with a larger buffer (SBP_ERR_CAPACITY).  Arm guard-band entries are settled with the
// host libm before returning.
extern "C" int32_t sbp_nodes_near_visible_host(sbp_nodes *s, sbp_map *map, int32_t arm, double L0,
                                               double L1, const double *x, double r,
                                               int32_t metric, int32_t closed, int32_t direction,
                                               int64_t *out_idx, uint8_t *out_vis, int32_t cap,
                                               int32_t *n_hits, int32_t *flags) {
    SBP_REQUIRE(s && map && x && out_idx && out_vis && n_hits, "sbp_nodes_near_visible_host: NULL argument");
    SBP_REQUIRE(cap >= 1, "sbp_nodes_near_visible_host: cap >= 1");
    sbp_ctx *ctx = s->ctx;
    SBP_DEVICE(ctx);
    void *dX, *dI, *dV;
    int32_t rc;
    if ((rc = sbp_ctx_scratch(ctx, 2, (size_t)s->d * 8, &dX))) return rc;
    if ((rc = sbp_ctx_scratch(ctx, 3, (size_t)cap * 8, &dI))) return rc;
    if ((rc = sbp_ctx_scratch(ctx, 4, (size_t)cap, &dV))) return rc;
    int32_t *dC = ctx->d_flags + 4, *dF = ctx->d_flags;        // count and flag words
    SBP_CUDA(cudaMemsetAsync(ctx->d_flags, 0, 16 * sizeof(int32_t), ctx->stream));
    SBP_CUDA(cudaMemcpyAsync(dX, x, (size_t)s->d * 8, cudaMemcpyHostToDevice, ctx->stream));
    rc = sbp_nodes_near_visible(s, map, arm, L0, L1, (const double *)dX, 1, r, metric, closed,
                                direction, (int64_t *)dI, (uint8_t *)dV, cap, dC, dF);
    if (rc) return rc;
    SBP_CUDA(cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, 8 * sizeof(int32_t), cudaMemcpyDeviceToHost,
                             ctx->stream));
    SBP_CUDA(cudaStreamSynchronize(ctx->stream));
    int32_t hits = ctx->h_flags[4];
    *n_hits = hits;
    if (flags) *flags = ctx->h_flags[0];
    if (hits > cap) {
        sbp_set_error("sbp_nodes_near_visible_host: %d hits exceed the buffer of %d", hits, cap);
        return SBP_ERR_CAPACITY;
    }
    if (hits == 0) return SBP_OK;
    std::vector<int64_t> hi(hits);
    std::vector<uint8_t> hv(hits);
    SBP_CUDA(cudaMemcpyAsync(hi.data(), dI, (size_t)hits * 8, cudaMemcpyDeviceToHost, ctx->stream));
    SBP_CUDA(cudaMemcpyAsync(hv.data(), dV, (size_t)hits, cudaMemcpyDeviceToHost, ctx->stream));
    SBP_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<int32_t> order(hits);
    for (int i = 0; i < hits; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return hi[a] < hi[b]; });
    for (int i = 0; i < hits; ++i) { out_idx[i] = hi[order[i]]; out_vis[i] = hv[order[i]]; }
    if (arm) {
        for (int i = 0; i < hits; ++i) {
            if (out_vis[i] != SBP_AMBIGUOUS) continue;
            double node[4];
            rc = sbp_nodes_read(s, out_idx[i], 1, node);
            if (rc) return rc;
            const double *c1 = direction == 0 ? x : node, *c2 = direction == 0 ? node : x;
            rc = sbp_arm_resolve_host(map, L0, L1, c1, c2, 1, out_vis + i, nullptr);
            if (rc) return rc;
        }
    }
    return SBP_OK;
}
