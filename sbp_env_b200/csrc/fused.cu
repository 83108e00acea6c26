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

// =====================================================================================
// sbp_tree: device-resident cost / parent arrays next to a node store, and the RRT* inner loop
// of one accepted sample as ONE launch:
//   choose_least_cost_parent (linear path)   rrtPlanner.py:173-196
//   add_newnode                              rrtPlanner.py:106-113
//   rewire (linear variant)                  rrtPlanner.py:250-266
// Stage 1 (all CTAs) is k_near_visible's scan: a thread that finds a node inside the radius walks
// the edge new -> node and appends (index, visible, c = sqrt_rn(dx*dx + dy*dy)) to a candidate list.
// The CTA that finishes last runs stage 2: arg-min of cost[p] + c over the visible candidates in the
// reference's order (the current nearest node first, then ascending index, strict <), the new node's
// cost and parent, the append of its pose, and the rewire pass (new.cost + c < cost[n] -> re-parent),
// all on the device arrays; the outcome goes to a pinned host struct.
//
// Exactness.  The reference evaluates c with engine.dist = sqrt(dx**2 + dy**2) through libm pow,
// which differs from the correctly rounded product used here by <= 1 ulp for ~0.1 % of inputs
// (DESIGN.md 4).  A comparison the device takes is therefore only trusted when its two sides are
// more than 64 ulp apart (and a candidate only when its d^2 is more than 4 ulp from the radius
// threshold); otherwise NOTHING is committed and the status is SBP_EXTEND_AMBIGUOUS: the caller
// decides that sample with the host path (planner_ops) and pushes the result with sbp_tree_set.
// Costs the device derives (cost[parent] + c) may be 1 ulp off the reference's: the caller, who
// knows the exact value from one engine.dist call, hands it to the NEXT extend call as a `fix`
// (applied before any cost is read), so the device arrays hold the reference's values bit for bit.
#define EXT_CAND 4096

struct sbp_tree {
    sbp_nodes *nodes = nullptr;
    double *d_cost = nullptr;
    int32_t *d_parent = nullptr;
    int64_t capacity = 0;
    int32_t *d_cidx = nullptr;     // candidates of the running call
    double *d_cc = nullptr;
    uint8_t *d_cvis = nullptr;
    int32_t *d_ctr = nullptr;      // [0] candidates, [1] finished CTAs, [2] near-threshold flag
    sbp_extend_result *h_res = nullptr;   // pinned, mapped
};

struct ExtFixes { int32_t n; int32_t idx[SBP_EXTEND_MAX_FIXES]; double cost[SBP_EXTEND_MAX_FIXES]; };

__device__ __forceinline__ double ext_ulps(double a, double b, double scale) {
    // |a - b| in units of ulp(scale)
    const double u = fabs(scale) * 2.220446049250313e-16 + 4.9e-324;
    return fabs(a - b) / u;
}

__global__ void __launch_bounds__(NV_TPB)
k_tree_extend(double *__restrict__ soa, float *__restrict__ soa32, double *__restrict__ maxabs, int64_t capacity,
              int64_t n, MapView mv, double qx, double qy, double thr, int closed, double thr_lo, double thr_hi,
              int64_t nn, int do_append, int do_rewire, double *__restrict__ cost, int32_t *__restrict__ parent,
              int32_t *__restrict__ cidx, double *__restrict__ cc, uint8_t *__restrict__ cvis,
              int32_t *__restrict__ ctr, ExtFixes fixes, sbp_extend_result *__restrict__ res) {
    __shared__ int s_last;
    __shared__ double s_key[NV_TPB / 32];
    __shared__ int32_t s_who[NV_TPB / 32];
    __shared__ int s_flag, s_cnt, s_chk;
    // ---- stage 1: scan, walk, collect
    int myflag = 0;
    for (int64_t t = (int64_t)blockIdx.x * NV_TPB + threadIdx.x; t < n; t += (int64_t)gridDim.x * NV_TPB) {
        const double px = soa[t], py = soa[capacity + t];
        const double a = __dsub_rn(qx, px), b = __dsub_rn(qy, py);          // engine.dist: p1 - p2
        const double d2 = __dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b));
        if (d2 >= thr_lo && d2 <= thr_hi) atomicOr(ctr + 2, 1);             // too close to the radius to call
        if (!near_in(d2, thr, closed)) continue;
        const int vis = edge_visible_seq<false>(mv.words, mv, qx, qy, px, py, myflag) ? 1 : 0;
        const int pos = atomicAdd(ctr, 1);
        if (pos < EXT_CAND) { cidx[pos] = (int32_t)t; cc[pos] = __dsqrt_rn(d2); cvis[pos] = (uint8_t)vis; }
    }
    if (myflag) atomicOr(ctr + 2, 2);                                        // NaN / inf coordinates: host path
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ctr + 1, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // ---- stage 2 (one CTA)
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int m = *(volatile int32_t *)ctr;
    const int near_flag = *(volatile int32_t *)(ctr + 2);
    for (int i = tid; i < fixes.n; i += NV_TPB) cost[fixes.idx[i]] = fixes.cost[i];
    if (tid == 0) { s_flag = 0; s_cnt = 0; s_chk = 0; }
    __syncthreads();
    int status = SBP_EXTEND_OK;
    if (near_flag || m > EXT_CAND) status = SBP_EXTEND_AMBIGUOUS;
    // the current nearest node opens the reference's loop (nn.cost + c_nn, rrtPlanner.py:178-188)
    double c_nn = 0.0, key_nn = INFINITY;
    if (nn >= 0) {
        const double a = __dsub_rn(qx, soa[nn]), b = __dsub_rn(qy, soa[capacity + nn]);
        c_nn = __dsqrt_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)));
        key_nn = __dadd_rn(cost[nn], c_nn);
    }
    const int mm = m < EXT_CAND ? m : EXT_CAND;
    // best = smallest key; the nearest node wins ties, then the lower index (strict < in index order)
    double bk = INFINITY;
    int32_t bw = 0x7fffffff;
    for (int j = tid; j < mm; j += NV_TPB) {
        if (!cvis[j] || cidx[j] == (int32_t)nn) continue;
        const double key = __dadd_rn(cost[cidx[j]], cc[j]);
        if (key < bk || (key == bk && cidx[j] < bw)) { bk = key; bw = cidx[j]; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double ok = __shfl_down_sync(FULL, bk, o);
        const int32_t ow = __shfl_down_sync(FULL, bw, o);
        if (ok < bk || (ok == bk && ow < bw)) { bk = ok; bw = ow; }
    }
    if (lane == 0) { s_key[w] = bk; s_who[w] = bw; }
    __syncthreads();
    bk = s_key[0]; bw = s_who[0];
    for (int i = 1; i < NV_TPB / 32; ++i)
        if (s_key[i] < bk || (s_key[i] == bk && s_who[i] < bw)) { bk = s_key[i]; bw = s_who[i]; }
    __syncthreads();
    // the winner among the others against the nearest node; every comparison that decided must be clear
    int32_t best = bw;
    double bestkey = bk;
    if (nn >= 0 && !(bk < key_nn)) { best = (int32_t)nn; bestkey = key_nn; }
    if (nn >= 0 && bw != 0x7fffffff && ext_ulps(bk, key_nn, key_nn) <= 64.0 && bk != key_nn) status = SBP_EXTEND_AMBIGUOUS;
    {   // any other candidate within 64 ulp of the best one (and not equal: equal keys are ordered by index)
        int amb = 0;
        for (int j = tid; j < mm; j += NV_TPB) {
            if (!cvis[j] || cidx[j] == best || cidx[j] == (int32_t)nn) continue;
            const double key = __dadd_rn(cost[cidx[j]], cc[j]);
            if (key != bestkey && ext_ulps(key, bestkey, bestkey) <= 64.0) amb = 1;
        }
        if (amb) atomicOr(&s_flag, 1);
    }
    __syncthreads();
    if (s_flag) status = SBP_EXTEND_AMBIGUOUS;
    if (best == 0x7fffffff) status = status == SBP_EXTEND_OK ? SBP_EXTEND_NO_PARENT : status;
    double c_best = 0.0, new_cost = 0.0;
    if (status == SBP_EXTEND_OK) {
        // newnode.cost = nn.cost + dist(nn.pos, newnode.pos) (rrtPlanner.py:192-194)
        const double a = __dsub_rn(soa[best], qx), b = __dsub_rn(soa[capacity + best], qy);
        c_best = __dsqrt_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)));
        new_cost = __dadd_rn(cost[best], c_best);
    }
    // ---- rewire decisions (rrtPlanner.py:250-266): every comparison must be clear before anything is written
    const int32_t new_idx = (int32_t)n;
    int n_rew = 0;
    if (do_rewire && status == SBP_EXTEND_OK) {
        const double bx = soa[best], by = soa[capacity + best];
        int amb = 0;
        for (int j = tid; j < mm; j += NV_TPB) {
            const int32_t p = cidx[j];
            const double px = soa[p], py = soa[capacity + p];
            // Node equality is by position (utils/common.py:187-192): nodes at the new node's or at its
            // parent's position are skipped
            const bool skip = (px == qx && py == qy) || (px == bx && py == by);
            int rew = 0;
            if (!skip) atomicAdd(&s_chk, 1);                     // a visible(n, new) call of the reference
            if (!skip && cvis[j]) {
                const double via = __dadd_rn(new_cost, cc[j]);
                if (via != cost[p] && ext_ulps(via, cost[p], cost[p]) <= 64.0) amb = 1;
                rew = via < cost[p];
            }
            cvis[j] = (uint8_t)((cvis[j] & 1) | (rew ? 2 : 0));
            if (rew && atomicAdd(&s_cnt, 1) >= SBP_EXTEND_MAX_REWIRED) amb = 1;     // would not fit the result
        }
        if (amb) atomicOr(&s_flag, 2);
        __syncthreads();
        if (s_flag & 2) status = SBP_EXTEND_AMBIGUOUS;
    }
    // ---- commit
    if (status == SBP_EXTEND_OK) {
        if (do_append && tid == 0) {
            soa[n] = qx; soa[capacity + n] = qy;
            soa32[n] = (float)qx; soa32[capacity + n] = (float)qy;
            const double mx = fmax(fabs(qx), fabs(qy));
            atomicMax((unsigned long long *)maxabs, (unsigned long long)__double_as_longlong(mx));
            cost[n] = new_cost;
            parent[n] = best;
        }
        if (do_rewire) {
            for (int j = tid; j < mm; j += NV_TPB) {
                if (!(cvis[j] & 2)) continue;
                const int32_t p = cidx[j];
                parent[p] = new_idx;
                cost[p] = __dadd_rn(new_cost, cc[j]);
                const int slot = atomicAdd(&s_flag, 4) >> 2;
                if (slot < SBP_EXTEND_MAX_REWIRED) { res->rewired[slot] = p; res->rewired_c[slot] = cc[j]; }
            }
            __syncthreads();
            n_rew = s_flag >> 2;
        }
    }
    if (tid == 0) {
        int nvis = 0;
        res->status = status;
        res->new_idx = do_append && status == SBP_EXTEND_OK ? new_idx : -1;
        res->parent = status == SBP_EXTEND_OK ? best : -1;
        res->cost = new_cost;
        res->c_parent = c_best;
        res->n_candidates = m;
        res->n_rewired = n_rew;
        res->n_rewire_checked = s_chk;
        (void)nvis;
        ctr[0] = 0; ctr[1] = 0; ctr[2] = 0;                  // ready for the next call
        __threadfence_system();
    }
}

extern "C" int32_t sbp_tree_create(sbp_nodes *nodes, sbp_tree **out) {
    SBP_REQUIRE(nodes && out, "sbp_tree_create: NULL argument");
    SBP_REQUIRE(nodes->d == 2, "sbp_tree_create: the one-launch extend step exists for the 2-D image checker");
    sbp_ctx *ctx = nodes->ctx;
    SBP_DEVICE(ctx);
    sbp_tree *t = new sbp_tree();
    t->nodes = nodes;
    t->capacity = nodes->capacity;
    SBP_CUDA(cudaMalloc(&t->d_cost, (size_t)t->capacity * sizeof(double)));
    SBP_CUDA(cudaMalloc(&t->d_parent, (size_t)t->capacity * sizeof(int32_t)));
    SBP_CUDA(cudaMalloc(&t->d_cidx, EXT_CAND * sizeof(int32_t)));
    SBP_CUDA(cudaMalloc(&t->d_cc, EXT_CAND * sizeof(double)));
    SBP_CUDA(cudaMalloc(&t->d_cvis, EXT_CAND));
    SBP_CUDA(cudaMalloc(&t->d_ctr, 4 * sizeof(int32_t)));
    SBP_CUDA(cudaMemsetAsync(t->d_ctr, 0, 4 * sizeof(int32_t), ctx->stream));
    SBP_CUDA(cudaMemsetAsync(t->d_cost, 0, (size_t)t->capacity * sizeof(double), ctx->stream));
    SBP_CUDA(cudaMemsetAsync(t->d_parent, 0xff, (size_t)t->capacity * sizeof(int32_t), ctx->stream));
    SBP_CUDA(cudaHostAlloc(&t->h_res, sizeof(sbp_extend_result), cudaHostAllocMapped));
    *out = t;
    return SBP_OK;
}

extern "C" int32_t sbp_tree_destroy(sbp_tree *t) {
    if (!t) return SBP_OK;
    SbpDeviceGuard guard(t->nodes->ctx->device);
    cudaStreamSynchronize(t->nodes->ctx->stream);
    cudaFree(t->d_cost); cudaFree(t->d_parent); cudaFree(t->d_cidx); cudaFree(t->d_cc); cudaFree(t->d_cvis);
    cudaFree(t->d_ctr);
    cudaFreeHost(t->h_res);
    delete t;
    return SBP_OK;
}

static int32_t tree_grow(sbp_tree *t, int64_t need) {
    sbp_ctx *ctx = t->nodes->ctx;
    if (need <= t->capacity) return SBP_OK;
    int64_t cap = t->capacity;
    while (cap < need) cap *= 2;
    double *nc = nullptr;
    int32_t *np = nullptr;
    SBP_CUDA(cudaMalloc(&nc, (size_t)cap * sizeof(double)));
    SBP_CUDA(cudaMalloc(&np, (size_t)cap * sizeof(int32_t)));
    SBP_CUDA(cudaMemsetAsync(nc, 0, (size_t)cap * sizeof(double), ctx->stream));
    SBP_CUDA(cudaMemsetAsync(np, 0xff, (size_t)cap * sizeof(int32_t), ctx->stream));
    SBP_CUDA(cudaMemcpyAsync(nc, t->d_cost, (size_t)t->capacity * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    SBP_CUDA(cudaMemcpyAsync(np, t->d_parent, (size_t)t->capacity * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    SBP_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(t->d_cost); cudaFree(t->d_parent);
    t->d_cost = nc; t->d_parent = np; t->capacity = cap;
    return SBP_OK;
}

__global__ void k_tree_set(double *cost, int32_t *parent, ExtFixes f, int32_t par0, int set_parent) {
    const int i = threadIdx.x;
    if (i < f.n) { cost[f.idx[i]] = f.cost[i]; }
    if (i == 0 && set_parent) parent[f.idx[0]] = par0;
}

/* cost / parent of node idx (stream ordered; parent < -1 leaves the parent untouched) */
extern "C" int32_t sbp_tree_set(sbp_tree *t, int64_t idx, double cost, int64_t parent) {
    SBP_REQUIRE(t, "sbp_tree_set: NULL argument");
    SBP_REQUIRE(idx >= 0 && idx < ((int64_t)1 << 31), "sbp_tree_set: bad index");
    sbp_ctx *ctx = t->nodes->ctx;
    SBP_DEVICE(ctx);
    int32_t rc = tree_grow(t, idx + 1 > t->nodes->capacity ? idx + 1 : t->nodes->capacity);
    if (rc) return rc;
    ExtFixes f{};
    f.n = 1; f.idx[0] = (int32_t)idx; f.cost[0] = cost;
    k_tree_set<<<1, 32, 0, ctx->stream>>>(t->d_cost, t->d_parent, f, (int32_t)parent, parent >= -1);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

extern "C" int32_t sbp_tree_read(sbp_tree *t, int64_t first, int64_t count, double *cost_host, int32_t *parent_host) {
    SBP_REQUIRE(t && first >= 0 && count >= 0 && first + count <= t->capacity, "sbp_tree_read: bad range");
    sbp_ctx *ctx = t->nodes->ctx;
    SBP_DEVICE(ctx);
    if (cost_host) SBP_CUDA(cudaMemcpyAsync(cost_host, t->d_cost + first, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (parent_host) SBP_CUDA(cudaMemcpyAsync(parent_host, t->d_parent + first, (size_t)count * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SBP_CUDA(cudaStreamSynchronize(ctx->stream));
    return SBP_OK;
}

extern "C" int32_t sbp_tree_write(sbp_tree *t, int64_t first, int64_t count, const double *cost_host,
                                  const int32_t *parent_host) {
    SBP_REQUIRE(t && first >= 0 && count >= 0 && cost_host && parent_host, "sbp_tree_write: bad argument");
    sbp_ctx *ctx = t->nodes->ctx;
    SBP_DEVICE(ctx);
    const int64_t need = first + count > t->nodes->capacity ? first + count : t->nodes->capacity;
    int32_t rc = tree_grow(t, need);
    if (rc) return rc;
    SBP_CUDA(cudaMemcpyAsync(t->d_cost + first, cost_host, (size_t)count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    SBP_CUDA(cudaMemcpyAsync(t->d_parent + first, parent_host, (size_t)count * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    SBP_CUDA(cudaStreamSynchronize(ctx->stream));           // (pageable sources may be reused by the caller)
    return SBP_OK;
}

extern "C" int32_t sbp_tree_extend_host(sbp_tree *t, sbp_map *map, const double *newpos, int64_t nn, double radius,
                                        int32_t do_append, int32_t do_rewire, const int64_t *fix_idx,
                                        const double *fix_cost, int32_t n_fixes, sbp_extend_result *out) {
    SBP_NVTX("sbp_tree_extend_host");
    SBP_REQUIRE(t && map && newpos && out, "sbp_tree_extend_host: NULL argument");
    SBP_REQUIRE(n_fixes >= 0 && n_fixes <= SBP_EXTEND_MAX_FIXES && (n_fixes == 0 || (fix_idx && fix_cost)),
                "sbp_tree_extend_host: at most SBP_EXTEND_MAX_FIXES fixes");
    sbp_nodes *s = t->nodes;
    SBP_REQUIRE(s->ctx == map->ctx, "sbp_tree_extend_host: store and map live in different contexts");
    SBP_REQUIRE(nn >= -1 && nn < s->n, "sbp_tree_extend_host: nn out of range");
    sbp_ctx *ctx = s->ctx;
    SBP_DEVICE(ctx);
    int32_t rc;
    if (do_append) {
        if ((rc = nodes_reserve(s, s->n + 1))) return rc;
    }
    if ((rc = tree_grow(t, s->capacity))) return rc;
    double thr;
    int cl;
    near_threshold(radius, 1, &thr, &cl);                     // engine.dist(...) <= radius
    // candidates whose d^2 lies within 4 ulp of the threshold are not decided on the device
    double lo = thr, hi = thr;
    for (int i = 0; i < 4; ++i) { lo = nextafter(lo, -INFINITY); hi = nextafter(hi, INFINITY); }
    ExtFixes f{};
    f.n = n_fixes;
    for (int i = 0; i < n_fixes; ++i) { f.idx[i] = (int32_t)fix_idx[i]; f.cost[i] = fix_cost[i]; }
    sbp_extend_result *dres = nullptr;
    SBP_CUDA(cudaHostGetDevicePointer((void **)&dres, t->h_res, 0));
    const int64_t n = s->n;
    int64_t want = (n + NV_TPB * 4 - 1) / (NV_TPB * 4);
    int grid = (int)(want < 1 ? 1 : want > 2 * (int64_t)ctx->sm_count ? 2 * (int64_t)ctx->sm_count : want);
    k_tree_extend<<<grid, NV_TPB, 0, ctx->stream>>>(s->d_soa, s->d_soa32, s->d_maxabs, s->capacity, n, map->view,
                                                    newpos[0], newpos[1], thr, cl, lo, hi, nn, do_append, do_rewire,
                                                    t->d_cost, t->d_parent, t->d_cidx, t->d_cc, t->d_cvis, t->d_ctr,
                                                    f, dres);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    SBP_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = *t->h_res;
    if (do_append && out->status == SBP_EXTEND_OK) s->n += 1;
    return SBP_OK;
}

// =====================================================================================
// find_nearest_neighbour_idx (rrtPlanner.py:268-279) for ONE host query in ONE launch: every CTA
// takes the arg-min of np.linalg.norm(poses - pos, axis=1) over its share of the nodes -- on the
// rounded roots, first index among equals, and np.argmin's rule that the first NaN wins -- and the
// CTA that finishes last folds the partial results and stores the answer in pinned host memory
// (the general sbp_nodes_nearest path costs two launches, two copies and their latencies: ~70 us
// per query on a 1M-node tree against ~25 us here).
struct Nearest1Result { int64_t idx; double dist; };

struct Nearest1State {
    double *d_pd = nullptr;        // per-CTA best distance
    int64_t *d_pi = nullptr;       // per-CTA best index (or first NaN index, see k_nearest1)
    int32_t *d_ticket = nullptr;
    Nearest1Result *h_res = nullptr;
    int blocks = 0;
};

template <int D>
__global__ void __launch_bounds__(NV_TPB)
k_nearest1(const double *__restrict__ soa, int64_t capacity, int64_t n, double q0, double q1, double q2, double q3,
           double q4, double q5, double *__restrict__ pd, int64_t *__restrict__ pi, int32_t *__restrict__ ticket,
           Nearest1Result *__restrict__ res) {
    __shared__ double s_d[NV_TPB / 32];
    __shared__ long long s_i[NV_TPB / 32], s_nan[NV_TPB / 32];
    __shared__ int s_last;
    const double qa[6] = {q0, q1, q2, q3, q4, q5};
    double qv[D];
#pragma unroll
    for (int j = 0; j < D; ++j) qv[j] = qa[j];
    const unsigned FULL = 0xffffffffu;
    const long long NONE = 0x7fffffffffffffffLL;
    double bd = INFINITY;
    long long bi = NONE, nan_i = NONE;
    auto better = [](double d, long long i, double bd_, long long bi_) { return d < bd_ || (d == bd_ && i < bi_); };
    for (int64_t t = (int64_t)blockIdx.x * NV_TPB + threadIdx.x; t < n; t += (int64_t)gridDim.x * NV_TPB) {
        double pv[D];
#pragma unroll
        for (int j = 0; j < D; ++j) pv[j] = soa[(int64_t)j * capacity + t];
        const double d2 = dist2_full<D>(pv, qv);
        if (d2 != d2) { nan_i = t < nan_i ? t : nan_i; continue; }
        const double dd = __dsqrt_rn(d2);
        if (better(dd, t, bd, bi)) { bd = dd; bi = t; }
    }
    auto fold = [&](double &d, long long &i, long long &ni) {
        for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_down_sync(FULL, d, o);
            const long long oi = __shfl_down_sync(FULL, i, o), on = __shfl_down_sync(FULL, ni, o);
            if (better(od, oi, d, i)) { d = od; i = oi; }
            ni = on < ni ? on : ni;
        }
    };
    fold(bd, bi, nan_i);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { s_d[w] = bd; s_i[w] = bi; s_nan[w] = nan_i; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < NV_TPB / 32; ++k) {
            if (better(s_d[k], s_i[k], bd, bi)) { bd = s_d[k]; bi = s_i[k]; }
            nan_i = s_nan[k] < nan_i ? s_nan[k] : nan_i;
        }
        // a NaN distance anywhere: np.argmin returns the FIRST NaN (encoded as dist = NaN)
        pd[blockIdx.x] = nan_i != NONE ? __longlong_as_double(0x7ff8000000000000LL) : bd;
        pi[blockIdx.x] = nan_i != NONE ? nan_i : bi;
        __threadfence();
        s_last = atomicAdd(ticket, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    bd = INFINITY; bi = NONE; nan_i = NONE;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += NV_TPB) {
        const double d = __ldcg(pd + b);
        const long long i = __ldcg(pi + b);
        if (d != d) nan_i = i < nan_i ? i : nan_i;
        else if (better(d, i, bd, bi)) { bd = d; bi = i; }
    }
    fold(bd, bi, nan_i);
    __syncthreads();
    if (lane == 0) { s_d[w] = bd; s_i[w] = bi; s_nan[w] = nan_i; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < NV_TPB / 32; ++k) {
            if (better(s_d[k], s_i[k], bd, bi)) { bd = s_d[k]; bi = s_i[k]; }
            nan_i = s_nan[k] < nan_i ? s_nan[k] : nan_i;
        }
        res->idx = nan_i != NONE ? nan_i : (bi == NONE ? -1 : bi);
        res->dist = nan_i != NONE ? __longlong_as_double(0x7ff8000000000000LL) : bd;
        *ticket = 0;
        __threadfence_system();
    }
}

static Nearest1State *nearest1_state(sbp_ctx *ctx) {
    // one per context (the queries of a context are stream ordered)
    static thread_local std::vector<std::pair<sbp_ctx *, Nearest1State *>> states;
    for (auto &p : states)
        if (p.first == ctx) return p.second;
    Nearest1State *st = new Nearest1State();
    st->blocks = 2 * ctx->sm_count;
    if (cudaMalloc(&st->d_pd, st->blocks * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&st->d_pi, st->blocks * sizeof(int64_t)) != cudaSuccess ||
        cudaMalloc(&st->d_ticket, sizeof(int32_t)) != cudaSuccess ||
        cudaMemset(st->d_ticket, 0, sizeof(int32_t)) != cudaSuccess ||
        cudaHostAlloc(&st->h_res, sizeof(Nearest1Result), cudaHostAllocMapped) != cudaSuccess) {
        delete st;
        return nullptr;
    }
    states.emplace_back(ctx, st);
    return st;
}

extern "C" int32_t sbp_nodes_nearest1_host(sbp_nodes *s, const double *x, int64_t *idx, double *dist) {
    SBP_NVTX("sbp_nodes_nearest1_host");
    SBP_REQUIRE(s && x && idx, "sbp_nodes_nearest1_host: NULL argument");
    sbp_ctx *ctx = s->ctx;
    SBP_DEVICE(ctx);
    if (s->n == 0) { *idx = -1; if (dist) *dist = INFINITY; return SBP_OK; }
    Nearest1State *st = nearest1_state(ctx);
    SBP_REQUIRE(st != nullptr, "sbp_nodes_nearest1_host: allocation failed");
    Nearest1Result *dres = nullptr;
    SBP_CUDA(cudaHostGetDevicePointer((void **)&dres, st->h_res, 0));
    double q[6] = {0, 0, 0, 0, 0, 0};
    for (int j = 0; j < s->d; ++j) q[j] = x[j];
    int64_t want = (s->n + NV_TPB * 4 - 1) / (NV_TPB * 4);
    const int grid = (int)(want < 1 ? 1 : want > st->blocks ? st->blocks : want);
#define N1_GO(DD)                                                                                                  \
    k_nearest1<DD><<<grid, NV_TPB, 0, ctx->stream>>>(s->d_soa, s->capacity, s->n, q[0], q[1], q[2], q[3], q[4], q[5], \
                                                     st->d_pd, st->d_pi, st->d_ticket, dres)
    switch (s->d) {
        case 2: N1_GO(2); break;
        case 3: N1_GO(3); break;
        case 4: N1_GO(4); break;
        default: N1_GO(6); break;
    }
#undef N1_GO
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    SBP_CUDA(cudaStreamSynchronize(ctx->stream));
    *idx = st->h_res->idx;
    if (dist) *dist = st->h_res->dist;
    return SBP_OK;
}
