// nodes.cu -- device-resident append-only node store and the batched planner primitives
// that gather candidate edges from it.  The queries live in knn.cu (K4) and near.cu (K5).
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
    // the query kernels are instantiated for these dimensions (2-D image space, the 4-D arm, and
    // 3 / 6 for free-flying bodies); a store of any other dimension could only fail at query time
    SBP_REQUIRE(d == 2 || d == 3 || d == 4 || d == 6, "sbp_nodes_create: dimension must be 2, 3, 4 or 6");
    SBP_REQUIRE(capacity >= 0 && capacity < ((int64_t)1 << 31), "sbp_nodes_create: bad capacity");
    SBP_DEVICE(ctx);
    sbp_nodes *s = new sbp_nodes();
    s->ctx = ctx;
    s->d = d;
    // a multiple of 1024 nodes: every per-dimension array stays 16-byte aligned and a tile
    // tail rounded up to 4 floats stays inside the allocation (bulk copies of k_knn_filter)
    s->capacity = capacity < 1024 ? 1024 : (capacity + 1023) / 1024 * 1024;
    cudaError_t e = cudaMalloc(&s->d_soa, (size_t)s->capacity * d * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&s->d_soa32, (size_t)s->capacity * d * sizeof(float));
    // unused fp32 slots read as 3.4e38 (never inside a threshold)
    if (e == cudaSuccess)
        e = cudaMemsetAsync(s->d_soa32, 0x7f, (size_t)s->capacity * d * sizeof(float), ctx->stream);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_maxabs, sizeof(double));
    if (e == cudaSuccess) e = cudaMemsetAsync(s->d_maxabs, 0, sizeof(double), ctx->stream);
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
    SbpDeviceGuard guard(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    if (s->d_soa) cudaFree(s->d_soa);
    if (s->d_soa32) cudaFree(s->d_soa32);
    if (s->d_maxabs) cudaFree(s->d_maxabs);
    delete s;
    return SBP_OK;
}

__global__ void k_aos_to_soa(const double *__restrict__ X, int64_t m, int d, double *__restrict__ soa,
                             float *__restrict__ soa32, double *__restrict__ maxabs,
                             int64_t capacity, int64_t offset) {
    int64_t total = m * d;
    double mx = 0.0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        int64_t row = t / d;
        int j = (int)(t - row * d);
        double v = X[t];
        soa[(int64_t)j * capacity + offset + row] = v;
        soa32[(int64_t)j * capacity + offset + row] = (float)v;
        double av = (v != v) ? INFINITY : fabs(v);
        mx = av > mx ? av : mx;
    }
    for (int o = 16; o > 0; o >>= 1) {
        double other = __shfl_down_sync(0xffffffffu, mx, o);
        mx = other > mx ? other : mx;
    }
    // non-negative doubles order like their bit patterns
    if ((threadIdx.x & 31) == 0 && mx > 0.0)
        atomicMax((unsigned long long *)maxabs, (unsigned long long)__double_as_longlong(mx));
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
    float *nw32 = nullptr;
    SBP_CUDA(cudaMalloc(&nw, (size_t)cap * s->d * sizeof(double)));
    SBP_CUDA(cudaMalloc(&nw32, (size_t)cap * s->d * sizeof(float)));
    SBP_CUDA(cudaMemsetAsync(nw32, 0x7f, (size_t)cap * s->d * sizeof(float), s->ctx->stream));
    for (int j = 0; j < s->d; ++j) {
        SBP_CUDA(cudaMemcpyAsync(nw + (int64_t)j * cap, s->d_soa + (int64_t)j * s->capacity,
                                 (size_t)s->n * sizeof(double), cudaMemcpyDeviceToDevice,
                                 s->ctx->stream));
        SBP_CUDA(cudaMemcpyAsync(nw32 + (int64_t)j * cap, s->d_soa32 + (int64_t)j * s->capacity,
                                 (size_t)s->n * sizeof(float), cudaMemcpyDeviceToDevice,
                                 s->ctx->stream));
    }
    SBP_CUDA(cudaStreamSynchronize(s->ctx->stream));
    SBP_CUDA(cudaFree(s->d_soa));
    SBP_CUDA(cudaFree(s->d_soa32));
    s->d_soa = nw;
    s->d_soa32 = nw32;
    s->capacity = cap;
    return SBP_OK;
}

int32_t nodes_reserve(sbp_nodes *s, int64_t need) { return nodes_grow(s, need); }

extern "C" int32_t sbp_nodes_append(sbp_nodes *s, const double *X, int64_t m, int32_t on_device) {
    SBP_NVTX("sbp_nodes_append");
    SBP_REQUIRE(s && (m == 0 || X), "sbp_nodes_append: NULL argument");
    SBP_REQUIRE(m >= 0, "sbp_nodes_append: m < 0");
    if (m == 0) return SBP_OK;
    sbp_ctx *ctx = s->ctx;
    SBP_DEVICE(ctx);
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
    k_aos_to_soa<<<blocks, 256, 0, ctx->stream>>>(src, m, s->d, s->d_soa, s->d_soa32, s->d_maxabs,
                                                  s->capacity, s->n);
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
    SBP_DEVICE(ctx);
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
    SBP_DEVICE(ctx);
    int64_t total = count * s->d;
    int blocks = (int)((total + 255) / 256);
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    k_soa_to_aos<<<blocks, 256, 0, ctx->stream>>>(s->d_soa, s->capacity, first, count, s->d,
                                                  out_device);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
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
    SBP_NVTX("sbp_nodes_gather_edges");
    SBP_REQUIRE(s && (m == 0 || (nbr && P && Q)), "sbp_nodes_gather_edges: NULL argument");
    SBP_REQUIRE(m >= 0 && k >= 1 && first_row >= 0, "sbp_nodes_gather_edges: bad sizes");
    if (m == 0) return SBP_OK;
    sbp_ctx *ctx = s->ctx;
    SBP_DEVICE(ctx);
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
    SBP_NVTX("sbp_nodes_knn_edges");
    SBP_REQUIRE(s && (m == 0 || (nbr_in && nbr_out && P && Q && valid)),
                "sbp_nodes_knn_edges: NULL argument");
    int k_out = X ? k_in : k_in - 1;
    SBP_REQUIRE(m >= 0 && k_out >= 1, "sbp_nodes_knn_edges: bad sizes");
    SBP_REQUIRE(X || (first_row >= 0 && first_row + m <= s->n), "sbp_nodes_knn_edges: row range");
    if (m == 0) return SBP_OK;
    sbp_ctx *ctx = s->ctx;
    SBP_DEVICE(ctx);
    int64_t total = m * k_out;
    int blocks = (int)((total + 255) / 256);
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    k_knn_edges<<<blocks, 256, 0, ctx->stream>>>(s->d_soa, s->capacity, s->n, s->d, nbr_in, m, k_in,
                                                 k_out, first_row, X, nbr_out, P, Q, valid);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}
