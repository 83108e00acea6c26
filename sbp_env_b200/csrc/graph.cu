// graph.cu -- K7: the PRM roadmap as a device-resident CSR graph and batched hop-count
// shortest paths on it.
//
// Reference behaviour restated (file:line under /root/reference/sbp_env/planners/):
//   PRMPlanner.build_graph    prmPlanner.py:91-101   directed edges m_g -> v into a networkx
//                                                    DiGraph for every visible candidate pair
//   PRMPlanner.get_solution   prmPlanner.py:128-154  nx.has_path / nx.shortest_path(graph,
//                                                    start, goal): UNWEIGHTED, fewest hops
//                                                    (no weight= argument, :143)
//
// Layout: row_ptr int64 [n + 1], col int32 [E]; row u lists the heads v of the edges u -> v
// (order inside a row is arbitrary: the search below does not depend on it).
//
// Search: one CTA per (start, goal) query, level-synchronous BFS over a CTA-private work-space
// (level and parent arrays, FIFO of touched nodes; only touched entries are reset between
// queries).  Among the fewest-hop paths networkx returns whichever its bidirectional search
// meets first (adjacency-order dependent); here the path is made deterministic instead: the
// parent of v is the LOWEST-INDEX node of the previous level with an edge to v (atomicMin), so
// the result does not depend on thread timing or on the order of `col`.  Hop counts and
// reachability are exactly networkx's; the path is one of the shortest ones.
#include "common.cuh"

struct sbp_graph {
    sbp_ctx *ctx = nullptr;
    int64_t n = 0, e = 0, e_capacity = 0;
    int64_t *d_row_ptr = nullptr;    // [n + 1]
    int32_t *d_col = nullptr;        // [e_capacity]
    int32_t *d_deg = nullptr;        // [n + 1] degree counts, then scatter cursors
    int64_t *d_total = nullptr;      // device scalar: number of arcs of the last build
    int64_t *d_tsums = nullptr;      // per-tile degree sums of the row_ptr scan
    int64_t tiles = 0;
    // weakly connected components of the last build (union-find over the arcs), computed on
    // the first query after a build: comp[u] != comp[v]  =>  no path either way
    int32_t *d_comp = nullptr;       // [n]
    bool comp_valid = false;
    // BFS work-space, grown on demand: per CTA slot level [n], parent [n], queue [n]
    int32_t *d_work = nullptr;
    int64_t work_slots = 0;
    bool symmetric = false;          // last build added both directions of every arc (undirected)
};

#define GSCAN_TILE 4096

extern "C" int32_t sbp_graph_create(sbp_ctx *ctx, int64_t n_nodes, sbp_graph **out) {
    SBP_REQUIRE(ctx && out, "sbp_graph_create: NULL argument");
    SBP_REQUIRE(n_nodes >= 0 && n_nodes < ((int64_t)1 << 31), "sbp_graph_create: bad node count");
    SBP_DEVICE(ctx);
    sbp_graph *g = new sbp_graph();
    g->ctx = ctx;
    g->n = n_nodes;
    g->tiles = (n_nodes + GSCAN_TILE - 1) / GSCAN_TILE;
    if (g->tiles < 1) g->tiles = 1;
    cudaError_t e = cudaMalloc(&g->d_row_ptr, (size_t)(n_nodes + 1) * sizeof(int64_t));
    if (e == cudaSuccess) e = cudaMalloc(&g->d_deg, (size_t)(n_nodes + 1) * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&g->d_total, sizeof(int64_t));
    if (e == cudaSuccess) e = cudaMalloc(&g->d_tsums, (size_t)g->tiles * sizeof(int64_t));
    if (e == cudaSuccess) e = cudaMalloc(&g->d_comp, (size_t)(n_nodes + 1) * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemsetAsync(g->d_row_ptr, 0, (size_t)(n_nodes + 1) * sizeof(int64_t), ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(g->d_total, 0, sizeof(int64_t), ctx->stream);
    if (e != cudaSuccess) {
        sbp_set_error("sbp_graph_create: %s", cudaGetErrorString(e));
        delete g;
        return SBP_ERR_CUDA;
    }
    *out = g;
    return SBP_OK;
}

extern "C" int32_t sbp_graph_destroy(sbp_graph *g) {
    if (!g) return SBP_OK;
    SbpDeviceGuard guard(g->ctx->device);
    cudaStreamSynchronize(g->ctx->stream);
    if (g->d_row_ptr) cudaFree(g->d_row_ptr);
    if (g->d_col) cudaFree(g->d_col);
    if (g->d_deg) cudaFree(g->d_deg);
    if (g->d_total) cudaFree(g->d_total);
    if (g->d_tsums) cudaFree(g->d_tsums);
    if (g->d_comp) cudaFree(g->d_comp);
    if (g->d_work) cudaFree(g->d_work);
    delete g;
    return SBP_OK;
}

// ----------------------------------------------------------------- build ----
// Candidate t of a kNN block: row v = first_row + t / k (the node whose neighbours were
// searched), neighbour m_g = nbr[t]; the reference adds the arc m_g -> v (prmPlanner.py:99-101).
// The block comes as (nbr int64, vis uint8) or as the packed adjacency code
// (nbr + 1) * 2 + vis (int32, what the visibility kernel / the multi-GPU exchange produce).
// Candidate t of an explicit list: src[t] -> dst[t].  `keep` (optional) masks candidates out.
//
// undirected: every kept arc u -> v also yields v -> u -- the graph the reference's radius rule
// builds, where v in near(u) <=> u in near(v) puts both (m_g -> v) and (v -> m_g) into the DiGraph
// (prmPlanner.py:91-101); a k-NN candidate set is not symmetric by itself.  For kNN / packed
// blocks the mirrored arc is skipped when row u of the same block already lists v as a kept
// candidate (the arc v -> u then exists in its own right), so that no arc is stored twice.
struct EdgeSource {
    const int64_t *nbr;      // kNN block, or src list
    const int32_t *packed;   // packed kNN block (then nbr == NULL)
    const int64_t *dst;      // NULL for a kNN block
    const uint8_t *keep;     // optional
    int64_t total;
    int32_t k;
    int64_t first_row, n, m;
    int32_t undirected;
    // rows in an explicit order (sbp_prm_connect2d_sorted): block row r belongs to node row_ids[r],
    // node u's row is row_pos[u] (both NULL: row r = node first_row + r)
    const int32_t *row_ids = nullptr;
    const int32_t *row_pos = nullptr;
    __device__ __forceinline__ bool cand(int64_t t, int64_t &u) const {
        if (packed) {
            const int32_t c = packed[t];
            u = (int64_t)(c >> 1) - 1;
            return (c & 1) != 0;
        }
        u = nbr[t];
        return !keep || keep[t] != 0;
    }
    __device__ __forceinline__ bool edge(int64_t t, int64_t &u, int64_t &v) const {
        if (!cand(t, u)) return false;
        v = dst ? dst[t] : (row_ids ? (int64_t)row_ids[t / k] : first_row + t / k);
        return u >= 0 && u < n && v >= 0 && v < n;
    }
    // the mirrored arc v -> u of kept candidate (row v, neighbour u) duplicates an arc the block
    // already holds iff row u lists v as a kept candidate
    __device__ __forceinline__ bool mirror_is_duplicate(int64_t u, int64_t v) const {
        if (dst) return false;
        const int64_t r = row_pos ? (int64_t)row_pos[u] : u - first_row;
        if (r < 0 || r >= m) return false;
        for (int c = 0; c < k; ++c) {
            int64_t w;
            if (cand(r * k + c, w) && w == v) return true;
        }
        return false;
    }
};

__global__ void k_graph_degree(EdgeSource es, int32_t *__restrict__ deg) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < es.total;
         t += (int64_t)gridDim.x * blockDim.x) {
        int64_t u, v;
        if (!es.edge(t, u, v)) continue;
        atomicAdd(deg + u, 1);
        if (es.undirected && u != v && !es.mirror_is_duplicate(u, v)) atomicAdd(deg + v, 1);
    }
}

// Exclusive scan deg[n] -> row_ptr[n + 1] in three launches (tile sums, scan of the tile sums
// by one CTA, per-tile scan + offset); deg is turned into the scatter cursor (zeros again).
__global__ void __launch_bounds__(1024)
k_graph_tile_sums(const int32_t *__restrict__ deg, int64_t n, int64_t *__restrict__ tsums) {
    __shared__ int32_t ws[32];
    const int64_t i0 = (int64_t)blockIdx.x * GSCAN_TILE + (int64_t)threadIdx.x * 4;
    int32_t v = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) v += i0 + j < n ? deg[i0 + j] : 0;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = ws[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) tsums[blockIdx.x] = v;
    }
}

__global__ void __launch_bounds__(1024)
k_graph_scan_tiles(int64_t *__restrict__ tsums, int64_t nt, int64_t *__restrict__ total,
                   int64_t *__restrict__ row_ptr_n) {
    __shared__ int64_t ws[32];
    __shared__ int64_t s_carry;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < nt; base += 1024) {
        const int64_t i = base + tid;
        const int64_t x = i < nt ? tsums[i] : 0;
        int64_t inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int64_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
        if (lane == 31) ws[w] = inc;
        __syncthreads();
        if (w == 0) {
            int64_t y = ws[lane], yi = y;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int64_t u = __shfl_up_sync(0xffffffffu, yi, o); if (lane >= o) yi += u; }
            ws[lane] = yi - y;
        }
        __syncthreads();
        const int64_t excl = s_carry + ws[w] + inc - x;
        if (i < nt) tsums[i] = excl;
        __syncthreads();
        if (tid == 1023) s_carry = excl + x;
        __syncthreads();
    }
    if (tid == 0) { *total = s_carry; *row_ptr_n = s_carry; }
}

__global__ void __launch_bounds__(1024)
k_graph_scan_apply(int32_t *__restrict__ deg, int64_t n, const int64_t *__restrict__ tsums,
                   int64_t *__restrict__ row_ptr) {
    __shared__ int32_t ws[32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t i0 = (int64_t)blockIdx.x * GSCAN_TILE + (int64_t)tid * 4;
    int32_t v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = i0 + j < n ? deg[i0 + j] : 0;
    const int32_t tsum = v[0] + v[1] + v[2] + v[3];
    int32_t inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
    if (lane == 31) ws[w] = inc;
    __syncthreads();
    if (w == 0) {
        int32_t y = ws[lane], yi = y;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int32_t u = __shfl_up_sync(0xffffffffu, yi, o); if (lane >= o) yi += u; }
        ws[lane] = yi - y;
    }
    __syncthreads();
    int64_t run = tsums[blockIdx.x] + (int64_t)(ws[w] + inc - tsum);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (i0 + j < n) { row_ptr[i0 + j] = run; deg[i0 + j] = 0; }
        run += v[j];
    }
}

__global__ void k_graph_scatter(EdgeSource es, const int64_t *__restrict__ row_ptr,
                                int32_t *__restrict__ cursor, int32_t *__restrict__ col) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < es.total;
         t += (int64_t)gridDim.x * blockDim.x) {
        int64_t u, v;
        if (!es.edge(t, u, v)) continue;
        col[row_ptr[u] + atomicAdd(cursor + u, 1)] = (int32_t)v;
        if (es.undirected && u != v && !es.mirror_is_duplicate(u, v))
            col[row_ptr[v] + atomicAdd(cursor + v, 1)] = (int32_t)u;
    }
}

static int32_t graph_build(sbp_graph *g, EdgeSource es) {
    sbp_ctx *ctx = g->ctx;
    SBP_DEVICE(ctx);
    es.n = g->n;
    g->comp_valid = false;
    g->symmetric = es.undirected != 0;
    // every candidate may become an arc (two when undirected): size `col` for all of them (no
    // host round trip)
    const int64_t need = es.total * (es.undirected ? 2 : 1);
    if (need > g->e_capacity) {
        SBP_CUDA(cudaStreamSynchronize(ctx->stream));
        if (g->d_col) SBP_CUDA(cudaFree(g->d_col));
        g->d_col = nullptr;
        g->e_capacity = 0;
        SBP_CUDA(cudaMalloc(&g->d_col, (size_t)(need > 0 ? need : 1) * sizeof(int32_t)));
        g->e_capacity = need;
    }
    SBP_CUDA(cudaMemsetAsync(g->d_deg, 0, (size_t)(g->n + 1) * sizeof(int32_t), ctx->stream));
    int blocks = (int)((es.total + 255) / 256);
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    if (blocks < 1) blocks = 1;
    k_graph_degree<<<blocks, 256, 0, ctx->stream>>>(es, g->d_deg);
    k_graph_tile_sums<<<(unsigned)g->tiles, 1024, 0, ctx->stream>>>(g->d_deg, g->n, g->d_tsums);
    k_graph_scan_tiles<<<1, 1024, 0, ctx->stream>>>(g->d_tsums, g->tiles, g->d_total, g->d_row_ptr + g->n);
    k_graph_scan_apply<<<(unsigned)g->tiles, 1024, 0, ctx->stream>>>(g->d_deg, g->n, g->d_tsums, g->d_row_ptr);
    k_graph_scatter<<<blocks, 256, 0, ctx->stream>>>(es, g->d_row_ptr, g->d_deg, g->d_col);
    ctx->launches += 5;
    SBP_CUDA(cudaGetLastError());
    g->e = -1;                       // known on the device; fetched lazily by sbp_graph_info
    return SBP_OK;
}

extern "C" int32_t sbp_graph_build_knn(sbp_graph *g, const int64_t *nbr, const uint8_t *vis, int64_t m,
                                       int32_t k, int64_t first_row, int32_t undirected) {
    SBP_NVTX("sbp_graph_build_knn");
    SBP_REQUIRE(g && (m == 0 || nbr), "sbp_graph_build_knn: NULL argument");
    SBP_REQUIRE(m >= 0 && k >= 1 && first_row >= 0, "sbp_graph_build_knn: bad sizes");
    EdgeSource es{nbr, nullptr, nullptr, vis, m * k, k, first_row, g->n, m, undirected ? 1 : 0};
    return graph_build(g, es);
}

extern "C" int32_t sbp_graph_build_packed(sbp_graph *g, const int32_t *packed, int64_t m, int32_t k,
                                          int64_t first_row, int32_t undirected) {
    SBP_NVTX("sbp_graph_build_packed");
    SBP_REQUIRE(g && (m == 0 || packed), "sbp_graph_build_packed: NULL argument");
    SBP_REQUIRE(m >= 0 && k >= 1 && first_row >= 0, "sbp_graph_build_packed: bad sizes");
    EdgeSource es{nullptr, packed, nullptr, nullptr, m * k, k, first_row, g->n, m, undirected ? 1 : 0};
    return graph_build(g, es);
}

__global__ void k_graph_row_pos(const int32_t *__restrict__ row_ids, int64_t m, int64_t n, int32_t *__restrict__ row_pos) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += (int64_t)gridDim.x * blockDim.x) {
        const int32_t v = row_ids[r];
        if (v >= 0 && (int64_t)v < n) row_pos[v] = (int32_t)r;
    }
}

extern "C" int32_t sbp_graph_build_packed_rows(sbp_graph *g, const int32_t *packed, const int32_t *row_ids,
                                               int64_t m, int32_t k, int32_t undirected) {
    SBP_NVTX("sbp_graph_build_packed_rows");
    SBP_REQUIRE(g && (m == 0 || (packed && row_ids)), "sbp_graph_build_packed_rows: NULL argument");
    SBP_REQUIRE(m >= 0 && k >= 1, "sbp_graph_build_packed_rows: bad sizes");
    sbp_ctx *ctx = g->ctx;
    SBP_DEVICE(ctx);
    EdgeSource es{nullptr, packed, nullptr, nullptr, m * k, k, 0, g->n, m, undirected ? 1 : 0};
    es.row_ids = row_ids;
    if (undirected) {
        // the inverse of row_ids (nodes without a row: -1), for the duplicate test of mirrored arcs
        void *pos = nullptr;
        int32_t rc = sbp_ctx_scratch(ctx, 15, (size_t)(g->n > 0 ? g->n : 1) * sizeof(int32_t), &pos);
        if (rc) return rc;
        SBP_CUDA(cudaMemsetAsync(pos, 0xff, (size_t)g->n * sizeof(int32_t), ctx->stream));
        int blocks = (int)((m + 255) / 256);
        if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
        if (blocks < 1) blocks = 1;
        k_graph_row_pos<<<blocks, 256, 0, ctx->stream>>>(row_ids, m, g->n, (int32_t *)pos);
        ctx->launches++;
        es.row_pos = (const int32_t *)pos;
    }
    return graph_build(g, es);
}

extern "C" int32_t sbp_graph_build_edges(sbp_graph *g, const int64_t *src, const int64_t *dst,
                                         const uint8_t *keep, int64_t e, int32_t undirected) {
    SBP_NVTX("sbp_graph_build_edges");
    SBP_REQUIRE(g && (e == 0 || (src && dst)), "sbp_graph_build_edges: NULL argument");
    SBP_REQUIRE(e >= 0, "sbp_graph_build_edges: e < 0");
    EdgeSource es{src, nullptr, dst, keep, e, 1, 0, g->n, 0, undirected ? 1 : 0};
    return graph_build(g, es);
}

extern "C" int32_t sbp_graph_info(sbp_graph *g, int64_t *n_nodes, int64_t *n_edges) {
    SBP_REQUIRE(g, "sbp_graph_info: NULL argument");
    if (g->e < 0) {
        SBP_DEVICE(g->ctx);
        SBP_CUDA(cudaMemcpyAsync(&g->e, g->d_total, sizeof(int64_t), cudaMemcpyDeviceToHost,
                                 g->ctx->stream));
        SBP_CUDA(cudaStreamSynchronize(g->ctx->stream));
    }
    if (n_nodes) *n_nodes = g->n;
    if (n_edges) *n_edges = g->e;
    return SBP_OK;
}

// Copy the CSR arrays out (device buffers: row_ptr [n + 1], col [>= n_edges]).
extern "C" int32_t sbp_graph_csr(sbp_graph *g, int64_t *row_ptr, int32_t *col, int64_t col_capacity) {
    SBP_REQUIRE(g && row_ptr, "sbp_graph_csr: NULL argument");
    int64_t e = 0;
    int32_t rc = sbp_graph_info(g, nullptr, &e);
    if (rc) return rc;
    SBP_REQUIRE(!col || col_capacity >= e, "sbp_graph_csr: col buffer too small");
    SBP_CUDA(cudaMemcpyAsync(row_ptr, g->d_row_ptr, (size_t)(g->n + 1) * sizeof(int64_t),
                             cudaMemcpyDeviceToDevice, g->ctx->stream));
    if (col && e > 0)
        SBP_CUDA(cudaMemcpyAsync(col, g->d_col, (size_t)e * sizeof(int32_t), cudaMemcpyDeviceToDevice,
                                 g->ctx->stream));
    return SBP_OK;
}

// ------------------------------------------------- connected components ----
// Weakly connected components by lock-free union-find over the arcs (hook the larger root
// under the smaller with atomicCAS, path halving on the way), then one flattening pass:
// comp[v] = smallest node index of v's component.  nx.has_path(u, v) needs comp[u] == comp[v]
// (sufficient too when the graph was built undirected), so the search kernel answers every
// query with different labels at once: on a fragmented roadmap (cfg5: ~2 % of the random
// start / goal pairs are connected) almost all of them.
__device__ __forceinline__ int32_t uf_find(int32_t *__restrict__ parent, int32_t v) {
    int32_t p = __ldcg(parent + v);
    while (p != v) {
        const int32_t gp = __ldcg(parent + p);
        if (gp != p) parent[v] = gp;                          // path halving (benign race)
        v = p;
        p = gp;
    }
    return v;
}

__global__ void k_graph_cc_init(int32_t *__restrict__ parent, int64_t n) {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n;
         v += (int64_t)gridDim.x * blockDim.x)
        parent[v] = (int32_t)v;
}

__global__ void k_graph_cc_hook(const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                                int64_t n, int32_t *__restrict__ parent) {
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < n;
         u += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e0 = row_ptr[u], e1 = row_ptr[u + 1];
        for (int64_t e = e0; e < e1; ++e) {
            int32_t a = uf_find(parent, (int32_t)u), b = uf_find(parent, col[e]);
            while (a != b) {
                if (a < b) { const int32_t t = a; a = b; b = t; }      // a > b: hook a under b
                const int32_t old = atomicCAS(parent + a, a, b);
                if (old == a) break;
                a = uf_find(parent, old);
                b = uf_find(parent, b);
            }
        }
    }
}

__global__ void k_graph_cc_flatten(int32_t *__restrict__ parent, int64_t n) {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n;
         v += (int64_t)gridDim.x * blockDim.x) {
        int32_t r = (int32_t)v;
        while (true) {
            const int32_t p = __ldcg(parent + r);
            if (p == r) break;
            r = p;
        }
        parent[v] = r;                   // roots keep themselves; only non-roots are rewritten
    }
}

static int32_t graph_ensure_components(sbp_graph *g) {
    if (g->comp_valid) return SBP_OK;
    sbp_ctx *ctx = g->ctx;
    int blocks = (int)((g->n + 255) / 256);
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    if (blocks < 1) blocks = 1;
    k_graph_cc_init<<<blocks, 256, 0, ctx->stream>>>(g->d_comp, g->n);
    k_graph_cc_hook<<<blocks, 256, 0, ctx->stream>>>(g->d_row_ptr, g->d_col, g->n, g->d_comp);
    k_graph_cc_flatten<<<blocks, 256, 0, ctx->stream>>>(g->d_comp, g->n);
    ctx->launches += 3;
    SBP_CUDA(cudaGetLastError());
    g->comp_valid = true;
    return SBP_OK;
}

// labels: device int32 [n] (nullable): smallest node index of each node's weakly connected
// component.  n_components (nullable, HOST): number of components (synchronises).
__global__ void k_graph_cc_count(const int32_t *__restrict__ comp, int64_t n, unsigned long long *cnt) {
    unsigned long long c = 0;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n;
         v += (int64_t)gridDim.x * blockDim.x)
        c += comp[v] == (int32_t)v ? 1ull : 0ull;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(cnt, c);
}

extern "C" int32_t sbp_graph_components(sbp_graph *g, int32_t *labels, int64_t *n_components) {
    SBP_NVTX("sbp_graph_components");
    SBP_REQUIRE(g, "sbp_graph_components: NULL argument");
    sbp_ctx *ctx = g->ctx;
    SBP_DEVICE(ctx);
    if (g->n == 0) { if (n_components) *n_components = 0; return SBP_OK; }
    int32_t rc = graph_ensure_components(g);
    if (rc) return rc;
    if (labels)
        SBP_CUDA(cudaMemcpyAsync(labels, g->d_comp, (size_t)g->n * sizeof(int32_t), cudaMemcpyDeviceToDevice,
                                 ctx->stream));
    if (n_components) {
        unsigned long long *cnt = (unsigned long long *)(ctx->d_flags + 12);     // 8-byte aligned slot
        SBP_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), ctx->stream));
        int blocks = (int)((g->n + 255) / 256);
        if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
        k_graph_cc_count<<<blocks, 256, 0, ctx->stream>>>(g->d_comp, g->n, cnt);
        ctx->launches++;
        SBP_CUDA(cudaMemcpyAsync(ctx->h_flags + 12, cnt, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                 ctx->stream));
        SBP_CUDA(cudaStreamSynchronize(ctx->stream));
        *n_components = (int64_t) * (unsigned long long *)(ctx->h_flags + 12);
    }
    return SBP_OK;
}

// ------------------------------------------------------------------- BFS ----
#define BFS_TPB 1024
#define BFS_UNSET 0x7fffffff

// Queries are taken from a device-side ticket counter (the ones the component labels settle
// cost nothing, so a static split would leave most CTAs idle); work-space slot = blockIdx.x.
//   level[v]  = -1 untouched, else BFS level of v (start = 0)
//   parent[v] = lowest-index predecessor on the previous level (BFS_UNSET until touched)
//   queue     = touched nodes in level order
__global__ void __launch_bounds__(BFS_TPB)
k_graph_bfs(const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ col, int64_t n,
            const int32_t *__restrict__ comp, const int64_t *__restrict__ start,
            const int64_t *__restrict__ goal, int64_t nq, int32_t *__restrict__ work,
            int32_t *__restrict__ hops, int64_t *__restrict__ paths, int32_t max_len,
            unsigned long long *__restrict__ ticket) {
    int32_t *level = work + (int64_t)blockIdx.x * 3 * n, *parent = level + n, *queue = parent + n;
    __shared__ int s_tail, s_found;
    __shared__ long long s_q;
    while (true) {
        // next query that needs a search (a whole warp scans 32 tickets at a time would be
        // overkill: the test is two loads per query)
        if (threadIdx.x == 0) {
            long long q;
            while (true) {
                q = (long long)atomicAdd(ticket, 1ull);
                if (q >= nq) break;
                const int64_t s = start[q], t = goal[q];
                const bool ok = s >= 0 && s < n && t >= 0 && t < n;
                if (ok && s != t && comp[s] == comp[t]) break;          // needs the search
                hops[q] = ok && s == t ? 0 : -1;
                if (paths) {
                    int64_t *p = paths + q * (int64_t)max_len;
                    for (int i = 0; i < max_len; ++i) p[i] = -1;
                    if (ok && s == t) p[0] = s;
                }
            }
            s_q = q;
        }
        __syncthreads();
        const int64_t q = s_q;
        if (q >= nq) return;
        const int64_t s = start[q], t = goal[q];
        int head = 0, tail = 0, lvl = 0;
        if (threadIdx.x == 0) {
            s_found = 0;
            level[s] = 0;
            queue[0] = (int32_t)s;
            s_tail = 1;
        }
        __syncthreads();
        tail = s_tail;
        bool found = false;
        __syncthreads();
        // expand level by level; the level in which the goal is touched is completed, so that
        // its parent is the minimum over the whole previous level.  s_tail / s_found are read
        // into registers between two barriers: every thread takes the same decision.
        while (head < tail && !found) {
            // a warp per frontier node, its lanes over the node's arcs: a level is a handful of
            // nodes with ~20 arcs each, and a thread walking them one by one would chain ~20
            // dependent global atomics
            for (int i = head + (threadIdx.x >> 5); i < tail; i += BFS_TPB / 32) {
                const int32_t u = queue[i];
                const int64_t e0 = row_ptr[u], e1 = row_ptr[u + 1];
                for (int64_t e = e0 + (threadIdx.x & 31); e < e1; e += 32) {
                    const int32_t v = col[e];
                    int32_t lv = __ldcg(level + v);
                    if (lv < 0) {
                        lv = atomicCAS(level + v, -1, lvl + 1);
                        if (lv < 0) {                            // first touch: enqueue
                            queue[atomicAdd(&s_tail, 1)] = v;
                            lv = lvl + 1;
                            if (v == t) s_found = 1;
                        }
                    }
                    if (lv == lvl + 1) atomicMin(parent + v, u);
                }
            }
            __syncthreads();
            head = tail;
            tail = s_tail;
            found = s_found != 0;
            ++lvl;
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const int32_t h = found ? __ldcg(level + t) : -1;
            hops[q] = h;
            if (paths && h >= 0 && h + 1 <= max_len) {
                int64_t *p = paths + q * (int64_t)max_len;
                int32_t v = (int32_t)t;
                for (int i = h; i >= 0; --i) {
                    p[i] = v;
                    v = __ldcg(parent + v);
                }
                for (int i = h + 1; i < max_len; ++i) p[i] = -1;
            } else if (paths) {
                for (int i = 0; i < max_len; ++i) paths[q * (int64_t)max_len + i] = -1;
            }
        }
        __syncthreads();
        // reset only what this query touched
        for (int i = threadIdx.x; i < tail; i += BFS_TPB) {
            const int32_t v = queue[i];
            level[v] = -1;
            parent[v] = BFS_UNSET;
        }
        __syncthreads();
    }
}

// Hop counts only, on a SYMMETRIC graph (built undirected): the search runs from both ends at once.
// Every step expands the current frontier of the start side AND of the goal side (a warp per
// frontier node, lanes over its arcs); a node belongs to the side that claims it first
// (level = depth * 2 + side).  An arc from a frontier node of one side (depth d) into a node of the
// other side (depth d') closes a path of d + 1 + d' arcs; the minimum over the arcs seen in the
// first step that closes any is the fewest-hop count -- every shorter path would have closed in an
// earlier step -- so a search of L hops takes ~L / 2 steps (the maze roadmap's paths are ~1000
// hops long and a step is a handful of nodes: the step count is the cost).
// Work-space: level [n]; the start side's FIFO in `queue`, the goal side's in the `parent` array
// (unused without paths; restored to BFS_UNSET afterwards).
__global__ void __launch_bounds__(BFS_TPB)
k_graph_bfs_bidir(const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ col, int64_t n,
                  const int32_t *__restrict__ comp, const int64_t *__restrict__ start,
                  const int64_t *__restrict__ goal, int64_t nq, int32_t *__restrict__ work,
                  int32_t *__restrict__ hops, unsigned long long *__restrict__ ticket) {
    int32_t *level = work + (int64_t)blockIdx.x * 3 * n, *qt = level + n, *qs = qt + n;
    __shared__ int s_tail[2], s_best;
    __shared__ long long s_q;
    while (true) {
        if (threadIdx.x == 0) {
            long long q;
            while (true) {
                q = (long long)atomicAdd(ticket, 1ull);
                if (q >= nq) break;
                const int64_t s = start[q], t = goal[q];
                const bool ok = s >= 0 && s < n && t >= 0 && t < n;
                if (ok && s != t && comp[s] == comp[t]) break;          // needs the search
                hops[q] = ok && s == t ? 0 : -1;
            }
            s_q = q;
        }
        __syncthreads();
        const int64_t q = s_q;
        if (q >= nq) return;
        const int64_t s = start[q], t = goal[q];
        if (threadIdx.x == 0) {
            level[s] = 0;                                   // depth 0, side 0
            level[t] = 1;                                   // depth 0, side 1
            qs[0] = (int32_t)s;
            qt[0] = (int32_t)t;
            s_tail[0] = 1; s_tail[1] = 1;
            s_best = 0x7fffffff;
        }
        __syncthreads();
        int head0 = 0, head1 = 0, tail0 = 1, tail1 = 1, depth = 0;
        int best = 0x7fffffff;
        while (best == 0x7fffffff && head0 < tail0 && head1 < tail1) {
            const int n0 = tail0 - head0, total = n0 + (tail1 - head1);
            for (int i = threadIdx.x >> 5; i < total; i += BFS_TPB / 32) {
                const int side = i < n0 ? 0 : 1;
                const int32_t u = side == 0 ? qs[head0 + i] : qt[head1 + (i - n0)];
                const int64_t e0 = row_ptr[u], e1 = row_ptr[u + 1];
                const int32_t mine = ((depth + 1) << 1) | side;
                for (int64_t e = e0 + (threadIdx.x & 31); e < e1; e += 32) {
                    const int32_t v = col[e];
                    int32_t lv = __ldcg(level + v);
                    if (lv < 0) {
                        lv = atomicCAS(level + v, -1, mine);
                        if (lv < 0) {                            // claimed for this side: enqueue
                            if (side == 0) qs[atomicAdd(&s_tail[0], 1)] = v;
                            else qt[atomicAdd(&s_tail[1], 1)] = v;
                            continue;
                        }
                    }
                    if ((lv & 1) != side) atomicMin(&s_best, depth + 1 + (lv >> 1));   // the sides meet on this arc
                }
            }
            __syncthreads();
            head0 = tail0; head1 = tail1;
            tail0 = s_tail[0]; tail1 = s_tail[1];
            best = s_best;
            ++depth;
            __syncthreads();
        }
        if (threadIdx.x == 0) hops[q] = best == 0x7fffffff ? -1 : best;
        // reset only what this query touched
        for (int i = threadIdx.x; i < tail0; i += BFS_TPB) level[qs[i]] = -1;
        for (int i = threadIdx.x; i < tail1; i += BFS_TPB) { level[qt[i]] = -1; }
        __syncthreads();
        for (int i = threadIdx.x; i < tail1; i += BFS_TPB) qt[i] = BFS_UNSET;      // (the parent array again)
        __syncthreads();
    }
}

__global__ void k_graph_work_init(int32_t *__restrict__ work, int64_t n, int64_t slots) {
    const int64_t total = slots * 3 * n;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t % (3 * n);
        work[t] = r < n ? -1 : BFS_UNSET;        // level | parent, queue (queue content is irrelevant)
    }
}

// For every query q: fewest-hop path start[q] -> goal[q] along the arcs.
//   hops[q]  = number of arcs, 0 when start == goal, -1 when unreachable / out of range
//   paths    = optional [nq][max_len] node indices start..goal, -1 padded; all -1 when the path
//              does not fit (hops[q] + 1 > max_len) or does not exist
// start / goal / hops / paths are device pointers.
extern "C" int32_t sbp_graph_bfs(sbp_graph *g, const int64_t *start, const int64_t *goal, int64_t nq,
                                 int32_t *hops, int64_t *paths, int32_t max_len) {
    SBP_NVTX("sbp_graph_bfs");
    SBP_REQUIRE(g && (nq == 0 || (start && goal && hops)), "sbp_graph_bfs: NULL argument");
    SBP_REQUIRE(nq >= 0 && (!paths || max_len >= 1), "sbp_graph_bfs: bad sizes");
    if (nq == 0) return SBP_OK;
    sbp_ctx *ctx = g->ctx;
    SBP_DEVICE(ctx);
    SBP_REQUIRE(g->n > 0, "sbp_graph_bfs: empty graph");
    int32_t rc = graph_ensure_components(g);
    if (rc) return rc;
    // two CTAs per SM hide the latency of the dependent loads; bounded by 8 GiB of work-space
    int64_t slots = 2 * (int64_t)ctx->sm_count;
    if (slots > nq) slots = nq;
    const int64_t per_slot = 3 * g->n * (int64_t)sizeof(int32_t);
    while (slots > 1 && slots * per_slot > ((int64_t)8 << 30)) slots /= 2;
    if (slots > g->work_slots) {
        SBP_CUDA(cudaStreamSynchronize(ctx->stream));
        if (g->d_work) SBP_CUDA(cudaFree(g->d_work));
        g->d_work = nullptr;
        g->work_slots = 0;
        SBP_CUDA(cudaMalloc(&g->d_work, (size_t)(slots * per_slot)));
        g->work_slots = slots;
        k_graph_work_init<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(g->d_work, g->n, slots);
        ctx->launches++;
    }
    unsigned long long *ticket = (unsigned long long *)(ctx->d_flags + 14);      // 8-byte aligned slot
    SBP_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned long long), ctx->stream));
    if (!paths && g->symmetric)
        k_graph_bfs_bidir<<<(unsigned)slots, BFS_TPB, 0, ctx->stream>>>(g->d_row_ptr, g->d_col, g->n, g->d_comp, start,
                                                                       goal, nq, g->d_work, hops, ticket);
    else
        k_graph_bfs<<<(unsigned)slots, BFS_TPB, 0, ctx->stream>>>(g->d_row_ptr, g->d_col, g->n, g->d_comp, start,
                                                                 goal, nq, g->d_work, hops, paths, max_len, ticket);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

// --------------------------------------- adjacency all-gather over peer memory ----
// Multi-GPU roadmap assembly: every rank packs its block of candidate
// edges -- (neighbour index + 1) * 2 + visible, one int32 per candidate -- and stores it
// straight into the symmetric buffer of EVERY rank (its own included) at its row offset.
// One kernel does the packing and the NVLink transfer (16-byte peer stores); the ranks then
// meet at a signal-pad barrier (torch symmetric memory) instead of running an NCCL all-gather
// (two kernels + ~40-70 us of collective latency for 2 MB per rank).
struct PeerBuffers { int32_t *p[16]; };

__global__ void __launch_bounds__(256)
k_adjacency_peer_scatter(const int64_t *__restrict__ nbr, const uint8_t *__restrict__ vis,
                         int64_t total, PeerBuffers peers, int world, int64_t offset) {
    const int64_t quads = total >> 2;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < quads;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t << 2;
        int4 c;
        c.x = (int32_t)((nbr[i] + 1) << 1) | (vis[i] ? 1 : 0);
        c.y = (int32_t)((nbr[i + 1] + 1) << 1) | (vis[i + 1] ? 1 : 0);
        c.z = (int32_t)((nbr[i + 2] + 1) << 1) | (vis[i + 2] ? 1 : 0);
        c.w = (int32_t)((nbr[i + 3] + 1) << 1) | (vis[i + 3] ? 1 : 0);
        for (int r = 0; r < world; ++r)
            *reinterpret_cast<int4 *>(peers.p[r] + offset + i) = c;
    }
    // tail (total not a multiple of 4)
    for (int64_t i = (quads << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t c = (int32_t)((nbr[i] + 1) << 1) | (vis[i] ? 1 : 0);
        for (int r = 0; r < world; ++r) peers.p[r][offset + i] = c;
    }
}

// nbr / vis: this rank's block (device, `total` candidates); peer_bufs: HOST array of `world`
// device pointers to the ranks' int32 buffers (peer-mapped); the block lands at element
// `offset` of each (offset a multiple of 4).  The caller synchronises the ranks afterwards.
extern "C" int32_t sbp_adjacency_peer_scatter(sbp_ctx *ctx, const int64_t *nbr, const uint8_t *vis,
                                              int64_t total, void *const *peer_bufs, int32_t world,
                                              int64_t offset) {
    SBP_NVTX("sbp_adjacency_peer_scatter");
    SBP_REQUIRE(ctx && (total == 0 || (nbr && vis)) && peer_bufs, "sbp_adjacency_peer_scatter: NULL argument");
    SBP_REQUIRE(world >= 1 && world <= 16, "sbp_adjacency_peer_scatter: 1 <= world <= 16");
    SBP_REQUIRE(total >= 0 && offset >= 0 && (offset & 3) == 0, "sbp_adjacency_peer_scatter: bad offset / size");
    if (total == 0) return SBP_OK;
    SBP_DEVICE(ctx);
    PeerBuffers pb{};
    for (int r = 0; r < world; ++r) {
        SBP_REQUIRE(peer_bufs[r] != nullptr, "sbp_adjacency_peer_scatter: NULL peer buffer");
        pb.p[r] = (int32_t *)peer_bufs[r];
    }
    int blocks = (int)(((total >> 2) + 255) / 256);
    if (blocks > ctx->sm_count * 4) blocks = ctx->sm_count * 4;
    if (blocks < 1) blocks = 1;
    k_adjacency_peer_scatter<<<blocks, 256, 0, ctx->stream>>>(nbr, vis, total, pb, world, offset);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

// Same exchange through the NVSwitch MULTICAST mapping of the symmetric buffer (NVLS): one
// `multimem.st` per 16 bytes is replicated by the switch into every rank's buffer, so a rank
// sends its block once instead of once per peer.
__global__ void __launch_bounds__(256)
k_adjacency_multicast(const int64_t *__restrict__ nbr, const uint8_t *__restrict__ vis, int64_t total,
                      int32_t *__restrict__ mc, int64_t offset) {
    const int64_t quads = total >> 2;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < quads;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t << 2;
        const float c0 = __int_as_float((int32_t)((nbr[i] + 1) << 1) | (vis[i] ? 1 : 0));
        const float c1 = __int_as_float((int32_t)((nbr[i + 1] + 1) << 1) | (vis[i + 1] ? 1 : 0));
        const float c2 = __int_as_float((int32_t)((nbr[i + 2] + 1) << 1) | (vis[i + 2] ? 1 : 0));
        const float c3 = __int_as_float((int32_t)((nbr[i + 3] + 1) << 1) | (vis[i + 3] ? 1 : 0));
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + offset + i),
                     "f"(c0), "f"(c1), "f"(c2), "f"(c3)
                     : "memory");
    }
    for (int64_t i = (quads << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float c = __int_as_float((int32_t)((nbr[i] + 1) << 1) | (vis[i] ? 1 : 0));
        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc + offset + i), "f"(c) : "memory");
    }
}

// mc_buf: the multicast (NVLS) address of the ranks' symmetric int32 buffer; see
// sbp_adjacency_peer_scatter for the rest.
extern "C" int32_t sbp_adjacency_multicast(sbp_ctx *ctx, const int64_t *nbr, const uint8_t *vis,
                                           int64_t total, void *mc_buf, int64_t offset) {
    SBP_NVTX("sbp_adjacency_multicast");
    SBP_REQUIRE(ctx && (total == 0 || (nbr && vis)) && mc_buf, "sbp_adjacency_multicast: NULL argument");
    SBP_REQUIRE(total >= 0 && offset >= 0 && (offset & 3) == 0, "sbp_adjacency_multicast: bad offset / size");
    if (total == 0) return SBP_OK;
    SBP_DEVICE(ctx);
    int blocks = (int)(((total >> 2) + 255) / 256);
    if (blocks > ctx->sm_count * 4) blocks = ctx->sm_count * 4;
    if (blocks < 1) blocks = 1;
    k_adjacency_multicast<<<blocks, 256, 0, ctx->stream>>>(nbr, vis, total, (int32_t *)mc_buf, offset);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}
