// prm.cu -- the PRM roadmap connection step as ONE C-ABI call (sbp_prm_connect2d): exact kNN of
// the stored rows among all stored nodes + visibility of the k candidate edges of every row.
//
// Reference behaviour restated (file:line under /root/reference/sbp_env/):
//   PRMPlanner.build_graph, the body of the double loop     planners/prmPlanner.py:91-101
//       for v in nodes: for m_g in near(v): if m_g is v: continue
//           cc.visible(m_g.pos, v.pos)  ->  edge m_g -> v
//   ImgCollisionChecker.visible / get_line (Bresenham)       collisionChecker.py:134-145, :155-215
// with the k nearest neighbours as the candidate set instead of the PRM* radius.
//
// Why a dedicated kernel (profiles/r2_visible_cfg5_before_full.txt: the general kernel spent 48
// instructions per pixel step with 17 of 32 lanes active on this workload):
//   * rows are processed in CELL-SORTED order (the permutation the kNN's cell grid built anyway):
//     the 32 rows of a CTA are spatial neighbours, their edges share map sectors (L1 hits);
//   * a CTA is 32 rows, a warp walks one COLUMN of them (lane = row): the j-th nearest neighbours
//     of 32 neighbouring rows are edges of nearly equal length, so the warp's lock-step walk wastes
//     few lanes (a warp of mixed columns idles behind its longest edge); warp w takes columns w and
//     k - 1 - w one after the other, which gives every warp of the CTA the same amount of pixels;
//   * the walk reads row-major / column-major copies of the map (x-major lines: word =
//     y * WR + (x >> 5); y-major lines: word = x * HR + (y >> 5)): the major axis always runs along
//     the bits of a word and a minor step is one add, ~16 instructions per pixel instead of 48 in
//     the general kernel on the tiled layout (whose 16 x 16-pixel sector locality the cell-sorted
//     row order replaces: the CTA's working window of the map stays in L1).  Edges with all four
//     endpoint coordinates inside [0, W) x [0, H) -- no numpy wrap-around to apply -- take this
//     path, anything else the general sequential walker;
//   * every lane stores its codes itself (local buffer, pinned host memory, or the NVSwitch
//     multicast mapping of the ranks' symmetric buffer: the multi-GPU exchange of the adjacency
//     overlaps the walk); no barrier after the row setup.
#include "common.cuh"
#include "device_fns.cuh"

// ------------------------------------------------- column-major map copy -----
// Transposed twin of the row-major copy (sbp_map::rows): word = x * HR + (y >> 5), bit = y & 31.
__global__ void k_cols_from_tiles(MapView mv, int HR, uint32_t n_words, uint32_t *__restrict__ out) {
    for (uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x; wi < n_words; wi += gridDim.x * blockDim.x) {
        const int x = (int)(wi / (uint32_t)HR), y0 = (int)(wi % (uint32_t)HR) * 32;
        uint32_t w = 0;
        if (x < mv.W)
            for (int b = 0; b < 32; ++b)
                if (y0 + b < mv.H && map_free<true>(mv.words, mv.SX, x, y0 + b)) w |= 1u << b;
        out[wi] = w;
    }
}

// The column-major copy lives in ONE allocation with the row-major copy, right behind it: the
// connection kernel then addresses both through one base pointer and a 32-bit word offset.
int32_t sbp_map_ensure_cols(sbp_map *m) {
    if (m->d_cols) return SBP_OK;
    sbp_ctx *ctx = m->ctx;
    int32_t rc = sbp_map_ensure_rows(m);
    if (rc) return rc;
    const int HR = ((m->view.H + 31) / 32) | 1;
    const uint64_t nw = ((uint64_t)HR * (uint64_t)m->view.W + 3ull) & ~3ull;
    const uint64_t nr = m->rows.n_words;
    SBP_REQUIRE(nr + nw <= 0xffffffffull / 4ull, "row / column-major map copies too large for 32-bit word indexing");
    uint32_t *both = nullptr;
    SBP_CUDA(cudaMalloc(&both, (size_t)(nr + nw) * 4));
    SBP_CUDA(cudaMemcpyAsync(both, m->d_rows, (size_t)nr * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    SBP_CUDA(cudaStreamSynchronize(ctx->stream));      // nothing in flight reads the old copy any more
    SBP_CUDA(cudaFree(m->d_rows));
    m->d_rows = both;
    m->rows.words = both;
    m->d_cols = both + nr;                             // a view: freed with d_rows
    m->cols = m->view;
    m->cols.words = m->d_cols;
    m->cols.SX = HR;
    m->cols.n_words = (uint32_t)nw;
    int blocks = (int)((nw + 255) / 256);
    if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
    k_cols_from_tiles<<<blocks, 256, 0, ctx->stream>>>(m->view, HR, (uint32_t)nw, m->d_cols);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

// ------------------------------------------------------------ the kernel ----
#define PRM_ROWS 32          // rows per CTA (= lanes of a warp)
#define PRM_KMAX 31          // k <= 31 (the kNN lists hold k + 1 <= 32 entries)

struct PrmOut {
    int64_t *nbr;            // [n][k] neighbour index (nullable)
    uint8_t *vis;            // [n][k] 1 = slot filled and edge free (nullable)
    int32_t *packed;         // [n][k] (nbr + 1) * 2 + vis (nullable)
    int packed_mc;           // packed is the multicast (NVLS) address of a symmetric buffer
};

__device__ __noinline__ bool prm_edge_slow(const MapView &tiled, double2 p, double2 q) {
    int f2 = 0;
    return edge_visible_seq<false>(tiled.words, tiled, p.x, p.y, q.x, q.y, f2);
}

// One candidate edge (row lane, column col): its adjacency code.
//   x-major lines walk the row-major copy (word = y * WR + (x >> 5), bit = x & 31), y-major lines
//   the column-major copy (word = x * HR + (y >> 5), bit = y & 31): in both the major axis runs
//   along the bits of a word and a minor step moves one row of words, so the loop is the same for
//   every lane: wi = rowoff + (a >> 5), rowoff += +-WR on a minor step.
__device__ __forceinline__ int32_t prm_edge_code(const MapView &tiled, const uint32_t *__restrict__ words, int WR,
                                                 uint32_t cols_off, int HR,
                                                 const double *__restrict__ soa, int64_t capacity, int64_t n,
                                                 int32_t v, int32_t nb, double2 q, int &myflag) {
    const unsigned FULL = 0xffffffffu;
    const bool exists = v >= 0 && nb >= 0 && (int64_t)nb < n;
    double2 p = q;
    if (exists) p = make_double2(soa[nb], soa[capacity + nb]);
    bool result = false;
    // visible(m_g.pos, v.pos): start = neighbour, end = row node (the 2-D line is direction
    // symmetric, but the argument order is the reference's)
    Edge2D ed = make_edge(p.x, p.y, q.x, q.y, tiled.W, tiled.H);
    if (v >= 0) myflag |= ed.flag;
    const bool fast = exists && ed.walk && ed.x1 >= 0 && ed.y1 >= 0 && ed.x2 >= 0 && ed.y2 >= 0;
    if (exists && ed.walk && !fast)                   // negative coordinates: numpy wrap-around
        result = prm_edge_slow(tiled, p, q);
    const int stride = ed.steep ? HR : WR;
    const int da = (int)ed.da;
    int adb = (int)ed.adb;
    asm volatile("" : "+r"(adb));                                    // (keep it in a register: no re-derivation per step)
    const int dstep = ed.ystep * stride;                             // row offset of one minor step
    int a = ed.a1, err = da >> 1;                                    // :198
    int left = fast ? da + 1 : 0;                                    // pixels still to test
    uint32_t rowoff = (uint32_t)(ed.b1 * stride) + (ed.steep ? cols_off : 0u);
    bool freeline = true;
    while (__any_sync(FULL, left > 0)) {
#pragma unroll
        for (int st = 0; st < 4; ++st) {
            if (left > 0) {
                const uint32_t w = __ldg(&words[rowoff + ((uint32_t)a >> 5)]);
                const bool fr = (w >> (a & 31)) & 1u;                // :141, :151
                err -= adb;                                          // :207-210
                const int m = err >> 31;
                err += da & m;
                rowoff += (uint32_t)(dstep & m);
                ++a;
                --left;
                if (!fr) { freeline = false; left = 0; }
            }
        }
    }
    if (fast) result = freeline;
    return ((nb + 1) << 1) | ((result && exists) ? 1 : 0);
}

__global__ void __launch_bounds__(512, 3)
k_prm_visible2d(MapView tiled, MapView rowsv, MapView colsv, const double *__restrict__ soa, int64_t capacity,
                int64_t n, const int64_t *__restrict__ idx, int K, const int32_t *__restrict__ order, int64_t p0,
                int64_t count, PrmOut out, int32_t *__restrict__ flags) {
    __shared__ int32_t s_nbr[PRM_ROWS][PRM_KMAX + 1];
    __shared__ double2 s_q[PRM_ROWS];
    __shared__ int32_t s_v[PRM_ROWS];
    const int k = K - 1;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t pb = (int64_t)blockIdx.x * PRM_ROWS;
    // ---- phase 0: the CTA's 32 rows -- node id, pose, candidate list with the self entry dropped
    // (`if m_g is v: continue`, prmPlanner.py:94-95: the column equal to the row index if present,
    // else the last column, so that rows always hold k candidates)
    if (threadIdx.x < PRM_ROWS) {
        const int64_t p = pb + threadIdx.x;
        int32_t v = -1;
        if (p < count) v = order ? order[p0 + p] : (int32_t)(p0 + p);
        s_v[threadIdx.x] = v;
        if (v >= 0) {
            s_q[threadIdx.x] = make_double2(soa[v], soa[capacity + v]);
            const int64_t *row = idx + (int64_t)v * K;
            int drop = K - 1;
            for (int c = 0; c < K; ++c)
                if (row[c] == (int64_t)v) { drop = c; break; }
            for (int j = 0; j < k; ++j) s_nbr[threadIdx.x][j] = (int32_t)row[j < drop ? j : j + 1];
        }
    }
    __syncthreads();
    // ---- phase 1: warp w walks columns w and k - 1 - w of the 32 rows, one after the other: the
    // near and the far neighbours of a row pair up, so every warp of the CTA carries about the same
    // number of pixels (a warp per column would leave the near columns' warps waiting for the far
    // ones until the CTA retires)
    const int32_t v = s_v[lane];
    const double2 q = s_q[lane];
    int myflag = 0;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const int col = pass == 0 ? w : k - 1 - w;
        if (pass == 1 && col <= w) break;
        const int32_t nb = v >= 0 ? s_nbr[lane][col] : -1;
        const int32_t code = prm_edge_code(tiled, rowsv.words, rowsv.SX, rowsv.n_words, colsv.SX, soa, capacity, n,
                                           v, nb, q, myflag);
        if (v >= 0) {
            const int64_t e = (int64_t)v * k + col;
            if (out.packed) {
                if (out.packed_mc)
                    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(out.packed + e),
                                 "f"(__int_as_float(code)) : "memory");
                else
                    out.packed[e] = code;
            }
            if (out.nbr) out.nbr[e] = (int64_t)nb;
            if (out.vis) out.vis[e] = (uint8_t)(code & 1);
        }
    }
    if (myflag && flags) atomicOr(flags, myflag);
}

__global__ void k_rows_aos(const double *__restrict__ soa, int64_t capacity, int64_t n, double2 *__restrict__ X) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        X[t] = make_double2(soa[t], soa[capacity + t]);
}

// Rows = stored nodes.  The rows are split into `parts` contiguous runs of the CELL-SORTED
// order (spatially compact; `part` in [0, parts)): the multi-GPU shard of a rank.  Outputs are
// [n][k] in NODE order; only the rows of this part are written.
extern "C" int32_t sbp_prm_connect2d(sbp_map *map, sbp_nodes *nodes, int32_t k, int32_t part, int32_t parts,
                                     int64_t *nbr_out, uint8_t *vis, int32_t *packed,
                                     int32_t packed_multicast, int32_t *flags) {
    SBP_NVTX("sbp_prm_connect2d");
    SBP_REQUIRE(map && nodes, "sbp_prm_connect2d: NULL argument");
    SBP_REQUIRE(nbr_out || vis || packed, "sbp_prm_connect2d: no output requested");
    SBP_REQUIRE(nodes->d == 2, "sbp_prm_connect2d: the 2-D checker needs a 2-D node store");
    SBP_REQUIRE(map->ctx == nodes->ctx, "sbp_prm_connect2d: map and node store live in different contexts");
    SBP_REQUIRE(k >= 1 && k <= PRM_KMAX, "sbp_prm_connect2d: 1 <= k <= 31");
    SBP_REQUIRE(parts >= 1 && part >= 0 && part < parts, "sbp_prm_connect2d: bad part");
    SBP_REQUIRE(!packed_multicast || packed, "sbp_prm_connect2d: multicast needs the packed output");
    sbp_ctx *ctx = map->ctx;
    SBP_DEVICE(ctx);
    const int64_t n = nodes->n;
    if (n == 0) return SBP_OK;
    int32_t rc = sbp_map_ensure_rows(map);
    if (rc) return rc;
    if ((rc = sbp_map_ensure_cols(map))) return rc;
    const int K = k + 1;
    void *xbuf = nullptr, *ibuf = nullptr;
    if ((rc = sbp_ctx_scratch(ctx, 12, (size_t)n * 16, &xbuf))) return rc;
    if ((rc = sbp_ctx_scratch(ctx, 13, (size_t)n * K * sizeof(int64_t), &ibuf))) return rc;
    int blocks = (int)((n + 255) / 256);
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    k_rows_aos<<<blocks, 256, 0, ctx->stream>>>(nodes->d_soa, nodes->capacity, n, (double2 *)xbuf);
    ctx->launches++;
    // this part's run of the cell-sorted order (balanced like distributed.shard_range)
    const int64_t base = n / parts, rem = n % parts;
    KnnRowsPart rp;
    rp.p0 = (int64_t)part * base + (part < rem ? part : rem);
    rp.count = base + (part < rem ? 1 : 0);
    rp.deterministic = parts > 1;
    if ((rc = sbp_knn_rows_part(nodes, (const double *)xbuf, K, (int64_t *)ibuf, &rp))) return rc;
    if (rp.count == 0) return SBP_OK;
    PrmOut out{nbr_out, vis, packed, packed_multicast};
    {
        KernelTimer timer(ctx, 3);
        k_prm_visible2d<<<(unsigned)((rp.count + PRM_ROWS - 1) / PRM_ROWS), 32 * ((k + 1) / 2), 0, ctx->stream>>>(
            map->view, map->rows, map->cols, nodes->d_soa, nodes->capacity, n, (const int64_t *)ibuf, K, rp.ids,
            rp.p0, rp.count, out, flags);
    }
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}
