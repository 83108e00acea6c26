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
//   * a CTA is 32 rows x k columns, warp j owns column j: the lanes of a warp walk the j-th
//     nearest neighbours of 32 neighbouring rows -- edges of nearly equal length, so the warp's
//     lock-step walk wastes few lanes (a warp of mixed columns idles behind its longest edge);
//   * the walk reads a second copy of the map in 32 x 8-pixel sectors of row words
//     (word = (((y >> 3) * SX + (x >> 5)) << 3) + (y & 7), bit = x & 31): 5 integer operations per
//     pixel address instead of ~14 for the 16 x 16 tiled layout, the same 32-byte sector locality
//     in both directions; edges with all four endpoint coordinates inside [0, W) x [0, H) -- no
//     numpy wrap-around to apply -- take this path, anything else the general sequential walker;
//   * results leave through shared memory: every row's k codes are written as one contiguous
//     run (local buffer, pinned host memory, or the NVSwitch multicast mapping of the ranks'
//     symmetric buffer: the multi-GPU exchange of the adjacency overlaps the walk).
#include "common.cuh"
#include "device_fns.cuh"

// ------------------------------------------------------- rows8 map copy -----
__global__ void k_rows8_from_tiles(MapView mv, int SX8, uint32_t n_words, uint32_t *__restrict__ out) {
    for (uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x; wi < n_words; wi += gridDim.x * blockDim.x) {
        const uint32_t t = wi >> 3;
        const int y = (int)(t / (uint32_t)SX8) * 8 + (int)(wi & 7u), x0 = (int)(t % (uint32_t)SX8) * 32;
        uint32_t w = 0;
        if (y < mv.H)
            for (int b = 0; b < 32; ++b)
                if (x0 + b < mv.W && map_free<true>(mv.words, mv.SX, x0 + b, y)) w |= 1u << b;
        out[wi] = w;
    }
}

int32_t sbp_map_ensure_rows8(sbp_map *m) {
    if (m->d_rows8) return SBP_OK;
    sbp_ctx *ctx = m->ctx;
    const int SX8 = (m->view.W + 31) / 32, SY8 = (m->view.H + 7) / 8;
    const uint64_t nw = (uint64_t)SX8 * (uint64_t)SY8 * 8ull;
    SBP_REQUIRE(nw <= 0xffffffffull / 4ull, "rows8 map copy too large for 32-bit word indexing");
    SBP_CUDA(cudaMalloc(&m->d_rows8, (size_t)nw * 4));
    m->rows8 = m->view;
    m->rows8.words = m->d_rows8;
    m->rows8.SX = SX8;
    m->rows8.SY = SY8;
    m->rows8.n_words = (uint32_t)nw;
    int blocks = (int)((nw + 255) / 256);
    if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
    k_rows8_from_tiles<<<blocks, 256, 0, ctx->stream>>>(m->view, SX8, (uint32_t)nw, m->d_rows8);
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

__global__ void __launch_bounds__(1024)
k_prm_visible2d(MapView tiled, MapView rows8, const double *__restrict__ soa, int64_t capacity, int64_t n,
                const int64_t *__restrict__ idx, int K, const int32_t *__restrict__ order, int64_t p0,
                int64_t count, PrmOut out, int32_t *__restrict__ flags) {
    __shared__ int32_t s_nbr[PRM_ROWS][PRM_KMAX + 1];     // candidates, then the result codes
    __shared__ double2 s_q[PRM_ROWS];
    __shared__ int32_t s_v[PRM_ROWS];
    const unsigned FULL = 0xffffffffu;
    const int k = K - 1;
    const int lane = threadIdx.x & 31, col = threadIdx.x >> 5;
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
    // ---- phase 1: warp `col` walks column col of the 32 rows
    const int32_t v = s_v[lane];
    int myflag = 0;
    int32_t code = 0;
    {
        const int32_t nb = v >= 0 ? s_nbr[lane][col] : -1;
        const bool exists = v >= 0 && nb >= 0 && (int64_t)nb < n;
        const double2 q = s_q[lane];
        double2 p = q;
        if (exists) p = make_double2(soa[nb], soa[capacity + nb]);
        bool result = false;
        // visible(m_g.pos, v.pos): start = neighbour, end = row node (the 2-D line is direction
        // symmetric, but the argument order is the reference's)
        Edge2D ed = make_edge(p.x, p.y, q.x, q.y, tiled.W, tiled.H);
        if (v >= 0) myflag |= ed.flag;
        const bool fast = exists && ed.walk && ed.x1 >= 0 && ed.y1 >= 0 && ed.x2 >= 0 && ed.y2 >= 0;
        if (exists && ed.walk && !fast) {
            int f2 = 0;                                   // negative coordinates: numpy wrap-around
            result = edge_visible_seq<false>(tiled.words, tiled, p.x, p.y, q.x, q.y, f2);
        }
        int a = ed.a1, b = ed.b1, err = (int)(ed.da >> 1);
        const int a2 = ed.a1 + (int)ed.da, da = (int)ed.da, adb = (int)ed.adb, ystep = ed.ystep;
        const bool steep = ed.steep;
        const uint32_t *__restrict__ words = rows8.words;
        const int SX8 = rows8.SX;
        bool active = fast, freeline = fast;
        while (__any_sync(FULL, active)) {
#pragma unroll
            for (int st = 0; st < 4; ++st) {
                if (active) {
                    const int x = steep ? b : a, y = steep ? a : b;                  // :205
                    const uint32_t wi = ((uint32_t)((y >> 3) * SX8 + (x >> 5)) << 3) + (uint32_t)(y & 7);
                    const uint32_t w = __ldg(words + wi);
                    const bool fr = (w >> (x & 31)) & 1u;
                    err -= adb;                                                       // :207-210
                    if (err < 0) { b += ystep; err += da; }
                    ++a;
                    if (!fr) { freeline = false; active = false; }
                    else if (a > a2) active = false;
                }
            }
        }
        if (fast) result = freeline;
        code = ((nb + 1) << 1) | ((result && exists) ? 1 : 0);
    }
    __syncthreads();                                       // every thread has read its candidate
    if (col < k) s_nbr[lane][col] = code;
    __syncthreads();
    // ---- phase 2: a row's k codes leave as one contiguous run
    for (int t = threadIdx.x; t < PRM_ROWS * k; t += blockDim.x) {
        const int r = t / k, j = t - r * k;
        const int32_t vr = s_v[r];
        if (vr < 0) continue;
        const int32_t c = s_nbr[r][j];
        const int64_t e = (int64_t)vr * k + j;
        if (out.packed) {
            if (out.packed_mc)
                asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(out.packed + e),
                             "f"(__int_as_float(c)) : "memory");
            else
                out.packed[e] = c;
        }
        if (out.nbr) out.nbr[e] = (int64_t)(c >> 1) - 1;
        if (out.vis) out.vis[e] = (uint8_t)(c & 1);
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
    int32_t rc = sbp_map_ensure_rows8(map);
    if (rc) return rc;
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
        k_prm_visible2d<<<(unsigned)((rp.count + PRM_ROWS - 1) / PRM_ROWS), 32 * k, 0, ctx->stream>>>(
            map->view, map->rows8, nodes->d_soa, nodes->capacity, n, (const int64_t *)ibuf, K, rp.ids, rp.p0,
            rp.count, out, flags);
    }
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}
