// prm.cu -- the PRM roadmap connection step as ONE C-ABI call (sbp_prm_connect2d): exact kNN of
// the stored rows among all stored nodes + visibility of the k candidate edges of every row.
//
// Reference behaviour restated (file:line under /root/reference/sbp_env/):
//   PRMPlanner.build_graph, the body of the double loop     planners/prmPlanner.py:91-101
//       for v in nodes: for m_g in near(v): if m_g is v: continue
//           cc.visible(m_g.pos, v.pos)  ->  edge m_g -> v
//   nearest_neighbours: np.linalg.norm(poses - pos, axis=1)  planners/prmPlanner.py:28-44
//   ImgCollisionChecker.visible / get_line (Bresenham)       collisionChecker.py:134-145, :155-215
// with the k nearest neighbours as the candidate set instead of the PRM* radius.
//
// ONE kernel answers the rows (k_prm_knn<K, true>: the kNN below, then the row's candidate edges --
// free bounding boxes from the free-run planes, probes, a CTA-local walk of the rest -- and the output
// rows from shared memory; see the comment on its FUSE parameter).  The two-kernel form
// (k_prm_knn<K, false> + k_prm_walk_sorted / k_prm_walk, tune "prm_fuse" = 0) is kept beside it.
// A THREAD per row and the rows in CELL-SORTED order (the permutation of the kNN's counting sort,
// ascending node index inside a cell):
//   k_prm_knn   the row's k nearest other nodes from the cell grid.  The 32 rows of a warp are
//     spatial neighbours: they scan the same cells of the cell-sorted pose copy `pts` (one 16-byte
//     load per candidate, no ids[e] -> soa[id] chain).  The list is kept sorted on the fp32
//     round-down of the exact fp64 d^2 (dx*dx + dy*dy unfused: the reference's sum before its root)
//     with the cell-sorted position as payload -- 2 registers and 5 instructions per slot and
//     insertion, no square root; equal keys are decided on the exact d^2 and the node index
//     (NearList).  The reference orders by (sqrt_rn(d^2), index); that differs from (d^2, index) only
//     where two DIFFERENT d^2 round to one root, i.e. inside a relative band of 2^-51: a row that
//     meets such a pair (band 2^-49) is not answered here (done = 0).  Coverage: with the list full
//     every node that can still enter lies in the cell rectangle of q +- sqrt(hi (1 + 2^-49))
//     (1 + 1e-9); cells of it outside the scanned block are scanned too, rectangles of
//     more than `limit` nodes are left to the exact path.
//   k_prm_walk  the row's k Bresenham walks, one after the other: in column j all 32 lanes walk
//     the edge to their j-th nearest neighbour -- nearly equal lengths, few idle lanes.  The walk is
//     TWO-LEVEL (device_fns.cuh:edge_free_tiles): chunks of up to 8 pixels are decided by the 2-bit
//     codes of one or two 8x8 tiles (all free / all blocked / mixed; 4 MiB for the 32k^2 map: L1 / L2
//     resident where the map itself is not); only chunks touching a mixed tile are tested pixel by
//     pixel.  Edges with an endpoint outside [0, W) x [0, H) (numpy wrap-around, out of range) take
//     the general sequential walker.  Output rows in node order (4-byte stores per lane) or in
//     cell-sorted order (`sorted_rows`: a warp's 32 rows are one contiguous block, staged in shared
//     memory and stored coalesced -- what pinned host memory and the NVSwitch multicast mapping need).
// Rows k_prm_knn does not settle (non-finite or huge coordinates, too few nodes, near ties, oversized
// rectangles) go through the exact pipeline of knn.cu (seeded filter + refine + fallback scan, all
// returning at once when nothing is pending) and k_prm_visible2d.
#include "common.cuh"
#include "device_fns.cuh"

#define PRM_KMAX 31          // k <= 31 (the kNN lists hold k + 1 <= 32 entries)
#ifndef PRM_TPB
#define PRM_TPB 128
#endif

struct PrmOut {
    int64_t *nbr;            // [n][k] neighbour index (nullable)
    uint8_t *vis;            // [n][k] 1 = slot filled and edge free (nullable)
    int32_t *packed;         // [n][k] (nbr + 1) * 2 + vis (nullable)
    int packed_mc;           // packed is the multicast (NVLS) address of a symmetric buffer
    int sorted_rows;         // rows of all three outputs in cell-sorted order (row = position in `ids`)
};

struct PrmFusedArgs {
    MapView tiled;
    TileCodeView af;
    PrmOut out;
    int32_t *flags;
    // the row poses as [n][2] for the exact path: written only when k_prm_knn left rows pending
    const double *soa;
    int64_t capacity;
    double2 *X;
    int host_out;            // `packed` is (pinned) host memory: the walk's stores are PCIe bound
    // the row permutation for the caller (nullable): positions [order_p0, order_p0 + order_count)
    int32_t *order_out;
    int64_t order_p0, order_count;
    int order_host;          // order_out is host memory: copied by the DMA engine beside the kNN
    int order_in_kernel;     // (set by the launcher) k_prm_knn<., true> writes order_out[p] for its rows itself ...
    int order_mc;            // ... through the multicast mapping
};

__device__ __noinline__ bool prm_edge_slow(const MapView &tiled, double px, double py, double qx, double qy,
                                           int *flag) {
    int f2 = 0;
    const bool r = edge_visible_seq<false>(tiled.words, tiled, px, py, qx, qy, f2);
    *flag |= f2;
    return r;
}

__device__ __forceinline__ void prm_store_packed(const PrmOut &out, int64_t e, int32_t code) {
    if (out.packed_mc)
        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(out.packed + e),
                     "f"(__int_as_float(code)) : "memory");
    else
        out.packed[e] = code;
}

// visible(m_g.pos, v.pos) of one candidate: start = neighbour p, end = row node q (the 2-D line is
// direction symmetric, but the argument order is the reference's)
template <bool PRETEST = true>
__device__ __forceinline__ bool prm_edge(const MapView &tiled, const TileCodeView &af, double px, double py,
                                         double qx, double qy, int &flag) {
    const double W = (double)tiled.W, H = (double)tiled.H;
    if (px >= 0.0 && px < W && py >= 0.0 && py < H && qx >= 0.0 && qx < W && qy >= 0.0 && qy < H) {
        // all four coordinates truncate into the index range, nothing wraps (:173-174)
        const int x1 = __double2int_rz(px), y1 = __double2int_rz(py), x2 = __double2int_rz(qx), y2 = __double2int_rz(qy);
        const int dx = x2 - x1, dy = y2 - y1;
        const bool steep = abs(dy) > abs(dx);                            // :179
        int a1 = steep ? y1 : x1, b1 = steep ? x1 : y1;                  // :182-184
        int a2 = steep ? y2 : x2, b2 = steep ? x2 : y2;
        if (a1 > a2) { int t = a1; a1 = a2; a2 = t; t = b1; b1 = b2; b2 = t; }   // :188-191
        return edge_free_tiles<PRETEST>(tiled.words, tiled.SX, af, a1, b1, (uint32_t)(a2 - a1), (uint32_t)abs(b2 - b1),
                                        b1 < b2 ? 1 : -1, steep);
    }
    return prm_edge_slow(tiled, px, py, qx, qy, &flag);
}

#define PW_ROWS 64
#define PW_BINS 32
struct PwEdge {                // 12 bytes: all coordinates of a fast-path edge lie in [0, 65536)
    uint16_t a1, b1;
    uint16_t da, adb;
    uint16_t slot;           // col * PW_ROWS + row
    uint16_t flags;          // bit 0 steep, bit 1 minor step +1
};


// ------------------------------------------- neighbour list on float keys -----
// K nearest by (d^2, node index), the row itself excluded.  The list is sorted on kf =
// round-down-to-fp32(d^2) with the cell-sorted position as payload: two registers and five
// instructions per slot and insertion.  Rounding is monotone, so entries with different keys are
// in the order of their exact d^2; whenever a candidate meets an EQUAL key the slow branch
// recomputes the exact fp64 d^2 of those entries from `pts` and orders by (d^2, node index) like
// the reference (a stable argsort of the rounded roots: equal d^2 have equal roots).  What the
// keys cannot decide is a pair of DIFFERENT d^2 inside the relative band 2^-49, whose roots may
// round to one value (then the lower index wins regardless of d^2): such a pair always shares its
// float key, is seen by the slow branch and marks the row ambiguous (-> exact path).
#define PRM_BAND 1.0000000000000018      // 1 + 2^-49
__device__ __forceinline__ double prm_d2(double2 c, double qx, double qy) {
    const double dx = __dsub_rn(c.x, qx), dy = __dsub_rn(c.y, qy);     // poses - pos (prmPlanner.py:39)
    return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));            // norm: unfused sum of squares
}

template <int K>
struct NearList {
    float kf[K];
    int32_t e[K];            // cell-sorted position, -1 = empty (kf = +inf)
    double hi;               // candidates with d^2 < hi can still enter (+inf until the list is full)
    bool amb;                // a near tie was seen: the order needs the rounded roots
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int t = 0; t < K; ++t) { kf[t] = INFINITY; e[t] = -1; }
        hi = INFINITY;
        amb = false;
    }
    // key = round-down-to-fp32 of the candidate's exact d^2 (recomputed here where it is needed)
    __device__ __forceinline__ void insert(float key, int32_t pos, double qx, double qy,
                                           const double2 *__restrict__ pts, const int32_t *__restrict__ ids) {
        bool anyeq = false;
#pragma unroll
        for (int t = 0; t < K; ++t) anyeq |= kf[t] == key;
        if (!anyeq) {
            bool r_cur = kf[K - 1] < key;                    // r_t: entry t ranks before the new one
#pragma unroll
            for (int t = K - 1; t >= 1; --t) {
                const bool r_prev = kf[t - 1] < key;
                const float nk = r_prev ? key : kf[t - 1];
                const int32_t ne = r_prev ? pos : e[t - 1];
                kf[t] = r_cur ? kf[t] : nk;
                e[t] = r_cur ? e[t] : ne;
                r_cur = r_prev;
            }
            kf[0] = r_cur ? kf[0] : key;
            e[0] = r_cur ? e[0] : pos;
        } else {
            const int32_t id = ids[pos];
            const double dd = prm_d2(pts[pos], qx, qy);
            bool r[K];
#pragma unroll
            for (int t = 0; t < K; ++t) {
                const bool eq = kf[t] == key;
                bool before = kf[t] < key;
                if (eq) {
                    const double dt = prm_d2(pts[e[t]], qx, qy);
                    before = dt < dd || (dt == dd && ids[e[t]] < id);
                    const double lo = dt < dd ? dt : dd, up = dt < dd ? dd : dt;
                    amb |= (dt != dd) & (up <= lo * PRM_BAND);
                }
                r[t] = before;
            }
#pragma unroll
            for (int t = K - 1; t >= 1; --t) {
                const float nk = r[t - 1] ? key : kf[t - 1];
                const int32_t ne = r[t - 1] ? pos : e[t - 1];
                kf[t] = r[t] ? kf[t] : nk;
                e[t] = r[t] ? e[t] : ne;
            }
            kf[0] = r[0] ? kf[0] : key;
            e[0] = r[0] ? e[0] : pos;
        }
    }
    // everything whose key is <= the K-th key may still enter: d^2 < next fp32 above that key
    // (after a step's insertions: the step's candidates were filtered with the bound before it)
    __device__ __forceinline__ void update_hi() {
        const float kth = kf[K - 1];
        hi = kth < INFINITY ? (double)__uint_as_float(__float_as_uint(kth) + 1u) : INFINITY;
    }
    __device__ __forceinline__ bool full() const { return e[K - 1] >= 0; }
};

// The rows of cells [ya, yb] x [xa, xb] minus the already scanned block [sy0, sy1] x [sx0, sx1]
// (sx1 < sx0: nothing cut out), the 32 lanes in lock-step: per row one contiguous range of `pts`
// (two when the block is cut out), U candidates per lane and step, then the lanes insert their
// passing candidates round by round.  centre_out: rows nearest to `mid` first, so that the list
// fills with near nodes and its bound tightens early.  `self`: the row's own position (skipped).
template <int K, int U>
__device__ __forceinline__ void prm_scan_rows(const double2 *__restrict__ pts, const int32_t *__restrict__ ids,
                                              const int32_t *__restrict__ start, int G, double qx, double qy,
                                              int32_t self, NearList<K> &L, bool active, int xa, int xb, int ya,
                                              int yb, int sx0, int sx1, int sy0, int sy1, int mid,
                                              bool centre_out) {
    const unsigned FULL = 0xffffffffu;
    const int nrows = active ? yb - ya + 1 : 0;
    const int maxrows = __reduce_max_sync(FULL, nrows);
    const bool cut = sx1 >= sx0;
    const int nseg = __any_sync(FULL, active && cut) ? 2 : 1;
    for (int j = 0; j < maxrows; ++j) {
        int row = ya + j;
        if (centre_out) {
            const int up = mid + ((j + 1) >> 1), dn = mid - ((j + 1) >> 1);
            const int lo_left = mid - ya, hi_left = yb - mid;      // rows below / above mid
            const int paired = 2 * (lo_left < hi_left ? lo_left : hi_left);
            if (j <= paired) row = (j & 1) ? up : dn;
            else row = lo_left < hi_left ? mid + (j - lo_left) : mid - (j - hi_left);
        }
        const bool rowok = j < nrows;
        const bool in_cut = cut && row >= sy0 && row <= sy1;
        for (int seg = 0; seg < nseg; ++seg) {
            // outside the cut rows: [xa, xb]; inside: [xa, sx0 - 1] and [sx1 + 1, xb]
            const int c0 = seg == 0 ? xa : (sx1 + 1 > xa ? sx1 + 1 : xa);
            const int c1 = seg == 0 ? (in_cut ? (sx0 - 1 < xb ? sx0 - 1 : xb) : xb) : xb;
            int32_t e = 0, e1 = 0;
            if (rowok && c1 >= c0 && (seg == 0 || in_cut)) {
                e = start[(int64_t)row * G + c0];
                e1 = start[(int64_t)row * G + c1 + 1];
            }
            while (__any_sync(FULL, e < e1)) {
                float key[U];
                unsigned mask = 0;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool ok = e + u < e1 && e + u != self;
                    const double d2 = prm_d2(pts[e + u < e1 ? e + u : 0], qx, qy);
                    key[u] = __double2float_rd(d2);
                    mask |= (ok && d2 < L.hi) ? (1u << u) : 0u;
                }
                const int32_t base = e;
                e += U;
                // round r: every lane inserts its r-th passing candidate
                const int rounds = __reduce_max_sync(FULL, __popc(mask));
                for (int r = 0; r < rounds; ++r) {
                    if (mask) {
                        const int u = __ffs(mask) - 1;
                        mask &= mask - 1;
                        float kk = key[0];
#pragma unroll
                        for (int v = 1; v < U; ++v) kk = u == v ? key[v] : kk;
                        L.insert(kk, base + u, qx, qy, pts, ids);
                    }
                }
                if (rounds) L.update_hi();
            }
        }
    }
}

// kq = candidates per row; K = list length >= kq.  cand [count][kq]: cell-sorted positions of the
// row's candidates in the reference's order (rows with done = 0: unspecified).
// FUSE: the kernel also decides the row's candidate edges and writes the output rows itself.  An edge
// whose bounding box of tiles is all free is visible (two words of the free-run planes,
// device_fns.cuh:tiles_rect_free -- most edges of a roadmap); the others (near walls, most of them
// blocked) are collected per CTA in shared memory and walked side by side over its threads (two-level
// walk on the tile codes).  No candidate list in global memory, no second pass over all edges.
template <int K, bool FUSE>
__global__ void __launch_bounds__(PRM_TPB, (FUSE && K <= 10) ? 8 : 1)
k_prm_knn(CellGridView g, int kq, int r_init, int limit, int64_t p0, int64_t count, double *__restrict__ seed2,
          uint8_t *__restrict__ done, int32_t *__restrict__ pending, int32_t *__restrict__ cand, PrmFusedArgs a) {
    const int64_t slot = (int64_t)blockIdx.x * PRM_TPB + threadIdx.x;
    const bool live = slot < count;
    const int64_t p = p0 + (live ? slot : 0);
    const int32_t v = g.ids[p];
    const double2 q = g.pts[p];
    const double mnx = g.b4[0], mxx = g.b4[1], mny = g.b4[2], mxy = g.b4[3];
    const int G = g.G;
    const int band0 = g.band ? g.band[0] : 0, band1 = g.band ? g.band[1] : G - 1;
    // finite coordinates whose squared distances stay inside the fp32 range, and at least K other
    // nodes: otherwise the order involves inf / NaN distances or unfilled slots (pad_nan_row)
    const bool usable = live && *g.maxabs < 1e18 && g.n > K && mnx <= mxx && mny <= mxy;
    NearList<K> L;
    L.init();
    int x0 = 0, x1 = -1, y0 = 0, y1 = -1, cy_q = 0;
    bool have_block = false;
    if (usable) {
        const int cx = cell_of(q.x, mnx, mxx, G), cy = cell_of(q.y, mny, mxy, G);
        cy_q = cy;
        for (int r = r_init;; r = r < 8 ? r + 1 : 2 * r) {
            x0 = cx - r < 0 ? 0 : cx - r; x1 = cx + r > G - 1 ? G - 1 : cx + r;
            y0 = cy - r < 0 ? 0 : cy - r; y1 = cy + r > G - 1 ? G - 1 : cy + r;
            int32_t c = 0;
            for (int row = y0; row <= y1; ++row)
                c += g.start[(int64_t)row * G + x1 + 1] - g.start[(int64_t)row * G + x0];
            have_block = c > K;                                   // K besides the row itself
            if (have_block || (x0 == 0 && y0 == 0 && x1 == G - 1 && y1 == G - 1)) break;
        }
        // (a rank's grid covers a band of cell rows: a search that leaves it goes to the exact path)
        have_block = have_block && y0 >= band0 && y1 <= band1;
    }
    prm_scan_rows<K, 4>(g.pts, g.ids, g.start, G, q.x, q.y, (int32_t)p, L, have_block, x0, x1, y0, y1, 0, -1, 0, -1,
                        cy_q, true);
    const bool full = have_block && L.full();
    bool handled = false, pass2 = false;
    int xa = 0, xb = -1, ya = 0, yb = -1;
    if (full) {
        // every node that can still enter has d^2 < hi, or lies in the rounding band above it
        const double r = sqrt(L.hi * PRM_BAND) * (1.0 + 1e-9) + 1e-300;
        xa = cell_of(__dsub_rd(q.x, r), mnx, mxx, G); xb = cell_of(__dadd_ru(q.x, r), mnx, mxx, G);
        ya = cell_of(__dsub_rd(q.y, r), mny, mxy, G); yb = cell_of(__dadd_ru(q.y, r), mny, mxy, G);
        if (xa >= x0 && xb <= x1 && ya >= y0 && yb <= y1) {
            handled = true;                                       // the block already covers the rectangle
        } else if (yb - ya < 64 && ya >= band0 && yb <= band1) {
            int32_t c = 0;
            for (int row = ya; row <= yb; ++row)
                c += g.start[(int64_t)row * G + xb + 1] - g.start[(int64_t)row * G + xa];
            pass2 = c <= limit;
            handled = pass2;
        }
    }
    // pass 2: the cells of the rectangle outside the block (the block's nodes are in the list already;
    // the list only improves, so the rectangle of the first bound stays a superset)
    prm_scan_rows<K, 4>(g.pts, g.ids, g.start, G, q.x, q.y, (int32_t)p, L, pass2, xa, xb, ya, yb, x0, x1, y0, y1, 0,
                        false);
    if (full) {
        // near ties the equal-key test of the insertion cannot see: a pair whose d^2 straddle an fp32
        // rounding boundary (adjacent keys) -- inside the list, or the K-th entry against a node that
        // stayed outside (every such node has d^2 >= hi)
        bool amb2 = prm_d2(g.pts[L.e[K - 1]], q.x, q.y) * PRM_BAND >= L.hi;
#pragma unroll
        for (int t = 0; t + 1 < K; ++t)
            if (L.kf[t + 1] == __uint_as_float(__float_as_uint(L.kf[t]) + 1u))
                amb2 |= prm_d2(g.pts[L.e[t + 1]], q.x, q.y) <= prm_d2(g.pts[L.e[t]], q.x, q.y) * PRM_BAND;
        L.amb |= amb2;
    }
    handled = handled && !L.amb;
    if (live) {
        if (!handled) {
            // the exact path has work: its bound for K + 1 entries, the row itself included (the K-th
            // other node).  done[] is 1 for every row since the grid build, and the bound is read for
            // rows with done = 0 only, so a settled row writes neither (1 M scattered stores less)
            seed2[v] = full ? L.hi * (1.0 + 1e-9) + 1e-300 : INFINITY;
            done[v] = 0;
            atomicAdd(pending, 1);
        }
        if (handled && !FUSE) {
#pragma unroll
            for (int t = 0; t < K; ++t)
                if (t < kq) cand[slot * kq + t] = L.e[t];
        }
    }
    if (FUSE) {
        __shared__ __align__(16) int32_t s_out[PRM_TPB * K];      // [row of the CTA][kq]: the output block
        __shared__ uint32_t s_xy[PRM_TPB * K];                    // neighbour pixel x | y << 16 of every edge
        __shared__ uint32_t s_q[PRM_TPB];                         // row pixel x | y << 16
        __shared__ uint16_t s_list[PRM_TPB * K], s_list2[PRM_TPB * K];   // edges left after phase A / phase B
        __shared__ int s_wsum[PRM_TPB / 32], s_n2;
        const unsigned FULL = 0xffffffffu;
        const PrmOut &out = a.out;
        const bool work = live && handled;
        const bool stage = out.sorted_rows && out.packed;
        const int tid = threadIdx.x;
        if (tid == 0) s_n2 = 0;
        if (live && a.order_in_kernel) {                          // the row permutation: node of cell-sorted position p
            if (a.order_mc)
                asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(a.order_out + p), "f"(__int_as_float(v)) : "memory");
            else
                a.order_out[p] = v;
        }
        const double W = (double)a.tiled.W, H = (double)a.tiled.H;
        const bool q_in = q.x >= 0.0 && q.x < W && q.y >= 0.0 && q.y < H;
        const int qxi = q_in ? __double2int_rz(q.x) : 0, qyi = q_in ? __double2int_rz(q.y) : 0;
        s_q[tid] = (uint32_t)qxi | ((uint32_t)qyi << 16);
        int myflag = 0;
        // ---- phase A: all edges of the row; bounding box of tiles all free -> visible
        unsigned walk = 0;
        if (work) {
#pragma unroll
            for (int t = 0; t < K; ++t) {
                if (t < kq) {
                    const int32_t e = L.e[t];
                    const double2 pp = g.pts[e];
                    int32_t code = (g.ids[e] + 1) << 1;
                    if (q_in && pp.x >= 0.0 && pp.x < W && pp.y >= 0.0 && pp.y < H) {
                        // start = neighbour, end = row node (:173-174)
                        const int x1 = __double2int_rz(pp.x), y1 = __double2int_rz(pp.y);
                        s_xy[tid * kq + t] = (uint32_t)x1 | ((uint32_t)y1 << 16);
                        if (tiles_rect_free(a.af, min(x1, qxi) >> 3, max(x1, qxi) >> 3, min(y1, qyi) >> 3, max(y1, qyi) >> 3))
                            code |= 1;
                        else
                            walk |= 1u << t;
                    } else {
                        code |= prm_edge_slow(a.tiled, pp.x, pp.y, q.x, q.y, &myflag) ? 1 : 0;   // wrap-around / out of range
                    }
                    s_out[tid * kq + t] = code;
                }
            }
        }
        // the undecided edges (near walls; most are blocked) as one list of the CTA (warp-local lists were
        // measured: no barriers, but a walk round per warp instead of one or two per CTA: 0.345 instead of 0.330 ms)
        int total;
        {
            const int mine = __popc(walk);
            int inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(FULL, inc, o); if ((tid & 31) >= o) inc += u; }
            if ((tid & 31) == 31) s_wsum[tid >> 5] = inc;
            __syncthreads();
            int base = inc - mine;
            total = 0;
#pragma unroll
            for (int w = 0; w < PRM_TPB / 32; ++w) { const int c = s_wsum[w]; base += w < (tid >> 5) ? c : 0; total += c; }
            while (walk) {
                const int t = __ffs(walk) - 1;
                walk &= walk - 1;
                s_list[base++] = (uint16_t)(tid * kq + t);
            }
        }
        __syncthreads();
        const uint32_t inv_kq = (1u << 20) / (uint32_t)kq + 1u;    // slot / kq = (slot * inv_kq) >> 20 for slot < 4096, kq <= 31
        // ---- phase B: probes.  Line pixels 16 apart along the major axis: one of them in an all-blocked
        // tile (or a blocked pixel of a mixed one) settles the edge as blocked -- a line that crosses a wall
        // of that thickness cannot miss them; what is left goes to the walk.
        for (int i = tid; i < total; i += PRM_TPB) {
            const int sl = s_list[i];
            const uint32_t qq = s_q[((uint32_t)sl * inv_kq) >> 20], xy = s_xy[sl];
            const int x1 = (int)(xy & 0xffffu), y1 = (int)(xy >> 16), x2 = (int)(qq & 0xffffu), y2 = (int)(qq >> 16);
            const int dx = x2 - x1, dy = y2 - y1;
            const bool steep = abs(dy) > abs(dx);                        // :179-191
            int a1 = steep ? y1 : x1, b1 = steep ? x1 : y1;
            int a2 = steep ? y2 : x2, b2 = steep ? x2 : y2;
            if (a1 > a2) { int sw = a1; a1 = a2; a2 = sw; sw = b1; b1 = b2; b2 = sw; }
            const uint32_t da = (uint32_t)(a2 - a1), adb = (uint32_t)abs(b2 - b1);
            const int ystep = b1 < b2 ? 1 : -1;
            const uint32_t dd = da ? da : 1u, c = da ? da - 1u - (da >> 1) : 0u;
            const float rcp = __fdividef(0.999998f, (float)dd);
            bool blocked = false;
            for (uint32_t pi = 8; pi < da && !blocked; pi += 16) {
                const uint32_t num = pi * adb + c;
                uint32_t qe = (uint32_t)(__uint2float_rz(num) * rcp);
                if (num - qe * dd >= dd) ++qe;
                const int pa = a1 + (int)pi, pb = b1 + ystep * (int)qe;
                const int px = steep ? pb : pa, py = steep ? pa : pb;
                const uint32_t word = __ldg(a.af.words + ((uint32_t)py >> 5) * (uint32_t)a.af.AW + ((uint32_t)px >> 5));
                const uint32_t cd = (word >> (((((uint32_t)py >> 3) & 3u) << 3) | ((((uint32_t)px >> 3) & 3u) << 1))) & 3u;
                blocked = cd == 2u || (cd == 0u && !map_free<true>(a.tiled.words, a.tiled.SX, px, py));
            }
            if (!blocked) s_list2[atomicAdd(&s_n2, 1)] = (uint16_t)sl;
        }
        __syncthreads();
        // ---- phase C: the two-level walk of the rest, side by side over the threads
        const int total2 = s_n2;
        for (int i = tid; i < total2; i += PRM_TPB) {
            const int sl = s_list2[i];
            const uint32_t qq = s_q[((uint32_t)sl * inv_kq) >> 20], xy = s_xy[sl];
            const int x1 = (int)(xy & 0xffffu), y1 = (int)(xy >> 16), x2 = (int)(qq & 0xffffu), y2 = (int)(qq >> 16);
            const int dx = x2 - x1, dy = y2 - y1;
            const bool steep = abs(dy) > abs(dx);
            int a1 = steep ? y1 : x1, b1 = steep ? x1 : y1;
            int a2 = steep ? y2 : x2, b2 = steep ? x2 : y2;
            if (a1 > a2) { int sw = a1; a1 = a2; a2 = sw; sw = b1; b1 = b2; b2 = sw; }
            if (edge_free_tiles<false>(a.tiled.words, a.tiled.SX, a.af, a1, b1, (uint32_t)(a2 - a1), (uint32_t)abs(b2 - b1),
                                       b1 < b2 ? 1 : -1, steep))
                s_out[sl] |= 1;
        }
        __syncthreads();
        if (myflag && a.flags) atomicOr(a.flags, myflag);
        if (work && (!stage || out.nbr || out.vis)) {
            const int64_t orow = out.sorted_rows ? p : (int64_t)v;
            for (int t = 0; t < kq; ++t) {
                const int32_t code = s_out[tid * kq + t];
                const int64_t o = orow * kq + t;
                if (!stage && out.packed) prm_store_packed(out, o, code);
                if (out.nbr) out.nbr[o] = (int64_t)(code >> 1) - 1;
                if (out.vis) out.vis[o] = (uint8_t)(code & 1);
            }
        }
        if (stage) {
            // the warp's 32 rows are one contiguous block of 32 * kq codes, in shared memory as in the output
            // (rows not settled here are written later by the exact path)
            const unsigned okmask = __ballot_sync(FULL, work);
            const int lane = tid & 31, wbase = tid & ~31;
            const int64_t slot0 = (int64_t)blockIdx.x * PRM_TPB + wbase;
            if (slot0 < count) {
                const int rows = (int)(count - slot0 < 32 ? count - slot0 : 32);
                const int64_t base = (p0 + slot0) * kq;
                const int32_t *src = s_out + wbase * kq;                  // (32 * kq * 4 bytes per warp: 16-byte aligned)
                int32_t *dst = out.packed + base;
                if (okmask == FULL && rows == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                    const int4 *src4 = reinterpret_cast<const int4 *>(src);
                    for (int t = lane; t < 8 * kq; t += 32) {
                        const int4 c = src4[t];
                        if (out.packed_mc)      // through the switch in 16-byte pieces
                            asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * t),
                                         "f"(__int_as_float(c.x)), "f"(__int_as_float(c.y)), "f"(__int_as_float(c.z)),
                                         "f"(__int_as_float(c.w)) : "memory");
                        else
                            reinterpret_cast<int4 *>(dst)[t] = c;
                    }
                } else {
                    int row = 0, col = lane;                             // word i = row * kq + col, i = lane, lane + 32, ...
                    while (col >= kq) { col -= kq; ++row; }
                    for (int i = lane; i < rows * kq; i += 32) {
                        if ((okmask >> row) & 1u) prm_store_packed(out, base + i, src[i]);
                        col += 32;
                        while (col >= kq) { col -= kq; ++row; }
                    }
                }
            }
        }
    }
}

// One thread per row: its kq candidate edges, one after the other (in column j all lanes of the
// warp walk the edge to their j-th nearest neighbour: nearly equal lengths).
__global__ void __launch_bounds__(PRM_TPB)
k_prm_walk(CellGridView g, int kq, int64_t p0, int64_t count, const uint8_t *__restrict__ done,
           const int32_t *__restrict__ cand, PrmFusedArgs a) {
    __shared__ int32_t s_code[PRM_KMAX][PRM_TPB];
    const int tid = threadIdx.x;
    const int64_t slot = (int64_t)blockIdx.x * PRM_TPB + tid;
    const bool live = slot < count;
    const int64_t p = p0 + (live ? slot : 0);
    const int32_t v = g.ids[p];
    const bool work = live && done[v];
    const double2 q = g.pts[p];
    const PrmOut &out = a.out;
    const int64_t orow = out.sorted_rows ? p : (int64_t)v;
    const bool stage = out.sorted_rows && out.packed;            // packed codes leave through shared memory
    int myflag = 0;
    if (work) {
#pragma unroll 1
        for (int col = 0; col < kq; ++col) {
            const int32_t e = cand[slot * kq + col];
            const int32_t nb = g.ids[e];
            const double2 pp = g.pts[e];
            const bool free_edge = prm_edge(a.tiled, a.af, pp.x, pp.y, q.x, q.y, myflag);
            const int32_t code = ((nb + 1) << 1) | (free_edge ? 1 : 0);
            const int64_t o = orow * kq + col;
            if (stage) s_code[col][tid] = code;
            else if (out.packed) prm_store_packed(out, o, code);
            if (out.nbr) out.nbr[o] = (int64_t)nb;
            if (out.vis) out.vis[o] = (uint8_t)(code & 1);
        }
    }
    if (stage) {
        // the warp's 32 rows are one contiguous block of 32 * kq codes: lane l stores words l, l + 32, ...
        // (rows the kNN kernel did not settle are written later by the exact path)
        const unsigned FULL = 0xffffffffu;
        const unsigned okmask = __ballot_sync(FULL, work);
        const int lane = tid & 31, wbase = tid & ~31;
        const int64_t slot0 = (int64_t)blockIdx.x * PRM_TPB + wbase;
        if (slot0 < count) {
            const int rows = (int)(count - slot0 < 32 ? count - slot0 : 32);
            const int64_t base = (p0 + slot0) * kq;
            if (out.packed_mc && okmask == 0xffffffffu && rows == 32 && (base & 3) == 0) {
                // through the switch in 16-byte pieces (the block is 16-byte aligned and whole)
                for (int i4 = lane * 4; i4 < 32 * kq; i4 += 128) {
                    float f[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int i = i4 + j, row = i / kq, col = i - row * kq;
                        f[j] = __int_as_float(s_code[col][wbase + row]);
                    }
                    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out.packed + base + i4),
                                 "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3]) : "memory");
                }
            } else {
                for (int i = lane; i < rows * kq; i += 32) {
                    const int row = i / kq, col = i - row * kq;
                    if ((okmask >> row) & 1u) prm_store_packed(out, base + i, s_code[col][wbase + row]);
                }
            }
        }
    }
    if (myflag && a.flags) atomicOr(a.flags, myflag);
}

__global__ void k_copy_order(const int32_t *__restrict__ ids, int64_t p0, int64_t count, int32_t *__restrict__ order_out,
                             int multicast);

// A finished slice of the rank's rows, device staging -> every rank's copy through the NVSwitch
// multicast mapping, 16 bytes per store (src / dst 16-byte aligned, `words` a multiple of 4 except
// for the tail).
__global__ void __launch_bounds__(256)
k_mc_copy(const int32_t *__restrict__ src, int32_t *__restrict__ dst_mc, int64_t words) {
    const int64_t quads = words >> 2;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < quads; t += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = *reinterpret_cast<const float4 *>(src + 4 * t);
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst_mc + 4 * t), "f"(v.x),
                     "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
    for (int64_t i = (quads << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < words;
         i += (int64_t)gridDim.x * blockDim.x)
        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(dst_mc + i), "f"(__int_as_float(src[i])) : "memory");
}

__global__ void k_rows_aos(const double *__restrict__ soa, int64_t capacity, int64_t n, double2 *__restrict__ X,
                           const int32_t *__restrict__ pending);

// The same walk with the CTA's edges SORTED BY LENGTH first (the default): in k_prm_walk a warp's 32
// lanes walk the j-th edges of 32 rows in lock-step and wait for the longest one (20.5 of 32 lanes
// active in the chunk loop, profiles/r2_prm_walk_full.txt).  Here a CTA of 64 rows sets up its 64 k
// edges (thread = row), counting-sorts them in shared memory by their number of 8-pixel chunks,
// and the threads then walk the sorted list: the 32 edges of a warp step have the same length.
// Results return to their (row, column) slot in shared memory; output as in k_prm_walk.
__global__ void __launch_bounds__(PW_ROWS)
k_prm_walk_sorted(CellGridView g, int kq, int64_t p0, int64_t count, const uint8_t *__restrict__ done,
                  const int32_t *__restrict__ cand, PrmFusedArgs a) {
    extern __shared__ __align__(16) unsigned char pw_smem[];
    int32_t *s_code = reinterpret_cast<int32_t *>(pw_smem);                            // [PW_ROWS][kq]: the output block
    PwEdge *s_edge = reinterpret_cast<PwEdge *>(s_code + kq * PW_ROWS);                // [PW_ROWS * kq]
    uint16_t *s_idx = reinterpret_cast<uint16_t *>(s_edge + kq * PW_ROWS);             // [PW_ROWS * kq]
    __shared__ int s_cnt[PW_BINS], s_off[PW_BINS], s_total;
    const int tid = threadIdx.x;
    const int64_t slot = (int64_t)blockIdx.x * PW_ROWS + tid;
    const bool live = slot < count;
    const int64_t p = p0 + (live ? slot : 0);
    const int32_t v = g.ids[p];
    const bool work = live && done[v];
    const double2 q = g.pts[p];
    const PrmOut &out = a.out;
    if (tid < PW_BINS) s_cnt[tid] = 0;
    __syncthreads();
    // ---- phase 1: set up the row's edges; an edge whose bounding box of tiles is all free is decided
    // here (two words of the free-run planes), the others are counted per length bin.  The loads of
    // the next columns are issued before the current one is worked on (candidate two columns ahead,
    // its pose one ahead): one memory round trip per column instead of three.
    const double W = (double)a.tiled.W, H = (double)a.tiled.H;
    const bool q_in = q.x >= 0.0 && q.x < W && q.y >= 0.0 && q.y < H;
    const int qxi = q_in ? __double2int_rz(q.x) : 0, qyi = q_in ? __double2int_rz(q.y) : 0;
    int myflag = 0;
    if (work) {
        const int32_t *__restrict__ crow = cand + slot * kq;
        int32_t e_ahead = kq > 1 ? crow[1] : 0;
        int32_t nb_next;
        double2 pp_next;
        {
            const int32_t e0 = crow[0];
            nb_next = g.ids[e0];
            pp_next = g.pts[e0];
        }
#pragma unroll 1
        for (int col = 0; col < kq; ++col) {
            const int32_t nb = nb_next;
            const double2 pp = pp_next;
            if (col + 1 < kq) { nb_next = g.ids[e_ahead]; pp_next = g.pts[e_ahead]; }
            if (col + 2 < kq) e_ahead = crow[col + 2];
            const int sl = tid * kq + col;
            int32_t code = (nb + 1) << 1;
            int bin = -1;
            if (q_in && pp.x >= 0.0 && pp.x < W && pp.y >= 0.0 && pp.y < H) {
                // start = neighbour, end = row node (:173-174)
                const int x1 = __double2int_rz(pp.x), y1 = __double2int_rz(pp.y);
                const int xl = min(x1, qxi) >> 3, xh = max(x1, qxi) >> 3, yl = min(y1, qyi) >> 3, yh = max(y1, qyi) >> 3;
                if (tiles_rect_free(a.af, xl, xh, yl, yh)) {
                    code |= 1;
                } else {
                    // steep / order swaps (:179-191)
                    const int dx = qxi - x1, dy = qyi - y1;
                    const bool steep = abs(dy) > abs(dx);
                    int a1 = steep ? y1 : x1, b1 = steep ? x1 : y1;
                    int a2 = steep ? qyi : qxi, b2 = steep ? qxi : qyi;
                    if (a1 > a2) { int t = a1; a1 = a2; a2 = t; t = b1; b1 = b2; b2 = t; }
                    PwEdge ed;
                    ed.a1 = (uint16_t)a1; ed.b1 = (uint16_t)b1;
                    ed.da = (uint16_t)(a2 - a1); ed.adb = (uint16_t)abs(b2 - b1);
                    ed.slot = (uint16_t)sl;
                    ed.flags = (uint16_t)((steep ? 1 : 0) | (b1 < b2 ? 2 : 0));
                    s_edge[sl] = ed;
                    const int nch = (a2 >> 3) - (a1 >> 3) + 1;
                    bin = nch < PW_BINS ? nch - 1 : PW_BINS - 1;
                    atomicAdd(&s_cnt[bin], 1);
                }
            } else {
                code |= prm_edge_slow(a.tiled, pp.x, pp.y, q.x, q.y, &myflag) ? 1 : 0;   // wrap-around / out of range
            }
            s_code[sl] = code;
            s_idx[sl] = (uint16_t)(bin + 1);          // (bin + 1 until the placement below; 0 = not walked here)
        }
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int b = PW_BINS - 1; b >= 0; --b) { s_off[b] = acc; acc += s_cnt[b]; }   // longest edges first
        s_total = acc;
    }
    __syncthreads();
    // ---- placement: the sorted order of the edges (read the bins back before anything is overwritten)
    const int total = s_total;
    if (total > 0) {
        int mybins[PRM_FUSED_KMAX - 1];
#pragma unroll
        for (int col = 0; col < PRM_FUSED_KMAX - 1; ++col) mybins[col] = (work && col < kq) ? (int)s_idx[tid * kq + col] - 1 : -1;
        __syncthreads();
#pragma unroll
        for (int col = 0; col < PRM_FUSED_KMAX - 1; ++col)
            if (mybins[col] >= 0) s_idx[atomicAdd(&s_off[mybins[col]], 1)] = (uint16_t)(tid * kq + col);
        __syncthreads();
        // ---- phase 2: walk the sorted edges
        for (int i = tid; i < total; i += PW_ROWS) {
            const PwEdge ed = s_edge[s_idx[i]];
            const bool fr = edge_free_tiles<false>(a.tiled.words, a.tiled.SX, a.af, (int)ed.a1, (int)ed.b1, ed.da, ed.adb,
                                                   (ed.flags & 2) ? 1 : -1, (ed.flags & 1) != 0);
            if (fr) s_code[ed.slot] |= 1;
        }
        __syncthreads();
    }
    // ---- output
    const int64_t orow = out.sorted_rows ? p : (int64_t)v;
    const bool stage = out.sorted_rows && out.packed;
    if (work && (!stage || out.nbr || out.vis)) {
        for (int col = 0; col < kq; ++col) {
            const int32_t code = s_code[tid * kq + col];
            const int64_t o = orow * kq + col;
            if (!stage && out.packed) prm_store_packed(out, o, code);
            if (out.nbr) out.nbr[o] = (int64_t)(code >> 1) - 1;
            if (out.vis) out.vis[o] = (uint8_t)(code & 1);
        }
    }
    if (stage) {
        // the warp's 32 rows are one contiguous block of 32 * kq codes, in shared memory as in the output
        // (rows the kNN kernel did not settle are written later by the exact path)
        const unsigned FULL = 0xffffffffu;
        const unsigned okmask = __ballot_sync(FULL, work);
        const int lane = tid & 31, wbase = tid & ~31;
        const int64_t slot0 = (int64_t)blockIdx.x * PW_ROWS + wbase;
        if (slot0 < count) {
            const int rows = (int)(count - slot0 < 32 ? count - slot0 : 32);
            const int64_t base = (p0 + slot0) * kq;
            const int32_t *src = s_code + wbase * kq;                    // (32 * kq * 4 bytes per warp: 16-byte aligned)
            int32_t *dst = out.packed + base;
            if (okmask == FULL && rows == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                const int4 *src4 = reinterpret_cast<const int4 *>(src);
                for (int t = lane; t < 8 * kq; t += 32) {
                    const int4 c = src4[t];
                    if (out.packed_mc)      // through the switch in 16-byte pieces
                        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * t),
                                     "f"(__int_as_float(c.x)), "f"(__int_as_float(c.y)), "f"(__int_as_float(c.z)),
                                     "f"(__int_as_float(c.w)) : "memory");
                    else
                        reinterpret_cast<int4 *>(dst)[t] = c;
                }
            } else {
                int row = 0, col = lane;                                 // word i = row * kq + col, i = lane, lane + 32, ...
                while (col >= kq) { col -= kq; ++row; }
                for (int i = lane; i < rows * kq; i += 32) {
                    if ((okmask >> row) & 1u) prm_store_packed(out, base + i, src[i]);
                    col += 32;
                    while (col >= kq) { col -= kq; ++row; }
                }
            }
        }
    }
    if (myflag && a.flags) atomicOr(a.flags, myflag);
}

// K = list length of knn.cu's dispatch (entries asked for, rounded up to an instantiated size),
// k = the entries asked for = candidates per row + the row itself.
int32_t sbp_prm_fused_launch(sbp_ctx *ctx, const CellGridView &grid, int K, int k, int r_init, int limit,
                             int64_t p0, int64_t count, double *seed2, uint8_t *done, int32_t *pending,
                             PrmFusedArgs *args) {
    if (count <= 0) return SBP_OK;
    const int kq = k - 1;
    // fused form (default): the candidates never leave the kNN kernel (k_prm_knn<., true>); coordinates
    // of a walked edge travel as 16-bit pixels
    const bool fuse = ctx->tune.prm_fuse && args->tiled.W <= 65536 && args->tiled.H <= 65536;
    void *cbuf = nullptr;
    int32_t rc = fuse ? SBP_OK : sbp_ctx_scratch(ctx, 16, (size_t)count * (kq > 0 ? kq : 1) * sizeof(int32_t), &cbuf);
    if (rc) return rc;
    // Optionally (tune "prm_slices", default 1 = off) the rows go through the two kernels in SLICES,
    // the walk of slice i on a second stream beside the kNN of slice i + 1.  Measured on cfg5 and
    // rejected as the default: 0.59 / 0.61 / 0.63 / 0.66 / 0.74 ms per step with 1 / 2 / 3 / 4 / 8 slices
    // -- each kernel already fills the SMs, so the second stream only adds launch tails.
    // When the packed rows go to HOST memory the slices serve another purpose: the kernels write a
    // device staging buffer and every slice's block is copied out by the DMA engine (second stream)
    // while the next slices are computed -- 40 MB take 0.70 ms over PCIe against 0.5 ms of compute,
    // so the step costs the copy plus one slice instead of compute plus copy (zero-copy stores from
    // the walk kernel reached only ~45 GB/s and 1.14 ms for the call; staged: 0.8x ms).
    // The same staging serves the multicast mapping (several ranks): stores through the switch issued
    // by the walk kernel itself made it wait for NVLink (89 us for a rank's 125 000 rows at N = 8 where
    // the walk alone takes 34); instead the walk runs in slices into the staging buffer and a copy
    // kernel on the second stream pushes every finished slice through the mapping meanwhile.
    int slices = ctx->tune.prm_slices;
    const bool stage_host = args->host_out && args->out.packed && args->out.sorted_rows && ctx->tune.time_kernels == 0;
    const bool stage_mc = args->out.packed_mc && args->out.packed && args->out.sorted_rows &&
                          ctx->tune.time_kernels == 0 && ctx->tune.prm_mc_stage && count >= 32768 &&
                          (((p0 * kq) & 3) == 0) && ((reinterpret_cast<uintptr_t>(args->out.packed) & 15) == 0);
    if (stage_host) slices = 8;
    if (stage_mc) slices = fuse ? 1 : 4;
    if (fuse && !stage_host) slices = 1;
    if (slices < 1 || ctx->tune.time_kernels != 0 || (count < 65536 && !stage_mc)) slices = 1;
    if (slices > 16) slices = 16;
    if (slices > 1 || stage_host || stage_mc || (args->order_out && args->order_host)) {
        if (!ctx->aux_stream) SBP_CUDA(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
        while ((int)ctx->aux_events.size() < slices + 1) {
            cudaEvent_t e;
            SBP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->aux_events.push_back(e);
        }
    }
    if (args->order_out && args->order_count > 0) {
        // the permutation is known once the grid is built: out it goes before the kernels start
        if (args->order_host) {
            if (!ctx->aux_stream) SBP_CUDA(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
            if (ctx->aux_events.empty()) {
                cudaEvent_t e;
                SBP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                ctx->aux_events.push_back(e);
            }
            SBP_CUDA(cudaEventRecord(ctx->aux_events[0], ctx->stream));
            SBP_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_events[0], 0));
            SBP_CUDA(cudaMemcpyAsync(args->order_out + args->order_p0, grid.ids + args->order_p0,
                                     (size_t)args->order_count * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->aux_stream));
        } else if (!(fuse && args->order_p0 == p0 && args->order_count == count)) {
            int ob = (int)((args->order_count + 255) / 256);
            if (ob > ctx->sm_count * 8) ob = ctx->sm_count * 8;
            k_copy_order<<<ob, 256, 0, ctx->stream>>>(grid.ids, args->order_p0, args->order_count, args->order_out,
                                                      args->out.packed_mc);
            ctx->launches++;
        }
    }
    PrmFusedArgs wargs = *args;
    // (fused form: the kernel's rows are the rows of the permutation asked for, it writes them itself)
    wargs.order_in_kernel = fuse && args->order_out && args->order_count > 0 && !args->order_host &&
                            args->order_p0 == p0 && args->order_count == count;
    wargs.order_mc = args->out.packed_mc;
    int32_t *host_packed = nullptr;
    if (stage_host || stage_mc) {
        void *stg = nullptr;
        if ((rc = sbp_ctx_scratch(ctx, 19, (size_t)grid.n * kq * sizeof(int32_t), &stg))) return rc;
        host_packed = args->out.packed;                       // (stage_mc: the multicast address)
        wargs.out.packed = (int32_t *)stg;                    // [n][kq] like the caller's buffer
        wargs.out.packed_mc = 0;
    }
    if (kq > 20) { sbp_set_error("sbp_prm_fused_launch: %d candidates per row not instantiated", kq); return SBP_ERR_ARG; }
    const int64_t per = ((count + slices - 1) / slices + PRM_TPB - 1) / PRM_TPB * PRM_TPB;   // whole CTAs (and warps)
    if (fuse) {
        for (int sl = 0; sl < slices; ++sl) {
            const int64_t off = (int64_t)sl * per, cnt = count - off < per ? count - off : per;
            if (cnt <= 0) break;
            const unsigned kblocks = (unsigned)((cnt + PRM_TPB - 1) / PRM_TPB);
#define PRM_KNN(KK)                                                                                            \
    k_prm_knn<KK, true><<<kblocks, PRM_TPB, 0, ctx->stream>>>(grid, kq, r_init, limit, p0 + off, cnt, seed2, done, \
                                                              pending, nullptr, wargs)
            {
                KernelTimer timer(ctx, ctx->tune.time_kernels == 3 ? 3 : 2);
                if (kq <= 4) PRM_KNN(4);
                else if (kq <= 8) PRM_KNN(8);
                else if (kq <= 10) PRM_KNN(10);
                else if (kq <= 16) PRM_KNN(16);
                else PRM_KNN(20);
            }
#undef PRM_KNN
            ctx->launches += 1;
            const int64_t r0 = (p0 + off) * kq;
            if (stage_host) {
                // this slice's block of rows: device staging -> the caller's host buffer, on the copy stream
                SBP_CUDA(cudaEventRecord(ctx->aux_events[sl], ctx->stream));
                SBP_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_events[sl], 0));
                SBP_CUDA(cudaMemcpyAsync(host_packed + r0, wargs.out.packed + r0, (size_t)cnt * kq * sizeof(int32_t),
                                         cudaMemcpyDeviceToHost, ctx->aux_stream));
            } else if (stage_mc) {
                k_mc_copy<<<ctx->sm_count, 256, 0, ctx->stream>>>(wargs.out.packed + r0, host_packed + r0, cnt * kq);
                ctx->launches++;
            }
        }
        if (stage_host || (args->order_out && args->order_host)) {   // join: everything behind sees all rows
            SBP_CUDA(cudaEventRecord(ctx->aux_events[slices], ctx->aux_stream));
            SBP_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->aux_events[slices], 0));
        }
        int ab = (int)((grid.n + 255) / 256);
        if (ab > ctx->sm_count * 8) ab = ctx->sm_count * 8;
        k_rows_aos<<<ab, 256, 0, ctx->stream>>>(args->soa, args->capacity, grid.n, args->X, pending);
        ctx->launches += 1;
        SBP_CUDA(cudaGetLastError());
        return SBP_OK;
    }
    for (int sl = 0; sl < slices; ++sl) {
        const int64_t off = (int64_t)sl * per, cnt = count - off < per ? count - off : per;
        if (cnt <= 0) break;
        const unsigned blocks = (unsigned)((cnt + PRM_TPB - 1) / PRM_TPB);
        int32_t *cand = (int32_t *)cbuf + off * kq;
        // (multicast staging: the kNN of ALL the rank's rows in one launch -- a rank's share is too
        // small to be cut further --, only the walk is sliced)
        const int64_t koff = stage_mc ? 0 : off, kcnt = stage_mc ? count : cnt;
        const unsigned kblocks = (unsigned)((kcnt + PRM_TPB - 1) / PRM_TPB);
        int32_t *kcand = (int32_t *)cbuf + koff * kq;
#define PRM_KNN(KK)                                                                                          \
    k_prm_knn<KK, false><<<kblocks, PRM_TPB, 0, ctx->stream>>>(grid, kq, r_init, limit, p0 + koff, kcnt, seed2, done, pending, kcand, wargs)
        if (!stage_mc || sl == 0) {
            KernelTimer timer(ctx, 2);
            if (kq <= 4) PRM_KNN(4);
            else if (kq <= 8) PRM_KNN(8);
            else if (kq <= 10) PRM_KNN(10);
            else if (kq <= 16) PRM_KNN(16);
            else PRM_KNN(20);
        }
#undef PRM_KNN
        cudaStream_t ws = ctx->stream;
        if (slices > 1 && !stage_host && !stage_mc) {
            SBP_CUDA(cudaEventRecord(ctx->aux_events[sl], ctx->stream));
            SBP_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_events[sl], 0));
            ws = ctx->aux_stream;
        }
        {
            KernelTimer timer(ctx, 3);
            if (ctx->tune.prm_walk_sorted && args->tiled.W <= 65536 && args->tiled.H <= 65536) {   // (16-bit pixel coordinates)
                const size_t sm = (size_t)kq * PW_ROWS * (sizeof(int32_t) + sizeof(PwEdge) + sizeof(uint16_t));
                k_prm_walk_sorted<<<(unsigned)((cnt + PW_ROWS - 1) / PW_ROWS), PW_ROWS, sm, ws>>>(grid, kq, p0 + off, cnt,
                                                                                                  done, cand, wargs);
            } else {
                k_prm_walk<<<blocks, PRM_TPB, 0, ws>>>(grid, kq, p0 + off, cnt, done, cand, wargs);
            }
        }
        if (stage_host || stage_mc) {
            // this slice's block of rows: device staging -> the caller's host buffer (DMA engine) or the
            // ranks' buffers (multicast stores), on the second stream
            SBP_CUDA(cudaEventRecord(ctx->aux_events[sl], ctx->stream));
            SBP_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_events[sl], 0));
            const int64_t r0 = (p0 + off) * kq;
            if (stage_host) {
                SBP_CUDA(cudaMemcpyAsync(host_packed + r0, wargs.out.packed + r0, (size_t)cnt * kq * sizeof(int32_t),
                                         cudaMemcpyDeviceToHost, ctx->aux_stream));
            } else {
                k_mc_copy<<<ctx->sm_count, 256, 0, ctx->aux_stream>>>(wargs.out.packed + r0, host_packed + r0, cnt * kq);
                ctx->launches++;
            }
        }
        ctx->launches += 2;
    }
    (void)K;
    if (slices > 1 || stage_host || stage_mc || (args->order_out && args->order_host)) {   // join: everything behind sees all rows
        SBP_CUDA(cudaEventRecord(ctx->aux_events[slices], ctx->aux_stream));
        SBP_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->aux_events[slices], 0));
    }
    // the exact path behind reads the row poses as [n][2]: converted only if it has work
    int ab = (int)((grid.n + 255) / 256);
    if (ab > ctx->sm_count * 8) ab = ctx->sm_count * 8;
    k_rows_aos<<<ab, 256, 0, ctx->stream>>>(args->soa, args->capacity, grid.n, args->X, pending);
    ctx->launches += 1;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

// ------------------------------ rows the fused kernel left to the exact path ----
// One thread per row with done[v] == 0 (all rows when `done` is NULL): the candidate list comes
// from the kNN rows idx [n][K] (node order) of knn.cu's exact pipeline.
__global__ void __launch_bounds__(PRM_TPB)
k_prm_visible2d(MapView tiled, TileCodeView af, const double *__restrict__ soa,
                int64_t capacity, int64_t n, const int64_t *__restrict__ idx, int K,
                const int32_t *__restrict__ order, int64_t p0, int64_t count, const uint8_t *__restrict__ done,
                const int32_t *__restrict__ pending, PrmOut out, int32_t *__restrict__ flags) {
    if (pending && *pending == 0) return;
    const int k = K - 1;
    int myflag = 0;
    for (int64_t slot = (int64_t)blockIdx.x * PRM_TPB + threadIdx.x; slot < count; slot += (int64_t)gridDim.x * PRM_TPB) {
    const int32_t v = order ? order[p0 + slot] : (int32_t)(p0 + slot);
    if (done && done[v]) continue;
    const double qx = soa[v], qy = soa[capacity + v];
    const int64_t *__restrict__ row = idx + (int64_t)v * K;
    int drop = K - 1;
    for (int c = K - 2; c >= 0; --c)
        if (row[c] == (int64_t)v) drop = c;
    const int64_t orow = out.sorted_rows ? p0 + slot : (int64_t)v;
#pragma unroll 1
    for (int col = 0; col < k; ++col) {
        const int32_t nb = (int32_t)row[col < drop ? col : col + 1];
        bool result = false;
        if (nb >= 0 && (int64_t)nb < n) result = prm_edge(tiled, af, soa[nb], soa[capacity + nb], qx, qy, myflag);
        const int32_t code = ((nb + 1) << 1) | (result ? 1 : 0);
        const int64_t o = orow * k + col;
        if (out.packed) prm_store_packed(out, o, code);
        if (out.nbr) out.nbr[o] = (int64_t)nb;
        if (out.vis) out.vis[o] = (uint8_t)(code & 1);
    }
    }
    if (myflag && flags) atomicOr(flags, myflag);
}

__global__ void k_rows_aos(const double *__restrict__ soa, int64_t capacity, int64_t n, double2 *__restrict__ X,
                           const int32_t *__restrict__ pending) {
    if (pending && *pending == 0) return;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        X[t] = make_double2(soa[t], soa[capacity + t]);
}

// order_out[p] = node of cell-sorted position p, for p in [p0, p0 + count) (through the multicast
// mapping when the rows leave that way: the other ranks' slices arrive the same way)
__global__ void k_copy_order(const int32_t *__restrict__ ids, int64_t p0, int64_t count, int32_t *__restrict__ order_out,
                             int multicast) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x) {
        const int32_t v = ids ? ids[p0 + t] : (int32_t)(p0 + t);
        if (multicast)
            asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(order_out + p0 + t),
                         "f"(__int_as_float(v)) : "memory");
        else
            order_out[p0 + t] = v;
    }
}

// Rows = stored nodes.  The rows are split into `parts` contiguous runs of the CELL-SORTED
// order (spatially compact; `part` in [0, parts)): the multi-GPU shard of a rank.  Only the rows of
// this part are written.
static int32_t prm_connect(sbp_map *map, sbp_nodes *nodes, int32_t k, int32_t part, int32_t parts,
                           int32_t sorted_rows, int64_t *nbr_out, uint8_t *vis, int32_t *packed,
                           int32_t packed_multicast, int32_t *order_out, int32_t *flags) {
    SBP_REQUIRE(map && nodes, "sbp_prm_connect2d: NULL argument");
    SBP_REQUIRE(nbr_out || vis || packed, "sbp_prm_connect2d: no output requested");
    SBP_REQUIRE(nodes->d == 2, "sbp_prm_connect2d: the 2-D checker needs a 2-D node store");
    SBP_REQUIRE(map->ctx == nodes->ctx, "sbp_prm_connect2d: map and node store live in different contexts");
    SBP_REQUIRE(k >= 1 && k <= PRM_KMAX, "sbp_prm_connect2d: 1 <= k <= 31");
    SBP_REQUIRE(parts >= 1 && part >= 0 && part < parts, "sbp_prm_connect2d: bad part");
    SBP_REQUIRE(!packed_multicast || packed, "sbp_prm_connect2d: multicast needs the packed output");
    SBP_REQUIRE(!sorted_rows || order_out, "sbp_prm_connect2d_sorted: order_out is NULL");
    sbp_ctx *ctx = map->ctx;
    SBP_DEVICE(ctx);
    const int64_t n = nodes->n;
    if (n == 0) return SBP_OK;
    int32_t rc = sbp_map_ensure_allfree(map);
    if (rc) return rc;
    const int K = k + 1;
    void *xbuf = nullptr, *ibuf = nullptr;
    if ((rc = sbp_ctx_scratch(ctx, 12, (size_t)n * 16, &xbuf))) return rc;
    if ((rc = sbp_ctx_scratch(ctx, 13, (size_t)n * K * sizeof(int64_t), &ibuf))) return rc;
    int blocks = (int)((n + 255) / 256);
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    const bool will_fuse = sbp_knn_rows_use_grid(nodes) && K <= PRM_FUSED_KMAX;
    if (!will_fuse) {
        k_rows_aos<<<blocks, 256, 0, ctx->stream>>>(nodes->d_soa, nodes->capacity, n, (double2 *)xbuf, nullptr);
        ctx->launches++;
    }
    // this part's run of the cell-sorted order (balanced like distributed.shard_range)
    const int64_t base = n / parts, rem = n % parts;
    PrmFusedArgs fa{map->view, sbp_map_codes(map),
                    PrmOut{nbr_out, vis, packed, packed_multicast, sorted_rows}, flags,
                    nodes->d_soa, nodes->capacity, (double2 *)xbuf, 0, nullptr, 0, 0, 0, 0, 0};
    if (packed && !packed_multicast) {
        cudaPointerAttributes attr{};
        if (cudaPointerGetAttributes(&attr, packed) == cudaSuccess && attr.type == cudaMemoryTypeHost) fa.host_out = 1;
        else cudaGetLastError();
    }
    KnnRowsPart rp;
    rp.p0 = (int64_t)part * base + (part < rem ? part : rem);
    rp.count = base + (part < rem ? 1 : 0);
    rp.deterministic = parts > 1 || sorted_rows;
    rp.fused = &fa;
    rp.banded = parts > 1;
    if (order_out && rp.count > 0) {
        // (several parts: a rank's grid -- and so its permutation -- covers its own rows only)
        fa.order_out = order_out;
        fa.order_p0 = parts > 1 ? rp.p0 : 0;
        fa.order_count = parts > 1 ? rp.count : n;
        cudaPointerAttributes attr{};
        if (!packed_multicast && cudaPointerGetAttributes(&attr, order_out) == cudaSuccess && attr.type == cudaMemoryTypeHost)
            fa.order_host = 1;
        else cudaGetLastError();
    }
    if ((rc = sbp_knn_rows_part(nodes, (const double *)xbuf, K, (int64_t *)ibuf, &rp))) return rc;
    if (order_out && rp.count > 0 && !rp.fused_ran) {
        // (several parts: a rank's grid -- and so its permutation -- covers its own rows only)
        const int64_t o0 = parts > 1 ? rp.p0 : 0, oc = parts > 1 ? rp.count : n;
        int ob = (int)((oc + 255) / 256);
        if (ob > ctx->sm_count * 8) ob = ctx->sm_count * 8;
        k_copy_order<<<ob, 256, 0, ctx->stream>>>(rp.ids, o0, oc, order_out, packed_multicast);
        ctx->launches++;
    }
    if (rp.count == 0) return SBP_OK;
    {
        KernelTimer timer(ctx, rp.fused_ran ? -1 : 3);
        const int64_t vb = (rp.count + PRM_TPB - 1) / PRM_TPB, vcap = (int64_t)ctx->sm_count * 16;
        // (after the fused kernels this one only has the rows they left: a capped grid that loops)
        k_prm_visible2d<<<(unsigned)(rp.fused_ran && vb > vcap ? vcap : vb), PRM_TPB, 0, ctx->stream>>>(
            map->view, sbp_map_codes(map), nodes->d_soa, nodes->capacity, n, (const int64_t *)ibuf, K, rp.ids,
            rp.p0, rp.count, rp.fused_ran ? rp.done : nullptr, rp.fused_ran ? rp.pending : nullptr, fa.out, flags);
    }
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

extern "C" int32_t sbp_prm_connect2d(sbp_map *map, sbp_nodes *nodes, int32_t k, int32_t part, int32_t parts,
                                     int64_t *nbr_out, uint8_t *vis, int32_t *packed,
                                     int32_t packed_multicast, int32_t *flags) {
    SBP_NVTX("sbp_prm_connect2d");
    return prm_connect(map, nodes, k, part, parts, 0, nbr_out, vis, packed, packed_multicast, nullptr, flags);
}

extern "C" int32_t sbp_prm_connect2d_sorted(sbp_map *map, sbp_nodes *nodes, int32_t k, int32_t part,
                                            int32_t parts, int64_t *nbr_out, uint8_t *vis, int32_t *packed,
                                            int32_t packed_multicast, int32_t *order_out, int32_t *flags) {
    SBP_NVTX("sbp_prm_connect2d_sorted");
    return prm_connect(map, nodes, k, part, parts, 1, nbr_out, vis, packed, packed_multicast, order_out, flags);
}
