// collide2d.cu -- K1 point feasibility, K2 segment-vs-grid visibility (integer
// Bresenham), K2b first-blocked-pixel, for the 2-D image-space checker.
//
// Reference behaviour restated by these kernels (file:line under
// /root/reference/sbp_env/):
//   ImgCollisionChecker.feasible                     collisionChecker.py:147-153
//   ImgCollisionChecker.visible                      collisionChecker.py:134-145
//   ImgCollisionChecker.get_line (Bresenham)         collisionChecker.py:155-215
//   ImgCollisionChecker.get_coor_before_collision    collisionChecker.py:118-132
//
// Bresenham without the serial loop.  After the reference's steep swap (:179-184)
// and order swap (:187-191) the i-th pixel of the canonical traversal is
//     a_i = a1 + i,   b_i = b1 + ystep * floor((i*|db| + c) / da),
//     c = da - 1 - floor(da / 2),  da = a2 - a1 >= |db|
// (error starts at floor(da/2) (:198), loses |db| per step and gains da whenever
// it went negative (:207-210), so it stays in [0, da) and the number of minor
// steps before pixel i is the smallest k with floor(da/2) - i*|db| + k*da >= 0).
// A group of G lanes owns one edge; lane j tests pixels j, j+G, ... and carries
// (k, rem) = divmod(i*|db| + c, da) forward by adding divmod(G*|db|, da), so the
// inner loop has no division.  A ballot per step gives the early exit.
#include "common.cuh"
#include "device_fns.cuh"

// ---------------------------------------------------------------- K1 -------
template <bool SMEM>
__global__ void __launch_bounds__(1024)
k_feasible2d(MapView mv, const double2 *__restrict__ P, int64_t n, uint8_t *__restrict__ out,
             int32_t *__restrict__ flags) {
    extern __shared__ __align__(16) uint32_t smap[];
    __shared__ uint64_t bar;
    if (SMEM) stage_map_to_smem(smap, mv, &bar);
    const uint32_t *words = SMEM ? smap : mv.words;
    int myflag = 0;
    // a pure stream (16 B in, 1 B out per point): 4 independent loads in flight per thread
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
        double2 p[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * stride;
            p[u] = i < n ? P[i] : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * stride;
            if (i >= n) break;
            TruncResult tx = trunc_index(p[u].x, mv.W), ty = trunc_index(p[u].y, mv.H);
            int f = tx.flag ? tx.flag : ty.flag;
            // numpy quirk pinned by tests/golden/img_cc.npz: an index in [2^63, 2^64)
            // raises OverflowError instead of IndexError
            if (!f && ((p[u].x >= 9223372036854775808.0 && p[u].x < 18446744073709551616.0) ||
                       (p[u].y >= 9223372036854775808.0 && p[u].y < 18446744073709551616.0)))
                f = SBP_FLAG_OVERFLOW;
            myflag |= f;
            bool ok = false;
            if (!f && tx.in_range && ty.in_range)
                ok = map_free<!SMEM>(words, mv.SX, wrap_index(tx.i, mv.W), wrap_index(ty.i, mv.H));
            out[i] = ok ? 1 : 0;
        }
    }
    if (myflag && flags) atomicOr(flags, myflag);
}

// ---------------------------------------------------------------- K2 -------
template <int G, bool SMEM, bool COUNT>
__global__ void __launch_bounds__(1024)
k_visible2d(MapView mv, const double2 *__restrict__ P, const double2 *__restrict__ Q, int64_t n,
            uint8_t *__restrict__ out, int32_t *__restrict__ flags,
            unsigned long long *__restrict__ px_tested) {
    extern __shared__ __align__(16) uint32_t smap[];
    __shared__ uint64_t bar;
    if (SMEM) stage_map_to_smem(smap, mv, &bar);
    const uint32_t *words = SMEM ? smap : mv.words;

    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int j = lane & (G - 1);               // lane within the edge group
    const int gshift = lane & ~(G - 1);         // first lane of my group
    const unsigned gmask_bits = G == 32 ? FULL : ((1u << G) - 1u);
    const int64_t groups_per_grid = ((int64_t)gridDim.x * blockDim.x) / G;
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    // all lanes of a warp run the same number of outer iterations
    const int64_t n_round = ((n + (32 / G) - 1) / (32 / G)) * (32 / G);
    int myflag = 0;
    unsigned long long tested = 0;

    for (int64_t eidx = gid; eidx < n_round; eidx += groups_per_grid) {
        const bool valid = eidx < n;
        double2 p = make_double2(0.0, 0.0), q = make_double2(0.0, 0.0);
        if (valid) { p = P[eidx]; q = Q[eidx]; }
        Edge2D e = make_edge(p.x, p.y, q.x, q.y, mv.W, mv.H);
        myflag |= valid ? e.flag : 0;
        bool result = valid && e.walk;          // stays true until a blocked pixel shows up
        const uint32_t npx = e.da + 1u;         // max(|dx|,|dy|) + 1 pixels (:204)
        // lane j starts at pixel j:  (k, rem) = divmod(j*|db| + c, da)
        uint32_t k = 0, rem = 0, sq = 0, sr = 0;
        if (e.da != 0u) {
            uint32_t c = e.da - 1u - (e.da >> 1);
            uint32_t u = (uint32_t)j * e.adb + c;
            k = u / e.da; rem = u - k * e.da;
            uint32_t s = (uint32_t)G * e.adb;
            sq = s / e.da; sr = s - sq * e.da;
        }
        uint32_t i = (uint32_t)j;
        bool done = !result;
        while (true) {
            bool act = !done && i < npx;
            bool blocked = false;
            if (act) {
                int a = e.a1 + (int)i;
                int b = e.b1 + e.ystep * (int)k;
                int x = e.steep ? b : a, y = e.steep ? a : b;       // :205
                x = wrap_index(x, mv.W); y = wrap_index(y, mv.H);
                blocked = !map_free<!SMEM>(words, mv.SX, x, y);     // :141, :151
            }
            unsigned bal = __ballot_sync(FULL, blocked);
            if (COUNT) tested += act ? 1ull : 0ull;
            if ((bal >> gshift) & gmask_bits) { result = false; done = true; }
            i += (uint32_t)G;
            k += sq; rem += sr;
            if (rem >= e.da && e.da != 0u) { rem -= e.da; ++k; }
            if (i >= npx) done = true;
            if (!__any_sync(FULL, !done)) break;
        }
        if (valid && j == 0) out[eidx] = result ? 1 : 0;
    }
    if (myflag && flags) atomicOr(flags, myflag);
    if (COUNT) {
        for (int o = 16; o > 0; o >>= 1) tested += __shfl_down_sync(FULL, tested, o);
        if (lane == 0 && tested && px_tested) atomicAdd(px_tested, tested);
    }
}

// ------------------------------------------------------ K2 (adaptive) -------
// Lane-per-edge first, warp-per-edge for the stragglers.
//   phase A  every lane owns one edge of a 32-edge batch: coalesced endpoint loads,
//            truncation / range check / canonicalisation once per edge (not once per
//            lane group), the END pixel tested first (a blocked endpoint is the most
//            likely early exit), then the classic error-accumulating walk from the
//            canonical start -- no division, ~16 instructions per pixel per lane.
//   phase B  as soon as fewer than COOP_BELOW lanes are still walking (the rest have
//            finished or hit an obstacle) the survivors are finished one at a time by
//            the whole warp: their walk state is broadcast with shuffles and lane j
//            tests pixel j, j+32, ... of the remainder via the closed form
//            m_j = floor((j*|db| + da-1-err) / da), ballot for the early exit.
// Batches of equal-length edges (PRM k-NN candidates) therefore run entirely in the
// divergence-free phase A; heavy-tailed batches (uniform random edges) hand their few
// long edges to phase B instead of idling 31 lanes behind them.
template <bool SMEM, bool ROWS>
__device__ __forceinline__ bool px_free(const uint32_t *words, const MapView &mv, int x, int y) {
    return ROWS ? map_free_rows<!SMEM>(words, mv.SX, x, y) : map_free<!SMEM>(words, mv.SX, x, y);
}

// ROWS: `mv` is the row-major copy of the map (sbp_map::rows; shared-memory resident maps:
// 3 integer operations per pixel index instead of ~12 for the tiled layout, which only pays off
// through its sector locality when the map lives in L2 / HBM).
// Where the edges come from: two endpoint arrays, or the rows of a kNN answer resolved against
// the node store on the fly (fused edge gather: no P / Q arrays are materialised).
struct EdgeArrays {
    const double2 *P, *Q;
    __device__ __forceinline__ bool load(int64_t e, double2 &p, double2 &q, int32_t &tag) const {
        p = P[e]; q = Q[e]; tag = 0;
        return true;
    }
    __device__ __forceinline__ void store(int64_t, int32_t, bool) const {}
    __device__ __forceinline__ void finish(int32_t *) const {}
};

// Candidate edge e = (row i, column j) of a kNN answer nbr_in [m][k_in] (prmPlanner.py:91-101):
// the row's own entry is dropped -- the column equal to the row index if present, else the last
// column -- and the edge is visible(m_g.pos, v.pos) = (neighbour pose -> row pose).  X != NULL:
// row i is the external query X[i] and no column is dropped.  Writes the neighbour index.
struct EdgeKnnRows {
    const double *soa;
    int64_t capacity, n_nodes;
    const int64_t *nbr_in;
    int k_in, k_out;
    int64_t first_row;
    const double2 *X;
    int64_t *nbr_out;
    // optional packed adjacency (neighbour + 1) * 2 + visible, one int32 per edge: a local buffer,
    // or (packed_mc) the NVSwitch multicast address of the ranks' symmetric buffer -- the result
    // is then replicated into every rank's copy by the switch as the edges are decided, so the
    // multi-GPU exchange of the adjacency overlaps the visibility walk (distributed.py)
    int32_t *packed;
    int packed_mc;
    // optional rendezvous of the ranks inside this kernel (instead of a barrier launch after it):
    // sig_mc / sig_local = multicast and local address of a symmetric int32 [world] signal array,
    // ticket / epoch = two local counters.  The last CTA of every rank publishes its epoch through
    // the multicast mapping and waits until every rank's entry has reached it.
    int32_t *sig_mc;
    const int32_t *sig_local;
    int32_t *ticket, *epoch;
    int rank, world;
    __device__ __forceinline__ void finish(int32_t *flags) const {
        if (!sig_mc) return;
        __shared__ int s_last, s_epoch;
        __threadfence_system();                               // this thread's multicast stores
        __syncthreads();
        if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1) == (int)gridDim.x - 1 ? 1 : 0;
        __syncthreads();
        if (!s_last) return;
        if (threadIdx.x == 0) {
            *ticket = 0;
            const int32_t ep = *epoch + 1;
            *epoch = ep;
            s_epoch = ep;
            __threadfence_system();
            asm volatile("multimem.st.release.sys.global.f32 [%0], %1;" ::"l"(sig_mc + rank),
                         "f"(__int_as_float(ep)) : "memory");
        }
        __syncthreads();
        if ((int)threadIdx.x < world) {
            const int32_t ep = s_epoch;
            int32_t seen;
            long long spins = 0;
            do {
                asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(seen) : "l"(sig_local + threadIdx.x) : "memory");
            } while (seen - ep < 0 && ++spins < 4000000ll);
            if (seen - ep < 0 && flags) atomicOr(flags, SBP_FLAG_PEER_TIMEOUT);   // never hang the GPU
        }
        __syncthreads();
    }
    __device__ __forceinline__ void store(int64_t e, int32_t tag, bool vis) const {
        if (!packed) return;
        const int32_t code = ((tag + 1) << 1) | (vis ? 1 : 0);
        if (packed_mc)
            asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(packed + e),
                         "f"(__int_as_float(code)) : "memory");
        else
            packed[e] = code;
    }
    __device__ __forceinline__ bool load(int64_t e, double2 &p, double2 &q, int32_t &tag) const {
        const int64_t i = e / k_out;
        const int j = (int)(e - i * k_out);
        const int64_t *row = nbr_in + i * k_in;
        int col = j;
        if (X) {
            q = X[i];
        } else {
            const int64_t self = first_row + i;
            int drop = k_in - 1;
            for (int c = 0; c < k_in; ++c)
                if (row[c] == self) { drop = c; break; }
            col = j < drop ? j : j + 1;
            q = make_double2(soa[self], soa[capacity + self]);
        }
        const int64_t nb = row[col];
        nbr_out[e] = nb;
        tag = (int32_t)nb;
        const bool ok = nb >= 0 && nb < n_nodes;
        p = ok ? make_double2(soa[nb], soa[capacity + nb]) : q;
        return ok;
    }
};

template <bool SMEM, bool COUNT, bool ROWS, class Edges = EdgeArrays>
__global__ void __launch_bounds__(1024)
k_visible2d_adaptive(MapView mv, Edges edges,
                     int64_t n, uint8_t *__restrict__ out, int32_t *__restrict__ flags,
                     unsigned long long *__restrict__ px_tested, int COOP_BELOW) {
    extern __shared__ __align__(16) uint32_t smap[];
    __shared__ uint64_t bar;
    if (SMEM) stage_map_to_smem(smap, mv, &bar);
    const uint32_t *words = SMEM ? smap : mv.words;

    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_round = (n + 31) & ~(int64_t)31;       // whole warps iterate together
    int myflag = 0;
    unsigned long long tested = 0;

    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double2 p = make_double2(0.0, 0.0), q = make_double2(0.0, 0.0);
    bool have = false;                                      // the edge exists (neighbour slot filled)
    int32_t tag = 0;                                        // what store() needs besides the result
    if (e < n) have = edges.load(e, p, q, tag);
    for (; e < n_round; e += stride) {
        const bool valid = e < n;
        const bool exists = have;
        const int32_t ctag = tag;
        const double2 cp = p, cq = q;
        const int64_t en = e + stride;                      // prefetch the next batch
        if (en < n) have = edges.load(en, p, q, tag);

        Edge2D ed = make_edge(cp.x, cp.y, cq.x, cq.y, mv.W, mv.H);
        if (valid) myflag |= ed.flag;
        bool result = valid && ed.walk;
        // end pixel first (:204-206: it is a member of the list)
        if (result) {
            result = px_free<SMEM, ROWS>(words, mv, wrap_index(ed.x2, mv.W), wrap_index(ed.y2, mv.H));
            if (COUNT) tested += 1ull;
        }
        int a = ed.a1, b = ed.b1;
        const int a2 = ed.a1 + (int)ed.da;
        int err = (int)(ed.da >> 1);                        // :198
        bool active = result;
        unsigned bal;
        // ---- phase A ---------------------------------------------------------
        // Hand-over rule: finishing sequentially costs ~18 instructions per step for
        // the longest remaining walk; finishing the walkers one by one with the whole
        // warp costs ~45 + 25*ceil(L/32) each.  The break-even number of walkers grows
        // with the remaining length (measured: 1 at 4 px, ~5 at 16 px, 12-16 from
        // 64 px on), approximated by L/4 capped at COOP_BELOW; re-evaluated every 4
        // steps with one REDUX.
        while (true) {
            bal = __ballot_sync(FULL, active);
            const int nact = __popc(bal);
            if (nact == 0) break;
            const int lmax = __reduce_max_sync(FULL, active ? a2 - a + 1 : 0);
            int thr = lmax >> 2;
            thr = thr < 1 ? 1 : thr > COOP_BELOW ? COOP_BELOW : thr;
            if (nact < thr) break;
#pragma unroll
            for (int st = 0; st < 4; ++st) {
                if (active) {
                    int x = ed.steep ? b : a, y = ed.steep ? a : b;          // :205
                    bool fr = px_free<SMEM, ROWS>(words, mv, wrap_index(x, mv.W), wrap_index(y, mv.H));
                    if (COUNT) tested += 1ull;
                    err -= (int)ed.adb;                                      // :207-210
                    if (err < 0) { b += ed.ystep; err += (int)ed.da; }
                    ++a;
                    if (!fr) { result = false; active = false; }
                    else if (a > a2) active = false;
                }
            }
        }
        // ---- phase B ---------------------------------------------------------
        while (bal) {
            const int src = __ffs(bal) - 1;
            bal &= bal - 1;
            const int s_a = __shfl_sync(FULL, a, src), s_b = __shfl_sync(FULL, b, src);
            const int s_err = __shfl_sync(FULL, err, src);
            const uint32_t s_da = __shfl_sync(FULL, ed.da, src), s_adb = __shfl_sync(FULL, ed.adb, src);
            const int s_pack = __shfl_sync(FULL, (ed.ystep > 0 ? 1 : 0) | (ed.steep ? 2 : 0), src);
            const int s_a2 = __shfl_sync(FULL, a2, src);
            const int ystep = (s_pack & 1) ? 1 : -1;
            const bool steep = (s_pack & 2) != 0;
            const uint32_t left = (uint32_t)(s_a2 - s_a + 1);           // pixels still to test (>= 1)
            // lane j: m = floor((j*adb + da-1-err) / da); stride 32: (sq, sr) = divmod(32*adb, da)
            uint32_t m = 0, rem = 0, sq = 0, sr = 0;
            if (s_da != 0u) {
                const double inv = 1.0 / (double)s_da;
                unsigned long long u = (unsigned long long)lane * s_adb + (s_da - 1u - (uint32_t)s_err);
                unsigned long long qq = (unsigned long long)((double)u * inv);
                long long r = (long long)u - (long long)qq * (long long)s_da;
                if (r < 0) { --qq; r += s_da; } else if (r >= (long long)s_da) { ++qq; r -= s_da; }
                m = (uint32_t)qq; rem = (uint32_t)r;
                uint32_t s32 = 32u * s_adb;
                sq = s32 / s_da; sr = s32 - sq * s_da;
            }
            bool hit = false;
            for (uint32_t base = 0; base < left; base += 32u) {
                const uint32_t j = base + (uint32_t)lane;
                bool blk = false;
                if (j < left) {
                    int aa = s_a + (int)j, bb = s_b + ystep * (int)m;
                    int x = steep ? bb : aa, y = steep ? aa : bb;
                    blk = !px_free<SMEM, ROWS>(words, mv, wrap_index(x, mv.W), wrap_index(y, mv.H));
                    if (COUNT) tested += 1ull;
                }
                if (__any_sync(FULL, blk)) { hit = true; break; }
                m += sq; rem += sr;
                if (rem >= s_da && s_da != 0u) { rem -= s_da; ++m; }
            }
            if (lane == src) result = !hit;
        }
        if (valid) {
            out[e] = (result && exists) ? 1 : 0;
            edges.store(e, ctag, result && exists);
        }
    }
    if (myflag && flags) atomicOr(flags, myflag);
    if (COUNT) {
        for (int o = 16; o > 0; o >>= 1) tested += __shfl_down_sync(FULL, tested, o);
        if (lane == 0 && tested && px_tested) atomicAdd(px_tested, tested);
    }
    edges.finish(flags);
}

// --------------------------------------------------------------- K2b -------
// First blocked pixel in start->end order (or the end pixel).  One warp per edge,
// lanes take consecutive ORIGINAL indices o; the canonical index is o or
// npx-1-o when the order swap happened (points.reverse(), :213-214).  Unlike
// `visible`, the walk must be followed across the edge of the index range (the
// reported pixel is the first one that is out of range or blocked, un-wrapped as
// get_line produced it), so endpoints are truncated to wide integers here;
// |coordinate| >= 2^30 is reported as unsupported (INT32_MIN sentinel) -- the
// reference would have to materialise a list of >= 2^30 tuples for such a call.
template <bool SMEM>
__global__ void __launch_bounds__(1024)
k_first_blocked2d(MapView mv, const double2 *__restrict__ P, const double2 *__restrict__ Q,
                  int64_t n, int32_t *__restrict__ out_xy, int32_t *__restrict__ flags) {
    extern __shared__ __align__(16) uint32_t smap[];
    __shared__ uint64_t bar;
    if (SMEM) stage_map_to_smem(smap, mv, &bar);
    const uint32_t *words = SMEM ? smap : mv.words;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double LIM = 1073741824.0;  // 2^30
    int myflag = 0;
    for (int64_t eidx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; eidx < n;
         eidx += warps) {
        double2 p = P[eidx], q = Q[eidx];
        double v[4] = {p.x, p.y, q.x, q.y};
        int flag = 0;
        bool wide_ok = true;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (!flag) flag = v[t] != v[t] ? SBP_FLAG_NAN : isinf(v[t]) ? SBP_FLAG_OVERFLOW : 0;
            wide_ok = wide_ok && (fabs(v[t]) < LIM);
        }
        myflag |= flag;
        int rx = INT32_MIN, ry = INT32_MIN;
        if (!flag && wide_ok) {
            int x1 = __double2int_rz(v[0]), y1 = __double2int_rz(v[1]);
            int x2 = __double2int_rz(v[2]), y2 = __double2int_rz(v[3]);
            int dx = x2 - x1, dy = y2 - y1;
            bool steep = abs(dy) > abs(dx);
            int a1 = steep ? y1 : x1, b1 = steep ? x1 : y1;
            int a2 = steep ? y2 : x2, b2 = steep ? x2 : y2;
            bool swapped = a1 > a2;
            if (swapped) { int t = a1; a1 = a2; a2 = t; t = b1; b1 = b2; b2 = t; }
            uint32_t da = (uint32_t)(a2 - a1), adb = (uint32_t)abs(b2 - b1);
            int ystep = b1 < b2 ? 1 : -1;
            const int64_t npx = (int64_t)da + 1;
            const uint32_t c = da ? da - 1u - (da >> 1) : 0u;
            rx = x2; ry = y2;                      // free line: endPos = last pixel (:127-132)
            for (int64_t base = 0; base < npx; base += 32) {
                int64_t o = base + lane;
                bool blocked = false;
                int x = 0, y = 0;
                if (o < npx) {
                    uint64_t ci = swapped ? (uint64_t)(npx - 1 - o) : (uint64_t)o;
                    uint32_t k = da ? (uint32_t)((ci * adb + c) / da) : 0u;
                    int a = a1 + (int)ci, b = b1 + ystep * (int)k;
                    x = steep ? b : a; y = steep ? a : b;
                    bool inr = x >= -mv.W && x < mv.W && y >= -mv.H && y < mv.H;
                    blocked = !inr || !map_free<!SMEM>(words, mv.SX, wrap_index(x, mv.W),
                                                       wrap_index(y, mv.H));
                }
                unsigned bal = __ballot_sync(FULL, blocked);
                if (bal) {
                    int src = __ffs(bal) - 1;
                    rx = __shfl_sync(FULL, x, src);
                    ry = __shfl_sync(FULL, y, src);
                    break;
                }
            }
        }
        if (lane == 0) { out_xy[2 * eidx] = rx; out_xy[2 * eidx + 1] = ry; }
    }
    if (myflag && flags) atomicOr(flags, myflag);
}

// ------------------------------------------------------------ launchers -----
struct LaunchCfg { int grid, threads; size_t smem; };

static LaunchCfg cfg_for(const sbp_map *m, bool smem_path) {
    LaunchCfg c;
    const sbp_ctx *ctx = m->ctx;
    if (smem_path) {
        c.smem = (size_t)m->packed_bytes;
        int per_sm = (int)((size_t)(228 * 1024) / (c.smem + 2048));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 4) per_sm = 4;
        c.threads = per_sm <= 2 ? 1024 : 512;
        c.grid = ctx->sm_count * per_sm;
    } else {
        c.smem = 0;
        c.threads = 256;
        c.grid = ctx->sm_count * 8;
    }
    return c;
}

template <typename K>
static int32_t set_smem(K kernel, size_t smem) {
    if (smem > 48 * 1024)
        SBP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return SBP_OK;
}

extern "C" int32_t sbp_feasible2d(sbp_map *map, const double *P, int64_t n, uint8_t *out,
                                  int32_t *flags) {
    SBP_NVTX("sbp_feasible2d");
    SBP_REQUIRE(map && (n == 0 || (P && out)), "sbp_feasible2d: NULL argument");
    SBP_REQUIRE(n >= 0, "sbp_feasible2d: n < 0");
    if (n == 0) return SBP_OK;
    sbp_ctx *ctx = map->ctx;
    SBP_DEVICE(ctx);
    // staging the map pays off only when the batch is large compared to the map
    bool smem = map->smem_resident && n * 16 >= map->packed_bytes * (int64_t)ctx->sm_count;
    LaunchCfg c = cfg_for(map, smem);
    int64_t need = (n + c.threads - 1) / c.threads;
    if (need < c.grid) c.grid = (int)need;
    if (smem) {
        int32_t rc = set_smem(k_feasible2d<true>, c.smem);
        if (rc) return rc;
        k_feasible2d<true><<<c.grid, c.threads, c.smem, ctx->stream>>>(map->view, (const double2 *)P, n, out, flags);
    } else {
        k_feasible2d<false><<<c.grid, c.threads, 0, ctx->stream>>>(map->view, (const double2 *)P, n, out, flags);
    }
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

// Sum and count of the event pairs recorded since the last call (synchronises the stream).
extern "C" int32_t sbp_ctx_kernel_time(sbp_ctx *ctx, double *total_ms, int64_t *launches) {
    SBP_REQUIRE(ctx && total_ms && launches, "sbp_ctx_kernel_time: NULL argument");
    SBP_DEVICE(ctx);
    SBP_CUDA(cudaStreamSynchronize(ctx->stream));
    double tot = 0.0;
    for (auto &pr : ctx->timed) {
        float ms = 0.f;
        SBP_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
        tot += ms;
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    *total_ms = tot;
    *launches = (int64_t)ctx->timed.size();
    ctx->timed.clear();
    return SBP_OK;
}

extern "C" int32_t sbp_ctx_tune(sbp_ctx *ctx, const char *key, int64_t value) {
    SBP_REQUIRE(ctx && key, "sbp_ctx_tune: NULL argument");
    sbp_tuning &t = ctx->tune;
    struct { const char *name; int *slot; } table[] = {
        {"visible_group", &t.visible_group}, {"coop_below", &t.coop_below},
        {"force_global", &t.force_global}, {"visible_tiles", &t.visible_tiles}, {"prm_slices", &t.prm_slices}, {"prm_walk_sorted", &t.prm_walk_sorted}, {"prm_mc_stage", &t.prm_mc_stage}, {"tile_rect", &t.tile_rect}, {"prm_fuse", &t.prm_fuse}, {"prm_build_ctas", &t.prm_build_ctas}, {"arm_group", &t.arm_group}, {"arm_flat", &t.arm_flat},
        {"near_filter", &t.near_filter}, {"knn_seed", &t.knn_seed}, {"knn_chunks", &t.knn_chunks},
        {"knn_prefilter", &t.knn_prefilter}, {"knn_filter", &t.knn_filter},
        {"knn_filter_chunks", &t.knn_filter_chunks}, {"knn_filter_variant", &t.knn_filter_variant},
        {"knn_cells", &t.knn_cells}, {"knn_cells_limit", &t.knn_cells_limit},
        {"knn_cells_qpw", &t.knn_cells_qpw}, {"knn_cells_rinit", &t.knn_cells_rinit},
        {"knn_cells_order", &t.knn_cells_order}, {"knn_cell_density", &t.knn_cell_density},
        {"time_kernels", &t.time_kernels}};
    for (auto &e : table)
        if (!strcmp(key, e.name)) { *e.slot = (int)value; return SBP_OK; }
    sbp_set_error("sbp_ctx_tune: unknown key %s", key);
    return SBP_ERR_ARG;
}

template <int G>
static int32_t launch_visible(sbp_map *map, const double *P, const double *Q, int64_t n,
                              uint8_t *out, int32_t *flags, unsigned long long *px, bool smem) {
    sbp_ctx *ctx = map->ctx;
    LaunchCfg c = cfg_for(map, smem);
    int64_t need = (n * G + c.threads - 1) / c.threads;
    if (need < c.grid) c.grid = (int)need;
    const double2 *p2 = (const double2 *)P, *q2 = (const double2 *)Q;
#define SBP_LAUNCH_VIS(SM, CT)                                                                   \
    do {                                                                                          \
        int32_t rc = set_smem(k_visible2d<G, SM, CT>, c.smem);                                    \
        if (rc) return rc;                                                                        \
        k_visible2d<G, SM, CT><<<c.grid, c.threads, c.smem, ctx->stream>>>(map->view, p2, q2, n,  \
                                                                            out, flags, px);      \
    } while (0)
    if (smem) { if (px) SBP_LAUNCH_VIS(true, true); else SBP_LAUNCH_VIS(true, false); }
    else      { if (px) SBP_LAUNCH_VIS(false, true); else SBP_LAUNCH_VIS(false, false); }
#undef SBP_LAUNCH_VIS
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

// Maps that are read from L2 / HBM (too large for shared memory): one thread per edge, two-level
// walk on the tile codes (device_fns.cuh:edge_free_tiles) -- an 8-pixel chunk of the line costs one
// or two cached 2-bit lookups instead of 8 pixel tests on a 128 MiB array.  Same booleans as the
// pixel walk for every map (the codes only summarise tiles that are uniformly free / blocked).
__global__ void __launch_bounds__(256)
k_visible2d_tiles(MapView mv, TileCodeView af, const double2 *__restrict__ P,
                  const double2 *__restrict__ Q, int64_t n, uint8_t *__restrict__ out, int32_t *__restrict__ flags) {
    int myflag = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double2 p = P[i], q = Q[i];
        const Edge2D ed = make_edge(p.x, p.y, q.x, q.y, mv.W, mv.H);
        myflag |= ed.flag;
        bool r = false;
        if (ed.walk) {
            if (ed.x1 >= 0 && ed.y1 >= 0 && ed.x2 >= 0 && ed.y2 >= 0)
                r = edge_free_tiles(mv.words, mv.SX, af, ed.a1, ed.b1, ed.da, ed.adb, ed.ystep, ed.steep);
            else {                                       // negative coordinates: numpy wrap-around
                int f2 = 0;
                r = edge_visible_seq<false>(mv.words, mv, p.x, p.y, q.x, q.y, f2);
            }
        }
        out[i] = r ? 1 : 0;
    }
    if (myflag && flags) atomicOr(flags, myflag);
}

extern "C" int32_t sbp_visible2d(sbp_map *map, const double *P, const double *Q, int64_t n,
                                 uint8_t *out, int32_t *flags, uint64_t *px_tested) {
    SBP_NVTX("sbp_visible2d");
    SBP_REQUIRE(map && (n == 0 || (P && Q && out)), "sbp_visible2d: NULL argument");
    SBP_REQUIRE(n >= 0, "sbp_visible2d: n < 0");
    if (n == 0) return SBP_OK;
    sbp_ctx *ctx = map->ctx;
    SBP_DEVICE(ctx);
    bool smem = !ctx->tune.force_global && map->smem_resident &&
                n * 32 >= map->packed_bytes * (int64_t)ctx->sm_count / 4;
    unsigned long long *px = (unsigned long long *)px_tested;
    if (!map->smem_resident && ctx->tune.visible_tiles && ctx->tune.visible_group == 0 && !px) {
        int32_t rc = sbp_map_ensure_allfree(map);
        if (rc) return rc;
        int64_t need = (n + 255) / 256;
        const int grid = (int)(need < (int64_t)ctx->sm_count * 32 ? need : (int64_t)ctx->sm_count * 32);
        KernelTimer timer(ctx, 3);
        k_visible2d_tiles<<<grid, 256, 0, ctx->stream>>>(map->view, sbp_map_codes(map), (const double2 *)P,
                                                         (const double2 *)Q, n, out, flags);
        ctx->launches++;
        SBP_CUDA(cudaGetLastError());
        return SBP_OK;
    }
    if (ctx->tune.visible_group == 0) {
        LaunchCfg c;
        if (smem) {                 // shared-memory resident: the row-major copy (cheaper index)
            int32_t rc = sbp_map_ensure_rows(map);
            if (rc) return rc;
            smem = map->rows_smem_resident;
        }
        const MapView mview = smem ? map->rows : map->view;
        c.smem = smem ? (size_t)map->rows.n_words * 4 : 0;
        c.threads = (smem && c.smem > 110 * 1024) ? 1024 : 512;
        c.grid = 0;
        int coop = ctx->tune.coop_below < 1 ? 1 : ctx->tune.coop_below > 33 ? 33 : ctx->tune.coop_below;
        const double2 *p2 = (const double2 *)P, *q2 = (const double2 *)Q;
#define SBP_LAUNCH_ADP(SM, CT)                                                                    \
    do {                                                                                           \
        int32_t rc = set_smem(k_visible2d_adaptive<SM, CT, SM>, c.smem);                           \
        if (rc) return rc;                                                                         \
        int per_sm = 0;                                                                            \
        SBP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(                                    \
            &per_sm, k_visible2d_adaptive<SM, CT, SM>, c.threads, c.smem));                        \
        c.grid = ctx->sm_count * (per_sm < 1 ? 1 : per_sm);                                        \
        int64_t need = (n + c.threads - 1) / c.threads;                                            \
        if (need < c.grid) c.grid = (int)need;                                                     \
        k_visible2d_adaptive<SM, CT, SM><<<c.grid, c.threads, c.smem, ctx->stream>>>(              \
            mview, EdgeArrays{p2, q2}, n, out, flags, px, coop);                                   \
    } while (0)
        {
            KernelTimer timer(ctx, 3);
            if (smem) { if (px) SBP_LAUNCH_ADP(true, true); else SBP_LAUNCH_ADP(true, false); }
            else      { if (px) SBP_LAUNCH_ADP(false, true); else SBP_LAUNCH_ADP(false, false); }
        }
#undef SBP_LAUNCH_ADP
        ctx->launches++;
        SBP_CUDA(cudaGetLastError());
        return SBP_OK;
    }
    switch (ctx->tune.visible_group) {
        case 4: return launch_visible<4>(map, P, Q, n, out, flags, px, smem);
        case 8: return launch_visible<8>(map, P, Q, n, out, flags, px, smem);
        case 16: return launch_visible<16>(map, P, Q, n, out, flags, px, smem);
        case 32: return launch_visible<32>(map, P, Q, n, out, flags, px, smem);
        default: sbp_set_error("visible_group must be 0, 4, 8, 16 or 32"); return SBP_ERR_ARG;
    }
}

// Fused form of sbp_nodes_knn_edges + sbp_visible2d for the 2-D checker: the candidate edges of
// a kNN answer are resolved against the node store inside the visibility kernel.
// nbr_out [m][k_out] and vis [m][k_out] (1 = the neighbour slot is filled and the edge is free).
extern "C" int32_t sbp_knn_edges_visible2d(sbp_map *map, sbp_nodes *nodes, const int64_t *nbr_in,
                                           int64_t m, int32_t k_in, int64_t first_row, const double *X,
                                           int64_t *nbr_out, uint8_t *vis, int32_t *packed,
                                           int32_t packed_multicast, const int64_t *rendezvous,
                                           int32_t *flags) {
    SBP_NVTX("sbp_knn_edges_visible2d");
    SBP_REQUIRE(map && nodes && (m == 0 || (nbr_in && nbr_out && vis)), "sbp_knn_edges_visible2d: NULL argument");
    SBP_REQUIRE(nodes->d == 2, "sbp_knn_edges_visible2d: the 2-D checker needs a 2-D node store");
    SBP_REQUIRE(map->ctx == nodes->ctx, "sbp_knn_edges_visible2d: map and node store live in different contexts");
    const int k_out = X ? k_in : k_in - 1;
    SBP_REQUIRE(m >= 0 && k_out >= 1, "sbp_knn_edges_visible2d: bad sizes");
    SBP_REQUIRE(X || (first_row >= 0 && first_row + m <= nodes->n), "sbp_knn_edges_visible2d: row range");
    if (m == 0) return SBP_OK;
    sbp_ctx *ctx = map->ctx;
    SBP_DEVICE(ctx);
    const int64_t n = m * k_out;
    bool smem = !ctx->tune.force_global && map->smem_resident &&
                n * 32 >= map->packed_bytes * (int64_t)ctx->sm_count / 4;
    if (smem) {
        int32_t rc = sbp_map_ensure_rows(map);
        if (rc) return rc;
        smem = map->rows_smem_resident;
    }
    const MapView mview = smem ? map->rows : map->view;
    const size_t sbytes = smem ? (size_t)map->rows.n_words * 4 : 0;
    const int threads = (smem && sbytes > 110 * 1024) ? 1024 : 512;
    const int coop = ctx->tune.coop_below < 1 ? 1 : ctx->tune.coop_below > 33 ? 33 : ctx->tune.coop_below;
    EdgeKnnRows src{nodes->d_soa, nodes->capacity, nodes->n, nbr_in, k_in, k_out, first_row,
                    (const double2 *)X, nbr_out, packed, packed_multicast,
                    nullptr, nullptr, nullptr, nullptr, 0, 0};
    if (rendezvous) {       // host array: sig_mc, sig_local, ticket, epoch (addresses), rank, world
        SBP_REQUIRE(packed && packed_multicast, "sbp_knn_edges_visible2d: rendezvous needs the multicast output");
        SBP_REQUIRE(rendezvous[0] && rendezvous[1] && rendezvous[2] && rendezvous[3] && rendezvous[5] >= 1 &&
                        rendezvous[5] <= 64 && rendezvous[4] >= 0 && rendezvous[4] < rendezvous[5],
                    "sbp_knn_edges_visible2d: bad rendezvous block");
        src.sig_mc = (int32_t *)(uintptr_t)rendezvous[0];
        src.sig_local = (const int32_t *)(uintptr_t)rendezvous[1];
        src.ticket = (int32_t *)(uintptr_t)rendezvous[2];
        src.epoch = (int32_t *)(uintptr_t)rendezvous[3];
        src.rank = (int)rendezvous[4];
        src.world = (int)rendezvous[5];
    }
#define SBP_LAUNCH_FUSED(SM)                                                                       \
    do {                                                                                           \
        int32_t rc = set_smem(k_visible2d_adaptive<SM, false, SM, EdgeKnnRows>, sbytes);           \
        if (rc) return rc;                                                                         \
        int per_sm = 0;                                                                            \
        SBP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(                                    \
            &per_sm, k_visible2d_adaptive<SM, false, SM, EdgeKnnRows>, threads, sbytes));          \
        int grid = ctx->sm_count * (per_sm < 1 ? 1 : per_sm);                                      \
        int64_t need = (n + threads - 1) / threads;                                                \
        if (need < grid) grid = (int)need;                                                         \
        k_visible2d_adaptive<SM, false, SM, EdgeKnnRows><<<grid, threads, sbytes, ctx->stream>>>(  \
            mview, src, n, vis, flags, nullptr, coop);                                             \
    } while (0)
    {
        KernelTimer timer(ctx, 3);
        if (smem) SBP_LAUNCH_FUSED(true); else SBP_LAUNCH_FUSED(false);
    }
#undef SBP_LAUNCH_FUSED
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

extern "C" int32_t sbp_first_blocked2d(sbp_map *map, const double *P, const double *Q, int64_t n,
                                       int32_t *out_xy, int32_t *flags) {
    SBP_NVTX("sbp_first_blocked2d");
    SBP_REQUIRE(map && (n == 0 || (P && Q && out_xy)), "sbp_first_blocked2d: NULL argument");
    SBP_REQUIRE(n >= 0, "sbp_first_blocked2d: n < 0");
    if (n == 0) return SBP_OK;
    sbp_ctx *ctx = map->ctx;
    SBP_DEVICE(ctx);
    bool smem = !ctx->tune.force_global && map->smem_resident &&
                n * 32 >= map->packed_bytes * (int64_t)ctx->sm_count / 4;
    LaunchCfg c = cfg_for(map, smem);
    int64_t need = (n * 32 + c.threads - 1) / c.threads;
    if (need < c.grid) c.grid = (int)need;
    if (smem) {
        int32_t rc = set_smem(k_first_blocked2d<true>, c.smem);
        if (rc) return rc;
        k_first_blocked2d<true><<<c.grid, c.threads, c.smem, ctx->stream>>>(
            map->view, (const double2 *)P, (const double2 *)Q, n, out_xy, flags);
    } else {
        k_first_blocked2d<false><<<c.grid, c.threads, 0, ctx->stream>>>(
            map->view, (const double2 *)P, (const double2 *)Q, n, out_xy, flags);
    }
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}
