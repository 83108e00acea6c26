// knn.cu -- K4: k-nearest-neighbour queries on the node store.  Batches are answered exactly
// from a uniform cell grid built inside the call (k_knn_cells_t: thread per query, k_knn_cells:
// warp per query); behind it, for the queries the grid cannot settle, the brute-force path:
// seeded filter + exact refine (filter.cuh), thread-per-query scan with per-thread top-k as the
// general / fallback kernel; CTA-per-query scan with shuffle merge for few queries.
//
// Reference behaviour restated (file:line under /root/reference/sbp_env/):
//   RRTPlanner.find_nearest_neighbour_idx     planners/rrtPlanner.py:268-279
//   np.argsort(distances)[:20] candidates     planners/rrtPlanner.py:151-152, :226-227
//
// Distance semantics: d = sqrt_rn(((dx*dx + dy*dy) + dz*dz) + ...),
// every operation rounded separately (no FMA), which is bit-identical to
// np.linalg.norm(poses - q, axis=1) for d <= 8.  Results are ordered by
// (d, index) with d the ROUNDED square root: distinct d^2 may collapse to one d,
// in which case the lower index wins, exactly like np.argmin / a stable argsort.
// The hot loop compares d^2 against round-up(d_k * d_k): sqrt_rn is monotone, so
// a candidate with d^2 >= that bound cannot have sqrt < d_k; only candidates
// below it take the slow path that computes the square root.
#include <cooperative_groups.h>
#include <math.h>

#include "common.cuh"
#include "device_fns.cuh"
#include "filter.cuh"

namespace cg = cooperative_groups;

#define WIDE_TPB 256

// --------------------------------------------------------- distance core ----
// Per-thread sorted list of the K best (d, idx), d ascending, ties keep arrival
// order (callers feed indices in ascending order, so ties keep the lower index).
template <int K>
struct TopK {
    double d[K];
    int32_t i[K];
    double bound2;   // round-up(d[K-1]^2): candidates with d^2 > bound2 are rejected
    double seed2;    // a-priori bound used while the list is not full (inf = none)
    __device__ __forceinline__ void init(double seed = INFINITY) {
#pragma unroll
        for (int t = 0; t < K; ++t) { d[t] = INFINITY; i[t] = -1; }
        bound2 = seed2 = seed;
    }
    // Empty slots are (inf, -1) and rank after every real entry, including real entries
    // whose distance is +inf (fp64 overflow of d^2): those fill the list in index order,
    // as a stable argsort of an all-inf row does.
    //
    // Branch-free sorted insertion: r_t = "entry t ranks before the new one" (real and
    // d[t] <= dd; ties keep the earlier arrival = lower index) is a prefix of trues, and
    //     new[t] = r_t ? old[t] : (r_{t-1} ? (dd, idx) : old[t-1]).
    // Pure predicate/select code, ~8 instructions per slot (the earlier `if (!placed)`
    // formulation made ptxas copy the whole list per step: ~30 instructions per slot).
    __device__ __forceinline__ void insert_sorted(double dd, int32_t idx) {
        bool r_cur = (i[K - 1] >= 0) & (d[K - 1] <= dd);     // r_{K-1}: true => list full, no room
#pragma unroll
        for (int t = K - 1; t >= 1; --t) {
            const bool r_prev = (i[t - 1] >= 0) & (d[t - 1] <= dd);
            const double nd = r_prev ? dd : d[t - 1];
            const int32_t ni = r_prev ? idx : i[t - 1];
            d[t] = r_cur ? d[t] : nd;
            i[t] = r_cur ? i[t] : ni;
            r_cur = r_prev;
        }
        d[0] = r_cur ? d[0] : dd;
        i[0] = r_cur ? i[0] : idx;
    }
    __device__ __forceinline__ void offer(double d2, int32_t idx) {
        if (d2 <= bound2 && idx >= 0) {
            insert_sorted(__dsqrt_rn(d2), idx);
            bound2 = i[K - 1] < 0 ? seed2 : __dmul_ru(d[K - 1], d[K - 1]);
        }
    }
    // The same insertion with the order (d, index) decided explicitly, for callers that do not
    // feed indices in ascending order (k_knn_cells_t walks the nodes cell by cell).  bound2 is
    // inflated so that it also admits a d^2 just above d_K^2 whose root still rounds to d_K.
    // (Bitwise predicate logic on purpose: with && / || ptxas emitted two divergent branches
    // per slot, ~265 instructions per insertion instead of ~150.  An empty slot is (inf, -1):
    // as an unsigned index it never ranks before a real one, also when dd itself is +inf.)
    __device__ __forceinline__ bool before(int t, double dd, int32_t idx) const {
        const bool lt = d[t] < dd, eq = d[t] == dd, il = (uint32_t)i[t] < (uint32_t)idx;
        return lt | (eq & il);
    }
    __device__ __forceinline__ void insert_pair(double dd, int32_t idx) {
        bool r_cur = before(K - 1, dd, idx);
#pragma unroll
        for (int t = K - 1; t >= 1; --t) {
            const bool r_prev = before(t - 1, dd, idx);
            const double nd = r_prev ? dd : d[t - 1];
            const int32_t ni = r_prev ? idx : i[t - 1];
            d[t] = r_cur ? d[t] : nd;
            i[t] = r_cur ? i[t] : ni;
            r_cur = r_prev;
        }
        d[0] = r_cur ? d[0] : dd;
        i[0] = r_cur ? i[0] : idx;
    }
    __device__ __forceinline__ void offer_pair(double d2, int32_t idx) {
        if (d2 <= bound2 && idx >= 0) {
            // (a query that is itself a node: d^2 = 0 would send the whole warp through the
            // slow path of the square root; the root is taken of 1 instead and discarded)
            const bool pos = d2 > 0.0;
            const double root = __dsqrt_rn(pos ? d2 : 1.0);
            insert_pair(pos ? root : d2, idx);
            bound2 = i[K - 1] < 0 ? seed2 : __dmul_ru(d[K - 1], d[K - 1]) * (1.0 + 1e-9) + 1e-300;
        }
    }
    // insert an already square-rooted distance (merging partial lists)
    __device__ __forceinline__ void offer_sorted(double dd, int32_t idx) { insert_sorted(dd, idx); }
    // remove the head (used by the block merge)
    __device__ __forceinline__ void pop() {
#pragma unroll
        for (int t = 0; t + 1 < K; ++t) { d[t] = d[t + 1]; i[t] = i[t + 1]; }
        d[K - 1] = INFINITY; i[K - 1] = -1;
    }
};

// numpy sorts NaN distances last, in index order (np.argsort, stable): a row the scan could
// not fill -- fewer than k nodes at a non-NaN distance, e.g. a NaN query -- is completed
// with the lowest-index nodes whose distance to the query is NaN (distance reported as NaN).
// A difference is NaN iff one side is NaN or both are the same infinity; squares and sums of
// non-NaN terms never produce NaN.  Called by the thread that wrote the row; rare path.
__device__ __noinline__ void pad_nan_row(const double *__restrict__ soa, int64_t capacity, int64_t n,
                                         int d, const double *__restrict__ xq, int k,
                                         int64_t *irow, double *drow) {
    int cnt = 0;
    while (cnt < k && irow[cnt] >= 0) ++cnt;
    for (int64_t i = 0; i < n && cnt < k; ++i) {
        bool isnan_d = false;
        for (int j = 0; j < d; ++j) {
            const double df = soa[(int64_t)j * capacity + i] - xq[j];
            isnan_d |= df != df;
        }
        if (isnan_d) {
            irow[cnt] = i;
            if (drow) drow[cnt] = __longlong_as_double(0x7ff8000000000000LL);
            ++cnt;
        }
    }
}

// ---------------------------------------------- K4a: batched, thread/query ---
// grid = (ceil(m / TPB), S).  CTA (bx, s) scans node chunk s for its TPB queries; node
// tiles are staged in shared memory (SoA) and read by broadcast.
//
// Two-phase scan per block of 32 nodes.  Phase 1 is branch-free with the top-k state
// read-only (a fused scan+insert loop made ptxas copy the whole list between unrolled
// iterations: ~45 instructions per pair) and only FLAGS candidates in a 32-bit mask.
// Phase 2 revisits the flagged nodes in lock-step over the warp and decides them exactly
// in fp64 (offer()); stale or spurious flags are harmless because offer() re-checks.
//
// fp32 prefilter (default).  Phase 1 runs on an fp32 shadow of the nodes: d2f =
// fma(dy,dy,dx*dx) is compared with a per-query threshold thr_f that is CONSERVATIVE:
// with u = 2^-24, M = max |coordinate| and a = 4*M*u*sqrt(D) (conversion + subtraction
// error of every component), the fp32 value of a pair whose exact distance is r satisfies
// d2f <= (r + a)^2 (1+u)^D, so every pair with fp64 d^2 < bound2 has
// d2f <= thr_f := round_up((sqrt(bound2) + a)^2 (1 + 1e-6)).  The prefilter can therefore
// only add false positives (a relative band of ~1e-6), never lose a neighbour; results
// stay bit-exact while the scan moves from the fp64 pipe (64 lanes/SM) to the fp32 pipe
// (128 lanes/SM).  It is bypassed (pure fp64 scan) when coordinates are too large for
// fp32 squares (M >= 1e15) or when sbp_tune("knn_prefilter", 0) is set.
__device__ __forceinline__ float knn_thr32(double bound2, double a) {
    if (!(bound2 < INFINITY)) return INFINITY;
    double r = sqrt(bound2) + a;
    return __double2float_ru(r * r * (1.0 + 1e-6));
}

template <int D, int K, int TPB>
__global__ void __launch_bounds__(TPB)
k_knn_batched(const double *__restrict__ soa, const float *__restrict__ soa32,
              const double *__restrict__ maxabs, int prefilter, int64_t capacity, int64_t n, int S,
              const double *__restrict__ X, int64_t m, const double *__restrict__ seed2,
              double *__restrict__ pd, int32_t *__restrict__ pi,
              const uint8_t *__restrict__ redo, const int32_t *__restrict__ pending = nullptr) {
    // as the fallback of the filter path: only the flagged blocks of 32 queries (TPB == 32);
    // pending == 0: the cell-grid kernel answered every query, nothing can be flagged
    if (pending && *pending == 0) return;
    constexpr int TILE = TPB * 8;     // nodes staged in shared memory per step
    __shared__ double tile[D][TILE];  // fp64 tile, or the fp32 tile in the same storage
    float(*tile32)[TILE] = reinterpret_cast<float(*)[TILE]>(&tile[0][0]);
    // the grid is capped (a launch that finds nothing pending must cost next to nothing): a CTA
    // loops over the (query tile, node chunk) pairs bid = s * nbx + bx
    const int64_t nbx = (m + TPB - 1) / TPB;
    for (int64_t bid = blockIdx.x; bid < nbx * S; bid += gridDim.x) {
    const int64_t bx = bid % nbx;
    if (redo && !redo[bx]) continue;
    const int64_t q = bx * TPB + threadIdx.x;
    const int s = (int)(bid / nbx);
    const int64_t chunk = (n + S - 1) / S;
    const int64_t c0 = (int64_t)s * chunk;
    const int64_t c1 = c0 + chunk < n ? c0 + chunk : n;
    double qv[D];
    float qf[D];
    double qmax = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        qv[j] = q < m ? X[q * D + j] : 0.0;
        qf[j] = (float)qv[j];
        double av = qv[j] != qv[j] ? INFINITY : fabs(qv[j]);
        qmax = av > qmax ? av : qmax;
    }
    const double M = *maxabs;
    const bool use32 = prefilter && M < 1e15;              // block-uniform
    const double mm = M > qmax ? M : qmax;
    const bool lane32 = mm < 1e15;                         // this lane's fp32 squares are safe
    const double a_err = 4.0 * mm * 5.9604644775390625e-8 * sqrt((double)D) * (1.0 + 1e-6);
    TopK<K> top;
    top.init((seed2 && q < m) ? seed2[q] : INFINITY);
    float thr_f = (use32 && lane32) ? knn_thr32(top.bound2, a_err) : INFINITY;
    for (int64_t t0 = c0; t0 < c1; t0 += TILE) {
        int cnt = (int)(c1 - t0 < TILE ? c1 - t0 : TILE);
        __syncthreads();
        if (use32) {
#pragma unroll
            for (int j = 0; j < D; ++j)
                for (int t = threadIdx.x; t < cnt; t += TPB)
                    tile32[j][t] = soa32[(int64_t)j * capacity + t0 + t];
        } else {
#pragma unroll
            for (int j = 0; j < D; ++j)
                for (int t = threadIdx.x; t < cnt; t += TPB)
                    tile[j][t] = soa[(int64_t)j * capacity + t0 + t];
        }
        __syncthreads();
        for (int b0 = 0; b0 < cnt; b0 += 32) {
            const int bn = cnt - b0 < 32 ? cnt - b0 : 32;
            const double bound2 = top.bound2;
            uint32_t mask = 0;
            if (use32) {
                if (bn == 32) {
#pragma unroll
                    for (int t = 0; t < 32; ++t) {
                        float sf = 0.f;
#pragma unroll
                        for (int j = 0; j < D; ++j) {
                            float df = tile32[j][b0 + t] - qf[j];
                            sf = j == 0 ? df * df : fmaf(df, df, sf);
                        }
                        mask |= (sf <= thr_f) ? (1u << t) : 0u;
                    }
                } else {
                    for (int t = 0; t < bn; ++t) {
                        float sf = 0.f;
#pragma unroll
                        for (int j = 0; j < D; ++j) {
                            float df = tile32[j][b0 + t] - qf[j];
                            sf = j == 0 ? df * df : fmaf(df, df, sf);
                        }
                        mask |= (sf <= thr_f) ? (1u << t) : 0u;
                    }
                }
            } else {
                if (bn == 32) {
#pragma unroll
                    for (int t = 0; t < 32; ++t) {
                        double pv[D];
#pragma unroll
                        for (int j = 0; j < D; ++j) pv[j] = tile[j][b0 + t];
                        mask |= (dist2_full<D>(pv, qv) <= bound2) ? (1u << t) : 0u;
                    }
                } else {
                    for (int t = 0; t < bn; ++t) {
                        double pv[D];
#pragma unroll
                        for (int j = 0; j < D; ++j) pv[j] = tile[j][b0 + t];
                        mask |= (dist2_full<D>(pv, qv) <= bound2) ? (1u << t) : 0u;
                    }
                }
            }
            // lock-step over the warp: round r decides every lane's r-th flagged node, so
            // insertions of different queries share one pass through offer()
            while (__any_sync(0xffffffffu, mask != 0u)) {
                double d2 = INFINITY;
                int32_t id = -1;
                if (mask) {
                    const int t = __ffs(mask) - 1;
                    mask &= mask - 1;
                    double pv[D];
#pragma unroll
                    for (int j = 0; j < D; ++j)
                        pv[j] = use32 ? soa[(int64_t)j * capacity + t0 + b0 + t] : tile[j][b0 + t];
                    d2 = dist2_full<D>(pv, qv);
                    id = (int32_t)(t0 + b0 + t);
                }
                top.offer(d2, id);
            }
            if (use32 && top.bound2 != bound2) thr_f = lane32 ? knn_thr32(top.bound2, a_err) : INFINITY;
        }
    }
    if (q < m) {
        int64_t base = (q * S + s) * K;
#pragma unroll
        for (int t = 0; t < K; ++t) { pd[base + t] = top.d[t]; pi[base + t] = top.i[t]; }
    }
    }
}

// --------------------------------------- K4a': seeded filter + exact refine ---
// (scan, thresholds and the centred shadow: filter.cuh)
template <int D, int NW, int FQ>
__global__ void __launch_bounds__(NW * 32)
k_knn_filter(const float *__restrict__ sh, const double *__restrict__ info, int64_t cap32,
             int64_t n, int S, int64_t chunk, const double *__restrict__ X, int64_t m,
             const double *__restrict__ seed2, const uint8_t *__restrict__ done,
             const int32_t *__restrict__ pending, int C,
             int32_t *__restrict__ cand, int32_t *__restrict__ ccount) {
    if (pending && *pending == 0) return;          // every query already answered (k_knn_cells)
    constexpr int TPB = NW * 32;
    const int64_t qb = (int64_t)blockIdx.x * (TPB * FQ) + threadIdx.x;   // query r: qb + r * TPB
    const int s = blockIdx.y;
    const int64_t c0 = (int64_t)s * chunk;                       // multiple of the tile
    const int64_t c1 = c0 + chunk < n ? c0 + chunk : n;
    float nq2[FQ][D];        // -2 q^_j
    float thr[FQ];
    int cc[FQ];
    int32_t *cl[FQ];         // this (query, chunk)'s candidate list
    bool any_usable = false;
#pragma unroll
    for (int r = 0; r < FQ; ++r) {
        const int64_t q = qb + (int64_t)r * TPB;
        const bool live = q < m && !(done && done[q]);               // done: answered by k_knn_cells
        // a bound so loose that more nodes are expected inside it than the lists hold (C S groups)
        const bool usable = filter_query_setup<D>(info, X + (live ? q : 0) * D, live,
                                                  live ? seed2[q] : -1.0, nq2[r], thr[r], (double)n,
                                                  2.0 * (double)C * (double)S);
        cc[r] = (live && !usable) ? C + 1 : 0;                       // C + 1 = "not filterable"
        cl[r] = cand + ((live ? q : 0) * S + s) * (int64_t)C;
        any_usable |= usable;
    }
    // nothing to filter for this CTA's queries (loose seeds of an isotropic cloud in d > 2,
    // outliers, ...): they are all flagged; skip the scan
    if (!__syncthreads_or(any_usable ? 1 : 0)) {
#pragma unroll
        for (int r = 0; r < FQ; ++r) {
            const int64_t q = qb + (int64_t)r * TPB;
            if (q < m) ccount[q * S + s] = cc[r];
        }
        return;
    }
    filter_scan<D, NW, FQ>(sh, cap32, c0, c1, nq2, thr, [&](int r, int32_t g) {
        cl[r][cc[r] < C ? cc[r] : C - 1] = g;                        // an overflowing list is never read
        ++cc[r];
    });
#pragma unroll
    for (int r = 0; r < FQ; ++r) {
        const int64_t q = qb + (int64_t)r * TPB;
        if (q < m) ccount[q * S + s] = cc[r];
    }
}

// One thread per query: exact fp64 decision of the candidate groups, chunks and groups in
// ascending node order (ties keep the lower index).  A query that cannot be answered from
// its lists flags its block of 32 queries for the one-kernel scan.
template <int D, int K>
__global__ void __launch_bounds__(64)
k_knn_refine(const double *__restrict__ soa, int64_t capacity, int64_t n,
             const double *__restrict__ X, int64_t m, const double *__restrict__ seed2, int S,
             int C, const int32_t *__restrict__ cand, const int32_t *__restrict__ ccount, int k,
             int64_t *__restrict__ idx, double *__restrict__ dist, uint8_t *__restrict__ redo,
             const uint8_t *__restrict__ done, const int32_t *__restrict__ pending) {
    if (pending && *pending == 0) return;
    const unsigned FULL = 0xffffffffu;
    // (capped grid: a CTA loops over its blocks of 64 queries)
    for (int64_t q0 = (int64_t)blockIdx.x * blockDim.x; q0 < m; q0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = q0 + threadIdx.x;
    bool live = q < m && !(done && done[q]);                     // done: row already written
    if (!__any_sync(FULL, live)) continue;
    if (live) {
        bool over = false;
        for (int s = 0; s < S; ++s) over |= ccount[q * S + s] > C;
        if (over) { redo[q >> 5] = 1; live = false; }
    }
    double qv[D];
#pragma unroll
    for (int j = 0; j < D; ++j) qv[j] = live ? X[q * D + j] : 0.0;
    TopK<K> top;
    top.init(live ? seed2[q] : -1.0);
    // The warp walks its 32 candidate lists in lock-step: every lane evaluates the 8 nodes of
    // its t-th group, then the lanes insert their passing nodes round by round, so that the
    // ~100-instruction sorted insertion always runs with many lanes active.
    for (int s = 0; s < S; ++s) {
        const int c = live ? ccount[q * S + s] : 0;
        const int32_t *cl = cand + ((live ? q : 0) * S + s) * (int64_t)C;
        const int cmax = __reduce_max_sync(FULL, c);
        // the candidate id of the next iteration is fetched one iteration ahead (the loop is
        // bound by the latency of its dependent loads: list entry -> node coordinates)
        int32_t g_next = 0 < c ? cl[0] : -1;
        for (int t = 0; t < cmax; ++t) {
            const int64_t b = g_next >= 0 ? (int64_t)g_next * 8 : n;
            g_next = t + 1 < c ? cl[t + 1] : -1;
            double d2[8];
            uint32_t mask = 0;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const bool ok = b + u < n;
                const int64_t id = ok ? b + u : 0;
                double pv[D];
#pragma unroll
                for (int j = 0; j < D; ++j) pv[j] = soa[(int64_t)j * capacity + id];
                d2[u] = dist2_full<D>(pv, qv);
                mask |= (ok && d2[u] <= top.bound2) ? (1u << u) : 0u;
            }
            while (__any_sync(FULL, mask != 0u)) {
                const int u = mask ? __ffs(mask) - 1 : 8;
                mask &= mask - 1;
                double dd = INFINITY;
#pragma unroll
                for (int v = 0; v < 8; ++v) dd = u == v ? d2[v] : dd;
                top.offer(dd, u < 8 ? (int32_t)(b + u) : -1);
            }
        }
    }
    if (!live) continue;
#pragma unroll
    for (int t = 0; t < K; ++t)
        if (t < k) {
            idx[q * k + t] = top.i[t];
            if (dist) dist[q * k + t] = top.i[t] >= 0 ? top.d[t] : INFINITY;
        }
    int have = 0;                             // (static indexing keeps the list in registers)
#pragma unroll
    for (int t = 0; t < K; ++t) have += (t < k && top.i[t] >= 0) ? 1 : 0;
    if (have < k)
        pad_nan_row(soa, capacity, n, D, X + q * D, k, idx + q * k, dist ? dist + q * k : nullptr);
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
            d2[u] = dist2_full<D>(pv[u], qv);
            mask |= (t0 + (int64_t)u * WIDE_TPB < c1 && d2[u] <= top.bound2) ? (1u << u) : 0u;
        }
        if (mask) {
#pragma unroll
            for (int u = 0; u < UNR; ++u)
                if (mask & (1u << u)) top.offer(d2[u], (int32_t)(t0 + (int64_t)u * WIDE_TPB));
        }
    }
    if (S == 1) {
        block_merge_topk<K>(top, nullptr, nullptr, idx + q * k_out, dist ? dist + q * k_out : nullptr, k_out);
        if (threadIdx.x == 0 && idx[q * k_out + k_out - 1] < 0)
            pad_nan_row(soa, capacity, n, D, X + q * D, k_out, idx + q * k_out,
                        dist ? dist + q * k_out : nullptr);
    } else {
        block_merge_topk<K>(top, pd + (q * S + s) * K, pi + (q * S + s) * K, nullptr, nullptr, 0);
    }
}

// Second level of the wide path: one CTA per query merges its S partial lists
// (S up to 1024; thread t takes lists t, t+256, ...; ascending chunk order keeps
// the lower index first among equal distances).
template <int K>
__global__ void __launch_bounds__(WIDE_TPB)
k_knn_merge_wide(const double *__restrict__ pd, const int32_t *__restrict__ pi, int S, int k_out,
                 int64_t *__restrict__ idx, double *__restrict__ dist,
                 const double *__restrict__ soa, int64_t capacity, int64_t n, int d,
                 const double *__restrict__ X) {
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
    if (threadIdx.x == 0 && idx[q * k_out + k_out - 1] < 0)
        pad_nan_row(soa, capacity, n, d, X + q * d, k_out, idx + q * k_out,
                    dist ? dist + q * k_out : nullptr);
}

// ----------------------------------------------------- K4c: chunk merge ------
// One thread per query merges S sorted partial lists (chunk s holds indices below
// chunk s+1, so on equal d the lower chunk wins) into the final [m][k] rows.
__device__ __forceinline__ void knn_merge_row(const double *__restrict__ pd, const int32_t *__restrict__ pi,
                                              int64_t q, int S, int K, int k, int64_t *__restrict__ idx,
                                              double *__restrict__ dist) {
    const double *d0 = pd + q * S * K;
    const int32_t *i0 = pi + q * S * K;
    if (S == 1) {
        for (int t = 0; t < k; ++t) {
            idx[q * k + t] = i0[t];
            if (dist) dist[q * k + t] = i0[t] >= 0 ? d0[t] : INFINITY;
        }
        return;
    }
    if (S <= 4) {
        // few chunks (the seeded batched path uses 1-4): head pointers stay in registers and
        // each list's current head is cached, so a round is S compares and one reload
        int h[4] = {0, 0, 0, 0};
        double hd[4];
        int32_t hi[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            hi[s] = s < S ? i0[s * K] : -1;
            hd[s] = s < S ? d0[s * K] : INFINITY;
        }
        for (int t = 0; t < k; ++t) {
            int best = -1;
            double bd = INFINITY;
#pragma unroll
            for (int s = 0; s < 4; ++s)
                if (hi[s] >= 0 && (best < 0 || hd[s] < bd)) { best = s; bd = hd[s]; }
            int64_t oi = -1;
#pragma unroll
            for (int s = 0; s < 4; ++s)
                if (s == best) {
                    oi = hi[s];
                    ++h[s];
                    hi[s] = h[s] < K ? i0[s * K + h[s]] : -1;
                    hd[s] = h[s] < K ? d0[s * K + h[s]] : INFINITY;
                }
            idx[q * k + t] = oi;
            if (dist) dist[q * k + t] = best >= 0 ? bd : INFINITY;
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

__global__ void k_knn_merge(const double *__restrict__ pd, const int32_t *__restrict__ pi, int64_t m,
                            int S, int K, int k, int64_t *__restrict__ idx,
                            double *__restrict__ dist, const uint8_t *__restrict__ redo,
                            const double *__restrict__ soa, int64_t capacity, int64_t n, int d,
                            const double *__restrict__ X, const int32_t *__restrict__ pending = nullptr) {
    if (pending && *pending == 0) return;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < m; q += (int64_t)gridDim.x * blockDim.x) {
        if (redo && !redo[q >> 5]) continue;
        knn_merge_row(pd, pi, q, S, K, k, idx, dist);
        if (idx[q * k + k - 1] < 0)
            pad_nan_row(soa, capacity, n, d, X + q * d, k, idx + q * k, dist ? dist + q * k : nullptr);
    }
}

// ---------------------------------------------------- kNN seeding (d = 2) ----
// An a-priori upper bound of each query's K-th neighbour distance, so that the brute
// force scan does not spend K*ln(N/K) list insertions per query discovering it.  Nodes
// are counted into a uniform G x G grid over their bounding box; a query grows a square
// block of cells around its own cell until the block holds >= K nodes; every node of the
// block lies within B = distance from the query to the farthest corner of the block, so
// the K-th nearest distance is <= B and no node farther than B can be among the K nearest.
// B^2 is inflated by 1e-9 (cell-assignment rounding, and distinct d^2 that collapse to one
// rounded square root).  The scan still evaluates every (query, node) pair; only the
// number of candidates that reach the exact fp64 insertion path shrinks (~9.4K -> ~2K).
// per-block partial bounding boxes [BOUNDS_BLOCKS][4] = (min x, max x, min y, max y)
__global__ void __launch_bounds__(256)
k_bounds2d_partial(const double *__restrict__ soa, int64_t capacity, int64_t n,
                   double *__restrict__ partial) {
    __shared__ double sh[4][8];
    double mnx = INFINITY, mxx = -INFINITY, mny = INFINITY, mxy = -INFINITY;
    for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < n; t += (int64_t)gridDim.x * 256) {
        double x = soa[t], y = soa[capacity + t];
        mnx = fmin(mnx, x); mxx = fmax(mxx, x); mny = fmin(mny, y); mxy = fmax(mxy, y);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mnx = fmin(mnx, __shfl_down_sync(0xffffffffu, mnx, o));
        mxx = fmax(mxx, __shfl_down_sync(0xffffffffu, mxx, o));
        mny = fmin(mny, __shfl_down_sync(0xffffffffu, mny, o));
        mxy = fmax(mxy, __shfl_down_sync(0xffffffffu, mxy, o));
    }
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sh[0][w] = mnx; sh[1][w] = mxx; sh[2][w] = mny; sh[3][w] = mxy; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) {
            mnx = fmin(mnx, sh[0][i]); mxx = fmax(mxx, sh[1][i]);
            mny = fmin(mny, sh[2][i]); mxy = fmax(mxy, sh[3][i]);
        }
        double *o = partial + 4 * blockIdx.x;
        o[0] = mnx; o[1] = mxx; o[2] = mny; o[3] = mxy;
    }
}

// every block of the histogram kernel folds the partials itself (block 0 publishes b4)
__device__ __forceinline__ void fold_bounds(const double *__restrict__ partial, int nb, double *sh4) {
    if (threadIdx.x < 32) {
        double mnx = INFINITY, mxx = -INFINITY, mny = INFINITY, mxy = -INFINITY;
        for (int i = threadIdx.x; i < nb; i += 32) {
            mnx = fmin(mnx, partial[4 * i]); mxx = fmax(mxx, partial[4 * i + 1]);
            mny = fmin(mny, partial[4 * i + 2]); mxy = fmax(mxy, partial[4 * i + 3]);
        }
        for (int o = 16; o > 0; o >>= 1) {
            mnx = fmin(mnx, __shfl_down_sync(0xffffffffu, mnx, o));
            mxx = fmax(mxx, __shfl_down_sync(0xffffffffu, mxx, o));
            mny = fmin(mny, __shfl_down_sync(0xffffffffu, mny, o));
            mxy = fmax(mxy, __shfl_down_sync(0xffffffffu, mxy, o));
        }
        if (threadIdx.x == 0) { sh4[0] = mnx; sh4[1] = mxx; sh4[2] = mny; sh4[3] = mxy; }
    }
    __syncthreads();
}

__global__ void k_cell_hist(const double *__restrict__ soa, int64_t capacity, int64_t n,
                            const double *__restrict__ partial, int nb, double *__restrict__ b4,
                            int G, int32_t *__restrict__ cells) {
    __shared__ double sb[4];
    fold_bounds(partial, nb, sb);
    if (blockIdx.x == 0 && threadIdx.x < 4) b4[threadIdx.x] = sb[threadIdx.x];
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (int64_t)gridDim.x * blockDim.x) {
        int cx = cell_of(soa[t], sb[0], sb[1], G), cy = cell_of(soa[capacity + t], sb[2], sb[3], G);
        atomicAdd(cells + (int64_t)cy * G + cx, 1);
    }
}

// exclusive scan of the cell counts -> cell starts (+ a copy used as the scatter cursor),
// in three small launches: per-tile sums, scan of the <= 1024 tile sums, per-tile scan + offset
#define SCAN_TILE 4096
__global__ void __launch_bounds__(1024)
k_cell_tile_sums(const int32_t *__restrict__ cells, int64_t L, int32_t *__restrict__ tsums) {
    __shared__ int32_t ws[32];
    const int64_t i0 = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * 4;
    int32_t v = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) v += i0 + j < L ? cells[i0 + j] : 0;
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
k_cell_scan_sums(int32_t *__restrict__ tsums, int nt, int32_t *__restrict__ total) {
    __shared__ int32_t ws[32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    int32_t x = tid < nt ? tsums[tid] : 0, inc = x;
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
    const int32_t excl = ws[w] + inc - x;
    if (tid < nt) tsums[tid] = excl;
    if (tid == nt - 1) *total = excl + x;
}

__global__ void __launch_bounds__(1024)
k_cell_scan_apply(const int32_t *__restrict__ cells, int64_t L, const int32_t *__restrict__ tsums,
                  int32_t *__restrict__ start, int32_t *__restrict__ cursor) {
    __shared__ int32_t ws[32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t i0 = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)tid * 4;
    int32_t v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = i0 + j < L ? cells[i0 + j] : 0;
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
    int32_t run = tsums[blockIdx.x] + ws[w] + inc - tsum;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (i0 + j < L) { start[i0 + j] = run; cursor[i0 + j] = run; }
        run += v[j];
    }
}

// The whole scan in one launch for grids of up to SCAN_FUSED_TILES tiles: every tile sums the
// cells of the tiles before it itself (a few hundred KB of L2 reads) instead of waiting for two
// more launches.
#define SCAN_FUSED_TILES 48
__global__ void __launch_bounds__(1024)
k_cell_scan_fused(const int32_t *__restrict__ cells, int64_t L, int32_t *__restrict__ start,
                  int32_t *__restrict__ cursor) {
    __shared__ int32_t ws[32];
    __shared__ int32_t s_before;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t t0 = (int64_t)blockIdx.x * SCAN_TILE;
    int32_t acc = 0;
    for (int64_t i = (int64_t)tid * 4; i < t0; i += 4096) {      // t0 is a multiple of 4096
        const int4 v = *reinterpret_cast<const int4 *>(cells + i);
        acc += v.x + v.y + v.z + v.w;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (lane == 0) ws[w] = acc;
    __syncthreads();
    if (w == 0) {
        acc = ws[lane];
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0) s_before = acc;
    }
    __syncthreads();
    const int32_t before = s_before;
    __syncthreads();
    const int64_t i0 = t0 + (int64_t)tid * 4;
    int32_t v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = i0 + j < L ? cells[i0 + j] : 0;
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
    int32_t run = before + ws[w] + inc - tsum;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (i0 + j < L) { start[i0 + j] = run; cursor[i0 + j] = run; }
        run += v[j];
    }
    if (i0 <= L - 1 && L - 1 < i0 + 4) start[L] = run;            // total = start[L]
}

// band (nullable): [first, last] row of cells this rank needs (k_prm_band); nodes of other rows are
// not placed (their slice of `ids` stays undefined on this rank).
__global__ void k_cell_scatter(const double *__restrict__ soa, int64_t capacity, int64_t n,
                               const double *__restrict__ b4, int G, int32_t *__restrict__ cursor,
                               int32_t *__restrict__ ids, const int32_t *__restrict__ band) {
    const int r0 = band ? band[0] : 0, r1 = band ? band[1] : G - 1;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int cy = cell_of(soa[capacity + t], b4[2], b4[3], G);
        if (cy < r0 || cy > r1) continue;
        const int cx = cell_of(soa[t], b4[0], b4[1], G);
        ids[atomicAdd(cursor + (int64_t)cy * G + cx, 1)] = (int32_t)t;
    }
}

// The rows of cells that hold the cell-sorted positions [p0, p0 + count), widened by `halo` rows on
// both sides: all a rank that answers those rows needs of the grid (a row whose search leaves the
// band is left to the exact path).  The cell counts / starts are global on every rank; placing the
// nodes, ordering the cells and copying the poses is then done for the band only -- the part of the
// grid build that shrinks with the number of ranks.
__global__ void k_prm_band(const int32_t *__restrict__ start, int G, int64_t p0, int64_t count, int halo,
                           int32_t *__restrict__ band) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int lo = 0, hi = G - 1;                                   // first row that ends beyond p0
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((int64_t)start[(int64_t)(mid + 1) * G] > p0) hi = mid; else lo = mid + 1;
    }
    const int r0 = lo;
    lo = 0; hi = G - 1;                                       // last row that begins before p0 + count
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((int64_t)start[(int64_t)mid * G] < p0 + count) lo = mid; else hi = mid - 1;
    }
    band[0] = r0 - halo < 0 ? 0 : r0 - halo;
    band[1] = lo + halo > G - 1 ? G - 1 : lo + halo;
}

// Last step of the grid build for the fused PRM kernels, one thread per cell (of the band): order the
// cell's segment of `ids` by node index (the scatter places them in the order its atomics landed;
// every rank must see the same permutation) and copy the poses into the cell-sorted array `pts`.
__global__ void k_cell_finalize(const double *__restrict__ soa, int64_t capacity, const int32_t *__restrict__ start,
                                int G, const int32_t *__restrict__ band, int32_t *__restrict__ ids,
                                double2 *__restrict__ pts) {
    const int64_t c0 = band ? (int64_t)band[0] * G : 0, c1 = band ? (int64_t)(band[1] + 1) * G : (int64_t)G * G;
    for (int64_t c = c0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < c1; c += (int64_t)gridDim.x * blockDim.x) {
        const int32_t e0 = start[c], e1 = start[c + 1];
        for (int32_t i = e0 + 1; i < e1; ++i) {
            const int32_t v = ids[i];
            int32_t j = i - 1;
            while (j >= e0 && ids[j] > v) { ids[j + 1] = ids[j]; --j; }
            ids[j + 1] = v;
        }
        for (int32_t i = e0; i < e1; ++i) {
            const int32_t v = ids[i];
            pts[i] = make_double2(soa[v], soa[capacity + v]);
        }
    }
}

// The scatter above places the nodes of a cell in the order their atomics happened to land.  Where
// several devices must agree on the permutation (the rows of a roadmap split by position in `ids`
// over the ranks), every cell's segment is sorted by node index: one thread per cell, segments are
// a handful of entries.
__global__ void k_cell_sort_segments(const int32_t *__restrict__ start, int64_t L, int32_t *__restrict__ ids) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < L; c += (int64_t)gridDim.x * blockDim.x) {
        const int32_t e0 = start[c], e1 = start[c + 1];
        for (int32_t i = e0 + 1; i < e1; ++i) {
            const int32_t v = ids[i];
            int32_t j = i - 1;
            while (j >= e0 && ids[j] > v) { ids[j + 1] = ids[j]; --j; }
            ids[j + 1] = v;
        }
    }
}

// The whole grid build of sbp_prm_connect2d in ONE cooperative launch (the separate kernels above cost
// eight dependent launches, ~2 us of graph-node latency each on top of their work): bounding box and
// zeroed counts -> histogram -> scan of the cell counts (per-block chunks) -> band of cell rows this
// part needs (k_prm_band) -> placement of the band's nodes -> per-cell order + cell-sorted poses
// (k_cell_finalize), separated by grid-wide barriers.  Also sets done[] = 1 (rows of other parts count
// as settled for the exact path) and zeroes the pending counter.
struct GridBuildArgs {
    const double *soa;
    int64_t capacity, n;
    double *b4;               // out: bounding box
    double *partial;          // [gridDim.x][4]
    int G;
    int32_t *hdr4;            // [0] pending (zeroed), [1..2] band (out), [4..19] per-slice counters of prm.cu (zeroed)
    int32_t *cells, *start, *cursor, *ids;
    double2 *pts;
    uint8_t *done;
    int32_t *bsum;            // [gridDim.x]
    int64_t p0, count;
    int banded, halo;
};

#define GB_TPB 512
__global__ void __launch_bounds__(GB_TPB, 3)
k_prm_grid_build(GridBuildArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[4][GB_TPB / 32];
    __shared__ double sb[4];
    __shared__ int32_t ws[GB_TPB / 32];
    __shared__ int32_t s_off, s_carry, s_r0, s_r1;
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nb = gridDim.x, b = blockIdx.x;
    const int64_t gtid = (int64_t)b * GB_TPB + tid, gsz = (int64_t)nb * GB_TPB;
    const int G = a.G;
    const int64_t L = (int64_t)G * G;
    // ---- phase 0: bounding-box partials, zeroed counts, done = 1
    {
        double mnx = INFINITY, mxx = -INFINITY, mny = INFINITY, mxy = -INFINITY;
#pragma unroll 4
        for (int64_t t = gtid; t < a.n; t += gsz) {
            const double x = a.soa[t], y = a.soa[a.capacity + t];
            mnx = fmin(mnx, x); mxx = fmax(mxx, x); mny = fmin(mny, y); mxy = fmax(mxy, y);
            a.done[t] = 1;
        }
        for (int64_t i = gtid; i < L; i += gsz) a.cells[i] = 0;
        if (gtid < GRID_HDR_INTS) a.hdr4[gtid] = 0;
        for (int o = 16; o > 0; o >>= 1) {
            mnx = fmin(mnx, __shfl_down_sync(FULL, mnx, o)); mxx = fmax(mxx, __shfl_down_sync(FULL, mxx, o));
            mny = fmin(mny, __shfl_down_sync(FULL, mny, o)); mxy = fmax(mxy, __shfl_down_sync(FULL, mxy, o));
        }
        if (lane == 0) { sh[0][w] = mnx; sh[1][w] = mxx; sh[2][w] = mny; sh[3][w] = mxy; }
        __syncthreads();
        if (tid == 0) {
            for (int i = 1; i < GB_TPB / 32; ++i) {
                mnx = fmin(mnx, sh[0][i]); mxx = fmax(mxx, sh[1][i]); mny = fmin(mny, sh[2][i]); mxy = fmax(mxy, sh[3][i]);
            }
            double *o = a.partial + 4 * b;
            o[0] = mnx; o[1] = mxx; o[2] = mny; o[3] = mxy;
        }
    }
    grid.sync();
    fold_bounds(a.partial, nb, sb);                              // (ends with a __syncthreads)
    if (b == 0 && tid < 4) a.b4[tid] = sb[tid];
    const double b0 = sb[0], b1 = sb[1], b2 = sb[2], b3 = sb[3];
    // ---- phase 1: histogram (four nodes per thread and step: their loads are in flight together)
    for (int64_t t = gtid; t < a.n; t += 4 * gsz) {
        double x[4], y[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t tt = t + u * gsz;
            x[u] = tt < a.n ? a.soa[tt] : 0.0;
            y[u] = tt < a.n ? a.soa[a.capacity + tt] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (t + u * gsz < a.n)
                atomicAdd(a.cells + (int64_t)cell_of(y[u], b2, b3, G) * G + cell_of(x[u], b0, b1, G), 1);
    }
    grid.sync();
    // ---- phase 2: exclusive scan, a contiguous chunk of cells per block
    const int64_t chunk = (L + nb - 1) / nb, c0 = (int64_t)b * chunk, c1 = c0 + chunk < L ? c0 + chunk : L;
    {
        int32_t v = 0;
        for (int64_t i = c0 + tid; i < c1; i += GB_TPB) v += a.cells[i];
        v = __reduce_add_sync(FULL, v);
        if (lane == 0) ws[w] = v;
        __syncthreads();
        if (tid == 0) {
            int32_t t = 0;
            for (int i = 0; i < GB_TPB / 32; ++i) t += ws[i];
            a.bsum[b] = t;
        }
    }
    grid.sync();
    {
        int32_t v = 0;
        for (int j = tid; j < b; j += GB_TPB) v += a.bsum[j];
        v = __reduce_add_sync(FULL, v);
        __syncthreads();
        if (lane == 0) ws[w] = v;
        __syncthreads();
        if (tid == 0) {
            int32_t t = 0;
            for (int i = 0; i < GB_TPB / 32; ++i) t += ws[i];
            s_off = t;
            s_carry = 0;
        }
        __syncthreads();
        for (int64_t base = c0; base < c1; base += GB_TPB) {
            const int64_t i = base + tid;
            const int32_t x = i < c1 ? a.cells[i] : 0;
            int32_t inc = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int32_t u = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += u; }
            if (lane == 31) ws[w] = inc;
            __syncthreads();
            if (w == 0) {
                const int32_t y = lane < GB_TPB / 32 ? ws[lane] : 0;
                int32_t yi = y;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int32_t u = __shfl_up_sync(FULL, yi, o); if (lane >= o) yi += u; }
                if (lane < GB_TPB / 32) ws[lane] = yi - y;
            }
            __syncthreads();
            const int32_t excl = s_off + s_carry + ws[w] + inc - x;
            if (i < c1) { a.start[i] = excl; a.cursor[i] = excl; }
            __syncthreads();
            if (tid == GB_TPB - 1) s_carry = excl + x - s_off;
            __syncthreads();
        }
        if (b == nb - 1 && tid == 0) a.start[L] = (int32_t)a.n;   // every node has a cell
    }
    grid.sync();
    // ---- band of cell rows this part needs (see k_prm_band)
    if (tid == 0) { s_r0 = G - 1; s_r1 = 0; }
    __syncthreads();
    if (a.banded) {
        for (int r = tid; r < G; r += GB_TPB) {
            if ((int64_t)a.start[(int64_t)(r + 1) * G] > a.p0) atomicMin(&s_r0, r);
            if ((int64_t)a.start[(int64_t)r * G] < a.p0 + a.count) atomicMax(&s_r1, r);
        }
        __syncthreads();
    }
    const int r0 = a.banded ? (s_r0 - a.halo < 0 ? 0 : s_r0 - a.halo) : 0;
    const int r1 = a.banded ? (s_r1 + a.halo > G - 1 ? G - 1 : s_r1 + a.halo) : G - 1;
    if (b == 0 && tid == 0) { a.hdr4[1] = r0; a.hdr4[2] = r1; }
    // ---- phase 3: placement of the band's nodes
    for (int64_t t = gtid; t < a.n; t += 4 * gsz) {
        double x[4], y[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t tt = t + u * gsz;
            x[u] = tt < a.n ? a.soa[tt] : 0.0;
            y[u] = tt < a.n ? a.soa[a.capacity + tt] : 0.0;
        }
        int32_t slot[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int cy = cell_of(y[u], b2, b3, G);
            const bool in = t + u * gsz < a.n && cy >= r0 && cy <= r1;
            slot[u] = in ? atomicAdd(a.cursor + (int64_t)cy * G + cell_of(x[u], b0, b1, G), 1) : -1;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (slot[u] >= 0) a.ids[slot[u]] = (int32_t)(t + u * gsz);
    }
    grid.sync();
    // ---- phase 4: per-cell order by node index (thread per cell), then the cell-sorted poses (thread per
    // position: all lanes busy, coalesced stores); a CTA owns a contiguous chunk of the band's cells, so a
    // CTA barrier separates the two
    {
        const int64_t cb0 = (int64_t)r0 * G, cb1 = (int64_t)(r1 + 1) * G;
        const int64_t per = (cb1 - cb0 + nb - 1) / nb;
        const int64_t k0 = cb0 + (int64_t)b * per, k1 = k0 + per < cb1 ? k0 + per : cb1;
        for (int64_t c = k0 + tid; c < k1; c += GB_TPB) {
            const int32_t e0 = a.start[c], e1 = a.start[c + 1];
            for (int32_t i = e0 + 1; i < e1; ++i) {
                const int32_t v = a.ids[i];
                int32_t j = i - 1;
                while (j >= e0 && a.ids[j] > v) { a.ids[j + 1] = a.ids[j]; --j; }
                a.ids[j + 1] = v;
            }
        }
        __syncthreads();
        if (k0 < k1) {
            const int32_t e0 = a.start[k0], e1 = a.start[k1];
            for (int32_t i = e0 + tid; i < e1; i += GB_TPB) {
                const int32_t v = a.ids[i];
                a.pts[i] = make_double2(a.soa[v], a.soa[a.capacity + v]);
            }
        }
    }
}

// seed2[q] = K-th smallest exact fp64 d^2 among ALL nodes of a block of cells around the
// query's cell that holds at least K nodes -- a subset of the nodes, so an upper bound of the
// true K-th smallest -- inflated by 1e-9.  GS = 16 or 32 lanes per query.  The block grows ring by ring
// (then geometrically, until it covers the whole grid: sparse pockets, queries far outside the
// cloud); lanes own its rows (a row of cells is one contiguous range of `ids`).  The nodes
// are enumerated 32 at a time, flat over the rows, and merged into the running K smallest by
// rank counting over shuffles (the K smallest live in lanes 0..K-1, sorted, via a per-warp
// shared-memory scatter).  inf = fewer than K nodes at a non-NaN distance.
template <int D, int GS>      // GS = lanes per query: 16 (K <= 16) or 32
__global__ void __launch_bounds__(256)
k_knn_seed(const double *__restrict__ soa, int64_t capacity, const double *__restrict__ X, int64_t m,
           const double *__restrict__ b4, int G, const int32_t *__restrict__ start,
           const int32_t *__restrict__ ids, int K, double *__restrict__ seed2) {
    __shared__ double sbest[256 / GS][GS];
    const int lane = threadIdx.x & (GS - 1), grp = threadIdx.x / GS;
    // every warp-level primitive below is restricted to this query's lanes
    const unsigned SEG = GS == 32 ? 0xffffffffu : (0xffffu << (threadIdx.x & 16));
    const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / GS;
    if (q >= m) return;                                       // uniform over the GS lanes
    double qv[D];
    bool finite = true;
#pragma unroll
    for (int j = 0; j < D; ++j) { qv[j] = X[q * D + j]; finite &= qv[j] == qv[j]; }
    const double mnx = b4[0], mxx = b4[1], mny = b4[2], mxy = b4[3];
    double out = INFINITY;
    if (finite && mnx <= mxx && mny <= mxy) {
        const int cx = cell_of(qv[0], mnx, mxx, G), cy = cell_of(qv[1], mny, mxy, G);
        int x0 = 0, x1 = 0, y0 = 0, y1 = 0;
        int32_t cnt = 0;
        for (int r = 1;; r = r < 8 ? r + 1 : 2 * r) {
            x0 = cx - r < 0 ? 0 : cx - r; x1 = cx + r > G - 1 ? G - 1 : cx + r;
            y0 = cy - r < 0 ? 0 : cy - r; y1 = cy + r > G - 1 ? G - 1 : cy + r;
            int32_t c = 0;
            for (int row = y0 + lane; row <= y1; row += GS)
                c += start[(int64_t)row * G + x1 + 1] - start[(int64_t)row * G + x0];
            cnt = __reduce_add_sync(SEG, c);
            if (cnt >= K || (x0 == 0 && y0 == 0 && x1 == G - 1 && y1 == G - 1)) break;
        }
        if (cnt >= K) {
            double best = INFINITY;                           // lanes 0..K-1: the K smallest so far
            int nb = 0;                                       // valid entries of `best`
            // rows in passes of GS (one pass unless the block is taller than GS cells)
            for (int rb = y0; rb <= y1; rb += GS) {
                int32_t e0 = 0, len = 0;
                if (rb + lane <= y1) {
                    e0 = start[(int64_t)(rb + lane) * G + x0];
                    len = start[(int64_t)(rb + lane) * G + x1 + 1] - e0;
                }
                int32_t inc = len;
#pragma unroll
                for (int o = 1; o < GS; o <<= 1) {
                    int32_t u = __shfl_up_sync(SEG, inc, o, GS);
                    if (lane >= o) inc += u;
                }
                const int32_t pc = __shfl_sync(SEG, inc, GS - 1, GS);   // nodes in this pass
                const int nrows = y1 - rb + 1 < GS ? y1 - rb + 1 : GS;
                for (int32_t f0 = 0; f0 < pc; f0 += GS) {
                    // flat candidate f0 + lane lives in the first row whose inclusive prefix exceeds it
                    const int32_t f = f0 + lane;
                    int32_t e = -1;
                    for (int row = 0; row < nrows; ++row) {
                        const int32_t rinc = __shfl_sync(SEG, inc, row, GS), rlen = __shfl_sync(SEG, len, row, GS);
                        const int32_t re0 = __shfl_sync(SEG, e0, row, GS);
                        if (e < 0 && f < rinc) e = re0 + (f - (rinc - rlen));
                    }
                    double v = INFINITY;
                    if (f < pc && e >= 0) {
                        const int32_t t = ids[e];
                        double pv[D];
#pragma unroll
                        for (int j = 0; j < D; ++j) pv[j] = soa[(int64_t)j * capacity + t];
                        v = dist2_full<D>(pv, qv);
                        v = v != v ? INFINITY : v;          // a NaN node is never a neighbour
                    }
                    // ranks in the union of `best` (first on ties) and this chunk; only the
                    // nb valid entries of `best` and the nv valid entries of the chunk are
                    // compared against (the rest are +inf and rank behind every finite value)
                    const int nv = pc - f0 < GS ? pc - f0 : GS;
                    int rank_b = 0, rank_v = 0;
                    for (int src = 0; src < nb; ++src) {
                        const double a = __shfl_sync(SEG, best, src, GS);
                        rank_b += (a < best || (a == best && src < lane)) ? 1 : 0;
                        rank_v += a <= v ? 1 : 0;
                    }
                    for (int src = 0; src < nv; ++src) {
                        const double c = __shfl_sync(SEG, v, src, GS);
                        rank_b += c < best ? 1 : 0;
                        rank_v += (c < v || (c == v && src < lane)) ? 1 : 0;
                    }
                    sbest[grp][lane] = INFINITY;
                    __syncwarp(SEG);
                    if (lane < nb && rank_b < GS) sbest[grp][rank_b] = best;
                    if (lane < nv && rank_v < GS) sbest[grp][rank_v] = v;
                    __syncwarp(SEG);
                    best = lane < K ? sbest[grp][lane] : INFINITY;
                    __syncwarp(SEG);
                    nb = nb + nv < K ? nb + nv : K;
                }
            }
            const double kth = __shfl_sync(SEG, best, K - 1, GS);
            out = kth * (1.0 + 1e-9) + 1e-300;
        }
    }
    if (lane == 0) seed2[q] = out;
}

// ------------------------------------------- K4c: exact kNN on the cell grid ---
// One warp per query.  Phase 1 is the seed above with (d, index) pairs: the K best nodes of a
// block of cells holding >= K nodes, d_K = their K-th distance.  Every node that can enter the
// answer has sqrt_rn(d^2) <= d_K, hence |x - q_x|, |y - q_y| <= d_K (1 + 2^-52): it lies in
// the rectangle R of cells [cell_of(q - r), cell_of(q + r)] with r = d_K (1 + 1e-9) (cell_of
// is monotone, so the argument survives its rounding and the clamping at the grid border; for
// D > 2 the grid bounds x and y only, which keeps R a superset).  Phase 2 merges the nodes of R
// (skipped when R lies inside the block) -- if R holds at most `limit` nodes the row is
// written here and done[q] = 1, otherwise done[q] = 0 and the query keeps its seed for the
// brute-force filter.  The order is (sqrt_rn(d^2), index) decided explicitly, so the cells may
// be visited in any order; bound2 = ru(d_K^2) (1 + 1e-9) also admits the d^2 just above
// d_K^2 whose root still rounds to d_K (an index tie the lower index must win).

// rows [ya, yb] x cells [xa, xb] of the grid, 32 nodes at a time: f(e, nv) with e the position
// in `ids` of this lane's node (-1: none) and nv the nodes in the chunk; warp-uniform calls
template <class F>
__device__ __forceinline__ void cells_rect_chunks(const int32_t *__restrict__ start, int G, int xa,
                                                  int xb, int ya, int yb, int lane, F f) {
    const unsigned FULL = 0xffffffffu;
    for (int rb = ya; rb <= yb; rb += 32) {
        int32_t e0 = 0, len = 0;
        if (rb + lane <= yb) {
            e0 = start[(int64_t)(rb + lane) * G + xa];
            len = start[(int64_t)(rb + lane) * G + xb + 1] - e0;
        }
        int32_t inc = len;                                        // inclusive prefix over the rows
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t u = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += u;
        }
        const int32_t pc = __shfl_sync(FULL, inc, 31);            // nodes in this pass
        for (int32_t f0 = 0; f0 < pc; f0 += 32) {
            // flat node f0 + lane lives in the first row whose inclusive prefix exceeds it
            const int32_t fl = f0 + lane;
            int row = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const int32_t v = __shfl_sync(FULL, inc, row + step - 1);
                if (v <= fl) row += step;
            }
            const int32_t rinc = __shfl_sync(FULL, inc, row), rlen = __shfl_sync(FULL, len, row);
            const int32_t re0 = __shfl_sync(FULL, e0, row);
            f(fl < pc ? re0 + (fl - (rinc - rlen)) : -1, pc - f0 < 32 ? pc - f0 : 32);
        }
    }
}

// the K best (d, index) of a warp: lanes 0..nb-1 hold them in ascending order, the other lanes
// (inf, INT_MAX).  merge() inserts the lanes' candidates flagged `ok` (finite d, distinct
// indices not yet in the list) by rank counting: survivors are compared over shuffles, the
// sorted list is binary-searched.
struct WarpBest {
    double d;
    int32_t i;
    int nb;
    __device__ __forceinline__ void init() { d = INFINITY; i = 0x7fffffff; nb = 0; }
    __device__ __forceinline__ void merge(bool ok, double vd, int32_t vi, int K, int lane,
                                          double *sd, int32_t *si) {
        const unsigned FULL = 0xffffffffu;
        const unsigned mask = __ballot_sync(FULL, ok);
        if (!mask) return;
        int before_b = 0, before_v = 0;                           // survivors ranking before me
        for (unsigned mm = mask; mm; mm &= mm - 1) {
            const int src = __ffs(mm) - 1;
            const double cd = __shfl_sync(FULL, vd, src);
            const int32_t ci = __shfl_sync(FULL, vi, src);
            before_b += (cd < d || (cd == d && ci < i)) ? 1 : 0;
            before_v += (cd < vd || (cd == vd && ci < vi)) ? 1 : 0;
        }
        int pos = 0;                                              // list entries ranking before me
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const double bd = __shfl_sync(FULL, d, pos + step - 1);
            const int32_t bi = __shfl_sync(FULL, i, pos + step - 1);
            if (bd < vd || (bd == vd && bi < vi)) pos += step;
        }
        {   // the five steps count up to 31 entries: a full list of 32 needs its last one tested too
            const double bd = __shfl_sync(FULL, d, 31);
            const int32_t bi = __shfl_sync(FULL, i, 31);
            if (pos == 31 && (bd < vd || (bd == vd && bi < vi))) pos = 32;
        }
        const int rank_b = lane + before_b, rank_v = pos + before_v;
        sd[lane] = INFINITY;
        si[lane] = 0x7fffffff;
        __syncwarp(FULL);
        if (lane < nb && rank_b < 32) { sd[rank_b] = d; si[rank_b] = i; }
        if (ok && rank_v < 32) { sd[rank_v] = vd; si[rank_v] = vi; }
        __syncwarp(FULL);
        d = lane < K ? sd[lane] : INFINITY;
        i = lane < K ? si[lane] : 0x7fffffff;
        __syncwarp(FULL);
        nb += __popc(mask);
        nb = nb < K ? nb : K;
    }
};

template <int D>
__global__ void __launch_bounds__(256)
k_knn_cells(const double *__restrict__ soa, int64_t capacity, const double *__restrict__ X, int64_t m,
            const double *__restrict__ b4, int G, const int32_t *__restrict__ start,
            const int32_t *__restrict__ ids, int K, int limit, double *__restrict__ seed2,
            uint8_t *__restrict__ done, int32_t *__restrict__ pending, int64_t *__restrict__ idx,
            double *__restrict__ dist) {
    __shared__ double sd[8][32];
    __shared__ int32_t si[8][32];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= m) return;                                           // uniform over the warp
    double qv[D];
    bool finite = true;
#pragma unroll
    for (int j = 0; j < D; ++j) { qv[j] = X[q * D + j]; finite &= qv[j] == qv[j]; }
    const double mnx = b4[0], mxx = b4[1], mny = b4[2], mxy = b4[3];
    double out = INFINITY;
    bool handled = false;
    WarpBest best;
    best.init();
    if (finite && mnx <= mxx && mny <= mxy) {
        const int cx = cell_of(qv[0], mnx, mxx, G), cy = cell_of(qv[1], mny, mxy, G);
        int x0 = 0, x1 = 0, y0 = 0, y1 = 0;
        int32_t cnt = 0;
        for (int r = 1;; r = r < 8 ? r + 1 : 2 * r) {
            x0 = cx - r < 0 ? 0 : cx - r; x1 = cx + r > G - 1 ? G - 1 : cx + r;
            y0 = cy - r < 0 ? 0 : cy - r; y1 = cy + r > G - 1 ? G - 1 : cy + r;
            int32_t c = 0;
            for (int row = y0 + lane; row <= y1; row += 32)
                c += start[(int64_t)row * G + x1 + 1] - start[(int64_t)row * G + x0];
            cnt = __reduce_add_sync(FULL, c);
            if (cnt >= K || (x0 == 0 && y0 == 0 && x1 == G - 1 && y1 == G - 1)) break;
        }
        if (cnt >= K) {
            double bound2 = INFINITY;
            auto visit = [&](int32_t e, int) {
                bool ok = false;
                double vd = INFINITY;
                int32_t t = 0x7fffffff;
                if (e >= 0) {
                    t = ids[e];
                    double pv[D];
#pragma unroll
                    for (int j = 0; j < D; ++j) pv[j] = soa[(int64_t)j * capacity + t];
                    const double d2 = dist2_full<D>(pv, qv);
                    ok = d2 <= bound2 && d2 < INFINITY;           // NaN / overflow: not a candidate here
                    if (ok) vd = __dsqrt_rn(d2);
                }
                best.merge(ok, vd, t, K, lane, sd[w], si[w]);
                if (best.nb == K) {
                    const double kth = __shfl_sync(FULL, best.d, K - 1);
                    const double b2 = __dmul_ru(kth, kth) * (1.0 + 1e-9) + 1e-300;
                    bound2 = b2 < bound2 ? b2 : bound2;
                }
            };
            cells_rect_chunks(start, G, x0, x1, y0, y1, lane, visit);
            if (best.nb == K && bound2 < INFINITY) {
                out = bound2;
                const double r = sqrt(bound2) * (1.0 + 1e-9) + 1e-300;
                const int xa = cell_of(__dsub_rd(qv[0], r), mnx, mxx, G), xb = cell_of(__dadd_ru(qv[0], r), mnx, mxx, G);
                const int ya = cell_of(__dsub_rd(qv[1], r), mny, mxy, G), yb = cell_of(__dadd_ru(qv[1], r), mny, mxy, G);
                if (xa >= x0 && xb <= x1 && ya >= y0 && yb <= y1) {
                    handled = true;                               // the block already covers R
                } else {
                    int32_t c = 0;
                    for (int row = ya + lane; row <= yb; row += 32)
                        c += start[(int64_t)row * G + xb + 1] - start[(int64_t)row * G + xa];
                    c = __reduce_add_sync(FULL, c);
                    if (c <= limit) {
                        best.init();
                        cells_rect_chunks(start, G, xa, xb, ya, yb, lane, visit);
                        handled = best.nb == K;
                    }
                }
            }
        }
    }
    if (handled && lane < K) {
        idx[q * K + lane] = best.i;
        if (dist) dist[q * K + lane] = best.d;
    }
    if (lane == 0) {
        seed2[q] = out;
        done[q] = handled ? 1 : 0;
        if (!handled) atomicAdd(pending, 1);                      // the brute-force kernels have work
    }
}

// ---------------------------- K4c': the same search with one THREAD per query ---
// Same algorithm and guarantees as k_knn_cells with a per-thread sorted list (TopK<K>,
// explicit (d, index) order): ~10x fewer warp instructions per query, which is what bounds
// the warp-per-query form once there are enough queries to fill the SMs with one-warp CTAs.
// The 32 queries of a warp walk their rectangles row by row in lock-step: per row one
// contiguous range of `ids` (two when the seed block, already merged, is cut out of R), four
// nodes per lane and step, then the lanes insert their passing nodes round by round.
template <int D, int K, int U>
__device__ __forceinline__ void cells_scan_rows(const double *__restrict__ soa, int64_t capacity,
                                                const int32_t *__restrict__ start,
                                                const int32_t *__restrict__ ids, int G,
                                                const double (&qv)[D], TopK<K> &top, bool active,
                                                int xa, int xb, int ya, int yb, int sx0, int sx1,
                                                int sy0, int sy1, int mid, bool centre_out) {
    const unsigned FULL = 0xffffffffu;
    const int nrows = active ? yb - ya + 1 : 0;
    const int maxrows = __reduce_max_sync(FULL, nrows);
    const bool cut = sx1 >= sx0;
    const int nseg = __any_sync(FULL, active && cut) ? 2 : 1;
    // (a software pipeline that requests the ids of the next segment and the range of the one
    // after it ahead of time was measured: no gain -- the kernel is bound by the instructions
    // it issues, not by the latency of its dependent loads)
    for (int j = 0; j < maxrows; ++j) {
        // rows nearest to the query's own row first: the list fills with near nodes, its bound
        // tightens early and fewer of the later nodes reach the insertion (results unchanged)
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
                double d2[U];
                int32_t id[U];
                unsigned mask = 0;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool ok = e + u < e1;
                    id[u] = ok ? ids[e + u] : 0;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    double pv[D];
#pragma unroll
                    for (int c = 0; c < D; ++c) pv[c] = soa[(int64_t)c * capacity + id[u]];
                    d2[u] = dist2_full<D>(pv, qv);
                    mask |= (e + u < e1 && d2[u] <= top.bound2) ? (1u << u) : 0u;
                }
                e += U;
                while (__any_sync(FULL, mask != 0u)) {
                    const int u = mask ? __ffs(mask) - 1 : U;
                    mask &= mask - 1;
                    double dd = INFINITY;
                    int32_t ii = -1;
#pragma unroll
                    for (int v = 0; v < U; ++v) { dd = u == v ? d2[v] : dd; ii = u == v ? id[v] : ii; }
                    top.offer_pair(dd, ii);
                }
            }
        }
    }
}

template <int D, int K, int U>
__global__ void __launch_bounds__(32)
k_knn_cells_t(const double *__restrict__ soa, int64_t capacity, const double *__restrict__ X, int64_t m,
              const double *__restrict__ b4, int G, const int32_t *__restrict__ start,
              const int32_t *__restrict__ ids, int k, int r_init, int limit, int qpw,
              const int32_t *__restrict__ order, double *__restrict__ seed2,
              uint8_t *__restrict__ done, int32_t *__restrict__ pending,
              int64_t *__restrict__ idx, double *__restrict__ dist) {
    // qpw queries per warp (the other lanes idle; experiment knob, 32 is fastest).  `order`: a
    // permutation of the queries; the launcher passes the cell-sorted node ids when there are as
    // many queries as nodes, so that -- when the queries are the stored nodes themselves (PRM
    // build) -- the 32 queries of a warp are neighbours: same cells, same loads, similar lists.
    const int64_t slot = (int64_t)blockIdx.x * qpw + threadIdx.x;
    const bool live = slot < m && (int)threadIdx.x < qpw;
    const int64_t q = live ? (order ? (int64_t)order[slot] : slot) : m;
    double qv[D];
    bool finite = live;
#pragma unroll
    for (int j = 0; j < D; ++j) { qv[j] = live ? X[q * D + j] : 0.0; finite &= qv[j] == qv[j]; }
    const double mnx = b4[0], mxx = b4[1], mny = b4[2], mxy = b4[3];
    TopK<K> top;
    top.init(INFINITY);
    int x0 = 0, x1 = -1, y0 = 0, y1 = -1, cy_q = 0;
    bool have_block = false;
    if (finite && mnx <= mxx && mny <= mxy) {
        const int cx = cell_of(qv[0], mnx, mxx, G), cy = cell_of(qv[1], mny, mxy, G);
        cy_q = cy;
        for (int r = r_init;; r = r < 8 ? r + 1 : 2 * r) {
            x0 = cx - r < 0 ? 0 : cx - r; x1 = cx + r > G - 1 ? G - 1 : cx + r;
            y0 = cy - r < 0 ? 0 : cy - r; y1 = cy + r > G - 1 ? G - 1 : cy + r;
            int32_t c = 0;
            for (int row = y0; row <= y1; ++row)
                c += start[(int64_t)row * G + x1 + 1] - start[(int64_t)row * G + x0];
            have_block = c >= K;
            if (have_block || (x0 == 0 && y0 == 0 && x1 == G - 1 && y1 == G - 1)) break;
        }
    }
    cells_scan_rows<D, K, U>(soa, capacity, start, ids, G, qv, top, have_block, x0, x1, y0, y1, 0, -1, 0, -1,
                          cy_q, true);
    const bool full = have_block && top.i[K - 1] >= 0 && top.bound2 < INFINITY;
    bool handled = false, pass2 = false;
    int xa = 0, xb = -1, ya = 0, yb = -1;
    if (full) {
        const double r = sqrt(top.bound2) * (1.0 + 1e-9) + 1e-300;
        xa = cell_of(__dsub_rd(qv[0], r), mnx, mxx, G); xb = cell_of(__dadd_ru(qv[0], r), mnx, mxx, G);
        ya = cell_of(__dsub_rd(qv[1], r), mny, mxy, G); yb = cell_of(__dadd_ru(qv[1], r), mny, mxy, G);
        if (xa >= x0 && xb <= x1 && ya >= y0 && yb <= y1) {
            handled = true;                                       // the block already covers R
        } else if (yb - ya < 64) {
            int32_t c = 0;
            for (int row = ya; row <= yb; ++row)
                c += start[(int64_t)row * G + xb + 1] - start[(int64_t)row * G + xa];
            pass2 = c <= limit;
            handled = pass2;
        }
    }
    // pass 2: the cells of R outside the block (the block's nodes are in the list already)
    cells_scan_rows<D, K, U>(soa, capacity, start, ids, G, qv, top, pass2, xa, xb, ya, yb, x0, x1, y0, y1, 0, false);
    if (!live) return;
    if (handled) {
#pragma unroll
        for (int t = 0; t < K; ++t)
            if (t < k) {
                idx[q * k + t] = top.i[t];
                if (dist) dist[q * k + t] = top.d[t];
            }
    }
    seed2[q] = full ? top.bound2 : INFINITY;
    done[q] = handled ? 1 : 0;
    if (!handled) atomicAdd(pending, 1);                          // the brute-force kernels have work
}

template <int D, int K>
static int32_t knn_launch(sbp_nodes *s, const double *X, int64_t m, int32_t k, int64_t *idx,
                          double *dist, KnnRowsPart *part = nullptr) {
    sbp_ctx *ctx = s->ctx;
    const int64_t n = s->n;
    bool wide = m < 256;
    if (part) part->ids = nullptr;
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
        // node chunks: with seeded bounds a chunk restart is cheap, so fill the SMs with
        // ~28 warps each (measured: 1.23 -> 1.04 ms on 50k x 50k); without seeds every chunk
        // re-learns its bound (K ln(N/K) insertions), so only enough CTAs to cover the SMs
        const bool seeded = D >= 2 && ctx->tune.knn_seed && n >= 2048;
        int64_t want = seeded ? (28 * sm * 32 / tpb + tiles - 1) / tiles : (2 * sm + tiles - 1) / tiles;
        int64_t maxs = (n + 4095) / 4096;
        S = (int)(want < maxs ? want : maxs);
    }
    if (!wide && ctx->tune.knn_chunks > 0) S = ctx->tune.knn_chunks;
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
                (const double *)pd, (const int32_t *)pi, S, k, idx, dist, s->d_soa, s->capacity, n, D, X);
            ctx->launches++;
            SBP_CUDA(cudaGetLastError());
        }
        return SBP_OK;
    }
    const double *seed = nullptr, *seed_bounds = nullptr;
    void *donebuf = nullptr;                 // per query: 1 = row already written by k_knn_cells
    int32_t *pending_all = nullptr;          // [0]: queries k_knn_cells left to the brute-force kernels
    const int32_t *pending = nullptr;        // ... handed to those kernels when k_knn_cells ran
    if (D >= 2 && ctx->tune.knn_seed && n >= 2048) {
        int G = (int)ceil(sqrt(100.0 / (double)(ctx->tune.knn_cell_density < 1 ? 1 : ctx->tune.knn_cell_density) * (double)n));
        if (G > 2048) G = 2048;
        if (G < 1) G = 1;
        const int64_t L = (int64_t)G * G;
        void *cellbuf = nullptr, *seedbuf = nullptr;
        // [bounds: 4 + 4*BOUNDS_BLOCKS doubles][pending + counters: GRID_HDR_INTS ints][cells L][start L+1][cursor L][ids n][tile sums 1024]
        const size_t hdr = (4 + 4 * BOUNDS_BLOCKS) * sizeof(double);
        rc = sbp_ctx_scratch(ctx, 5, hdr + (size_t)(GRID_HDR_INTS + 3 * L + 1 + n + 1024) * sizeof(int32_t), &cellbuf);
        if (rc) return rc;
        rc = sbp_ctx_scratch(ctx, 6, (size_t)m * sizeof(double), &seedbuf);
        if (rc) return rc;
        double *b4 = (double *)cellbuf;
        pending_all = (int32_t *)((char *)cellbuf + hdr);        // zeroed with the cell counts
        int32_t *cells = pending_all + GRID_HDR_INTS, *start = cells + L, *cursor = start + L + 1,
                *ids = cursor + L, *tsums = ids + n;
        const bool fused_rows = part && part->fused && K <= PRM_FUSED_KMAX && ctx->tune.knn_filter &&
                                ctx->tune.knn_cells;
        int32_t *band = nullptr;
        void *ptsbuf = nullptr;
        int hb = (int)((n + 255) / 256);
        if (hb > ctx->sm_count * 8) hb = ctx->sm_count * 8;
        if (fused_rows) {
            // sbp_prm_connect2d: the whole build in one cooperative launch (k_prm_grid_build)
            rc = sbp_ctx_scratch(ctx, 14, (size_t)n * sizeof(double2), &ptsbuf);
            if (rc) return rc;
            rc = sbp_ctx_scratch(ctx, 11, (size_t)m, &donebuf);
            if (rc) return rc;
            int per_sm = 0;
            SBP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_prm_grid_build, GB_TPB, 0));
            if (per_sm < 1) per_sm = 1;
            if (per_sm > 3) per_sm = 3;
            if (ctx->tune.prm_build_ctas > 0 && ctx->tune.prm_build_ctas < per_sm) per_sm = ctx->tune.prm_build_ctas;
            // (a grid-wide barrier costs more the more CTAs meet at it: small stores get a small grid)
            int nblk = ctx->sm_count * per_sm;
            const int64_t want_blk = (n + 1023) / 1024;
            if (want_blk < nblk) nblk = want_blk < 16 ? 16 : (int)want_blk;
            void *gb = nullptr;
            rc = sbp_ctx_scratch(ctx, 18, (size_t)nblk * (4 * sizeof(double) + sizeof(int32_t)), &gb);
            if (rc) return rc;
            band = pending_all + 1;                                  // two of the spare header ints
            GridBuildArgs ga{s->d_soa, s->capacity, n, b4, (double *)gb, G, pending_all, cells, start, cursor, ids,
                             (double2 *)ptsbuf, (uint8_t *)donebuf, (int32_t *)((double *)gb + 4 * (size_t)nblk),
                             part->p0, part->count, part->banded, 12};
            void *kargs[] = {&ga};
            SBP_CUDA(cudaLaunchCooperativeKernel((void *)k_prm_grid_build, dim3(nblk), dim3(GB_TPB), kargs, 0,
                                                 ctx->stream));
            ctx->launches++;
        } else {
        SBP_CUDA(cudaMemsetAsync(pending_all, 0, (size_t)(L + GRID_HDR_INTS) * sizeof(int32_t), ctx->stream));
        double *partial = b4 + 4;                                  // needs 4 * BOUNDS_BLOCKS doubles
        k_bounds2d_partial<<<BOUNDS_BLOCKS, 256, 0, ctx->stream>>>(s->d_soa, s->capacity, n, partial);
        k_cell_hist<<<hb, 256, 0, ctx->stream>>>(s->d_soa, s->capacity, n, partial, BOUNDS_BLOCKS, b4, G,
                                                 cells);
        const int nt = (int)((L + SCAN_TILE - 1) / SCAN_TILE);            // <= 1024 tiles (G <= 2048)
        if (nt <= SCAN_FUSED_TILES) {
            k_cell_scan_fused<<<nt, 1024, 0, ctx->stream>>>(cells, L, start, cursor);
            ctx->launches -= 2;
        } else {
            k_cell_tile_sums<<<nt, 1024, 0, ctx->stream>>>(cells, L, tsums);
            k_cell_scan_sums<<<1, 1024, 0, ctx->stream>>>(tsums, nt, start + L);
            k_cell_scan_apply<<<nt, 1024, 0, ctx->stream>>>(cells, L, tsums, start, cursor);
        }
        k_cell_scatter<<<hb, 256, 0, ctx->stream>>>(s->d_soa, s->capacity, n, b4, G, cursor, ids, nullptr);
        if (part && part->deterministic) {
            int sbk = (int)((L + 255) / 256);
            if (sbk > ctx->sm_count * 8) sbk = ctx->sm_count * 8;
            k_cell_sort_segments<<<sbk, 256, 0, ctx->stream>>>(start, L, ids);
            ctx->launches++;
        }
        ctx->launches += 6;
        }
        // (16 lanes per query were measured too: 53 us instead of 49 us on cfg2 -- the kernel is
        // bound by its chain of dependent loads start -> ids -> node, not by the lane count)
        if (ctx->tune.knn_filter && ctx->tune.knn_cells) {
            rc = sbp_ctx_scratch(ctx, 11, (size_t)m, &donebuf);
            if (rc) return rc;
            // rows part: only the cell-sorted positions [p0, p0 + count) are answered here; every
            // other row counts as done for the kernels behind (their work list is `done == 0`)
            const bool rows_part = part != nullptr;
            const int64_t m_cells = rows_part ? part->count : m;
            const int32_t *order = rows_part ? ids + part->p0
                                             : ((ctx->tune.knn_cells_order && m == n) ? ids : nullptr);
            if (rows_part) {
                if (!fused_rows) SBP_CUDA(cudaMemsetAsync(donebuf, 1, (size_t)m, ctx->stream));
                part->ids = ids;
            }
            KernelTimer timer(ctx, fused_rows ? -1 : 2);
            // (measured on 50 000 nodes, k = 11, whole call: 4096 queries 44 us warp form / 66 us thread
            // form, 16 384 queries 73 / 74 us, 50 000 queries 137 / 71 us: a one-warp CTA's chain of
            // 32 queries takes ~50 us however few CTAs there are)
            if (fused_rows) {
                // sbp_prm_connect2d: the cell-grid answer of a row AND the visibility of its k candidate
                // edges in one kernel (prm.cu); rows it does not settle keep done = 0 like below
                const double per_cell = (double)n / ((double)G * (double)G);
                int r_init = (int)ceil((sqrt(2.0 * (double)K / per_cell) - 1.0) / 2.0);
                r_init = r_init < 1 ? 1 : r_init > 8 ? 8 : r_init;
                if (ctx->tune.knn_cells_rinit > 0) r_init = ctx->tune.knn_cells_rinit;
                CellGridView gv{b4, G, start, ids, (const double2 *)ptsbuf, n, s->d_maxabs, band};
                rc = sbp_prm_fused_launch(ctx, gv, K, k, r_init, ctx->tune.knn_cells_limit, part->p0, part->count,
                                          (double *)seedbuf, (uint8_t *)donebuf, pending_all, part->fused);
                if (rc) return rc;
                part->fused_ran = 1;
                part->done = (const uint8_t *)donebuf;
                part->pending = pending_all;
            } else if (rows_part || ctx->tune.knn_cells >= 3 || (ctx->tune.knn_cells == 2 && m >= 20000)) {
                // first ring radius whose block is expected to hold 2 K nodes (then the rectangle of
                // the K-th distance usually lies inside the block and there is no second pass)
                const double per_cell = (double)n / ((double)G * (double)G);
                int r_init = (int)ceil((sqrt(2.0 * (double)K / per_cell) - 1.0) / 2.0);
                r_init = r_init < 1 ? 1 : r_init > 8 ? 8 : r_init;
                if (ctx->tune.knn_cells_rinit > 0) r_init = ctx->tune.knn_cells_rinit;
                int qpw = ctx->tune.knn_cells_qpw;
                // (measured on cfg2: 16 / 8 queries per warp take 0.094 / 0.113 ms instead of 0.075:
                // the kernel is bound by the instructions it issues, not by the length of a warp's chain)
                if (qpw <= 0 || qpw > 32) qpw = 32;
                // (8 nodes per lane and step were measured too: 0.077 instead of 0.076 ms on cfg2)
                k_knn_cells_t<D, K, 4><<<(unsigned)((m_cells + qpw - 1) / qpw), 32, 0, ctx->stream>>>(
                    s->d_soa, s->capacity, X, m_cells, b4, G, start, ids, k, r_init, ctx->tune.knn_cells_limit / 4, qpw,
                    order, (double *)seedbuf, (uint8_t *)donebuf, pending_all, idx, dist);
            } else {
                k_knn_cells<D><<<(unsigned)((m * 32 + 255) / 256), 256, 0, ctx->stream>>>(
                    s->d_soa, s->capacity, X, m, b4, G, start, ids, k, ctx->tune.knn_cells_limit, (double *)seedbuf,
                    (uint8_t *)donebuf, pending_all, idx, dist);
            }
            pending = pending_all;
        } else {
            k_knn_seed<D, 32><<<(unsigned)((m * 32 + 255) / 256), 256, 0, ctx->stream>>>(
                s->d_soa, s->capacity, X, m, b4, G, start, ids, K, (double *)seedbuf);
        }
        if (!fused_rows) ctx->launches += 1;
        SBP_CUDA(cudaGetLastError());
        seed = (const double *)seedbuf;
        seed_bounds = b4;
    }
    const uint8_t *redo = nullptr;
    if (seed && ctx->tune.knn_filter) {
        // seeded filter + exact refine; the one-kernel scan below then only redoes the
        // blocks of 32 queries the refine step flagged (normally none)
        const int NW = (ctx->tune.knn_filter_variant & 3) >= 2 ? 4 : 2;    // warps per CTA
        const int FQv = ctx->tune.knn_filter_variant >= 8 ? 8 : ctx->tune.knn_filter_variant >= 4 ? 2 : FILTER_Q;
        const int QPB = NW * 32 * FQv;                             // queries per CTA
        const int64_t qblocks = (m + QPB - 1) / QPB;
        constexpr int FT = filter_tile(D);
        const int64_t tiles = (n + FT - 1) / FT;
        int64_t FS = ctx->tune.knn_filter_chunks > 0 ? ctx->tune.knn_filter_chunks
                                             : (16 * (int64_t)ctx->sm_count + qblocks - 1) / qblocks;
        if (FS > tiles) FS = tiles;
        if (FS > 64) FS = 64;
        if (FS < 1) FS = 1;
        const int64_t chunk = (tiles + FS - 1) / FS * FT;
        FS = (n + chunk - 1) / chunk;
        int C = 4 * K / (int)FS + 12;
        const int64_t cap32 = tiles * FT;
        void *cbuf = nullptr, *nbuf = nullptr, *shbuf = nullptr;
        rc = sbp_ctx_scratch(ctx, 8, (size_t)m * FS * C * sizeof(int32_t), &cbuf);
        if (rc) return rc;
        const size_t nflags = (size_t)((m + 31) / 32);
        rc = sbp_ctx_scratch(ctx, 9, (size_t)m * FS * sizeof(int32_t) + nflags, &nbuf);
        if (rc) return rc;
        const size_t shdr = 128 + 128 + (size_t)BOUNDS_BLOCKS * 2 * D * sizeof(double);   // info | bounds | partials
        rc = sbp_ctx_scratch(ctx, 10, shdr + (size_t)(D + 1) * cap32 * sizeof(float), &shbuf);
        if (rc) return rc;
        double *info = (double *)shbuf;                            // [1 + 2 D] doubles
        float *shadow = (float *)((char *)shbuf + shdr);
        if (D > 2) {               // the cell grid only bounds x and y: all D coordinates here
            double *bnd = (double *)((char *)shbuf + 128), *partial = (double *)((char *)shbuf + 256);
            k_bounds_partial<D><<<BOUNDS_BLOCKS, 256, 0, ctx->stream>>>(s->d_soa, s->capacity, n, partial,
                                                                        pending);
            k_bounds_fold<D><<<1, 32, 0, ctx->stream>>>(partial, BOUNDS_BLOCKS, bnd, pending);
            ctx->launches += 2;
            seed_bounds = bnd;
        }
        int32_t *ccount = (int32_t *)nbuf;
        uint8_t *flags = (uint8_t *)nbuf + (size_t)m * FS * sizeof(int32_t);
        SBP_CUDA(cudaMemsetAsync(flags, 0, nflags, ctx->stream));
        int sb = (int)((cap32 + 255) / 256);
        if (sb > ctx->sm_count * 8) sb = ctx->sm_count * 8;
        k_shadow_centered<D><<<sb, 256, 0, ctx->stream>>>(s->d_soa, s->capacity, n, seed_bounds, cap32,
                                                          shadow, info, pending);
        dim3 fgrid((unsigned)qblocks, (unsigned)FS);
#define SBP_FILTER_LAUNCH(NW_, FQ_)                                                         \
    k_knn_filter<D, NW_, FQ_><<<fgrid, NW_ * 32, 0, ctx->stream>>>(                          \
        shadow, info, cap32, n, (int)FS, chunk, X, m, seed, (const uint8_t *)donebuf, pending, \
        C, (int32_t *)cbuf, ccount)
        {
        KernelTimer timer(ctx);
        switch (ctx->tune.knn_filter_variant) {
            case 2: SBP_FILTER_LAUNCH(4, FILTER_Q); break;
            case 4: SBP_FILTER_LAUNCH(2, 2); break;
            case 6: SBP_FILTER_LAUNCH(4, 2); break;
            case 8: SBP_FILTER_LAUNCH(2, 8); break;
            case 10: SBP_FILTER_LAUNCH(4, 8); break;
            default: SBP_FILTER_LAUNCH(2, FILTER_Q); break;
        }
        }
#undef SBP_FILTER_LAUNCH
        ctx->launches++;
        const int64_t rb = (m + 63) / 64, rcap = (int64_t)ctx->sm_count * 32;
        k_knn_refine<D, K><<<(unsigned)(rb < rcap ? rb : rcap), 64, 0, ctx->stream>>>(
            s->d_soa, s->capacity, n, X, m, seed, (int)FS, C, (const int32_t *)cbuf, ccount, k, idx,
            dist, flags, (const uint8_t *)donebuf, pending);
        ctx->launches += 2;
        SBP_CUDA(cudaGetLastError());
        redo = flags;
        tpb = 32;
    }
    // (1-D grids capped at 64 CTAs per SM: the kernels loop over their blocks)
    const int64_t bcap = (int64_t)ctx->sm_count * 64;
    if (tpb == 32) {
        const int64_t nb = (m + 31) / 32 * S;
        const unsigned grid = (unsigned)(nb < bcap ? nb : bcap);
        k_knn_batched<D, K, 32><<<grid, 32, 0, ctx->stream>>>(s->d_soa, s->d_soa32, s->d_maxabs,
                                                              ctx->tune.knn_prefilter, s->capacity, n, S, X,
                                                              m, seed, (double *)pd, (int32_t *)pi,
                                                              redo, redo ? pending : nullptr);
    } else {
        const int64_t nb = (m + 127) / 128 * S;
        const unsigned grid = (unsigned)(nb < bcap ? nb : bcap);
        k_knn_batched<D, K, 128><<<grid, 128, 0, ctx->stream>>>(s->d_soa, s->d_soa32, s->d_maxabs,
                                                                ctx->tune.knn_prefilter, s->capacity, n, S,
                                                                X, m, seed, (double *)pd, (int32_t *)pi,
                                                                nullptr);
    }
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    const int64_t mb = (m + 127) / 128, mcap = (int64_t)ctx->sm_count * 16;
    k_knn_merge<<<(unsigned)(mb < mcap ? mb : mcap), 128, 0, ctx->stream>>>(
        (const double *)pd, (const int32_t *)pi, m, S, K, k, idx, dist, redo, s->d_soa, s->capacity, n, D, X,
        redo ? pending : nullptr);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

template <int D>
static int32_t knn_dispatch_k(sbp_nodes *s, const double *X, int64_t m, int32_t k, int64_t *idx,
                              double *dist, KnnRowsPart *part = nullptr) {
    // list length K >= k; 11 and 21 are the PRM / RRT* cases (k = 10 or 20 plus self)
    if (k <= 1) return knn_launch<D, 1>(s, X, m, k, idx, dist, part);
    if (k <= 2) return knn_launch<D, 2>(s, X, m, k, idx, dist, part);
    if (k <= 4) return knn_launch<D, 4>(s, X, m, k, idx, dist, part);
    if (k <= 8) return knn_launch<D, 8>(s, X, m, k, idx, dist, part);
    if (k <= 11) return knn_launch<D, 11>(s, X, m, k, idx, dist, part);
    if (k <= 16) return knn_launch<D, 16>(s, X, m, k, idx, dist, part);
    if (k <= 21) return knn_launch<D, 21>(s, X, m, k, idx, dist, part);
    return knn_launch<D, 32>(s, X, m, k, idx, dist, part);
}

// kNN of the stored rows themselves for sbp_prm_connect2d (2-D stores): X_all = the n row poses,
// idx [n][k]; see KnnRowsPart.  Without a cell grid (small stores, knobs off) every row is answered.
bool sbp_knn_rows_use_grid(const sbp_nodes *s) {
    const sbp_ctx *ctx = s->ctx;
    return s->d == 2 && ctx->tune.knn_seed && ctx->tune.knn_filter && ctx->tune.knn_cells && s->n >= 2048;
}

int32_t sbp_knn_rows_part(sbp_nodes *s, const double *X_all, int32_t k, int64_t *idx, KnnRowsPart *part) {
    return knn_dispatch_k<2>(s, X_all, s->n, k, idx, nullptr, sbp_knn_rows_use_grid(s) ? part : nullptr);
}

// np.argmin semantics for find_nearest_neighbour_idx (rrtPlanner.py:278-279): a NaN distance is
// returned as soon as it is seen, i.e. the FIRST node whose distance to the query is NaN wins
// (the argsort rule of sbp_nodes_knn ranks NaN last).  One CTA per query; returns at once when
// neither the store (maxabs, +inf once a non-finite coordinate was appended) nor the query can
// produce a NaN distance.
template <int D>
__global__ void __launch_bounds__(256)
k_nearest_first_nan(const double *__restrict__ soa, int64_t capacity, int64_t n,
                    const double *__restrict__ maxabs, const double *__restrict__ X,
                    int64_t *__restrict__ idx, double *__restrict__ dist) {
    const int64_t q = blockIdx.x;
    double qv[D];
    bool q_finite = true;
#pragma unroll
    for (int j = 0; j < D; ++j) { qv[j] = X[q * D + j]; q_finite &= isfinite(qv[j]); }
    if (q_finite && isfinite(*maxabs)) return;
    __shared__ unsigned long long s_first;
    if (threadIdx.x == 0) s_first = ~0ull;
    __syncthreads();
    unsigned long long mine = ~0ull;
    for (int64_t t = threadIdx.x; t < n && (unsigned long long)t < mine; t += blockDim.x) {
        double pv[D];
#pragma unroll
        for (int j = 0; j < D; ++j) pv[j] = soa[(int64_t)j * capacity + t];
        const double d2 = dist2_full<D>(pv, qv);
        if (d2 != d2) { mine = (unsigned long long)t; break; }
    }
    if (mine != ~0ull) atomicMin(&s_first, mine);
    __syncthreads();
    if (threadIdx.x == 0 && s_first != ~0ull) {
        idx[q] = (int64_t)s_first;
        if (dist) dist[q] = __longlong_as_double(0x7ff8000000000000LL);
    }
}

extern "C" int32_t sbp_nodes_nearest(sbp_nodes *s, const double *X, int64_t m, int64_t *idx, double *dist) {
    SBP_NVTX("sbp_nodes_nearest");
    SBP_REQUIRE(s && (m == 0 || (X && idx)), "sbp_nodes_nearest: NULL argument");
    SBP_REQUIRE(m >= 0, "sbp_nodes_nearest: m < 0");
    if (m == 0) return SBP_OK;
    int32_t rc = sbp_nodes_knn(s, X, m, 1, 64, idx, dist);
    if (rc || s->n == 0) return rc;
    sbp_ctx *ctx = s->ctx;
    SBP_DEVICE(ctx);
    switch (s->d) {
        case 2: k_nearest_first_nan<2><<<(unsigned)m, 256, 0, ctx->stream>>>(s->d_soa, s->capacity, s->n, s->d_maxabs, X, idx, dist); break;
        case 3: k_nearest_first_nan<3><<<(unsigned)m, 256, 0, ctx->stream>>>(s->d_soa, s->capacity, s->n, s->d_maxabs, X, idx, dist); break;
        case 4: k_nearest_first_nan<4><<<(unsigned)m, 256, 0, ctx->stream>>>(s->d_soa, s->capacity, s->n, s->d_maxabs, X, idx, dist); break;
        default: k_nearest_first_nan<6><<<(unsigned)m, 256, 0, ctx->stream>>>(s->d_soa, s->capacity, s->n, s->d_maxabs, X, idx, dist); break;
    }
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

__global__ void k_fill_empty_knn(int64_t total, int64_t *idx, double *dist) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        idx[t] = -1;
        if (dist) dist[t] = INFINITY;
    }
}

__global__ void k_dist_to_f32(const double *__restrict__ d64, int64_t total, float *__restrict__ d32) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
        d32[t] = __double2float_rn(d64[t]);
}

extern "C" int32_t sbp_nodes_knn(sbp_nodes *s, const double *X, int64_t m, int32_t k,
                                 int32_t precision, int64_t *idx, double *dist) {
    SBP_NVTX("sbp_nodes_knn");
    SBP_REQUIRE(s && (m == 0 || (X && idx)), "sbp_nodes_knn: NULL argument");
    SBP_REQUIRE(m >= 0, "sbp_nodes_knn: m < 0");
    SBP_REQUIRE(k >= 1 && k <= 32, "sbp_nodes_knn: 1 <= k <= 32");
    SBP_REQUIRE(precision == 64 || precision == 32, "sbp_nodes_knn: precision is 64 or 32");
    if (m == 0) return SBP_OK;
    SBP_DEVICE(s->ctx);
    // precision 32: `dist` is a float buffer.  The selection itself stays the exact one -- the cell-grid
    // search is bound by its list insertions and dependent loads, not by fp64 arithmetic, so an fp32
    // selection would buy nothing and lose the index-exactness at ties -- and the distances are
    // delivered rounded to fp32 (half the output bytes): |d32 - d64| <= 2^-24 d64.
    double *dist64 = dist;
    if (precision == 32 && dist) {
        void *tmp = nullptr;
        int32_t rc = sbp_ctx_scratch(s->ctx, 17, (size_t)m * k * sizeof(double), &tmp);
        if (rc) return rc;
        dist64 = (double *)tmp;
    }
    int32_t rc = SBP_OK;
    if (s->n == 0) {
        k_fill_empty_knn<<<64, 256, 0, s->ctx->stream>>>(m * k, idx, dist64);
        s->ctx->launches++;
        SBP_CUDA(cudaGetLastError());
    } else {
        switch (s->d) {
            case 2: rc = knn_dispatch_k<2>(s, X, m, k, idx, dist64); break;
            case 3: rc = knn_dispatch_k<3>(s, X, m, k, idx, dist64); break;
            case 4: rc = knn_dispatch_k<4>(s, X, m, k, idx, dist64); break;
            case 6: rc = knn_dispatch_k<6>(s, X, m, k, idx, dist64); break;
            default:
                sbp_set_error("sbp_nodes_knn: dimension %d not instantiated (2, 3, 4, 6)", s->d);
                return SBP_ERR_ARG;
        }
    }
    if (rc == SBP_OK && precision == 32 && dist) {
        int blocks = (int)((m * k + 255) / 256);
        if (blocks > s->ctx->sm_count * 8) blocks = s->ctx->sm_count * 8;
        k_dist_to_f32<<<blocks, 256, 0, s->ctx->stream>>>(dist64, m * k, (float *)dist);
        s->ctx->launches++;
        SBP_CUDA(cudaGetLastError());
    }
    return rc;
}

