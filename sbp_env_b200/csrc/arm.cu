// arm.cu -- K3: 4-DoF stick-arm link sweeps against the occupancy grid.
//
// Reference behaviour restated (file:line under /root/reference/sbp_env/):
//   RobotArm4dCollisionChecker.feasible               collisionChecker.py:404-426
//   .get_pt_from_angle_and_length                     collisionChecker.py:428-441
//   ._pt_feasible                                     collisionChecker.py:393-402
//   ._get_line (Bresenham, list variant)              collisionChecker.py:443-503
//   .create_ranges / ._interpolate_configs / .visible collisionChecker.py:346-391
//
// fp64 policy.  The angle lerp is three separately rounded fp64
// operations and is reproduced bit-for-bit with __dmul_rn/__dadd_rn.  The forward
// kinematics go through cos/sin: the reference calls glibc (<= 1 ulp), the device
// libdevice (<= 2 ulp), so the two can differ in the last bits of
// pt + L*cos(a).  That only matters when such a coordinate lies within a few ulp
// of an integer, where int() may truncate differently.  Every FK coordinate is
// therefore compared with a guard band ~100x wider than the worst-case
// discrepancy; a config inside the band is not decided on the device.  The result
// byte is SBP_AMBIGUOUS unless another, unambiguous config already blocks the
// edge, and sbp_arm_resolve_host() settles it with the host libm (the very libm
// CPython's math.cos/sin call).  On uniform random configs the band is hit
// ~1e-8 of the time; exact multiples of pi/2 on integer bases hit it always.
#include <math.h>

#include "common.cuh"
#include "device_fns.cuh"

template <bool SMEM>
__global__ void __launch_bounds__(1024)
k_arm_feasible(MapView mv, double L0, double L1, const double4 *__restrict__ C, int64_t n,
               uint8_t *__restrict__ out, int32_t *__restrict__ flags) {
    extern __shared__ __align__(16) uint32_t smap[];
    __shared__ uint64_t bar;
    if (SMEM) stage_map_to_smem(smap, mv, &bar);
    const uint32_t *words = SMEM ? smap : mv.words;
    int myflag = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        double4 c = C[i];
        out[i] = (uint8_t)arm_config<SMEM, true>(words, mv, L0, L1, c.x, c.y, c.z, c.w, myflag);
    }
    if (myflag && flags) atomicOr(flags, myflag);
}

// visible: a group of G lanes owns one edge; lane j evaluates interpolated
// configs j, j+G, ... (one config = base pixel o of the base Bresenham line in
// start->end order + lerped angles), ballot for the early exit.
template <int G, bool SMEM, bool COUNT>
__global__ void __launch_bounds__(1024)
k_arm_visible(MapView mv, double L0, double L1, const double4 *__restrict__ C1,
              const double4 *__restrict__ C2, int64_t n, uint8_t *__restrict__ out,
              int32_t *__restrict__ flags, unsigned long long *__restrict__ cfg_tested) {
    extern __shared__ __align__(16) uint32_t smap[];
    __shared__ uint64_t bar;
    if (SMEM) stage_map_to_smem(smap, mv, &bar);
    const uint32_t *words = SMEM ? smap : mv.words;

    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int j = lane & (G - 1);
    const int gshift = lane & ~(G - 1);
    const unsigned gmask_bits = G == 32 ? FULL : ((1u << G) - 1u);
    const int64_t groups_per_grid = ((int64_t)gridDim.x * blockDim.x) / G;
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int64_t n_round = ((n + (32 / G) - 1) / (32 / G)) * (32 / G);
    int myflag = 0;
    unsigned long long tested = 0;

    for (int64_t eidx = gid; eidx < n_round; eidx += groups_per_grid) {
        const bool valid = eidx < n;
        double4 a = make_double4(0, 0, 0, 0), b = make_double4(0, 0, 0, 0);
        if (valid) { a = C1[eidx]; b = C2[eidx]; }
        // base line between int()-truncated base points (:373, :461-462)
        TruncResult t0 = trunc_index(a.x, mv.W), t1 = trunc_index(a.y, mv.H);
        TruncResult t2 = trunc_index(b.x, mv.W), t3 = trunc_index(b.y, mv.H);
        int f = t0.flag ? t0.flag : t1.flag ? t1.flag : t2.flag ? t2.flag : t3.flag;
        if (!f && (isinf(a.z) || isinf(a.w) || isinf(b.z) || isinf(b.w) || a.z != a.z ||
                   a.w != a.w || b.z != b.z || b.w != b.w))
            f = SBP_FLAG_NAN;
        if (valid) myflag |= f;
        // config 0 / config N-1 sit on the base end pixels: an out-of-range base
        // endpoint blocks the edge
        bool walk = valid && !f && t0.in_range && t1.in_range && t2.in_range && t3.in_range;
        int x1 = t0.i, y1 = t1.i, x2 = t2.i, y2 = t3.i;
        int dx = x2 - x1, dy = y2 - y1;
        bool steep = abs(dy) > abs(dx);
        int a1 = steep ? y1 : x1, b1 = steep ? x1 : y1;
        int a2 = steep ? y2 : x2, b2 = steep ? x2 : y2;
        bool swapped = a1 > a2;
        if (swapped) { int t = a1; a1 = a2; a2 = t; t = b1; b1 = b2; b2 = t; }
        uint32_t da = (uint32_t)(a2 - a1), adb = (uint32_t)abs(b2 - b1);
        int ystep = b1 < b2 ? 1 : -1;
        const uint32_t npx = da + 1u;
        const uint32_t N = npx < 2u ? 2u : npx;                     // :375-377
        const uint32_t c = da ? da - 1u - (da >> 1) : 0u;
        // create_ranges (:363-364): steps = (1.0/(N-1))*(stop-start); rot = steps*i + start
        const double inv = __ddiv_rn(1.0, (double)(N - 1u));
        const double st0 = __dmul_rn(inv, __dsub_rn(b.z, a.z));
        const double st1 = __dmul_rn(inv, __dsub_rn(b.w, a.w));

        int state = walk ? 1 : 0;   // 1 free so far, 0 blocked, 2 ambiguous seen
        bool done = !walk;
        uint32_t o = (uint32_t)j;
        while (true) {
            bool act = !done && o < N;
            int r = 1;
            if (act) {
                uint32_t po = npx == 1u ? 0u : o;                   // duplicated single pixel
                uint64_t ci = swapped ? (uint64_t)(npx - 1u - po) : (uint64_t)po;
                uint32_t k = da ? (uint32_t)((ci * adb + c) / da) : 0u;
                int aa = a1 + (int)ci, bb = b1 + ystep * (int)k;
                double bx = (double)(steep ? bb : aa), by = (double)(steep ? aa : bb);
                double r0 = __dadd_rn(__dmul_rn(st0, (double)o), a.z);
                double r1 = __dadd_rn(__dmul_rn(st1, (double)o), a.w);
                r = arm_config<SMEM, true>(words, mv, L0, L1, bx, by, r0, r1, myflag);
                if (COUNT) tested += 1ull;
            }
            unsigned blocked = (__ballot_sync(FULL, act && r == 0) >> gshift) & gmask_bits;
            unsigned ambig = (__ballot_sync(FULL, act && r == 2) >> gshift) & gmask_bits;
            if (ambig && state == 1) state = 2;
            if (blocked) { state = 0; done = true; }     // an unambiguous block decides the edge
            o += (uint32_t)G;
            if (o >= N) done = true;
            if (!__any_sync(FULL, !done)) break;
        }
        if (valid && j == 0) out[eidx] = (uint8_t)state;
    }
    if (myflag && flags) atomicOr(flags, myflag);
    if (COUNT) {
        for (int o = 16; o > 0; o >>= 1) tested += __shfl_down_sync(FULL, tested, o);
        if (lane == 0 && tested && cfg_tested) atomicAdd(cfg_tested, tested);
    }
}

// visible, flattened form (the default): the lane-group kernel above leaves lanes idle whenever an
// edge's number of interpolated configs is not a multiple of the group (RRT-style edges have 7-11:
// 12.4 of 32 lanes active, profiles/r1_arm_visible_full.txt).  Here a CTA takes a block of
// AF_EDGES edges, sets them up once (thread = edge: truncation, base line, angle steps, N), prefix-
// sums their N, and then ALL threads walk the flattened list of (edge, config) pairs -- every lane
// has a config until the block's last ones.  A config looks up its edge by binary search in the
// prefix array; results are OR-ed into the edge's flags (bit 0 = some config blocked, bit 1 = some
// config ambiguous: blocked wins, as in the kernel above); configs of an edge already known to be
// blocked are skipped, which keeps most of the early exit.
#define AF_EDGES 256
struct ArmEdge {
    int32_t a1, b1;
    uint32_t da, adb, npx;
    int32_t flags;            // bit 0 steep, bit 1 swapped, bit 2 minor step +1
    double az, aw, st0, st1;
};

template <bool SMEM>
__global__ void __launch_bounds__(1024)
k_arm_visible_flat(MapView mv, double L0, double L1, const double4 *__restrict__ C1,
                   const double4 *__restrict__ C2, int64_t n, uint8_t *__restrict__ out,
                   int32_t *__restrict__ flags) {
    extern __shared__ __align__(16) uint32_t smap[];
    __shared__ uint64_t bar;
    __shared__ ArmEdge s_e[AF_EDGES];
    __shared__ uint32_t s_pref[AF_EDGES + 1];
    __shared__ int32_t s_f[AF_EDGES];
    __shared__ uint32_t s_ws[AF_EDGES / 32];
    if (SMEM) stage_map_to_smem(smap, mv, &bar);
    const uint32_t *words = SMEM ? smap : mv.words;
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    int myflag = 0;
    for (int64_t base = (int64_t)blockIdx.x * AF_EDGES; base < n; base += (int64_t)gridDim.x * AF_EDGES) {
        // ---- phase 1: set up the block's edges
        uint32_t N = 0;
        if (tid < AF_EDGES) {
            const int64_t eidx = base + tid;
            const bool valid = eidx < n;
            double4 a = make_double4(0, 0, 0, 0), b = make_double4(0, 0, 0, 0);
            if (valid) { a = C1[eidx]; b = C2[eidx]; }
            // base line between int()-truncated base points (:373, :461-462)
            TruncResult t0 = trunc_index(a.x, mv.W), t1 = trunc_index(a.y, mv.H);
            TruncResult t2 = trunc_index(b.x, mv.W), t3 = trunc_index(b.y, mv.H);
            int f = t0.flag ? t0.flag : t1.flag ? t1.flag : t2.flag ? t2.flag : t3.flag;
            if (!f && (isinf(a.z) || isinf(a.w) || isinf(b.z) || isinf(b.w) || a.z != a.z || a.w != a.w ||
                       b.z != b.z || b.w != b.w))
                f = SBP_FLAG_NAN;
            if (valid) myflag |= f;
            // an out-of-range base endpoint blocks the edge (configs 0 / N-1 sit on the base end pixels)
            const bool walk = valid && !f && t0.in_range && t1.in_range && t2.in_range && t3.in_range;
            int x1 = t0.i, y1 = t1.i, x2 = t2.i, y2 = t3.i;
            int dx = x2 - x1, dy = y2 - y1;
            const bool steep = abs(dy) > abs(dx);
            int a1 = steep ? y1 : x1, b1 = steep ? x1 : y1;
            int a2 = steep ? y2 : x2, b2 = steep ? x2 : y2;
            const bool swapped = a1 > a2;
            if (swapped) { int t = a1; a1 = a2; a2 = t; t = b1; b1 = b2; b2 = t; }
            ArmEdge e;
            e.a1 = a1; e.b1 = b1;
            e.da = (uint32_t)(a2 - a1); e.adb = (uint32_t)abs(b2 - b1);
            e.npx = e.da + 1u;
            e.flags = (steep ? 1 : 0) | (swapped ? 2 : 0) | (b1 < b2 ? 4 : 0);
            const uint32_t NN = e.npx < 2u ? 2u : e.npx;                 // :375-377
            // create_ranges (:363-364): steps = (1.0/(N-1))*(stop-start); rot = steps*i + start
            const double inv = __ddiv_rn(1.0, (double)(NN - 1u));
            e.az = a.z; e.aw = a.w;
            e.st0 = __dmul_rn(inv, __dsub_rn(b.z, a.z));
            e.st1 = __dmul_rn(inv, __dsub_rn(b.w, a.w));
            s_e[tid] = e;
            N = walk ? NN : 0u;
            s_f[tid] = walk ? 0 : 1;                                     // not walked = blocked
        }
        // exclusive prefix of N over the block's edges
        if (tid < AF_EDGES) {
            uint32_t inc = N;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += u; }
            if (lane == 31) s_ws[w] = inc;
            s_pref[tid + 1] = inc;                                        // (warp-local for now)
        }
        __syncthreads();
        if (tid < AF_EDGES) {
            uint32_t off = 0;
            for (int i = 0; i < w; ++i) off += s_ws[i];
            s_pref[tid + 1] += off;
            if (tid == 0) s_pref[0] = 0;
        }
        __syncthreads();
        const uint32_t T = s_pref[AF_EDGES];
        // ---- phase 2: all threads over the flattened (edge, config) pairs
        for (uint32_t c = tid; c < T; c += blockDim.x) {
            int lo = 0, hi = AF_EDGES;                                   // largest j with s_pref[j] <= c
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_pref[mid] <= c) lo = mid; else hi = mid;
            }
            if (s_f[lo] & 1) continue;                                   // edge already blocked
            const ArmEdge e = s_e[lo];
            const uint32_t o = c - s_pref[lo];
            const uint32_t po = e.npx == 1u ? 0u : o;                    // duplicated single pixel
            const bool steep = e.flags & 1, swapped = e.flags & 2;
            const uint64_t ci = swapped ? (uint64_t)(e.npx - 1u - po) : (uint64_t)po;
            const uint32_t cc = e.da ? e.da - 1u - (e.da >> 1) : 0u;
            const uint32_t k = e.da ? (uint32_t)((ci * e.adb + cc) / e.da) : 0u;
            const int aa = e.a1 + (int)ci, bb = e.b1 + ((e.flags & 4) ? 1 : -1) * (int)k;
            const double bx = (double)(steep ? bb : aa), by = (double)(steep ? aa : bb);
            const double r0 = __dadd_rn(__dmul_rn(e.st0, (double)o), e.az);
            const double r1 = __dadd_rn(__dmul_rn(e.st1, (double)o), e.aw);
            const int r = arm_config<SMEM, true>(words, mv, L0, L1, bx, by, r0, r1, myflag);
            if (r != 1) atomicOr(&s_f[lo], r == 0 ? 1 : 2);
        }
        __syncthreads();
        // ---- phase 3: an unambiguous block decides the edge, else an ambiguous config leaves it open
        if (tid < AF_EDGES && base + tid < n) {
            const int f = s_f[tid];
            out[base + tid] = (uint8_t)((f & 1) ? 0 : (f & 2) ? 2 : 1);
        }
        __syncthreads();
    }
    if (myflag && flags) atomicOr(flags, myflag);
}

// visible, flattened form WITH the free-box test of the 2-D path (the default where the map has tile
// codes and fits 16-bit coordinates).  k_arm_visible_flat spends its time in the two link sweeps of
// every config (<= 31 pixels each, lanes of different length, 18.4 of 32 active).  A link is a
// Bresenham line: when the bounding box of its end pixels is free of blocked pixels (two words of
// the free-run planes, device_fns.cuh:tiles_rect_free) it is free without a sweep -- 77 % of the
// configs of tree-like edges on maps/4d.png have both links settled that way.  Phase A (thread =
// config): lerp, forward kinematics, guard bands, ranges (arm_config_pre) and the two box tests;
// configs with a link left are listed in shared memory.  Phase B, whenever the list holds a
// block's worth of entries: the listed links are swept side by side (full lanes).  Same results
// as arm_config for every map: a free box proves a free link, everything else is swept.
struct ArmLink {               // 12 bytes; coordinates in [-W, W) fit 16 bits (W, H <= 32767)
    int16_t x0, y0, x1, y1;
    uint16_t edge, pad;      // edge of the CTA's block
};
#define AB_LIST 96             // entries per warp: fewer than 32 left over + at most 2 per lane and step

template <bool SMEM>
__global__ void __launch_bounds__(1024)
k_arm_visible_boxes(MapView mv, TileCodeView af, double L0, double L1, const double4 *__restrict__ C1,
                    const double4 *__restrict__ C2, int64_t n, uint8_t *__restrict__ out,
                    int32_t *__restrict__ flags) {
    extern __shared__ __align__(16) uint32_t smap[];
    __shared__ uint64_t bar;
    __shared__ ArmEdge s_e[AF_EDGES];
    __shared__ uint32_t s_pref[AF_EDGES + 1];
    __shared__ int32_t s_f[AF_EDGES];
    __shared__ uint32_t s_ws[AF_EDGES / 32];
    if (SMEM) stage_map_to_smem(smap, mv, &bar);
    const uint32_t *words = SMEM ? smap : mv.words;
    // [AB_LIST per warp] entries behind the staged map (dynamic shared memory; the map is padded to 16 bytes)
    ArmLink *s_todo = reinterpret_cast<ArmLink *>(smap + (SMEM ? mv.n_words : 0u));
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    int myflag = 0;
    for (int64_t base = (int64_t)blockIdx.x * AF_EDGES; base < n; base += (int64_t)gridDim.x * AF_EDGES) {
        // ---- phase 1: set up the block's edges (as in k_arm_visible_flat)
        uint32_t N = 0;
        if (tid < AF_EDGES) {
            const int64_t eidx = base + tid;
            const bool valid = eidx < n;
            double4 a = make_double4(0, 0, 0, 0), b = make_double4(0, 0, 0, 0);
            if (valid) { a = C1[eidx]; b = C2[eidx]; }
            TruncResult t0 = trunc_index(a.x, mv.W), t1 = trunc_index(a.y, mv.H);
            TruncResult t2 = trunc_index(b.x, mv.W), t3 = trunc_index(b.y, mv.H);
            int f = t0.flag ? t0.flag : t1.flag ? t1.flag : t2.flag ? t2.flag : t3.flag;
            if (!f && (isinf(a.z) || isinf(a.w) || isinf(b.z) || isinf(b.w) || a.z != a.z || a.w != a.w ||
                       b.z != b.z || b.w != b.w))
                f = SBP_FLAG_NAN;
            if (valid) myflag |= f;
            const bool walk = valid && !f && t0.in_range && t1.in_range && t2.in_range && t3.in_range;
            int x1 = t0.i, y1 = t1.i, x2 = t2.i, y2 = t3.i;
            int dx = x2 - x1, dy = y2 - y1;
            const bool steep = abs(dy) > abs(dx);
            int a1 = steep ? y1 : x1, b1 = steep ? x1 : y1;
            int a2 = steep ? y2 : x2, b2 = steep ? x2 : y2;
            const bool swapped = a1 > a2;
            if (swapped) { int t = a1; a1 = a2; a2 = t; t = b1; b1 = b2; b2 = t; }
            ArmEdge e;
            e.a1 = a1; e.b1 = b1;
            e.da = (uint32_t)(a2 - a1); e.adb = (uint32_t)abs(b2 - b1);
            e.npx = e.da + 1u;
            e.flags = (steep ? 1 : 0) | (swapped ? 2 : 0) | (b1 < b2 ? 4 : 0);
            const uint32_t NN = e.npx < 2u ? 2u : e.npx;                 // :375-377
            const double inv = __ddiv_rn(1.0, (double)(NN - 1u));        // create_ranges (:363-364)
            e.az = a.z; e.aw = a.w;
            e.st0 = __dmul_rn(inv, __dsub_rn(b.z, a.z));
            e.st1 = __dmul_rn(inv, __dsub_rn(b.w, a.w));
            s_e[tid] = e;
            N = walk ? NN : 0u;
            s_f[tid] = walk ? 0 : 1;                                     // not walked = blocked
        }
        if (tid < AF_EDGES) {
            uint32_t inc = N;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += u; }
            if (lane == 31) s_ws[w] = inc;
            s_pref[tid + 1] = inc;                                        // (warp-local for now)
        }
        __syncthreads();
        if (tid < AF_EDGES) {
            uint32_t off = 0;
            for (int i = 0; i < w; ++i) off += s_ws[i];
            s_pref[tid + 1] += off;
            if (tid == 0) s_pref[0] = 0;
        }
        __syncthreads();
        const uint32_t T = s_pref[AF_EDGES];
        // ---- phase 2: (edge, config) pairs, a warp's 32 at a time.  A config is blocked iff link 1 is
        // blocked, or what follows link 1 says so (`after` = 0), or link 2 decides (`after` = 3) and is
        // blocked; it is ambiguous iff link 1 is free and `after` = 2 -- and "blocked" wins for the edge,
        // so the two links can be judged independently of each other.  Links without a free box go to
        // the warp's own list and are swept, 32 at a time, whenever it holds that many: one sweep per
        // lane, no CTA barrier inside.
        ArmLink *wl = s_todo + w * AB_LIST;
        int wn = 0;                                                      // (warp uniform)
        for (uint32_t c0 = 0; c0 < T; c0 += blockDim.x) {
            const uint32_t c = c0 + tid;
            bool need1 = false, need2 = false;
            ArmLink t1, t2;
            if (c < T) {
                int lo = 0, hi = AF_EDGES;                               // largest j with s_pref[j] <= c
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (s_pref[mid] <= c) lo = mid; else hi = mid;
                }
                if (!(s_f[lo] & 1)) {                                    // (edge not yet known blocked)
                    const ArmEdge e = s_e[lo];
                    const uint32_t o = c - s_pref[lo];
                    const uint32_t po = e.npx == 1u ? 0u : o;            // duplicated single pixel
                    const bool steep = e.flags & 1, swapped = e.flags & 2;
                    const uint64_t ci = swapped ? (uint64_t)(e.npx - 1u - po) : (uint64_t)po;
                    const uint32_t cc = e.da ? e.da - 1u - (e.da >> 1) : 0u;
                    const uint32_t k = e.da ? (uint32_t)((ci * e.adb + cc) / e.da) : 0u;
                    const int aa = e.a1 + (int)ci, bb = e.b1 + ((e.flags & 4) ? 1 : -1) * (int)k;
                    const double bx = (double)(steep ? bb : aa), by = (double)(steep ? aa : bb);
                    const double r0 = __dadd_rn(__dmul_rn(e.st0, (double)o), e.az);
                    const double r1 = __dadd_rn(__dmul_rn(e.st1, (double)o), e.aw);
                    const ArmPre p = arm_config_pre(mv, L0, L1, bx, by, r0, r1, myflag);
                    int r = p.r;                                         // 0 / 2 final, -1: the links decide
                    if (r < 0) {
                        r = 1;
                        // the links' end pixels first (base and elbow of link 1, the tip of link 2): most
                        // blocked configs end here
                        const bool ends1 = map_free_rows<!SMEM>(words, mv.SX, wrap_index(p.x0, mv.W), wrap_index(p.y0, mv.H)) &&
                                           map_free_rows<!SMEM>(words, mv.SX, wrap_index(p.x1, mv.W), wrap_index(p.y1, mv.H));
                        const bool tip = p.after != 3 ||
                                         map_free_rows<!SMEM>(words, mv.SX, wrap_index(p.x2, mv.W), wrap_index(p.y2, mv.H));
                        if (!ends1 || !tip || p.after == 0) {
                            r = 0;
                        } else {
                            if (p.after == 2) r = 2;                      // (counts only if link 1 turns out free; else blocked wins)
                            // (boxes only for coordinates that do not wrap around: :393-402 index the image with
                            // negative values too, pixel by pixel)
                            need1 = !((p.x0 | p.y0 | p.x1 | p.y1) >= 0 &&
                                      tiles_rect_free(af, min(p.x0, p.x1) >> 3, max(p.x0, p.x1) >> 3, min(p.y0, p.y1) >> 3,
                                                      max(p.y0, p.y1) >> 3));
                            need2 = p.after == 3 && !((p.x1 | p.y1 | p.x2 | p.y2) >= 0 &&
                                                      tiles_rect_free(af, min(p.x1, p.x2) >> 3, max(p.x1, p.x2) >> 3,
                                                                      min(p.y1, p.y2) >> 3, max(p.y1, p.y2) >> 3));
                            t1.x0 = (int16_t)p.x0; t1.y0 = (int16_t)p.y0; t1.x1 = (int16_t)p.x1; t1.y1 = (int16_t)p.y1;
                            t2.x0 = (int16_t)p.x1; t2.y0 = (int16_t)p.y1; t2.x1 = (int16_t)p.x2; t2.y1 = (int16_t)p.y2;
                            t1.edge = t2.edge = (uint16_t)lo;
                            t1.pad = t2.pad = 0;
                        }
                    }
                    if (r == 0 || r == 2) atomicOr(&s_f[lo], r == 0 ? 1 : 2);
                }
            }
            const unsigned lt = (1u << lane) - 1u;
            const unsigned m1 = __ballot_sync(FULL, need1), m2 = __ballot_sync(FULL, need2);
            if (need1) wl[wn + __popc(m1 & lt)] = t1;
            wn += __popc(m1);
            if (need2) wl[wn + __popc(m2 & lt)] = t2;
            wn += __popc(m2);
            __syncwarp();
            const bool last = c0 + blockDim.x >= T;
            while (wn >= 32 || (last && wn > 0)) {
                // ---- one listed link per lane (the LAST 32 entries: nothing to move)
                const int take = wn < 32 ? wn : 32;
                if (lane < take) {
                    const ArmLink q = wl[wn - take + lane];
                    if (!(s_f[q.edge] & 1) && !link_free<SMEM, true>(words, mv, q.x0, q.y0, q.x1, q.y1))   // :413-415, :422-424
                        atomicOr(&s_f[q.edge], 1);
                }
                wn -= take;
                __syncwarp();
            }
        }
        __syncthreads();
        // ---- phase 3: an unambiguous block decides the edge, else an ambiguous config leaves it open
        if (tid < AF_EDGES && base + tid < n) {
            const int f = s_f[tid];
            out[base + tid] = (uint8_t)((f & 1) ? 0 : (f & 2) ? 2 : 1);
        }
        __syncthreads();
    }
    if (myflag && flags) atomicOr(flags, myflag);
}

// feasible with the same free-box test (thread = config, a warp's undecided links swept 32 at a time
// from its own list; out[i] is written by the config's lane first -- 1, or 2 when only the tip's guard
// band is open -- and set to 0 by whichever lane finds one of its links blocked: blocked wins, as in
// arm_config, where a blocked link 1 returns before the tip is looked at).
struct ArmLinkF {              // 12 bytes
    int16_t x0, y0, x1, y1;
    uint32_t idx;            // config
};

template <bool SMEM>
__global__ void __launch_bounds__(1024)
k_arm_feasible_boxes(MapView mv, TileCodeView af, double L0, double L1, const double4 *__restrict__ C, int64_t n,
                     uint8_t *__restrict__ out, int32_t *__restrict__ flags) {
    extern __shared__ __align__(16) uint32_t smap[];
    __shared__ uint64_t bar;
    if (SMEM) stage_map_to_smem(smap, mv, &bar);
    const uint32_t *words = SMEM ? smap : mv.words;
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31;
    ArmLinkF *wl = reinterpret_cast<ArmLinkF *>(smap + (SMEM ? mv.n_words : 0u)) + (tid >> 5) * AB_LIST;
    int wn = 0;                                                          // (warp uniform)
    int myflag = 0;
    const int64_t gsz = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + (tid & ~31); i0 < n; i0 += gsz) {
        const int64_t i = i0 + lane;
        bool need1 = false, need2 = false;
        ArmLinkF t1, t2;
        if (i < n) {
            const double4 c = C[i];
            const ArmPre p = arm_config_pre(mv, L0, L1, c.x, c.y, c.z, c.w, myflag);
            int r = p.r;
            if (r < 0) {
                r = 1;
                const bool ends1 = map_free_rows<!SMEM>(words, mv.SX, wrap_index(p.x0, mv.W), wrap_index(p.y0, mv.H)) &&
                                   map_free_rows<!SMEM>(words, mv.SX, wrap_index(p.x1, mv.W), wrap_index(p.y1, mv.H));
                const bool tip = p.after != 3 ||
                                 map_free_rows<!SMEM>(words, mv.SX, wrap_index(p.x2, mv.W), wrap_index(p.y2, mv.H));
                if (!ends1 || !tip || p.after == 0) {
                    r = 0;
                } else {
                    if (p.after == 2) r = 2;
                    need1 = !((p.x0 | p.y0 | p.x1 | p.y1) >= 0 &&
                              tiles_rect_free(af, min(p.x0, p.x1) >> 3, max(p.x0, p.x1) >> 3, min(p.y0, p.y1) >> 3,
                                              max(p.y0, p.y1) >> 3));
                    need2 = p.after == 3 && !((p.x1 | p.y1 | p.x2 | p.y2) >= 0 &&
                                              tiles_rect_free(af, min(p.x1, p.x2) >> 3, max(p.x1, p.x2) >> 3,
                                                              min(p.y1, p.y2) >> 3, max(p.y1, p.y2) >> 3));
                    t1.x0 = (int16_t)p.x0; t1.y0 = (int16_t)p.y0; t1.x1 = (int16_t)p.x1; t1.y1 = (int16_t)p.y1;
                    t2.x0 = (int16_t)p.x1; t2.y0 = (int16_t)p.y1; t2.x1 = (int16_t)p.x2; t2.y1 = (int16_t)p.y2;
                    t1.idx = t2.idx = (uint32_t)i;
                }
            }
            out[i] = (uint8_t)r;
        }
        const unsigned lt = (1u << lane) - 1u;
        const unsigned m1 = __ballot_sync(FULL, need1), m2 = __ballot_sync(FULL, need2);
        if (need1) wl[wn + __popc(m1 & lt)] = t1;
        wn += __popc(m1);
        if (need2) wl[wn + __popc(m2 & lt)] = t2;
        wn += __popc(m2);
        __syncwarp();
        const bool last = i0 + gsz >= n;
        while (wn >= 32 || (last && wn > 0)) {
            const int take = wn < 32 ? wn : 32;
            if (lane < take) {
                const ArmLinkF q = wl[wn - take + lane];
                if (!link_free<SMEM, true>(words, mv, q.x0, q.y0, q.x1, q.y1)) out[q.idx] = 0;     // :413-415, :422-424
            }
            wn -= take;
            __syncwarp();
        }
    }
    if (myflag && flags) atomicOr(flags, myflag);
}

// ------------------------------------------------------------ launchers -----

// The arm kernels read the ROW-MAJOR copy of the map (sbp_map::rows, common.cuh).
static void arm_cfg(const sbp_map *m, bool smem, int &grid, int &threads, size_t &bytes) {
    const sbp_ctx *ctx = m->ctx;
    if (smem) {
        bytes = (size_t)m->rows.n_words * 4;
        int per_sm = (int)((size_t)(228 * 1024) / (bytes + 2048));
        per_sm = per_sm < 1 ? 1 : per_sm > 4 ? 4 : per_sm;
        threads = per_sm <= 2 ? 1024 : 512;
        grid = ctx->sm_count * per_sm;
    } else {
        bytes = 0;
        threads = 256;
        grid = ctx->sm_count * 8;
    }
}

extern "C" int32_t sbp_arm_feasible(sbp_map *map, double L0, double L1, const double *C,
                                    int64_t n, uint8_t *out, int32_t *flags) {
    SBP_NVTX("sbp_arm_feasible");
    SBP_REQUIRE(map && (n == 0 || (C && out)), "sbp_arm_feasible: NULL argument");
    SBP_REQUIRE(n >= 0, "sbp_arm_feasible: n < 0");
    if (n == 0) return SBP_OK;
    sbp_ctx *ctx = map->ctx;
    SBP_DEVICE(ctx);
    int32_t rc_rows = sbp_map_ensure_rows(map);
    if (rc_rows) return rc_rows;
    bool smem = !ctx->tune.force_global && map->rows_smem_resident &&
                n * 64 >= (int64_t)map->rows.n_words * 4 * (int64_t)ctx->sm_count / 4;
    int grid, threads; size_t bytes;
    if (ctx->tune.arm_flat >= 2 && ctx->tune.tile_rect && map->view.W <= 32767 && map->view.H <= 32767 &&
        n < ((int64_t)1 << 32)) {
        // free-box form: tile codes / free-run planes of the map, a list of undecided links per warp
        int32_t rc = sbp_map_ensure_allfree(map);
        if (rc) return rc;
        const size_t map_bytes = smem ? (size_t)map->rows.n_words * 4 : 0, list_bytes = sizeof(ArmLinkF) * AB_LIST * 32;
        int per_sm = smem ? (int)((size_t)(227 * 1024) / (map_bytes + list_bytes + 2048)) : 2;
        per_sm = per_sm < 1 ? 1 : per_sm > 2 ? 2 : per_sm;
        threads = per_sm == 1 ? 1024 : 512;
        grid = ctx->sm_count * per_sm;
        const int64_t want = (n + threads - 1) / threads;
        if (want < grid) grid = (int)want;
        const TileCodeView af = sbp_map_codes(map);
        if (smem) {
            SBP_CUDA(cudaFuncSetAttribute(k_arm_feasible_boxes<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(map_bytes + list_bytes)));
            k_arm_feasible_boxes<true><<<grid, threads, map_bytes + list_bytes, ctx->stream>>>(map->rows, af, L0, L1,
                                                                                             (const double4 *)C, n, out, flags);
        } else {
            k_arm_feasible_boxes<false><<<grid, threads, list_bytes, ctx->stream>>>(map->rows, af, L0, L1, (const double4 *)C,
                                                                                   n, out, flags);
        }
        ctx->launches++;
        SBP_CUDA(cudaGetLastError());
        return SBP_OK;
    }
    arm_cfg(map, smem, grid, threads, bytes);
    int64_t need = (n + threads - 1) / threads;
    if (need < grid) grid = (int)need;
    if (smem) {
        if (bytes > 48 * 1024)
            SBP_CUDA(cudaFuncSetAttribute(k_arm_feasible<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        k_arm_feasible<true><<<grid, threads, bytes, ctx->stream>>>(map->rows, L0, L1, (const double4 *)C, n, out, flags);
    } else {
        k_arm_feasible<false><<<grid, threads, 0, ctx->stream>>>(map->rows, L0, L1, (const double4 *)C, n, out, flags);
    }
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

template <int G>
static int32_t launch_arm_visible(sbp_map *map, double L0, double L1, const double *C1,
                                  const double *C2, int64_t n, uint8_t *out, int32_t *flags,
                                  unsigned long long *cnt, bool smem) {
    sbp_ctx *ctx = map->ctx;
    int grid, threads; size_t bytes;
    arm_cfg(map, smem, grid, threads, bytes);
    int64_t need = (n * G + threads - 1) / threads;
    if (need < grid) grid = (int)need;
    const double4 *c1 = (const double4 *)C1, *c2 = (const double4 *)C2;
#define SBP_LAUNCH_ARM(SM, CT)                                                                    \
    do {                                                                                           \
        if (bytes > 48 * 1024)                                                                     \
            SBP_CUDA(cudaFuncSetAttribute(k_arm_visible<G, SM, CT>,                                \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes)); \
        k_arm_visible<G, SM, CT><<<grid, threads, bytes, ctx->stream>>>(map->rows, L0, L1, c1, c2, \
                                                                        n, out, flags, cnt);       \
    } while (0)
    if (smem) { if (cnt) SBP_LAUNCH_ARM(true, true); else SBP_LAUNCH_ARM(true, false); }
    else      { if (cnt) SBP_LAUNCH_ARM(false, true); else SBP_LAUNCH_ARM(false, false); }
#undef SBP_LAUNCH_ARM
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

extern "C" int32_t sbp_arm_visible(sbp_map *map, double L0, double L1, const double *C1,
                                   const double *C2, int64_t n, uint8_t *out, int32_t *flags,
                                   uint64_t *cfg_tested) {
    SBP_NVTX("sbp_arm_visible");
    SBP_REQUIRE(map && (n == 0 || (C1 && C2 && out)), "sbp_arm_visible: NULL argument");
    SBP_REQUIRE(n >= 0, "sbp_arm_visible: n < 0");
    if (n == 0) return SBP_OK;
    sbp_ctx *ctx = map->ctx;
    SBP_DEVICE(ctx);
    int32_t rc_rows = sbp_map_ensure_rows(map);
    if (rc_rows) return rc_rows;
    bool smem = !ctx->tune.force_global && map->rows_smem_resident &&
                n * 64 >= (int64_t)map->rows.n_words * 4 * (int64_t)ctx->sm_count / 16;
    unsigned long long *cnt = (unsigned long long *)cfg_tested;
    if (!cnt && ctx->tune.arm_group == 0 && ctx->tune.arm_flat) {
        size_t bytes = smem ? (size_t)map->rows.n_words * 4 : 0;
        // free-box form: needs the tile codes / free-run planes of the map and 16-bit coordinates
        const bool boxes = ctx->tune.arm_flat >= 2 && ctx->tune.tile_rect && map->view.W <= 32767 && map->view.H <= 32767;
        if (boxes) {
            int32_t rc = sbp_map_ensure_allfree(map);
            if (rc) return rc;
        }
        // static shared memory of the kernel (edge block, prefix, flags, todo list) next to the staged map
        const size_t fixed = sizeof(ArmEdge) * AF_EDGES + 4 * (2 * AF_EDGES + 16) + 1024 + (boxes ? sizeof(ArmLink) * AB_LIST * 32 + 16 : 0);
        const size_t todo_bytes = sizeof(ArmLink) * AB_LIST * 32;         // dynamic, behind the map (a list per warp)
        int per_sm = smem ? (int)((size_t)(227 * 1024) / (bytes + fixed)) : 2;
        per_sm = per_sm < 1 ? 1 : per_sm > 2 ? 2 : per_sm;
        const int threads = per_sm == 1 ? 1024 : 512;
        int grid = ctx->sm_count * per_sm;
        const int64_t need = (n + AF_EDGES - 1) / AF_EDGES;
        if (need < grid) grid = (int)need;
        const double4 *c1 = (const double4 *)C1, *c2 = (const double4 *)C2;
        if (boxes) {
            const TileCodeView af = sbp_map_codes(map);
            if (smem) {
                SBP_CUDA(cudaFuncSetAttribute(k_arm_visible_boxes<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)(bytes + todo_bytes)));
                k_arm_visible_boxes<true><<<grid, threads, bytes + todo_bytes, ctx->stream>>>(map->rows, af, L0, L1, c1, c2, n, out, flags);
            } else {
                SBP_CUDA(cudaFuncSetAttribute(k_arm_visible_boxes<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)todo_bytes));
                k_arm_visible_boxes<false><<<grid, threads, todo_bytes, ctx->stream>>>(map->rows, af, L0, L1, c1, c2, n, out, flags);
            }
            ctx->launches++;
            SBP_CUDA(cudaGetLastError());
            return SBP_OK;
        }
        if (smem) {
            if (bytes + fixed > 48 * 1024)
                SBP_CUDA(cudaFuncSetAttribute(k_arm_visible_flat<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            k_arm_visible_flat<true><<<grid, threads, bytes, ctx->stream>>>(map->rows, L0, L1, c1, c2, n, out, flags);
        } else {
            k_arm_visible_flat<false><<<grid, threads, 0, ctx->stream>>>(map->rows, L0, L1, c1, c2, n, out, flags);
        }
        ctx->launches++;
        SBP_CUDA(cudaGetLastError());
        return SBP_OK;
    }
    int G = ctx->tune.arm_group ? ctx->tune.arm_group : 8;
    switch (G) {
        case 4: return launch_arm_visible<4>(map, L0, L1, C1, C2, n, out, flags, cnt, smem);
        case 8: return launch_arm_visible<8>(map, L0, L1, C1, C2, n, out, flags, cnt, smem);
        case 16: return launch_arm_visible<16>(map, L0, L1, C1, C2, n, out, flags, cnt, smem);
        case 32: return launch_arm_visible<32>(map, L0, L1, C1, C2, n, out, flags, cnt, smem);
        default: sbp_set_error("arm_group must be 4, 8, 16 or 32"); return SBP_ERR_ARG;
    }
}

// ------------------------------------------- host guard-band resolution -----
// Exact re-evaluation of the few SBP_AMBIGUOUS entries with the host libm.  This
// is not a CPU implementation of the path: it decides only entries the device
// declined to decide, using the host copy of the same packed map.
namespace {

struct HostMap {
    const uint32_t *w;
    int W, H, SX;
    bool free_at(int64_t x, int64_t y) const {
        if (x < -W || x >= W || y < -H || y >= H) return false;
        if (x < 0) x += W;
        if (y < 0) y += H;
        uint32_t sec = (uint32_t)(y >> 4) * (uint32_t)SX + (uint32_t)(x >> 4);
        uint32_t wi = sec * 8u + (uint32_t)(((y >> 3) & 1) * 4 + ((x >> 3) & 1) * 2 + ((y >> 2) & 1));
        return (w[wi] >> (((y & 3) << 3) | (x & 7))) & 1u;
    }
};

bool host_trunc(double v, int64_t &o) {
    if (!(fabs(v) < 4.0e18)) return false;
    o = (int64_t)v;
    return true;
}

bool host_link_free(const HostMap &m, int64_t x1, int64_t y1, int64_t x2, int64_t y2) {
    if (!m.free_at(x1, y1) || !m.free_at(x2, y2)) return false;   // endpoints are pixels
    int64_t dx = x2 - x1, dy = y2 - y1;
    bool steep = llabs(dy) > llabs(dx);
    int64_t a1 = steep ? y1 : x1, b1 = steep ? x1 : y1, a2 = steep ? y2 : x2, b2 = steep ? x2 : y2;
    if (a1 > a2) { int64_t t = a1; a1 = a2; a2 = t; t = b1; b1 = b2; b2 = t; }
    int64_t da = a2 - a1, adb = llabs(b2 - b1), err = da / 2, ystep = b1 < b2 ? 1 : -1, b = b1;
    for (int64_t a = a1; a <= a2; ++a) {
        if (!m.free_at(steep ? b : a, steep ? a : b)) return false;
        err -= adb;
        if (err < 0) { b += ystep; err += da; }
    }
    return true;
}

bool host_config_free(const HostMap &m, double L0, double L1, double bx, double by, double r0,
                      double r1) {
    double p2x = bx + L0 * cos(r0), p2y = by + L0 * sin(r0);
    int64_t ax, ay, cx, cy, ex, ey;
    if (!host_trunc(bx, ax) || !host_trunc(by, ay) || !host_trunc(p2x, cx) || !host_trunc(p2y, cy))
        return false;
    if (!host_link_free(m, ax, ay, cx, cy)) return false;
    double p3x = p2x + L1 * cos(r1), p3y = p2y + L1 * sin(r1);
    if (!host_trunc(p3x, ex) || !host_trunc(p3y, ey)) return false;
    return host_link_free(m, cx, cy, ex, ey);
}

bool host_edge_free(const HostMap &m, double L0, double L1, const double *a, const double *b) {
    int64_t x1, y1, x2, y2;
    if (!host_trunc(a[0], x1) || !host_trunc(a[1], y1) || !host_trunc(b[0], x2) ||
        !host_trunc(b[1], y2))
        return false;
    if (!m.free_at(x1, y1) || !m.free_at(x2, y2)) return false;
    int64_t dx = x2 - x1, dy = y2 - y1;
    bool steep = llabs(dy) > llabs(dx);
    int64_t a1 = steep ? y1 : x1, b1 = steep ? x1 : y1, a2 = steep ? y2 : x2, b2 = steep ? x2 : y2;
    bool swapped = a1 > a2;
    if (swapped) { int64_t t = a1; a1 = a2; a2 = t; t = b1; b1 = b2; b2 = t; }
    int64_t da = a2 - a1, adb = llabs(b2 - b1), ystep = b1 < b2 ? 1 : -1;
    int64_t npx = da + 1, N = npx < 2 ? 2 : npx, c = da ? da - 1 - da / 2 : 0;
    double inv = 1.0 / (double)(N - 1);
    double st0 = inv * (b[2] - a[2]), st1 = inv * (b[3] - a[3]);
    for (int64_t o = 0; o < N; ++o) {
        int64_t po = npx == 1 ? 0 : o;
        int64_t ci = swapped ? npx - 1 - po : po;
        int64_t k = da ? (ci * adb + c) / da : 0;
        int64_t aa = a1 + ci, bb = b1 + ystep * k;
        double m0 = st0 * (double)o, m1 = st1 * (double)o;   // separately rounded (-fmad=false)
        double r0 = m0 + a[2], r1 = m1 + a[3];
        if (!host_config_free(m, L0, L1, (double)(steep ? bb : aa), (double)(steep ? aa : bb), r0, r1))
            return false;
    }
    return true;
}

}  // namespace

static int32_t ensure_host_words(sbp_map *map) {
    if (!map->h_words.empty()) return SBP_OK;
    map->h_words.resize(map->view.n_words);
    SBP_CUDA(cudaMemcpyAsync(map->h_words.data(), map->d_words, (size_t)map->view.n_words * 4,
                             cudaMemcpyDeviceToHost, map->ctx->stream));
    SBP_CUDA(cudaStreamSynchronize(map->ctx->stream));
    return SBP_OK;
}

extern "C" int32_t sbp_arm_resolve_host(sbp_map *map, double L0, double L1, const double *C1,
                                        const double *C2, int64_t n, uint8_t *inout,
                                        int64_t *resolved) {
    SBP_REQUIRE(map && (n == 0 || (C1 && inout)), "sbp_arm_resolve_host: NULL argument");
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (inout[i] != SBP_AMBIGUOUS) continue;
        if (cnt == 0) {
            int32_t rc = ensure_host_words(map);
            if (rc) return rc;
        }
        HostMap hm{map->h_words.data(), map->view.W, map->view.H, map->view.SX};
        const double *a = C1 + 4 * i;
        bool ok = C2 ? host_edge_free(hm, L0, L1, a, C2 + 4 * i)
                     : host_config_free(hm, L0, L1, a[0], a[1], a[2], a[3]);
        inout[i] = ok ? SBP_FREE : SBP_BLOCKED;
        ++cnt;
    }
    if (resolved) *resolved = cnt;
    return SBP_OK;
}
