// device_fns.cuh -- device functions shared by the kernels of several translation
// units: edge canonicalisation, sequential edge / arm-config checks, distance cores.
#pragma once
#include <math.h>

#include "common.cuh"

// ======================= 2-D edges (collisionChecker.py:134-215) ================
struct Edge2D {
    int a1, b1;        // canonical start (major, minor)
    uint32_t da, adb;  // major extent, |minor extent|
    int ystep;
    bool steep, swapped;
    bool walk;         // both endpoints inside the index range: pixels must be tested
    int flag;          // SBP_FLAG_* bits
    int x1, y1, x2, y2;
};

__device__ __forceinline__ Edge2D make_edge(double px, double py, double qx, double qy, int W,
                                            int H) {
    Edge2D e;
    TruncResult t0 = trunc_index(px, W), t1 = trunc_index(py, H);   // x1, y1 = map(int, start)
    TruncResult t2 = trunc_index(qx, W), t3 = trunc_index(qy, H);   // x2, y2 = map(int, end)
    // CPython raises at the first offending conversion (:173-174)
    e.flag = t0.flag ? t0.flag : t1.flag ? t1.flag : t2.flag ? t2.flag : t3.flag;
    // both endpoints are members of the pixel list (:204-206): an endpoint outside
    // the index range makes the all-of False without walking
    e.walk = t0.in_range && t1.in_range && t2.in_range && t3.in_range;
    e.x1 = t0.i; e.y1 = t1.i; e.x2 = t2.i; e.y2 = t3.i;
    int dx = e.x2 - e.x1, dy = e.y2 - e.y1;
    e.steep = abs(dy) > abs(dx);                                      // :179
    int a1 = e.steep ? e.y1 : e.x1, b1 = e.steep ? e.x1 : e.y1;       // :182-184
    int a2 = e.steep ? e.y2 : e.x2, b2 = e.steep ? e.x2 : e.y2;
    e.swapped = a1 > a2;                                              // :188-191
    if (e.swapped) { int t = a1; a1 = a2; a2 = t; t = b1; b1 = b2; b2 = t; }
    e.a1 = a1; e.b1 = b1;
    e.da = (uint32_t)(a2 - a1);                                       // :194
    e.adb = (uint32_t)abs(b2 - b1);                                   // :195, :207
    e.ystep = b1 < b2 ? 1 : -1;                                       // :199
    return e;
}


// One thread walks one whole edge (used where candidates are few and scattered, e.g.
// the fused near+visible query).  Same semantics as the batched kernels.
template <bool SMEM>
__device__ __forceinline__ bool edge_visible_seq(const uint32_t *words, const MapView &mv,
                                                 double px, double py, double qx, double qy,
                                                 int &flag) {
    Edge2D e = make_edge(px, py, qx, qy, mv.W, mv.H);
    flag |= e.flag;
    if (!e.walk) return false;
    int a = e.a1, b = e.b1, err = (int)(e.da >> 1);
    const int a2 = e.a1 + (int)e.da;
    for (; a <= a2; ++a) {
        int x = e.steep ? b : a, y = e.steep ? a : b;
        if (!map_free<!SMEM>(words, mv.SX, wrap_index(x, mv.W), wrap_index(y, mv.H))) return false;
        err -= (int)e.adb;
        if (err < 0) { b += e.ystep; err += (int)e.da; }
    }
    return true;
}

// ---- two-level walk for maps that do not fit shared memory (sbp_map::af, common.cuh) ----------
// The line is cut into CHUNKS of up to 8 pixels that share one aligned group of 8 major
// coordinates.  Between the first and the last pixel of a chunk the minor coordinate moves by at
// most 7, so the chunk's pixels lie in the 8x8 tile of its first pixel and the 8x8 tile of its
// last pixel (often the same), and each of the two holds a pixel of the line: when the coarse
// level says one of them is all blocked the edge is blocked, when it says both are all free the
// chunk is free, without touching the map; otherwise the chunk's pixels are tested one by one as
// the reference does (collisionChecker.py:141, :204-210).  The chunk ends come from the closed
// form of the header of collide2d.cu -- pixel i has made q_i = floor((i * |db| + c) / da) minor
// steps, c = da - 1 - floor(da / 2), and the reference's error term before pixel i is
// da - 1 - (i * |db| + c) mod da -- so the decision is the reference's for every map, whatever
// its structure.
// Requires: every pixel of the line inside [0, W) x [0, H) (no numpy wrap-around), da < 65536.
// Lookup of the 2-bit tile codes (map.cu:k_tile_codes) by (major tile u, minor tile v) in the plane
// of the line's orientation, with the last word kept in registers: consecutive chunks of an edge
// mostly stay inside one 32 x 32-pixel word.
struct TileCodes {
    const uint32_t *__restrict__ plane;
    int WW;
    uint32_t idx0, w0, idx1, w1;       // two cached words: the one being walked and the line's last one
    // AW x AH words in plane 0 (x-major lines); plane 1 (y-major lines) is its transpose
    __device__ __forceinline__ TileCodes(const uint32_t *af, int AW, int AH, bool steep)
        : plane(steep ? af + (size_t)AW * (size_t)AH : af), WW(steep ? AH : AW), idx0(0xffffffffu), w0(0u),
          idx1(0xffffffffu), w1(0u) {}
    // the words of the line's first and last tile, requested together before the walk starts (a line of
    // a few dozen pixels rarely touches a third one): one memory round trip per edge instead of one
    // per word
    __device__ __forceinline__ void prefetch(uint32_t u_s, int v_s, uint32_t u_e, int v_e) {
        idx0 = ((uint32_t)v_s >> 2) * (uint32_t)WW + (u_s >> 2);
        idx1 = ((uint32_t)v_e >> 2) * (uint32_t)WW + (u_e >> 2);
        w0 = __ldg(plane + idx0);
        w1 = __ldg(plane + idx1);
    }
    __device__ __forceinline__ uint32_t get(uint32_t uw, uint32_t ub2, int v) {   // uw = u >> 2, ub2 = (u & 3) * 2
        const uint32_t idx = ((uint32_t)v >> 2) * (uint32_t)WW + uw;
        if (idx != idx0 && idx != idx1) { idx0 = idx; w0 = __ldg(plane + idx); }  // (the older word goes)
        const uint32_t w = idx == idx0 ? w0 : w1;
        return (w >> ((((uint32_t)v & 3u) << 3) | ub2)) & 3u;
    }
};

// Rectangle of tiles [tu0, tu1] x [tv0, tv1] (x tiles, y tiles, inside the map): true = every pixel
// of it is free, proven by two words of the free-run planes (common.cuh:TileCodeView::runs; spans of
// up to 15 x 16 tiles); false = not decided.  A Bresenham line never leaves the bounding box of its
// end pixels, so a free box settles visible() = True for every map without walking the line.
__device__ __forceinline__ bool tiles_rect_free(const TileCodeView &af, int tu0, int tu1, int tv0, int tv1) {
    const int w = tu1 - tu0 + 1, h = tv1 - tv0 + 1;
    if (af.runs == nullptr || w > 15 || h > 16) return false;
    const int j = 31 - __clz(w);                               // two runs of 2^j tiles cover [tu0, tu1]
    const uint32_t mask = ((1u << h) - 1u) << (tv0 & 15);
    const uint32_t row = ((uint32_t)j * (uint32_t)af.RK + ((uint32_t)tv0 >> 4)) * (uint32_t)af.RU;   // (< 2^32 words)
    const uint32_t a = __ldg(af.runs + (row + (uint32_t)tu0)), b = __ldg(af.runs + (row + (uint32_t)(tu1 + 1) - (1u << j)));
    return (a & b & mask) == mask;
}

template <bool PRETEST = true>
__device__ __forceinline__ bool edge_free_tiles(const uint32_t *__restrict__ tiled, int SX, const TileCodeView &af,
                                                int a1, int b1, uint32_t da, uint32_t adb, int ystep, bool steep) {
    if (PRETEST) {
        const int a2 = a1 + (int)da, b2 = b1 + ystep * (int)adb;
        const int bl = (ystep > 0 ? b1 : b2) >> 3, bh = (ystep > 0 ? b2 : b1) >> 3;
        if (tiles_rect_free(af, steep ? bl : a1 >> 3, steep ? bh : a2 >> 3, steep ? a1 >> 3 : bl, steep ? a2 >> 3 : bh))
            return true;
    }
    const uint32_t dd = da ? da : 1u;
    const uint32_t c = da ? da - 1u - (da >> 1) : 0u;
    // quotient estimate from below (2e-6 covers the roundings of the conversion, the reciprocal
    // and the product: the estimate is q or q - 1, one correction step)
    const float rcp = __fdividef(0.999998f, (float)dd);
    TileCodes codes(af.words, af.AW, af.AH, steep);
    codes.prefetch((uint32_t)a1 >> 3, b1 >> 3, (uint32_t)(a1 + (int)da) >> 3, (b1 + ystep * (int)adb) >> 3);
    uint32_t i = 0, q = 0, r = c;
    while (i <= da) {
        const int a = a1 + (int)i, b = b1 + ystep * (int)q;
        uint32_t ie = i + (7u - ((uint32_t)a & 7u));
        ie = ie < da ? ie : da;
        const uint32_t num = ie * adb + c;
        uint32_t qe = (uint32_t)(__uint2float_rz(num) * rcp);
        uint32_t re = num - qe * dd;
        if (re >= dd) { re -= dd; ++qe; }
        const int be = b1 + ystep * (int)qe;
        const uint32_t ta = (uint32_t)a >> 3, uw = ta >> 2, ub2 = (ta & 3u) << 1;
        const uint32_t c0 = codes.get(uw, ub2, b >> 3);
        const uint32_t c1 = codes.get(uw, ub2, be >> 3);
        if ((c0 | c1) & 2u) return false;
        if (!(c0 & c1 & 1u)) {
            int E = (int)da - 1 - (int)r, aa = a, bb = b;
            for (uint32_t j = i; j <= ie; ++j) {
                if (!map_free<true>(tiled, SX, steep ? bb : aa, steep ? aa : bb)) return false;
                E -= (int)adb;
                if (E < 0) { bb += ystep; E += (int)da; }
                ++aa;
            }
        }
        i = ie + 1u;
        q = qe;
        r = re + adb;
        if (r >= dd) { r -= dd; ++q; }
    }
    return true;
}

// ======================= 4-DoF arm (collisionChecker.py:346-503) ================
// Sequential Bresenham walk of one link; all pixels free?  Order is irrelevant
// for the boolean, so the canonical traversal is walked directly.
// ROWS selects the row-major map copy (mv = sbp_map::rows, mv.SX = words per row).
template <bool SMEM, bool ROWS = false>
__device__ __forceinline__ bool link_free(const uint32_t *words, const MapView &mv, int x1, int y1,
                                          int x2, int y2) {
    int dx = x2 - x1, dy = y2 - y1;
    bool steep = abs(dy) > abs(dx);                     // :467
    int a1 = steep ? y1 : x1, b1 = steep ? x1 : y1;     // :470-472
    int a2 = steep ? y2 : x2, b2 = steep ? x2 : y2;
    if (a1 > a2) { int t = a1; a1 = a2; a2 = t; t = b1; b1 = b2; b2 = t; }  // :476-479
    int da = a2 - a1, adb = abs(b2 - b1);               // :482-483
    int err = da >> 1;                                  // int(dx / 2.0), :486
    int ystep = b1 < b2 ? 1 : -1;                       // :487
    int b = b1;
    // one loop for both major axes (selects, not branches): lanes of a warp walk links of
    // different orientation, and per-axis loops would serialise them
    for (int a = a1; a <= a2; ++a) {                    // :492-498
        int x = steep ? b : a, y = steep ? a : b;
        const int xw = wrap_index(x, mv.W), yw = wrap_index(y, mv.H);
        if (!(ROWS ? map_free_rows<!SMEM>(words, mv.SX, xw, yw) : map_free<!SMEM>(words, mv.SX, xw, yw)))
            return false;
        err -= adb;
        if (err < 0) { b += ystep; err += da; }
    }
    return true;
}

// |v - nearest integer| within the guard band?  The band covers the discrepancy
// between glibc and libdevice cos/sin propagated through pt + L*c (see header).
__device__ __forceinline__ bool near_pixel_boundary(double v, double L) {
    double guard = (fabs(v) + fabs(L) + 1.0) * 5.6843418860808015e-14;  // 2^-44
    return fabs(v - rint(v)) <= guard;
}

// One configuration: 1 free, 0 blocked, 2 ambiguous.  flag receives SBP_FLAG_*.
template <bool SMEM, bool ROWS = false>
__device__ __forceinline__ int arm_config(const uint32_t *words, const MapView &mv, double L0,
                                          double L1, double bx, double by, double r0, double r1,
                                          int &flag) {
    // math.cos(+-inf) raises ValueError("math domain error"); nan propagates to int(nan)
    if (isinf(r0) || isinf(r1) || r0 != r0 || r1 != r1) { flag |= SBP_FLAG_NAN; return 0; }
    double s0, c0, s1, c1;
    sincos(r0, &s0, &c0);
    sincos(r1, &s1, &c1);
    double p2x = __dadd_rn(bx, __dmul_rn(L0, c0));      // :438
    double p2y = __dadd_rn(by, __dmul_rn(L0, s0));      // :439
    double p3x = __dadd_rn(p2x, __dmul_rn(L1, c1));     // :417-419
    double p3y = __dadd_rn(p2y, __dmul_rn(L1, s1));
    TruncResult tbx = trunc_index(bx, mv.W), tby = trunc_index(by, mv.H);
    int f = tbx.flag ? tbx.flag : tby.flag;
    if (f) { flag |= f; return 0; }
    // the base pixel is exact: when it is outside the index range the first tested
    // pixel fails whatever the FK says
    if (!tbx.in_range || !tby.in_range) return 0;
    bool amb1 = near_pixel_boundary(p2x, L0) || near_pixel_boundary(p2y, L0);
    TruncResult t2x = trunc_index(p2x, mv.W), t2y = trunc_index(p2y, mv.H);
    if (t2x.flag || t2y.flag) { flag |= t2x.flag ? t2x.flag : t2y.flag; return 0; }
    if (amb1) return 2;
    if (!t2x.in_range || !t2y.in_range) return 0;       // end pixel of link 1 is out of range
    if (!link_free<SMEM, ROWS>(words, mv, tbx.i, tby.i, t2x.i, t2y.i)) return 0;   // :413-415
    bool amb2 = near_pixel_boundary(p3x, L1) || near_pixel_boundary(p3y, L1);
    TruncResult t3x = trunc_index(p3x, mv.W), t3y = trunc_index(p3y, mv.H);
    if (t3x.flag || t3y.flag) { flag |= t3x.flag ? t3x.flag : t3y.flag; return 0; }
    if (amb2) return 2;
    if (!t3x.in_range || !t3y.in_range) return 0;
    return link_free<SMEM, ROWS>(words, mv, t2x.i, t2y.i, t3x.i, t3y.i) ? 1 : 0;   // :422-424
}

// The same configuration WITHOUT its two link sweeps (k_arm_visible_boxes): everything arm_config
// decides before / between them.  r >= 0: final (0 blocked, 2 ambiguous).  r = -1: link 1 (base ->
// elbow) has to be free, and then the result is `after` -- 0 / 2, or 3 = link 2 (elbow -> tip) decides.
// (arm_config evaluates the tip's ambiguity / range only once link 1 is known free; taking them first
// changes nothing: they are consulted only in that case.)
struct ArmPre {
    int r, after;
    int x0, y0, x1, y1, x2, y2;
};
__device__ __forceinline__ ArmPre arm_config_pre(const MapView &mv, double L0, double L1, double bx, double by,
                                                 double r0, double r1, int &flag) {
    ArmPre p;
    p.r = 0; p.after = 0;
    p.x0 = p.y0 = p.x1 = p.y1 = p.x2 = p.y2 = 0;
    if (isinf(r0) || isinf(r1) || r0 != r0 || r1 != r1) { flag |= SBP_FLAG_NAN; return p; }
    double s0, c0, s1, c1;
    sincos(r0, &s0, &c0);
    sincos(r1, &s1, &c1);
    double p2x = __dadd_rn(bx, __dmul_rn(L0, c0));      // :438
    double p2y = __dadd_rn(by, __dmul_rn(L0, s0));      // :439
    double p3x = __dadd_rn(p2x, __dmul_rn(L1, c1));     // :417-419
    double p3y = __dadd_rn(p2y, __dmul_rn(L1, s1));
    TruncResult tbx = trunc_index(bx, mv.W), tby = trunc_index(by, mv.H);
    int f = tbx.flag ? tbx.flag : tby.flag;
    if (f) { flag |= f; return p; }
    if (!tbx.in_range || !tby.in_range) return p;
    bool amb1 = near_pixel_boundary(p2x, L0) || near_pixel_boundary(p2y, L0);
    TruncResult t2x = trunc_index(p2x, mv.W), t2y = trunc_index(p2y, mv.H);
    if (t2x.flag || t2y.flag) { flag |= t2x.flag ? t2x.flag : t2y.flag; return p; }
    if (amb1) { p.r = 2; return p; }
    if (!t2x.in_range || !t2y.in_range) return p;
    p.r = -1;
    p.x0 = tbx.i; p.y0 = tby.i; p.x1 = t2x.i; p.y1 = t2y.i;
    bool amb2 = near_pixel_boundary(p3x, L1) || near_pixel_boundary(p3y, L1);
    TruncResult t3x = trunc_index(p3x, mv.W), t3y = trunc_index(p3y, mv.H);
    // (a finite in-range elbow + a link length cannot give a non-finite tip: the flag case of
    // arm_config is unreachable here; it is kept as "blocked" for safety)
    if (t3x.flag || t3y.flag) { p.after = 0; return p; }
    if (amb2) { p.after = 2; return p; }
    if (!t3x.in_range || !t3y.in_range) { p.after = 0; return p; }
    p.after = 3;
    p.x2 = t3x.i; p.y2 = t3y.i;
    return p;
}


// One thread checks one whole arm edge: 1 free, 0 blocked, 2 ambiguous (arm.cu header).
template <bool SMEM>
__device__ __forceinline__ int arm_edge_seq(const uint32_t *words, const MapView &mv, double L0,
                                            double L1, const double (&c1)[4], const double (&c2)[4],
                                            int &flag) {
    TruncResult t0 = trunc_index(c1[0], mv.W), t1 = trunc_index(c1[1], mv.H);
    TruncResult t2 = trunc_index(c2[0], mv.W), t3 = trunc_index(c2[1], mv.H);
    int f = t0.flag ? t0.flag : t1.flag ? t1.flag : t2.flag ? t2.flag : t3.flag;
    if (!f && (isinf(c1[2]) || isinf(c1[3]) || isinf(c2[2]) || isinf(c2[3]) || c1[2] != c1[2] ||
               c1[3] != c1[3] || c2[2] != c2[2] || c2[3] != c2[3]))
        f = SBP_FLAG_NAN;
    flag |= f;
    if (f || !(t0.in_range && t1.in_range && t2.in_range && t3.in_range)) return 0;
    int x1 = t0.i, y1 = t1.i, x2 = t2.i, y2 = t3.i;
    int dx = x2 - x1, dy = y2 - y1;
    bool steep = abs(dy) > abs(dx);
    int a1 = steep ? y1 : x1, b1 = steep ? x1 : y1, a2 = steep ? y2 : x2, b2 = steep ? x2 : y2;
    bool swapped = a1 > a2;
    if (swapped) { int t = a1; a1 = a2; a2 = t; t = b1; b1 = b2; b2 = t; }
    uint32_t da = (uint32_t)(a2 - a1), adb = (uint32_t)abs(b2 - b1);
    int ystep = b1 < b2 ? 1 : -1;
    const uint32_t npx = da + 1u, N = npx < 2u ? 2u : npx;
    const uint32_t c = da ? da - 1u - (da >> 1) : 0u;
    const double inv = __ddiv_rn(1.0, (double)(N - 1u));
    const double st0 = __dmul_rn(inv, __dsub_rn(c2[2], c1[2]));
    const double st1 = __dmul_rn(inv, __dsub_rn(c2[3], c1[3]));
    int state = 1;
    for (uint32_t o = 0; o < N; ++o) {
        uint32_t po = npx == 1u ? 0u : o;
        uint64_t ci = swapped ? (uint64_t)(npx - 1u - po) : (uint64_t)po;
        uint32_t k = da ? (uint32_t)((ci * adb + c) / da) : 0u;
        int aa = a1 + (int)ci, bb = b1 + ystep * (int)k;
        double bx = (double)(steep ? bb : aa), by = (double)(steep ? aa : bb);
        double r0 = __dadd_rn(__dmul_rn(st0, (double)o), c1[2]);
        double r1 = __dadd_rn(__dmul_rn(st1, (double)o), c1[3]);
        int r = arm_config<SMEM>(words, mv, L0, L1, bx, by, r0, r1, flag);
        if (r == 0) return 0;
        if (r == 2) state = 2;
    }
    return state;
}

// cell of a coordinate in the uniform G x G grid over [mn, mx] (kNN cell grid; monotone in v)
__device__ __forceinline__ int cell_of(double v, double mn, double mx, int G) {
    double w = mx - mn;
    if (!(w > 0.0)) return 0;
    double c = floor((v - mn) * ((double)G / w));
    if (!(c >= 0.0)) return 0;                                      // negative or NaN
    return c > (double)(G - 1) ? G - 1 : (int)c;
}

// ======================= distances (rrtPlanner.py:278, prmPlanner.py:39, engine.py:13-24) =====
template <int D>
__device__ __forceinline__ double dist2_full(const double (&p)[D], const double (&q)[D]) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
        double df = __dsub_rn(p[j], q[j]);        // poses - pos
        double sq = __dmul_rn(df, df);
        acc = j == 0 ? sq : __dadd_rn(acc, sq);   // left-to-right, unfused
    }
    return acc;
}

template <int D, bool XY>
__device__ __forceinline__ double near_d2(const double (&p)[D], const double (&q)[D]) {
    if (XY) {
        // engine.euclidean_dist (engine.py:23-24): p = p1 - p2 (query - node), first
        // two coordinates only.  The reference squares with libm pow, which differs
        // from the correctly rounded product by 1 ulp for ~0.1% of inputs (SURVEY.md
        // D4): "parity unpinned at the last bit" for this metric.
        double a = __dsub_rn(q[0], p[0]), b = __dsub_rn(q[1], p[1]);
        return __dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b));
    }
    return dist2_full<D>(p, q);
}

__device__ __forceinline__ bool near_in(double d2, double thr, int closed) {
    return closed ? (d2 <= thr) : (d2 < thr);
}

// Smallest fp64 t with sqrt_rn(t) >= r  (d < r  <=>  d^2 < t), and largest t with
// sqrt_rn(t) <= r  (d <= r <=> d^2 <= t); SURVEY.md H5.  Host sqrt is IEEE
// correctly rounded, like __dsqrt_rn.
static inline void near_threshold(double r, int closed, double *thr, int *closed_out) {
    *closed_out = closed;
    if (r != r) { *thr = -1.0; *closed_out = 0; return; }            // nan: nothing matches
    if (r < 0) { *thr = -1.0; *closed_out = 0; return; }             // d >= 0 always
    if (isinf(r)) { *thr = INFINITY; *closed_out = 1; return; }      // finite d always matches
    double s = r * r;
    if (isinf(s)) { *thr = INFINITY; *closed_out = 1; return; }
    if (!closed) {
        while (s > 0 && sqrt(s) >= r) s = nextafter(s, -INFINITY);
        while (sqrt(nextafter(s, INFINITY)) < r) s = nextafter(s, INFINITY);
        // now sqrt(s) < r (or s == 0) and sqrt(next(s)) >= r
        *thr = (sqrt(s) >= r) ? s : nextafter(s, INFINITY);
    } else {
        while (sqrt(s) > r) s = nextafter(s, -INFINITY);
        while (sqrt(nextafter(s, INFINITY)) <= r) s = nextafter(s, INFINITY);
        *thr = s;
    }
}

