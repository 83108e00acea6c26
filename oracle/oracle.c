/*
 * oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, optional OpenMP) of the collision / nearest-
 * neighbour arithmetic of soraxas/sbp-env v2.0.2.  It is the checker the GPU
 * path is compared against and the CPU baseline bench.py times; it is never
 * linked, imported or called by the product package (sbp_env_b200/).
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function
 * here against (i) the reference's own known-answer tests
 * (tests/test_image_space_collision_checker.py:33-60,
 *  tests/test_4d_arm_collision_checker.py:23-75, tests/test_rrtPlanner.py:136-148,
 *  the get_line docstring collisionChecker.py:163-169) and (ii) fixtures under
 * tests/golden/ produced by oracle/gen_golden.py from the *imported, unmodified*
 * reference on seeded random + adversarial inputs.  The two sub-ulp caveats
 * (glibc cos/sin in the arm's forward kinematics, libm pow in engine.dist) are
 * inherited from the platform's libm exactly as CPython inherits them.
 *
 * The map is the reference's `_img` as a dense uint8 array img[x*H + y]
 * (1 = free), i.e. deliberately NOT the product's bit-packed tile layout, so
 * that the two implementations share no indexing code.
 *
 * All citations are file:line under /root/reference/sbp_env/.
 *
 * Build with -ffp-contract=off: every fp64 operation below must round
 * separately, as numpy / CPython evaluate them.
 */
#if defined(__FP_FAST_FMA) && !defined(ORC_ALLOW_FMA)
/* fine: contraction is still disabled by -ffp-contract=off in the Makefile */
#endif
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_OK 0
#define ORC_NAN 1      /* int(nan)  -> ValueError  in CPython */
#define ORC_INF 2      /* int(inf)  -> OverflowError in CPython */

typedef struct {
    const uint8_t *img; /* [W][H], 1 = free */
    int64_t W, H;
} orc_map;

void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* CPython int(float): truncation toward zero; nan raises ValueError, inf raises
 * OverflowError.  Finite values beyond the int64 range cannot index any array;
 * they are clamped to a value that is out of range for every map. */
static int py_int(double v, int64_t *out) {
    if (v != v) return ORC_NAN;
    if (isinf(v)) return ORC_INF;
    if (v >= 9223372036854775808.0) { *out = INT64_MAX / 2; return ORC_OK; }
    if (v < -9223372036854775808.0) { *out = INT64_MIN / 2; return ORC_OK; }
    *out = (int64_t)v; /* C conversion truncates toward zero */
    return ORC_OK;
}

/* numpy 2.3 quirk, pinned by tests/golden/img_cc.npz: an index in [2^63, 2^64)
 * raises OverflowError ("Python int too large to convert to C long") instead of
 * IndexError; every other out-of-range magnitude is a plain IndexError -> False. */
static int numpy_index_overflows(double v) {
    return v >= 9223372036854775808.0 && v < 18446744073709551616.0;
}

/* numpy indexing `_img[x, y] == 1` with IndexError -> False
 * (collisionChecker.py:147-153, :393-402).  Negative indices wrap once. */
static int cell_free(const orc_map *m, int64_t x, int64_t y) {
    if (x < -m->W || x >= m->W || y < -m->H || y >= m->H) return 0;
    if (x < 0) x += m->W;
    if (y < 0) y += m->H;
    return m->img[x * m->H + y] == 1;
}

static int in_index_range(const orc_map *m, int64_t x, int64_t y) {
    return !(x < -m->W || x >= m->W || y < -m->H || y >= m->H);
}

/* ImgCollisionChecker.feasible (collisionChecker.py:147-153).
 * status[i]: ORC_OK / ORC_NAN / ORC_INF (what CPython would raise). */
void orc_feasible2d(const uint8_t *img, int64_t W, int64_t H, const double *P,
                    int64_t n, uint8_t *out, uint8_t *status) {
    orc_map m = {img, W, H};
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        int64_t x, y;
        int s = py_int(P[2 * i], &x);
        if (!s) s = py_int(P[2 * i + 1], &y);
        if (!s && (numpy_index_overflows(P[2 * i]) || numpy_index_overflows(P[2 * i + 1]))) s = ORC_INF;
        if (status) status[i] = (uint8_t)s;
        out[i] = s ? 0 : (uint8_t)cell_free(&m, x, y);
    }
}

/* Number of pixels get_line emits: max(|dx|,|dy|)+1 (collisionChecker.py:204). */
int64_t orc_line_len(int64_t x1, int64_t y1, int64_t x2, int64_t y2) {
    int64_t dx = llabs(x2 - x1), dy = llabs(y2 - y1);
    return (dx > dy ? dx : dy) + 1;
}

/* ImgCollisionChecker.get_line / RobotArm4dCollisionChecker._get_line
 * (collisionChecker.py:155-215, :443-503) on already-truncated endpoints.
 * Writes orc_line_len() (x,y) pairs in start->end order. */
void orc_get_line(int64_t x1, int64_t y1, int64_t x2, int64_t y2, int64_t *xy) {
    int64_t dx = x2 - x1, dy = y2 - y1, t;
    int steep = llabs(dy) > llabs(dx);            /* :179 */
    if (steep) { t = x1; x1 = y1; y1 = t; t = x2; x2 = y2; y2 = t; } /* :182-184 */
    int swapped = 0;
    if (x1 > x2) {                                 /* :188-191 */
        t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; swapped = 1;
    }
    dx = x2 - x1; dy = y2 - y1;                    /* :194-195 */
    int64_t err = dx / 2;                          /* int(dx / 2.0), dx >= 0 (:198) */
    int64_t ystep = y1 < y2 ? 1 : -1;              /* :199 */
    int64_t ady = llabs(dy), y = y1, n = dx + 1, k = 0;
    for (int64_t x = x1; x <= x2; ++x, ++k) {      /* :204-210 */
        int64_t slot = swapped ? (n - 1 - k) : k;  /* points.reverse() (:213-214) */
        xy[2 * slot] = steep ? y : x;
        xy[2 * slot + 1] = steep ? x : y;
        err -= ady;
        if (err < 0) { y += ystep; err += dx; }
    }
}

/* Longest line the oracle materialises.  Longer lines can only arise from an
 * endpoint far outside the index range, for which the all-of is False (both
 * endpoints are members of the pixel list, :204-206) and the reference itself
 * would have to build a list of that many tuples first. */
#define ORC_MAX_LINE ((int64_t)1 << 24)

/* Walk the line in start->end order; returns 1 if all pixels are free.
 * *tested = pixels examined up to and including the first blocked one;
 * (*lastx,*lasty) = last pixel examined. */
static int walk_line(const orc_map *m, int64_t x1, int64_t y1, int64_t x2,
                     int64_t y2, int64_t *tested, int64_t *lastx, int64_t *lasty) {
    if (!in_index_range(m, x1, y1)) {          /* pixel 0 is the start (:140-142) */
        if (tested) *tested = 1;
        if (lastx) { *lastx = x1; *lasty = y1; }
        return 0;
    }
    int64_t n = orc_line_len(x1, y1, x2, y2);
    if (n > ORC_MAX_LINE) {                     /* end pixel is out of range */
        if (tested) *tested = -1;
        if (lastx) { *lastx = x2; *lasty = y2; }
        return 0;
    }
    int64_t stackbuf[2 * 512];
    int64_t *xy = n > 512 ? (int64_t *)malloc(sizeof(int64_t) * 2 * (size_t)n) : stackbuf;
    int ok = 1;
    int64_t k;
    orc_get_line(x1, y1, x2, y2, xy);           /* whole list first (:138) */
    for (k = 0; k < n; ++k) {                   /* :140-142 */
        if (lastx) { *lastx = xy[2 * k]; *lasty = xy[2 * k + 1]; }
        if (!cell_free(m, xy[2 * k], xy[2 * k + 1])) { ok = 0; ++k; break; }
    }
    if (tested) *tested = k;
    if (xy != stackbuf) free(xy);
    return ok;
}

/* ImgCollisionChecker.visible (collisionChecker.py:134-145).
 * status: ORC_NAN edges return False (ValueError is caught, :143-144);
 * ORC_INF is what propagates as OverflowError in CPython.
 * px_full / px_tested (nullable): per-edge full line length and pixels
 * examined until the first blocked one. */
void orc_visible2d(const uint8_t *img, int64_t W, int64_t H, const double *P,
                   const double *Q, int64_t n, uint8_t *out, uint8_t *status,
                   int64_t *px_full, int64_t *px_tested) {
    orc_map m = {img, W, H};
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; ++i) {
        int64_t x1 = 0, y1 = 0, x2 = 0, y2 = 0, tested = 0;
        int s = py_int(P[2 * i], &x1);             /* :173 */
        if (!s) s = py_int(P[2 * i + 1], &y1);
        if (!s) s = py_int(Q[2 * i], &x2);         /* :174 */
        if (!s) s = py_int(Q[2 * i + 1], &y2);
        if (status) status[i] = (uint8_t)s;
        if (s) { out[i] = 0; if (px_full) px_full[i] = 0; if (px_tested) px_tested[i] = 0; continue; }
        out[i] = (uint8_t)walk_line(&m, x1, y1, x2, y2, &tested, 0, 0);
        if (px_full) px_full[i] = orc_line_len(x1, y1, x2, y2);
        if (px_tested) px_tested[i] = tested;
    }
}

/* ImgCollisionChecker.get_coor_before_collision (collisionChecker.py:118-132):
 * the last pixel examined walking start->end (the first blocked one, or the
 * end pixel when the line is free).  out_xy[2*i..] */
void orc_coor_before_collision(const uint8_t *img, int64_t W, int64_t H,
                               const double *P, const double *Q, int64_t n,
                               int64_t *out_xy, uint8_t *status) {
    orc_map m = {img, W, H};
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; ++i) {
        int64_t x1 = 0, y1 = 0, x2 = 0, y2 = 0, t, lx = 0, ly = 0;
        int s = py_int(P[2 * i], &x1);
        if (!s) s = py_int(P[2 * i + 1], &y1);
        if (!s) s = py_int(Q[2 * i], &x2);
        if (!s) s = py_int(Q[2 * i + 1], &y2);
        if (status) status[i] = (uint8_t)s;
        if (s) { out_xy[2 * i] = out_xy[2 * i + 1] = 0; continue; }
        walk_line(&m, x1, y1, x2, y2, &t, &lx, &ly);
        out_xy[2 * i] = lx; out_xy[2 * i + 1] = ly;
    }
}

/* RobotArm4dCollisionChecker.feasible (collisionChecker.py:404-426) with
 * get_pt_from_angle_and_length (:428-441): absolute joint angles, un-truncated
 * intermediate points, link 1 fully before link 2.
 * px (nullable) accumulates pixel tests performed. */
static int arm_config_free(const orc_map *m, double L0, double L1, double bx,
                           double by, double r0, double r1, int *status,
                           int64_t *px) {
    double p2x = bx + L0 * cos(r0);                 /* :438 */
    double p2y = by + L0 * sin(r0);                 /* :439 */
    int64_t ax, ay, cx, cy, ex, ey, t = 0;
    int s = py_int(bx, &ax);                        /* :461 */
    if (!s) s = py_int(by, &ay);
    if (!s) s = py_int(p2x, &cx);                   /* :462 */
    if (!s) s = py_int(p2y, &cy);
    if (s) { *status = s; return 0; }
    int ok = walk_line(m, ax, ay, cx, cy, &t, 0, 0);   /* :413-415 */
    if (px) *px += t;
    if (!ok) return 0;
    double p3x = p2x + L1 * cos(r1);                /* :417-419 */
    double p3y = p2y + L1 * sin(r1);
    s = py_int(p3x, &ex);
    if (!s) s = py_int(p3y, &ey);
    if (s) { *status = s; return 0; }
    ok = walk_line(m, cx, cy, ex, ey, &t, 0, 0);    /* :422-424 */
    if (px) *px += t;
    return ok;
}

void orc_arm_feasible(const uint8_t *img, int64_t W, int64_t H, double L0,
                      double L1, const double *C, int64_t n, uint8_t *out,
                      uint8_t *status) {
    orc_map m = {img, W, H};
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; ++i) {
        int s = 0;
        const double *c = C + 4 * i;
        /* math.cos(nan) returns nan -> int(nan) ValueError; math.cos(inf) raises
         * ValueError("math domain error"): both surface as ORC_NAN-class. */
        if (isinf(c[2]) || isinf(c[3])) { s = ORC_NAN; out[i] = 0; }
        else out[i] = (uint8_t)arm_config_free(&m, L0, L1, c[0], c[1], c[2], c[3], &s, 0);
        if (s) out[i] = 0;
        if (status) status[i] = (uint8_t)s;
    }
}

/* RobotArm4dCollisionChecker.visible (collisionChecker.py:385-391) over
 * _interpolate_configs (:366-383) and create_ranges (:346-364).
 * n_cfg (nullable): number of interpolated configs N of each edge;
 * cfg_tested (nullable): configs examined until the first infeasible one. */
void orc_arm_visible(const uint8_t *img, int64_t W, int64_t H, double L0,
                     double L1, const double *C1, const double *C2, int64_t n,
                     uint8_t *out, uint8_t *status, int64_t *n_cfg,
                     int64_t *cfg_tested) {
    orc_map m = {img, W, H};
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t e = 0; e < n; ++e) {
        const double *a = C1 + 4 * e, *b = C2 + 4 * e;
        int64_t x1, y1, x2, y2;
        int s = py_int(a[0], &x1);
        if (!s) s = py_int(a[1], &y1);
        if (!s) s = py_int(b[0], &x2);
        if (!s) s = py_int(b[1], &y2);
        if (!s && (a[2] != a[2] || a[3] != a[3] || b[2] != b[2] || b[3] != b[3] ||
                   isinf(a[2]) || isinf(a[3]) || isinf(b[2]) || isinf(b[3])))
            s = ORC_NAN;
        if (status) status[e] = (uint8_t)s;
        if (s) { out[e] = 0; if (n_cfg) n_cfg[e] = 0; if (cfg_tested) cfg_tested[e] = 0; continue; }
        int64_t npx = orc_line_len(x1, y1, x2, y2);
        int64_t N = npx < 2 ? 2 : npx;             /* :375-377 */
        if (n_cfg) n_cfg[e] = N;
        /* config 0 / config N-1 start their first link on the base end pixels, so
         * an out-of-range base endpoint makes the all-of False */
        if (!in_index_range(&m, x1, y1)) { out[e] = 0; if (cfg_tested) cfg_tested[e] = 1; continue; }
        if (npx > ORC_MAX_LINE) { out[e] = 0; if (cfg_tested) cfg_tested[e] = -1; continue; }
        int64_t *xy = (int64_t *)malloc(sizeof(int64_t) * 2 * (size_t)N);
        orc_get_line(x1, y1, x2, y2, xy);          /* :373 */
        if (npx == 1) { xy[2] = xy[0]; xy[3] = xy[1]; }
        /* create_ranges: steps = (1.0/(N-1))*(stop-start); rot = steps*i + start,
         * three separately rounded fp64 operations (:363-364) */
        double inv = 1.0 / (double)(N - 1);
        double st0 = inv * (b[2] - a[2]);
        double st1 = inv * (b[3] - a[3]);
        int ok = 1;
        int64_t k;
        for (k = 0; k < N; ++k) {                  /* :388-390 */
            double r0 = st0 * (double)k + a[2];    /* built with -ffp-contract=off */
            double r1 = st1 * (double)k + a[3];
            int cs = 0;
            if (!arm_config_free(&m, L0, L1, (double)xy[2 * k], (double)xy[2 * k + 1],
                                 r0, r1, &cs, 0)) { ok = 0; ++k; break; }
        }
        free(xy);
        out[e] = (uint8_t)ok;
        if (cfg_tested) cfg_tested[e] = k;
    }
}

/* create_ranges (collisionChecker.py:346-364), endpoint=True: out[k*N + i]. */
void orc_create_ranges(const double *start, const double *stop, int64_t K,
                       int64_t N, double *out) {
    double inv = 1.0 / (double)(N - 1);
    for (int64_t k = 0; k < K; ++k) {
        double step = inv * (stop[k] - start[k]);
        for (int64_t i = 0; i < N; ++i) {
            out[k * N + i] = step * (double)i + start[k];
        }
    }
}

/* np.linalg.norm(poses - q, axis=1): sqrt of the left-to-right, unfused sum of
 * exact-rounded squares (rrtPlanner.py:278, prmPlanner.py:39). */
static inline double norm_full(const double *p, const double *q, int d) {
    double acc = 0.0;
    for (int j = 0; j < d; ++j) {
        double df = p[j] - q[j];
        double sq = df * df;
        acc = j == 0 ? sq : acc + sq;
    }
    return sqrt(acc);
}

/* engine.euclidean_dist (engine.py:13-24): first two coordinates only; the
 * squares go through libm pow exactly as `np.float64 ** 2` does. */
static double (*volatile libm_pow)(double, double) = pow; /* keep gcc from folding pow(x,2) into x*x */
static inline double norm_xy_pow(const double *p, const double *q) {
    double a = p[0] - q[0], b = p[1] - q[1];
    double s0 = libm_pow(a, 2.0), s1 = libm_pow(b, 2.0);
    return sqrt(s0 + s1);
}

/* find_nearest_neighbour_idx (rrtPlanner.py:268-279): argmin, first minimum wins. */
void orc_nearest(const double *poses, int64_t N, int d, const double *X,
                 int64_t m, int64_t *idx, double *dist) {
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < m; ++q) {
        int64_t best = -1; double bd = INFINITY;
        for (int64_t i = 0; i < N; ++i) {
            double v = norm_full(poses + i * d, X + q * d, d);
            /* np.argmin: a NaN is returned as soon as it is seen */
            if (v != v) { if (bd == bd) { bd = v; best = i; } continue; }
            if (bd == bd && (best < 0 || v < bd)) { bd = v; best = i; }
        }
        idx[q] = best; if (dist) dist[q] = bd;
    }
}

/* k nearest by (distance, index) -- the stable order np.argsort(kind="stable")
 * gives, and what argmin gives for k=1 on NaN-free rows.  numpy sorts NaN distances
 * last (after +inf), in index order.  idx/dist are [m][k]; rows are padded with
 * -1 / +inf when N < k. */
void orc_knn(const double *poses, int64_t N, int d, const double *X, int64_t m,
             int k, int64_t *idx, double *dist) {
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t q = 0; q < m; ++q) {
        int64_t *bi = idx + q * k; double *bd = dist + q * k;
        int64_t nan_i[64]; int nn = 0;
        int cnt = 0;
        for (int j = 0; j < k; ++j) { bi[j] = -1; bd[j] = INFINITY; }
        for (int64_t i = 0; i < N; ++i) {
            double v = norm_full(poses + i * d, X + q * d, d);
            if (v != v) { if (nn < k && nn < 64) nan_i[nn++] = i; continue; }
            if (cnt == k && !(v < bd[k - 1])) continue;
            int pos = cnt < k ? cnt : k - 1;
            while (pos > 0 && v < bd[pos - 1]) { bd[pos] = bd[pos - 1]; bi[pos] = bi[pos - 1]; --pos; }
            bd[pos] = v; bi[pos] = i;
            if (cnt < k) ++cnt;
        }
        for (int j = 0; cnt < k && j < nn; ++j, ++cnt) { bi[cnt] = nan_i[j]; bd[cnt] = NAN; }
    }
}

/* Radius query, index-ascending (prmPlanner.py:28-44 `<`; rrtPlanner.py:177-181
 * `<=` on engine.dist).  metric: 0 = all d dims via np.linalg.norm,
 * 1 = first two dims via engine.dist (pow).  closed: 0 `<`, 1 `<=`.
 * Pass idx = NULL to count only; row_ptr is [m+1]. */
void orc_near(const double *poses, int64_t N, int d, const double *X, int64_t m,
              double r, int metric, int closed, int64_t *row_ptr, int64_t *idx) {
    int64_t *cnt = (int64_t *)calloc((size_t)m + 1, sizeof(int64_t));
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t q = 0; q < m; ++q) {
        int64_t c = 0;
        for (int64_t i = 0; i < N; ++i) {
            double v = metric ? norm_xy_pow(X + q * d, poses + i * d)
                              : norm_full(poses + i * d, X + q * d, d);
            c += closed ? (v <= r) : (v < r);
        }
        cnt[q] = c;
    }
    row_ptr[0] = 0;
    for (int64_t q = 0; q < m; ++q) row_ptr[q + 1] = row_ptr[q] + cnt[q];
    free(cnt);
    if (!idx) return;
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t q = 0; q < m; ++q) {
        int64_t o = row_ptr[q];
        for (int64_t i = 0; i < N; ++i) {
            double v = metric ? norm_xy_pow(X + q * d, poses + i * d)
                              : norm_full(poses + i * d, X + q * d, d);
            if (closed ? (v <= r) : (v < r)) idx[o++] = i;
        }
    }
}

/* All distances (for dumping / small checks): out[q*N + i]. */
void orc_dist_matrix(const double *poses, int64_t N, int d, const double *X,
                     int64_t m, int metric, double *out) {
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < m; ++q)
        for (int64_t i = 0; i < N; ++i)
            out[q * N + i] = metric ? norm_xy_pow(X + q * d, poses + i * d)
                                    : norm_full(poses + i * d, X + q * d, d);
}
