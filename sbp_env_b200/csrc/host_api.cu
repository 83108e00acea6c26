// host_api.cu -- the *_host entry points: what the Python plugin calls when the
// user hands it host arrays (numpy).  Each one stages the inputs on the device,
// launches the same kernels as the device-pointer entry points, copies the result
// back and synchronises the context's stream.  There is no host computation of
// the path here; the only host arithmetic is sbp_arm_resolve_host (arm.cu).
#include "common.cuh"

namespace {

struct Staging {
    sbp_ctx *ctx;
    int32_t rc = SBP_OK;
    explicit Staging(sbp_ctx *c) : ctx(c) {}
    void *dev(int slot, size_t bytes) {
        void *p = nullptr;
        if (rc == SBP_OK) rc = sbp_ctx_scratch(ctx, slot, bytes, &p);
        return p;
    }
    void h2d(void *dst, const void *src, size_t bytes) {
        if (rc != SBP_OK || bytes == 0) return;
        cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) { sbp_set_error("H2D copy failed: %s", cudaGetErrorString(e)); rc = SBP_ERR_CUDA; }
    }
    void d2h(void *dst, const void *src, size_t bytes) {
        if (rc != SBP_OK || bytes == 0) return;
        cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) { sbp_set_error("D2H copy failed: %s", cudaGetErrorString(e)); rc = SBP_ERR_CUDA; }
    }
    void zero_flags() {
        if (rc != SBP_OK) return;
        cudaError_t e = cudaMemsetAsync(ctx->d_flags, 0, 16 * sizeof(int32_t), ctx->stream);
        if (e != cudaSuccess) { sbp_set_error("memset failed: %s", cudaGetErrorString(e)); rc = SBP_ERR_CUDA; }
    }
    void sync() {
        if (rc != SBP_OK) return;
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { sbp_set_error("stream sync failed: %s", cudaGetErrorString(e)); rc = SBP_ERR_CUDA; }
    }
};

}  // namespace

extern "C" int32_t sbp_feasible2d_host(sbp_map *map, const double *P, int64_t n, uint8_t *out,
                                       int32_t *flags) {
    SBP_REQUIRE(map && (n == 0 || (P && out)), "sbp_feasible2d_host: NULL argument");
    SBP_REQUIRE(n >= 0, "sbp_feasible2d_host: n < 0");
    if (flags) *flags = 0;
    if (n == 0) return SBP_OK;
    sbp_ctx *ctx = map->ctx;
    SBP_DEVICE(ctx);
    Staging st(ctx);
    double *dP = (double *)st.dev(2, (size_t)n * 16);
    uint8_t *dO = (uint8_t *)st.dev(3, (size_t)n);
    st.zero_flags();
    st.h2d(dP, P, (size_t)n * 16);
    if (st.rc == SBP_OK) st.rc = sbp_feasible2d(map, dP, n, dO, ctx->d_flags);
    st.d2h(out, dO, (size_t)n);
    st.d2h(ctx->h_flags, ctx->d_flags, sizeof(int32_t));
    st.sync();
    if (st.rc == SBP_OK && flags) *flags = ctx->h_flags[0];
    return st.rc;
}

extern "C" int32_t sbp_visible2d_host(sbp_map *map, const double *P, const double *Q, int64_t n,
                                      uint8_t *out, int32_t *flags) {
    SBP_REQUIRE(map && (n == 0 || (P && Q && out)), "sbp_visible2d_host: NULL argument");
    SBP_REQUIRE(n >= 0, "sbp_visible2d_host: n < 0");
    if (flags) *flags = 0;
    if (n == 0) return SBP_OK;
    sbp_ctx *ctx = map->ctx;
    SBP_DEVICE(ctx);
    Staging st(ctx);
    double *dP = (double *)st.dev(2, (size_t)n * 16);
    double *dQ = (double *)st.dev(3, (size_t)n * 16);
    uint8_t *dO = (uint8_t *)st.dev(4, (size_t)n);
    st.zero_flags();
    st.h2d(dP, P, (size_t)n * 16);
    st.h2d(dQ, Q, (size_t)n * 16);
    if (st.rc == SBP_OK) st.rc = sbp_visible2d(map, dP, dQ, n, dO, ctx->d_flags, nullptr);
    st.d2h(out, dO, (size_t)n);
    st.d2h(ctx->h_flags, ctx->d_flags, sizeof(int32_t));
    st.sync();
    if (st.rc == SBP_OK && flags) *flags = ctx->h_flags[0];
    return st.rc;
}

extern "C" int32_t sbp_first_blocked2d_host(sbp_map *map, const double *P, const double *Q,
                                            int64_t n, int32_t *out_xy, int32_t *flags) {
    SBP_REQUIRE(map && (n == 0 || (P && Q && out_xy)), "sbp_first_blocked2d_host: NULL argument");
    SBP_REQUIRE(n >= 0, "sbp_first_blocked2d_host: n < 0");
    if (flags) *flags = 0;
    if (n == 0) return SBP_OK;
    sbp_ctx *ctx = map->ctx;
    SBP_DEVICE(ctx);
    Staging st(ctx);
    double *dP = (double *)st.dev(2, (size_t)n * 16);
    double *dQ = (double *)st.dev(3, (size_t)n * 16);
    int32_t *dO = (int32_t *)st.dev(4, (size_t)n * 8);
    st.zero_flags();
    st.h2d(dP, P, (size_t)n * 16);
    st.h2d(dQ, Q, (size_t)n * 16);
    if (st.rc == SBP_OK) st.rc = sbp_first_blocked2d(map, dP, dQ, n, dO, ctx->d_flags);
    st.d2h(out_xy, dO, (size_t)n * 8);
    st.d2h(ctx->h_flags, ctx->d_flags, sizeof(int32_t));
    st.sync();
    if (st.rc == SBP_OK && flags) *flags = ctx->h_flags[0];
    return st.rc;
}

static int64_t count_ambiguous(const uint8_t *out, int64_t n) {
    int64_t c = 0;
    for (int64_t i = 0; i < n; ++i) c += out[i] == SBP_AMBIGUOUS;
    return c;
}

extern "C" int32_t sbp_arm_feasible_host(sbp_map *map, double L0, double L1, const double *C,
                                         int64_t n, uint8_t *out, int32_t *flags,
                                         int64_t *n_ambiguous) {
    SBP_REQUIRE(map && (n == 0 || (C && out)), "sbp_arm_feasible_host: NULL argument");
    SBP_REQUIRE(n >= 0, "sbp_arm_feasible_host: n < 0");
    if (flags) *flags = 0;
    if (n_ambiguous) *n_ambiguous = 0;
    if (n == 0) return SBP_OK;
    sbp_ctx *ctx = map->ctx;
    SBP_DEVICE(ctx);
    Staging st(ctx);
    double *dC = (double *)st.dev(2, (size_t)n * 32);
    uint8_t *dO = (uint8_t *)st.dev(4, (size_t)n);
    st.zero_flags();
    st.h2d(dC, C, (size_t)n * 32);
    if (st.rc == SBP_OK) st.rc = sbp_arm_feasible(map, L0, L1, dC, n, dO, ctx->d_flags);
    st.d2h(out, dO, (size_t)n);
    st.d2h(ctx->h_flags, ctx->d_flags, sizeof(int32_t));
    st.sync();
    if (st.rc != SBP_OK) return st.rc;
    if (flags) *flags = ctx->h_flags[0];
    int64_t amb = count_ambiguous(out, n);
    if (n_ambiguous) *n_ambiguous = amb;
    if (amb) return sbp_arm_resolve_host(map, L0, L1, C, nullptr, n, out, nullptr);
    return SBP_OK;
}

extern "C" int32_t sbp_arm_visible_host(sbp_map *map, double L0, double L1, const double *C1,
                                        const double *C2, int64_t n, uint8_t *out, int32_t *flags,
                                        int64_t *n_ambiguous) {
    SBP_REQUIRE(map && (n == 0 || (C1 && C2 && out)), "sbp_arm_visible_host: NULL argument");
    SBP_REQUIRE(n >= 0, "sbp_arm_visible_host: n < 0");
    if (flags) *flags = 0;
    if (n_ambiguous) *n_ambiguous = 0;
    if (n == 0) return SBP_OK;
    sbp_ctx *ctx = map->ctx;
    SBP_DEVICE(ctx);
    Staging st(ctx);
    double *d1 = (double *)st.dev(2, (size_t)n * 32);
    double *d2 = (double *)st.dev(3, (size_t)n * 32);
    uint8_t *dO = (uint8_t *)st.dev(4, (size_t)n);
    st.zero_flags();
    st.h2d(d1, C1, (size_t)n * 32);
    st.h2d(d2, C2, (size_t)n * 32);
    if (st.rc == SBP_OK) st.rc = sbp_arm_visible(map, L0, L1, d1, d2, n, dO, ctx->d_flags, nullptr);
    st.d2h(out, dO, (size_t)n);
    st.d2h(ctx->h_flags, ctx->d_flags, sizeof(int32_t));
    st.sync();
    if (st.rc != SBP_OK) return st.rc;
    if (flags) *flags = ctx->h_flags[0];
    int64_t amb = count_ambiguous(out, n);
    if (n_ambiguous) *n_ambiguous = amb;
    if (amb) return sbp_arm_resolve_host(map, L0, L1, C1, C2, n, out, nullptr);
    return SBP_OK;
}

extern "C" int32_t sbp_nodes_knn_host(sbp_nodes *s, const double *X, int64_t m, int32_t k,
                                      int32_t precision, int64_t *idx, double *dist) {
    SBP_REQUIRE(s && (m == 0 || (X && idx)), "sbp_nodes_knn_host: NULL argument");
    SBP_REQUIRE(m >= 0 && k >= 1, "sbp_nodes_knn_host: bad sizes");
    if (m == 0) return SBP_OK;
    sbp_ctx *ctx = s->ctx;
    SBP_DEVICE(ctx);
    Staging st(ctx);
    double *dX = (double *)st.dev(2, (size_t)m * s->d * 8);
    int64_t *dI = (int64_t *)st.dev(3, (size_t)m * k * 8);
    double *dD = (double *)st.dev(4, (size_t)m * k * 8);
    st.h2d(dX, X, (size_t)m * s->d * 8);
    if (st.rc == SBP_OK) st.rc = sbp_nodes_knn(s, dX, m, k, precision, dI, dist ? dD : nullptr);
    st.d2h(idx, dI, (size_t)m * k * 8);
    if (dist) st.d2h(dist, dD, (size_t)m * k * (precision == 32 ? 4 : 8));    // 32: a float buffer
    st.sync();
    return st.rc;
}

extern "C" int32_t sbp_nodes_nearest_host(sbp_nodes *s, const double *X, int64_t m, int64_t *idx,
                                          double *dist) {
    SBP_REQUIRE(s && (m == 0 || (X && idx)), "sbp_nodes_nearest_host: NULL argument");
    SBP_REQUIRE(m >= 0, "sbp_nodes_nearest_host: m < 0");
    if (m == 0) return SBP_OK;
    sbp_ctx *ctx = s->ctx;
    SBP_DEVICE(ctx);
    Staging st(ctx);
    double *dX = (double *)st.dev(2, (size_t)m * s->d * 8);
    int64_t *dI = (int64_t *)st.dev(3, (size_t)m * 8);
    double *dD = (double *)st.dev(4, (size_t)m * 8);
    st.h2d(dX, X, (size_t)m * s->d * 8);
    if (st.rc == SBP_OK) st.rc = sbp_nodes_nearest(s, dX, m, dI, dist ? dD : nullptr);
    st.d2h(idx, dI, (size_t)m * 8);
    if (dist) st.d2h(dist, dD, (size_t)m * 8);
    st.sync();
    return st.rc;
}

extern "C" int32_t sbp_nodes_near_host(sbp_nodes *s, const double *X, int64_t m, double r,
                                       int32_t metric, int32_t closed, int64_t *row_ptr,
                                       int64_t *idx, int64_t idx_capacity, int64_t *needed) {
    SBP_REQUIRE(s && row_ptr && needed && (m == 0 || X), "sbp_nodes_near_host: NULL argument");
    SBP_REQUIRE(m >= 0 && idx_capacity >= 0, "sbp_nodes_near_host: bad sizes");
    sbp_ctx *ctx = s->ctx;
    SBP_DEVICE(ctx);
    Staging st(ctx);
    double *dX = (double *)st.dev(2, (size_t)(m ? m : 1) * s->d * 8);
    int64_t *dR = (int64_t *)st.dev(3, (size_t)(m + 1) * 8);
    int64_t cap = idx ? idx_capacity : 0;
    int64_t *dI = (int64_t *)st.dev(4, (size_t)(cap ? cap : 1) * 8);
    int64_t *dN = (int64_t *)(ctx->d_flags + 8);          // 8-byte aligned slot of the flag block
    st.h2d(dX, X, (size_t)m * s->d * 8);
    if (st.rc == SBP_OK)
        st.rc = sbp_nodes_near(s, dX, m, r, metric, closed, dR, cap ? dI : nullptr, cap, dN);
    st.d2h(row_ptr, dR, (size_t)(m + 1) * 8);
    st.d2h(ctx->h_flags + 8, dN, 8);
    st.sync();
    if (st.rc != SBP_OK) return st.rc;
    int64_t need = *(int64_t *)(ctx->h_flags + 8);
    *needed = need;
    if (idx && need <= cap && need > 0) {
        st.d2h(idx, dI, (size_t)need * 8);
        st.sync();
    }
    return st.rc;
}
