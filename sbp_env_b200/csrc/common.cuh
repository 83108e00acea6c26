// common.cuh -- shared host/device plumbing of libsbpgpu.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/sbp_gpu.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libsbpgpu is written for sm_100a (B200) only"
#endif

// ---------------------------------------------------------------- errors ----
void sbp_set_error(const char *fmt, ...);

#define SBP_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess) {                                                   \
            sbp_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,              \
                          cudaGetErrorString(e__));                                 \
            return SBP_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

#define SBP_REQUIRE(cond, msg)                                                      \
    do {                                                                            \
        if (!(cond)) {                                                              \
            sbp_set_error("%s:%d %s", __FILE__, __LINE__, msg);                     \
            return SBP_ERR_ARG;                                                     \
        }                                                                           \
    } while (0)

// Every entry point runs on its context's device and leaves the caller's current device as it
// found it (torch keeps its own notion of the current device).
struct SbpDeviceGuard {
    int prev = -1;
    bool changed = false;
    cudaError_t err;
    explicit SbpDeviceGuard(int dev) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) {
            err = cudaSetDevice(dev);
            changed = err == cudaSuccess;
        }
    }
    ~SbpDeviceGuard() { if (changed) cudaSetDevice(prev); }
};
#define SBP_DEVICE(ctxp)                                                             \
    SbpDeviceGuard sbp_dev_guard__((ctxp)->device);                                  \
    do {                                                                             \
        if (sbp_dev_guard__.err != cudaSuccess) {                                    \
            sbp_set_error("%s:%d cudaSetDevice(%d) -> %s", __FILE__, __LINE__,       \
                          (ctxp)->device, cudaGetErrorString(sbp_dev_guard__.err));  \
            return SBP_ERR_CUDA;                                                     \
        }                                                                            \
    } while (0)

// NVTX range around a C-ABI entry point (header-only NVTX v3: a no-op unless a profiler is
// attached), so that timelines show the library calls by name.
struct SbpNvtxRange {
    explicit SbpNvtxRange(const char *name) { nvtxRangePushA(name); }
    ~SbpNvtxRange() { nvtxRangePop(); }
};
#define SBP_NVTX(name) SbpNvtxRange sbp_nvtx_range__(name)

// ------------------------------------------------------------- handles -----
// scratch slots: 0,1 kNN partial lists / near work-space; 2-4 staging of the *_host entry points
// and the fused query; 5,6 cell grid and seeds; 7 append / read staging; 8,9 kNN filter lists;
// 10,11 filter shadow / done flags; 12,13 sbp_prm_connect2d (row poses, kNN rows); 14 cell-sorted
// poses; 15 row positions of sbp_graph_build_packed_rows; 16 candidate positions of the fused PRM kernels;
// 17 fp64 distances of a precision-32 kNN call; 18 per-block partials of the cooperative grid build
#define SBP_SCRATCH_SLOTS 20
// Tuning / instrumentation knobs (sbp_ctx_tune): per context, so that two contexts -- or two
// threads driving different contexts -- never see each other's settings.  None of them changes
// a result; the defaults are the measured optima.
struct sbp_tuning {
    int visible_group = 0;        // 0 = adaptive kernel; 4/8/16/32 = lane-group kernel of that width
    int coop_below = 12;          // adaptive kernel: cap of the hand-over threshold (walkers)
    int force_global = 0;         // 1 = never stage the map in shared memory
    int visible_tiles = 1;        // maps read from L2 / HBM: two-level walk on the 8x8-tile codes
    int tile_rect = 1;            // tile walks: settle an edge whose bounding box of tiles is all free from the run planes
    int prm_build_ctas = 0;       // k_prm_grid_build: resident CTAs per SM (0 = as many as fit, at most 3)
    int prm_fuse = 1;             // sbp_prm_connect2d: the kNN kernel settles edges with an all-free bounding box and writes the rows (0: separate walk kernel over all edges)
    int prm_mc_stage = 0;         // sbp_prm_connect2d: multicast outputs leave through a staging buffer + copy kernel
    int prm_walk_sorted = 1;      // sbp_prm_connect2d: walk the CTA's edges sorted by length (0: a row's edges in order)
    int prm_slices = 1;           // sbp_prm_connect2d: row slices whose walk overlaps the next slice's kNN
                                  // (measured on cfg5: 0.59 / 0.61 / 0.63 / 0.66 / 0.74 ms with 1 / 2 / 3 / 4 / 8 slices)
    int arm_group = 0;            // lanes per arm edge (0 = default 8)
    int arm_flat = 2;             // arm visible: flattened (edge, config) pairs per CTA; 2 = with the free-box test of the links (0: lane groups per edge)
    int near_filter = 1;          // filtered scan of the batched radius query
    int knn_seed = 1;             // grid-seeded a-priori bounds
    int knn_chunks = 0;           // node chunks of the batched kNN kernel (0 = auto)
    int knn_prefilter = 1;        // fp32 prefilter of the batched kNN scan
    int knn_filter = 1;           // seeded filter + refine pipeline
    int knn_filter_chunks = 0;    // node chunks of the filter kernel (0 = auto)
    int knn_filter_variant = 0;   // CTA shape / unrolling of the filter kernel
    int knn_cells = 2;            // exact answers from the cell grid (0 = off, 1 = warp per query,
                                  // 2 = thread per query when >= 20000 queries, 3 = always)
    int knn_cells_limit = 1024;   // ... where the rectangle holds at most this many nodes
    int knn_cells_qpw = 0;        // queries per warp of k_knn_cells_t (0 = 32)
    int knn_cells_rinit = 0;      // first ring radius of k_knn_cells_t (0 = from the density)
    int knn_cells_order = 1;      // cell-sorted query order when m == n
    int knn_cell_density = 150;   // nodes per grid cell x 100
    int time_kernels = 0;         // event pairs around one kernel family (sbp_ctx_kernel_time):
                                  // 1 = k_knn_filter, 2 = k_knn_cells*, 3 = 2-D visibility
};

struct sbp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0;
    int max_smem_optin = 0;
    int64_t launches = 0;
    sbp_tuning tune;
    // growable scratch for the *_host entry points and internal work-spaces
    void *d_scratch[SBP_SCRATCH_SLOTS] = {};
    size_t d_scratch_bytes[SBP_SCRATCH_SLOTS] = {};
    int32_t *d_flags = nullptr;      // 16 x int32
    int32_t *h_flags = nullptr;      // pinned mirror
    cudaEvent_t switch_event = nullptr;   // orders a new stream after the previous one
    cudaStream_t aux_stream = nullptr;    // second stream of sbp_prm_connect2d (walk of slice i beside kNN of i + 1)
    std::vector<cudaEvent_t> aux_events;  // fork / join events of that overlap
    int64_t scratch_generation = 0;  // bumped whenever a scratch slot is reallocated
    // bench instrumentation (sbp_tune("time_kernels", 1)): CUDA events around the dominant
    // kernel of the kNN path, read back with sbp_ctx_kernel_time
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timed;
};

struct KernelTimer {          // RAII: records an event pair on the launching stream
    sbp_ctx *ctx;
    cudaEvent_t a = nullptr, b = nullptr;
    // which: 1 = k_knn_filter, 2 = k_knn_cells / k_knn_cells_t, 3 = the 2-D visibility kernel
    // (sbp_ctx_tune(ctx, "time_kernels", which))
    explicit KernelTimer(sbp_ctx *c, int which = 1) : ctx(c->tune.time_kernels == which ? c : nullptr) {
        if (!ctx) return;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, ctx->stream);
    }
    ~KernelTimer() {
        if (!ctx) return;
        cudaEventRecord(b, ctx->stream);
        ctx->timed.emplace_back(a, b);
    }
};

int sbp_ctx_scratch(sbp_ctx *ctx, int slot, size_t bytes, void **out);

// sbp_prm_connect2d -> knn.cu: the queries are the stored rows themselves (X = their poses, row
// major, all n of them), of which this call answers the cell-sorted positions [p0, p0 + count);
// on return `ids` is the cell-sorted permutation of the node ids (NULL when no cell grid was
// built: the rows were then all answered, in index order).
struct PrmFusedArgs;
struct KnnRowsPart {
    int64_t p0 = 0, count = 0;
    int deterministic = 0;           // order the ids inside every cell (all ranks see the same ids)
    const int32_t *ids = nullptr;    // out
    PrmFusedArgs *fused = nullptr;   // in: answer the rows with prm.cu's fused kNN + visibility kernel
    int banded = 0;                  // in (fused): build the grid only for the band of cell rows the part needs
    int fused_ran = 0;               // out: that kernel was launched (ids are then sorted inside every cell)
    const uint8_t *done = nullptr;   // out (fused_ran): per node, 1 = row settled by the fused kernel
    const int32_t *pending = nullptr;  // out (fused_ran): device counter of the rows it left
};
#define GRID_HDR_INTS 20             // ints ahead of the cell counts: pending, band[2], spare, 16 slice counters (prm.cu)
#define PRM_FUSED_KMAX 21            // list lengths (k + 1) the fused kernel is instantiated for
// the cell grid as the fused kernel sees it (knn.cu builds it, prm.cu walks it)
struct CellGridView {
    const double *b4;                // bounding box (min x, max x, min y, max y)
    int G;                           // cells per side
    const int32_t *start;            // [G * G + 1] first position of every cell (row major)
    const int32_t *ids;              // [n] node ids in cell order, ascending inside a cell
    const double2 *pts;              // [n] their poses
    int64_t n;
    const double *maxabs;            // device scalar: max |coordinate| in the store (inf = non-finite)
    const int32_t *band;             // device [2]: first / last row of cells that ids / pts cover (NULL: all)
};
int32_t sbp_prm_fused_launch(sbp_ctx *ctx, const CellGridView &grid, int K, int k, int r_init, int limit,
                             int64_t p0, int64_t count, double *seed2, uint8_t *done, int32_t *pending,
                             PrmFusedArgs *args);
int32_t sbp_knn_rows_part(sbp_nodes *s, const double *X_all, int32_t k, int64_t *idx, KnnRowsPart *part);
// ... answers the rows from the cell grid (and, for lists of up to PRM_FUSED_KMAX entries, with
// prm.cu's fused kernels) when this holds
bool sbp_knn_rows_use_grid(const sbp_nodes *s);

// Device view of the bit-packed occupancy map.
//   pixel (x, y), 0 <= x < W, 0 <= y < H  (the reference's `_img[x, y]`)
//   sector  = (y >> 4) * SX + (x >> 4)               16x16 px = 32 B
//   word32  = sector * 8 + ((y >> 3) & 1) * 4 + ((x >> 3) & 1) * 2 + ((y >> 2) & 1)
//   bit     = ((y & 3) << 3) | (x & 7)               1 = free
// i.e. each 64-bit pair of words is one 8x8-pixel tile, 4 tiles per sector.
struct MapView {
    const uint32_t *words;
    int32_t W, H;
    int32_t SX, SY;       // sectors per row / column
    uint32_t n_words;     // 32-bit words (multiple of 8)
};

struct TileCodeView {
    const uint32_t *words;    // two planes of AW x AH / AH x AW words (map.cu:k_tile_codes)
    int32_t AW, AH;
    // free-run planes (map.cu:k_tile_runs, lookup: device_fns.cuh:tiles_rect_free): plane j in 0..3,
    // word [j][k][u] bit b = "the 2^j tiles (u .. u + 2^j - 1, 16 k + b) are all free"; windows of 32
    // tile rows every 16 (any span of <= 16 rows lies inside one word); RK windows x RU tile columns.
    const uint32_t *runs;
    int32_t RU, RK;
};

struct sbp_map {
    sbp_ctx *ctx = nullptr;
    MapView view{};
    uint32_t *d_words = nullptr;
    std::vector<uint32_t> h_words;   // host mirror (arm guard-band resolution)
    int64_t packed_bytes = 0;
    bool smem_resident = false;      // whole map fits one CTA's shared memory
    // Second, row-major copy for the arm link sweeps (built on first use, sbp_map_ensure_rows):
    // word = y * WR + (x >> 5), bit = x & 31 -- two shifts and one multiply-add per pixel where
    // the tiled layout above needs ~12 integer operations (the arm sweep is bound by exactly
    // that arithmetic); `rows.SX` holds WR, the words per row.
    uint32_t *d_rows = nullptr;
    MapView rows{};
    bool rows_smem_resident = false;
    // Coarse level for edge walks on big maps (built on first use, sbp_map_ensure_allfree): a 2-bit
    // code per 8x8-pixel tile, 1 = all 64 pixels free, 2 = all blocked, 0 = mixed; a word = 4 x 4
    // tiles, words row major, in two planes (x-major and, transposed, y-major walks; layout:
    // map.cu:k_tile_codes, lookup: device_fns.cuh:TileCodes; af.AW x af.AH words per plane).  1/16
    // of the map's size (8 MiB for 32k^2): it stays in L1 / L2 where the map itself does not.
    uint32_t *d_allfree = nullptr;
    uint32_t *d_runs = nullptr;
    TileCodeView af{nullptr, 0, 0, nullptr, 0, 0};
};

// the tile-code view a kernel is handed (knob tile_rect = 0: without the free-run planes)
static inline TileCodeView sbp_map_codes(const sbp_map *m) {
    TileCodeView v = m->af;
    if (!m->ctx->tune.tile_rect) v.runs = nullptr;
    return v;
}

int32_t sbp_map_ensure_rows(sbp_map *map);
int32_t sbp_map_ensure_allfree(sbp_map *map);

struct sbp_nodes {
    sbp_ctx *ctx = nullptr;
    int d = 0;
    int64_t n = 0;
    int64_t capacity = 0;
    double *d_soa = nullptr;         // [d][capacity] fp64, structure of arrays
    float *d_soa32 = nullptr;        // fp32 shadow (kNN prefilter only; never decides a result)
    double *d_maxabs = nullptr;      // device scalar: max |coordinate| ever appended (inf if non-finite)
};

int32_t nodes_reserve(sbp_nodes *s, int64_t need);      // capacity >= need (nodes.cu)

// --------------------------------------------------------- device helpers ---
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t map_word_index(int SX, int x, int y) {
    uint32_t sec = (uint32_t)(y >> 4) * (uint32_t)SX + (uint32_t)(x >> 4);
    return sec * 8u + (uint32_t)(((y >> 3) & 1) * 4 + ((x >> 3) & 1) * 2 + ((y >> 2) & 1));
}

__device__ __forceinline__ uint32_t map_bit_index(int x, int y) {
    return (uint32_t)(((y & 3) << 3) | (x & 7));
}

// `words` may point to shared or global memory (generic address).
template <bool GLOBAL_LDG>
__device__ __forceinline__ bool map_free(const uint32_t *words, int SX, int x, int y) {
    uint32_t wi = map_word_index(SX, x, y);
    uint32_t w = GLOBAL_LDG ? __ldg(words + wi) : words[wi];
    return (w >> map_bit_index(x, y)) & 1u;
}

template <bool GLOBAL_LDG>
__device__ __forceinline__ bool map_free_rows(const uint32_t *words, int WR, int x, int y) {
    uint32_t wi = (uint32_t)y * (uint32_t)WR + ((uint32_t)x >> 5);
    uint32_t w = GLOBAL_LDG ? __ldg(words + wi) : words[wi];
    return (w >> ((uint32_t)x & 31u)) & 1u;
}

// numpy index semantics of `_img[x, y]`: valid iff -W <= i < W,
// negative indices wrap once.  Caller guarantees the range; this only wraps.
__device__ __forceinline__ int wrap_index(int i, int n) { return i < 0 ? i + n : i; }

// CPython int(float) on the device: truncation toward zero, with the
// non-finite classes reported.  `v` outside (-(n+1), n) is out of the index
// range of an axis of length n whatever its exact value, so the saturating
// conversion is only used for in-range values.
struct TruncResult {
    int i;          // truncated value (valid only when in_range)
    bool in_range;  // -n <= trunc(v) < n
    int flag;       // 0, SBP_FLAG_NAN, SBP_FLAG_OVERFLOW
};

__device__ __forceinline__ TruncResult trunc_index(double v, int n) {
    TruncResult r;
    r.flag = 0;
    if (v != v) r.flag = SBP_FLAG_NAN;
    else if (isinf(v)) r.flag = SBP_FLAG_OVERFLOW;
    r.in_range = (v > -(double)(n + 1)) && (v < (double)n);  // false for NaN
    r.i = r.in_range ? __double2int_rz(v) : 0;
    return r;
}

// ---- mbarrier + 1-D bulk async copy (TMA engine; SASS: UBLKCP) ------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// Stage the whole packed map into shared memory with the bulk-copy engine.
// Called by all threads of the CTA; returns after the bytes have landed.
__device__ __forceinline__ void stage_map_to_smem(uint32_t *smap, const MapView &mv,
                                                  uint64_t *bar) {
    const uint32_t total = mv.n_words * 4u;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, total);
        // one bulk copy moves at most 2^20-16 bytes; chunk to be safe
        const uint32_t CH = 32768u;
        for (uint32_t off = 0; off < total; off += CH) {
            uint32_t b = total - off < CH ? total - off : CH;
            bulk_g2s((char *)smap + off, (const char *)mv.words + off, b, bar);
        }
    }
    mbar_wait(bar, 0);
}

#endif  // __CUDACC__
