/*
 * sbp_gpu.h -- C ABI of libsbpgpu.so, the B200 (sm_100a) implementation of the
 * collision-checking / nearest-neighbour hot path of soraxas/sbp-env.
 *
 * Every entry point replaces one reference interface; the citation after
 * "replaces:" is file:line under /root/reference/sbp_env/.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types.
 *  - every function returns int32 status: 0 = SBP_OK, negative = error;
 *    sbp_last_error() returns a thread-local description of the last failure.
 *  - pointers marked "device" are raw CUDA addresses in the context's device
 *    (e.g. torch.Tensor.data_ptr()); the caller owns every buffer it passes.
 *  - device entry points are asynchronous on the context's stream and never
 *    synchronise; the *_host entry points copy in, launch, copy out and
 *    synchronise the stream before returning.
 *  - the library owns handles until the matching *_destroy.
 *  - there is no CPU implementation behind any of these calls: without a CUDA
 *    device sbp_ctx_create fails and nothing else can be called.
 */
#ifndef SBP_GPU_H
#define SBP_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SBP_ABI_VERSION 2

/* status codes */
#define SBP_OK 0
#define SBP_ERR_CUDA (-1)      /* a CUDA runtime call failed */
#define SBP_ERR_ARG (-2)       /* invalid argument */
#define SBP_ERR_NODEV (-3)     /* no usable CUDA device */
#define SBP_ERR_CAPACITY (-4)  /* caller buffer / node store too small */

/* bits OR-ed into the `flags` word by the collision kernels: what CPython would
 * raise for that batch (collisionChecker.py:151, :173-174, :400) */
#define SBP_FLAG_NAN 1       /* int(nan)           -> ValueError   */
#define SBP_FLAG_OVERFLOW 2  /* int(+-inf), or an index in [2^63,2^64) -> OverflowError */
#define SBP_FLAG_PEER_TIMEOUT 1024 /* sbp_knn_edges_visible2d rendezvous: a peer never signalled */

/* values written to the `out` bytes of the arm entry points */
#define SBP_FREE 1
#define SBP_BLOCKED 0
#define SBP_AMBIGUOUS 2 /* a forward-kinematics point lies within the fp64 guard band of a
                           pixel boundary: resolve with sbp_arm_resolve_host */

/* near(): which distance and which comparison */
#define SBP_METRIC_FULL 0 /* np.linalg.norm over all d dims (prmPlanner.py:39, rrtPlanner.py:278) */
#define SBP_METRIC_XY 1   /* first two dims only, engine.euclidean_dist (engine.py:13-24)        */
#define SBP_CMP_LT 0      /* d <  r  (prmPlanner.py:42, birrtPlanner.py:85, rrdtPlanner.py:713)  */
#define SBP_CMP_LE 1      /* d <= r  (rrtPlanner.py:179, :238, :259)                             */

typedef struct sbp_ctx sbp_ctx;
typedef struct sbp_map sbp_map;
typedef struct sbp_nodes sbp_nodes;
typedef struct sbp_graph sbp_graph;

int32_t sbp_abi_version(void);
const char *sbp_last_error(void);

/* ---- context ------------------------------------------------------------- */
/* stream_or_null: a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream);
 * NULL creates a private non-blocking stream owned by the context. */
int32_t sbp_ctx_create(int32_t device, void *stream_or_null, sbp_ctx **out);
int32_t sbp_ctx_destroy(sbp_ctx *ctx);
int32_t sbp_ctx_set_stream(sbp_ctx *ctx, void *stream);
int32_t sbp_ctx_sync(sbp_ctx *ctx);
/* number of SMs of the context's device and kernels launched so far */
int32_t sbp_ctx_info(sbp_ctx *ctx, int32_t *sm_count, int64_t *launches);
/* Counter that changes whenever one of the context's internal work buffers is reallocated
 * (a later call needed more room).  A CUDA graph captured from calls of this library froze the
 * old addresses: compare the value taken at capture time before every replay and recapture
 * when it moved (planning.CapturedStep does). */
int32_t sbp_ctx_scratch_generation(sbp_ctx *ctx, int64_t *generation);

/* ---- instrumentation (never changes a result) ----------------------------------
 * Tuning knobs, PER CONTEXT: key is one of visible_group, coop_below, force_global, arm_group,
 * near_filter, knn_seed, knn_chunks, knn_prefilter, knn_filter, knn_filter_chunks,
 * knn_filter_variant, knn_cells, knn_cells_limit, knn_cells_qpw, knn_cells_rinit,
 * knn_cells_order, knn_cell_density, visible_tiles, tile_rect, prm_fuse, prm_walk_sorted,
 * prm_slices, prm_mc_stage, arm_flat, time_kernels (csrc/common.cuh: sbp_tuning). */
int32_t sbp_ctx_tune(sbp_ctx *ctx, const char *key, int64_t value);
/* With time_kernels = 1 (k_knn_filter), 2 (k_knn_cells*) or 3 (2-D visibility kernel) every
 * launch of that kernel is bracketed by a CUDA event pair on the launching stream; this returns
 * the summed milliseconds and the number of pairs since the last call (synchronises). */
int32_t sbp_ctx_kernel_time(sbp_ctx *ctx, double *total_ms, int64_t *launches);
/* Measured denominators for the rooflines (csrc/microbench.cu): which = 0 shared-memory word
 * gathers, 1 random 32-byte sector gathers over `bytes` of global memory, 2 unfused fp64
 * mul/add, 3 packed fp32 FMA, 5-7 min/compare mixes.  *ms = time of `iters` iterations,
 * *units = operations performed. */
int32_t sbp_microbench(sbp_ctx *ctx, int32_t which, int64_t bytes, int32_t iters, double *ms,
                       double *units);

/* ---- occupancy map -------------------------------------------------------
 * replaces: ImgCollisionChecker.__init__ / RobotArm4dCollisionChecker.__init__
 * storage of `_img` (collisionChecker.py:101-108, :319-340).
 * free_xy: HOST bytes, the checker's `_img` laid out [W][H] row-major, non-zero
 * byte = free (`_img[x, y] == 1`).  The map is bit-packed on the device into
 * 8x8-pixel 64-bit tiles, 2x2 tiles (16x16 px) per 32-byte sector. */
int32_t sbp_map_create(sbp_ctx *ctx, const uint8_t *free_xy, int32_t W, int32_t H,
                       sbp_map **out);
/* same, from a DEVICE byte grid (for procedurally generated 8k-32k maps) */
int32_t sbp_map_create_device(sbp_ctx *ctx, const uint8_t *free_xy_device, int32_t W,
                              int32_t H, sbp_map **out);
int32_t sbp_map_destroy(sbp_map *map);
int32_t sbp_map_info(sbp_map *map, int32_t *W, int32_t *H, int64_t *packed_bytes,
                     int32_t *smem_resident);

/* ---- 2-D image-space checker ---------------------------------------------- */
/* replaces: ImgCollisionChecker.feasible (collisionChecker.py:147-153),
 * batched.  P: device [n][2] fp64.  out: device n bytes (1 free / 0 blocked).
 * flags: device int32 (nullable), SBP_FLAG_* bits are OR-ed in. */
int32_t sbp_feasible2d(sbp_map *map, const double *P, int64_t n, uint8_t *out,
                       int32_t *flags);
/* replaces: ImgCollisionChecker.visible + get_line (collisionChecker.py:134-145,
 * :155-215), batched: integer Bresenham between int()-truncated endpoints, all
 * pixels free.  An edge with a NaN endpoint yields 0 (ValueError is caught at
 * :143-144) and sets SBP_FLAG_NAN; +-inf sets SBP_FLAG_OVERFLOW.
 * px_tested: device uint64 (nullable) += pixel tests actually executed. */
int32_t sbp_visible2d(sbp_map *map, const double *P, const double *Q, int64_t n,
                      uint8_t *out, int32_t *flags, uint64_t *px_tested);
/* replaces: ImgCollisionChecker.get_coor_before_collision (collisionChecker.py:118-132),
 * batched: out_xy device [n][2] int32 = first blocked pixel walking start->end,
 * or the end pixel when the line is free. */
int32_t sbp_first_blocked2d(sbp_map *map, const double *P, const double *Q, int64_t n,
                            int32_t *out_xy, int32_t *flags);

/* ---- 4-DoF stick arm checker ---------------------------------------------- */
/* replaces: RobotArm4dCollisionChecker.feasible + get_pt_from_angle_and_length +
 * _pt_feasible (collisionChecker.py:393-441), batched.  C: device [n][4] fp64
 * (x, y, r0, r1).  out: SBP_FREE / SBP_BLOCKED / SBP_AMBIGUOUS per config. */
int32_t sbp_arm_feasible(sbp_map *map, double L0, double L1, const double *C, int64_t n,
                         uint8_t *out, int32_t *flags);
/* replaces: RobotArm4dCollisionChecker.visible + _interpolate_configs +
 * create_ranges (collisionChecker.py:346-391), batched.  C1, C2: device [n][4].
 * cfg_tested: device uint64 (nullable) += interpolated configs evaluated. */
int32_t sbp_arm_visible(sbp_map *map, double L0, double L1, const double *C1,
                        const double *C2, int64_t n, uint8_t *out, int32_t *flags,
                        uint64_t *cfg_tested);
/* Exact resolution of SBP_AMBIGUOUS entries with the host libm cos/sin (the same
 * libm CPython's math.cos/sin call, collisionChecker.py:437-440).  All pointers
 * HOST.  C2 == NULL resolves configs (feasible), otherwise edges (visible).
 * Only entries equal to SBP_AMBIGUOUS are touched.  *resolved (nullable) counts them. */
int32_t sbp_arm_resolve_host(sbp_map *map, double L0, double L1, const double *C1,
                             const double *C2, int64_t n, uint8_t *inout,
                             int64_t *resolved);

/* ---- host-buffer forms (what the Python plugin calls for user arrays) ------
 * Same semantics as above with HOST pointers; H2D copy, kernel(s), D2H copy and a
 * stream synchronise happen inside.  *flags receives the SBP_FLAG_* bits.
 * The arm forms also run sbp_arm_resolve_host, so out is 0/1 only. */
int32_t sbp_feasible2d_host(sbp_map *map, const double *P, int64_t n, uint8_t *out,
                            int32_t *flags);
int32_t sbp_visible2d_host(sbp_map *map, const double *P, const double *Q, int64_t n,
                           uint8_t *out, int32_t *flags);
int32_t sbp_first_blocked2d_host(sbp_map *map, const double *P, const double *Q, int64_t n,
                                 int32_t *out_xy, int32_t *flags);
int32_t sbp_arm_feasible_host(sbp_map *map, double L0, double L1, const double *C,
                              int64_t n, uint8_t *out, int32_t *flags, int64_t *n_ambiguous);
int32_t sbp_arm_visible_host(sbp_map *map, double L0, double L1, const double *C1,
                             const double *C2, int64_t n, uint8_t *out, int32_t *flags,
                             int64_t *n_ambiguous);

/* ---- node store: device-resident, append-only ------------------------------
 * replaces: the planners' pre-allocated `poses` arrays and their appends
 * (rrtPlanner.py:24-26, :106-113; birrtPlanner.py:24-31; rrdtPlanner.py:858-887). */
/* d: 2, 3, 4 or 6 (the dimensions the query kernels are instantiated for). */
int32_t sbp_nodes_create(sbp_ctx *ctx, int32_t d, int64_t capacity, sbp_nodes **out);
int32_t sbp_nodes_destroy(sbp_nodes *nodes);
/* X: [m][d] fp64, row-major; on_device selects the address space of X.
 * Grows the store (capacity doubling) when needed. */
int32_t sbp_nodes_append(sbp_nodes *nodes, const double *X, int64_t m, int32_t on_device);
int32_t sbp_nodes_size(sbp_nodes *nodes, int64_t *n);
/* keep only the first n nodes (n <= size) */
int32_t sbp_nodes_truncate(sbp_nodes *nodes, int64_t n);
/* copy rows [first, first+count) back to HOST as [count][d] */
int32_t sbp_nodes_read(sbp_nodes *nodes, int64_t first, int64_t count, double *out_host);
/* same into a DEVICE buffer [count][d] (asynchronous; e.g. to query the store with
 * its own rows, PRMPlanner.build_graph's outer loop, prmPlanner.py:91-92) */
int32_t sbp_nodes_rows_device(sbp_nodes *nodes, int64_t first, int64_t count, double *out_device);

/* replaces: RRTPlanner.find_nearest_neighbour_idx (rrtPlanner.py:268-279) for
 * k = 1 and the np.argsort(distances)[:k] candidate lists (rrtPlanner.py:151-152,
 * :226-227) for k > 1: the k nearest of ALL nodes, d = sqrt_rn of the
 * left-to-right unfused sum of squares over all dims, ordered by (d, index).
 * (How: batches of >= 256 queries over >= 2048 nodes are answered exactly from a cell
 * grid built inside the call, with the brute-force scan behind it for the queries the
 * grid cannot settle; fewer queries / nodes are scanned by brute force.  Same rows
 * either way -- DESIGN.md 3a.)
 * X: device [m][dim].  idx: device [m][k] int64, dist: device [m][k] fp64
 * (nullable); rows are padded with -1 / +inf when size < k.  1 <= k <= 32.
 * precision: 64 = fp64 distances, bit-exact.  32 = `dist` is a FLOAT buffer [m][k] (cast the
 * pointer): the same rows -- the selection stays the exact fp64 one, so the index sets are
 * bit-exact also at ties -- with the distances rounded to fp32, |d32 - d64| <= 2^-24 d64
 * (BASELINE.json north_star: "within 1e-6 relative in fp32"); half the output bytes. */
int32_t sbp_nodes_knn(sbp_nodes *nodes, const double *X, int64_t m, int32_t k,
                      int32_t precision, int64_t *idx, double *dist);
/* replaces: RRTPlanner.find_nearest_neighbour_idx (rrtPlanner.py:268-279) EXACTLY,
 * np.argmin(np.linalg.norm(poses - pos, axis=1)): the first minimum -- and, like np.argmin, the
 * FIRST node whose distance is NaN as soon as there is one (a NaN query, NaN / inf - inf node
 * coordinates), where the argsort rule of sbp_nodes_knn would rank NaN distances last.
 * X: device [m][dim]; idx device [m] (-1 for an empty store); dist device [m] (nullable). */
int32_t sbp_nodes_nearest(sbp_nodes *nodes, const double *X, int64_t m, int64_t *idx, double *dist);
int32_t sbp_nodes_nearest_host(sbp_nodes *nodes, const double *X, int64_t m, int64_t *idx,
                               double *dist);

/* replaces: prmPlanner.nearest_neighbours (prmPlanner.py:28-44; FULL, LT), the
 * radius gates of choose_least_cost_parent / rewire (rrtPlanner.py:177-181,
 * :236-239; XY, LE), birrtPlanner.py:82-85, rrdtPlanner.py:704-713.
 * CSR output, node indices ascending within each row.  row_ptr: device [m+1]
 * int64.  idx: device [idx_capacity] int64 (nullable -> count only).
 * *needed: device int64 = total number of neighbours.  When needed >
 * idx_capacity only row_ptr / needed are valid. */
int32_t sbp_nodes_near(sbp_nodes *nodes, const double *X, int64_t m, double r,
                       int32_t metric, int32_t closed, int64_t *row_ptr, int64_t *idx,
                       int64_t idx_capacity, int64_t *needed);

/* host-buffer forms of the queries */
int32_t sbp_nodes_knn_host(sbp_nodes *nodes, const double *X, int64_t m, int32_t k,
                           int32_t precision, int64_t *idx, double *dist);
/* two-call protocol: idx == NULL returns row_ptr and *needed only */
int32_t sbp_nodes_near_host(sbp_nodes *nodes, const double *X, int64_t m, double r,
                            int32_t metric, int32_t closed, int64_t *row_ptr, int64_t *idx,
                            int64_t idx_capacity, int64_t *needed);

/* ---- batched planner primitives ------------------------------------------- */
/* Gather the endpoints of candidate edges (i -> nbr[i][j]) from the node store
 * into P, Q ([m*k][d] device) for a following sbp_visible2d / sbp_arm_visible:
 * the batched form of PRMPlanner.build_graph's inner loop (prmPlanner.py:91-101).
 * Entries with nbr < 0 or nbr == self produce valid = 0 and a degenerate edge. */
int32_t sbp_nodes_gather_edges(sbp_nodes *nodes, const int64_t *nbr, int64_t m, int32_t k,
                               int64_t first_row, double *P, double *Q, uint8_t *valid);
/* kNN rows -> candidate-edge batch in one pass (all pointers device).
 *  X == NULL: row i is stored node first_row+i; its own entry (`m_g is v`,
 *    prmPlanner.py:94-95) is dropped from nbr_in [m][k_in] -- the column equal to the
 *    row index if present, else the last column -- leaving k_out = k_in-1 columns.
 *  X != NULL: row i is the external query X[i] ([m][d]); k_out = k_in columns
 *    (connecting new samples / start / goal to the roadmap, prmPlanner.py:131-138).
 * Writes nbr_out [m][k_out], P = neighbour pose, Q = row pose ([m*k_out][d]) and
 * valid [m*k_out] (0 where the neighbour slot is empty). */
int32_t sbp_nodes_knn_edges(sbp_nodes *nodes, const int64_t *nbr_in, int64_t m, int32_t k_in,
                            int64_t first_row, const double *X, int64_t *nbr_out, double *P,
                            double *Q, uint8_t *valid);

/* sbp_nodes_knn_edges + sbp_visible2d in ONE launch for the 2-D image checker: the candidate
 * edges of a kNN answer are resolved against the node store inside the visibility kernel, no
 * endpoint arrays are written.
 * replaces: the body of PRMPlanner.build_graph's double loop (prmPlanner.py:91-101) --
 * `if m_g is v: continue` and `cc.visible(m_g.pos, v.pos)` -- for a whole block of rows.
 * Row / column / X meaning as in sbp_nodes_knn_edges.  Writes nbr_out [m][k_out] and
 * vis [m][k_out]: 1 = the neighbour slot is filled and ImgCollisionChecker.visible is True.
 * packed (optional, [m][k_out] int32): the adjacency code (neighbour + 1) * 2 + visible of every
 * edge, written as the edge is decided -- to a local buffer, or with packed_multicast != 0 through
 * the NVSwitch multicast address of the ranks' symmetric buffer (multimem.st): the multi-GPU
 * assembly of the PRM adjacency (BASELINE.json north_star) then overlaps the visibility walk and
 * only a barrier follows (distributed.PeerAdjacencyGather.target / complete) -- or not even that:
 * rendezvous (optional HOST array of 6 int64: multicast and local address of a symmetric int32
 * [world] signal array, addresses of two local int32 counters (CTA ticket, epoch), rank, world)
 * makes the last CTA publish this rank's epoch through the switch and wait for every rank's, so
 * the kernel returns when the adjacency is complete on this rank (bounded spin: a missing peer
 * sets SBP_FLAG_PEER_TIMEOUT in *flags instead of hanging). */
int32_t sbp_knn_edges_visible2d(sbp_map *map, sbp_nodes *nodes, const int64_t *nbr_in, int64_t m,
                                int32_t k_in, int64_t first_row, const double *X,
                                int64_t *nbr_out, uint8_t *vis, int32_t *packed,
                                int32_t packed_multicast, const int64_t *rendezvous, int32_t *flags);

/* The whole roadmap connection step in ONE call, for the 2-D image checker and the rows of the
 * node store themselves.
 * replaces: PRMPlanner.build_graph's double loop (prmPlanner.py:91-101) with the k nearest
 * neighbours as candidate set: for every stored node v its k nearest OTHER nodes m_g (exact, fp64,
 * ordered by (distance, index); the row's own entry dropped like `if m_g is v: continue`) and
 * cc.visible(m_g.pos, v.pos) for each.
 * Rows are processed in the cell-sorted order of the kNN's grid and split into `parts` contiguous
 * runs of that order (spatially compact shards: one per rank; parts = 1: all rows); this call
 * handles run `part`.  Outputs (each nullable, at least one given) are [n][k] in NODE order and
 * only this part's rows are written: nbr_out int64, vis uint8 (1 = slot filled and edge free),
 * packed int32 = (neighbour + 1) * 2 + visible.  packed may be pinned host memory, or with
 * packed_multicast != 0 the NVSwitch multicast address of the ranks' symmetric [n][k] buffer:
 * every rank's rows then land in every rank's copy as the edges are decided. */
int32_t sbp_prm_connect2d(sbp_map *map, sbp_nodes *nodes, int32_t k, int32_t part, int32_t parts,
                          int64_t *nbr_out, uint8_t *vis, int32_t *packed, int32_t packed_multicast,
                          int32_t *flags);
/* The same with the rows of every output in CELL-SORTED order: row p of nbr_out / vis / packed belongs
 * to node order_out[p] (device int32 [n], written by the call: the permutation of the kNN's counting
 * sort, ascending node index inside a cell -- a function of the poses only, so every rank computes
 * the same one).  The rows of part i are then ONE contiguous block [p0, p0 + count) x k that the
 * kernel stores with coalesced 128-byte lines: the form to use when `packed` is pinned host memory
 * or the NVSwitch multicast mapping.  sbp_graph_build_packed_rows consumes it.
 * parts > 1: a rank builds the cell grid for the band of cell rows its part needs only (the part
 * of the build that shrinks with the number of ranks), so only order_out[p0, p0 + count) is written,
 * like the rows; with packed_multicast != 0 order_out is the multicast address of the ranks'
 * symmetric order buffer as well, and every rank ends up with the whole permutation. */
int32_t sbp_prm_connect2d_sorted(sbp_map *map, sbp_nodes *nodes, int32_t k, int32_t part, int32_t parts,
                                 int64_t *nbr_out, uint8_t *vis, int32_t *packed,
                                 int32_t packed_multicast, int32_t *order_out, int32_t *flags);

/* Fused radius query + edge visibility in ONE launch, for the sequential planners.
 * replaces: the O(N) Python loops of choose_least_cost_parent (rrtPlanner.py:173-196:
 * engine.dist(new,p) <= radius and cc.visible(new.pos, p.pos)), rewire
 * (rrtPlanner.py:250-266: cc.visible(n.pos, new.pos)), PRMPlanner.get_nearest_free
 * (prmPlanner.py:103-126) and the Bi-RRT join test (birrtPlanner.py:82-87).
 * X: device [m][d] queries.  For every stored node within r of X[i] (metric / closed as
 * in sbp_nodes_near) the edge is checked on the spot: direction 0 = visible(X[i] -> node),
 * 1 = visible(node -> X[i]); arm = 0 uses the 2-D image checker on the first two
 * coordinates, arm = 1 (d == 4) the stick-arm checker with link lengths L0, L1.
 * out_idx / out_vis: device [m][cap], hits of query i in ARRIVAL order (unordered);
 * out_vis holds SBP_FREE / SBP_BLOCKED / SBP_AMBIGUOUS.  count: device int32 [m], total
 * hits per query (may exceed cap: then only cap were stored). */
int32_t sbp_nodes_near_visible(sbp_nodes *nodes, sbp_map *map, int32_t arm, double L0, double L1,
                               const double *X, int64_t m, double r, int32_t metric,
                               int32_t closed, int32_t direction, int64_t *out_idx,
                               uint8_t *out_vis, int32_t cap, int32_t *count, int32_t *flags);
/* Host form for one query x (HOST [d]); hits come back sorted by node index (the order of
 * the reference's loops) with guard-band entries already resolved.  *n_hits > cap returns
 * SBP_ERR_CAPACITY: call again with a larger buffer. */
int32_t sbp_nodes_near_visible_host(sbp_nodes *nodes, sbp_map *map, int32_t arm, double L0,
                                    double L1, const double *x, double r, int32_t metric,
                                    int32_t closed, int32_t direction, int64_t *out_idx,
                                    uint8_t *out_vis, int32_t cap, int32_t *n_hits,
                                    int32_t *flags);

/* replaces: RRTPlanner.find_nearest_neighbour_idx (rrtPlanner.py:268-279) for ONE query x (HOST [d]) in one
 * launch: np.argmin(np.linalg.norm(poses - x, axis=1)) -- first index of the smallest rounded root; the
 * first NaN distance wins when there is one (dist then NaN).  idx -1 for an empty store.  Synchronises. */
int32_t sbp_nodes_nearest1_host(sbp_nodes *nodes, const double *x, int64_t *idx, double *dist);

/* ---- device-resident RRT* tree: cost / parent arrays and the inner loop of one accepted sample
 * in ONE launch.
 * replaces: RRTPlanner.choose_least_cost_parent, linear path (rrtPlanner.py:173-196), add_newnode
 * (:106-113) and the linear rewire (:250-266) for the 2-D image checker: radius scan of the tree
 * (engine.dist(new, p) <= radius), visible(new.pos, p.pos) per candidate, arg-min of p.cost + c in the
 * reference's order (nearest node first, ascending index, strict <), append of the new pose with its
 * cost / parent, re-parenting of every candidate with new.cost + c < n.cost -- on device arrays.
 * A comparison whose sides are within 64 ulp (the reference's libm-pow distance may differ from
 * the device's correctly rounded one in the last bit) or a candidate within 4 ulp of the radius is
 * not decided: status SBP_EXTEND_AMBIGUOUS, nothing written, the caller settles the sample with the
 * host path and sbp_tree_set.  `fix_*`: exact costs the caller derived for nodes of earlier calls
 * (engine.dist on the host), applied before any cost is read. */
typedef struct sbp_tree sbp_tree;
#define SBP_EXTEND_OK 0
#define SBP_EXTEND_AMBIGUOUS 1
#define SBP_EXTEND_NO_PARENT 2      /* nn == -1 and no visible node inside the radius (LookupError) */
#define SBP_EXTEND_MAX_FIXES 24
#define SBP_EXTEND_MAX_REWIRED 64
typedef struct sbp_extend_result {
    int32_t status;
    int32_t n_candidates;           /* nodes inside the radius = visible() calls of the reference's loop */
    int32_t n_rewired;
    int32_t n_rewire_checked;       /* candidates the rewire pass looked at (not at the new node's / its parent's position) */
    int64_t new_idx;                /* index of the appended node, -1 when nothing was appended */
    int64_t parent;
    double cost;                    /* cost[parent] + c_parent in device arithmetic */
    double c_parent;
    int32_t rewired[SBP_EXTEND_MAX_REWIRED];     /* unordered */
    double rewired_c[SBP_EXTEND_MAX_REWIRED];    /* their distance to the new node (device arithmetic) */
} sbp_extend_result;
int32_t sbp_tree_create(sbp_nodes *nodes, sbp_tree **out);
int32_t sbp_tree_destroy(sbp_tree *tree);
/* cost (and parent, unless parent < -1) of one node; stream ordered */
int32_t sbp_tree_set(sbp_tree *tree, int64_t idx, double cost, int64_t parent);
/* bulk upload of HOST cost / parent arrays into [first, first + count) (an existing tree); synchronises */
int32_t sbp_tree_write(sbp_tree *tree, int64_t first, int64_t count, const double *cost_host,
                       const int32_t *parent_host);
/* HOST copies of cost / parent [first, first + count); synchronises */
int32_t sbp_tree_read(sbp_tree *tree, int64_t first, int64_t count, double *cost_host, int32_t *parent_host);
/* newpos: HOST [2]; nn: current nearest node or -1; do_append / do_rewire: 0 / 1; synchronises */
int32_t sbp_tree_extend_host(sbp_tree *tree, sbp_map *map, const double *newpos, int64_t nn, double radius,
                             int32_t do_append, int32_t do_rewire, const int64_t *fix_idx,
                             const double *fix_cost, int32_t n_fixes, sbp_extend_result *out);


/* ---- roadmap graph ------------------------------- */
/* The PRM roadmap as a device-resident CSR graph.
 * replaces: the networkx DiGraph PRMPlanner.build_graph fills edge by edge
 * (prmPlanner.py:91-101, `graph.add_weighted_edges_from([(m_g, v, dist)])`) and the
 * nx.has_path / nx.shortest_path query of PRMPlanner.get_solution (prmPlanner.py:128-154;
 * unweighted: fewest hops).  Edge weights are engine.dist of the end poses and are
 * recomputed from the node store where a cost is needed. */
int32_t sbp_graph_create(sbp_ctx *ctx, int64_t n_nodes, sbp_graph **out);
int32_t sbp_graph_destroy(sbp_graph *g);
/* (Re)build from a candidate-edge block of sbp_nodes_knn_edges + the visibility bytes of
 * sbp_visible2d / sbp_arm_visible (device pointers, [m][k]): for row i and column j with
 * vis != 0 and nbr >= 0 the arc nbr[i][j] -> first_row + i (m_g -> v).
 * undirected != 0 also adds first_row + i -> nbr[i][j]: the reference's radius rule is symmetric
 * (v in near(u) <=> u in near(v)), so its DiGraph holds both (m_g -> v) and (v -> m_g)
 * (prmPlanner.py:91-101), which a k-NN candidate set does not give by itself.  A mirrored arc
 * that the block already holds in its own right is not stored twice. */
int32_t sbp_graph_build_knn(sbp_graph *g, const int64_t *nbr, const uint8_t *vis, int64_t m,
                            int32_t k, int64_t first_row, int32_t undirected);
/* The same from the packed adjacency (nbr + 1) * 2 + vis, one int32 per candidate [m][k]: what
 * sbp_knn_edges_visible2d / sbp_prm_connect2d write and what the multi-GPU exchange assembles
 * on every rank. */
int32_t sbp_graph_build_packed(sbp_graph *g, const int32_t *packed, int64_t m, int32_t k,
                               int64_t first_row, int32_t undirected);
/* The same for rows in an explicit order: row r of packed [m][k] belongs to node row_ids[r] (device
 * int32 [m], distinct) -- the cell-sorted rows and order_out of sbp_prm_connect2d_sorted. */
int32_t sbp_graph_build_packed_rows(sbp_graph *g, const int32_t *packed, const int32_t *row_ids,
                                    int64_t m, int32_t k, int32_t undirected);
/* (Re)build from explicit arc lists src[e] -> dst[e] (device), `keep` an optional mask;
 * undirected != 0 adds dst[e] -> src[e] too (duplicates in the list are kept). */
int32_t sbp_graph_build_edges(sbp_graph *g, const int64_t *src, const int64_t *dst,
                              const uint8_t *keep, int64_t e, int32_t undirected);
int32_t sbp_graph_info(sbp_graph *g, int64_t *n_nodes, int64_t *n_edges);   /* synchronises */
/* Copy the CSR arrays to device buffers row_ptr [n_nodes + 1], col [col_capacity >= n_edges]
 * (col may be NULL); the order of a row's entries is unspecified. */
int32_t sbp_graph_csr(sbp_graph *g, int64_t *row_ptr, int32_t *col, int64_t col_capacity);
/* Weakly connected components (union-find over the arcs): labels (device int32 [n], nullable) =
 * smallest node index of each node's component; *n_components (HOST, nullable; synchronises).
 * nx.has_path(u, v) (prmPlanner.py:141) needs label[u] == label[v] -- and that suffices when the
 * graph was built undirected -- so sbp_graph_bfs settles every query with different labels
 * without a search. */
int32_t sbp_graph_components(sbp_graph *g, int32_t *labels, int64_t *n_components);
/* Batched fewest-hop paths start[q] -> goal[q] (device int64 [nq]), one CTA per searched query.
 * hops [nq]: number of edges, 0 when start == goal, -1 when unreachable (nx.has_path False)
 * or an index is out of range.  paths (optional) [nq][max_len]: node indices start..goal,
 * -1 padded; all -1 when there is no path or it has more than max_len nodes.  Among the
 * fewest-hop paths the one whose every node has the lowest-index predecessor on the previous
 * BFS level is returned (deterministic; networkx returns an adjacency-order dependent one). */
int32_t sbp_graph_bfs(sbp_graph *g, const int64_t *start, const int64_t *goal, int64_t nq,
                      int32_t *hops, int64_t *paths, int32_t max_len);

/* Multi-GPU assembly of the roadmap adjacency over peer memory: packs
 * this rank's candidate block -- (nbr + 1) * 2 + vis, one int32 per candidate -- and stores it
 * into the peer-mapped int32 buffer of every rank (peer_bufs: HOST array of `world` device
 * pointers, e.g. torch symmetric memory) at element `offset` (a multiple of 4).  One kernel:
 * packing + NVLink stores; the caller runs the cross-rank barrier. */
int32_t sbp_adjacency_peer_scatter(sbp_ctx *ctx, const int64_t *nbr, const uint8_t *vis,
                                   int64_t total, void *const *peer_bufs, int32_t world,
                                   int64_t offset);

/* The same exchange through the NVSwitch multicast (NVLS) mapping of the symmetric buffer:
 * `multimem.st` stores, replicated by the switch into every rank's buffer. */
int32_t sbp_adjacency_multicast(sbp_ctx *ctx, const int64_t *nbr, const uint8_t *vis,
                                int64_t total, void *mc_buf, int64_t offset);

#ifdef __cplusplus
}
#endif
#endif /* SBP_GPU_H */
