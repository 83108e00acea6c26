// map.cu -- context and occupancy-map handles; K6 map bit-pack kernel.
//
// Replaces the storage side of ImgCollisionChecker.__init__ /
// RobotArm4dCollisionChecker.__init__ (/root/reference/sbp_env/collisionChecker.py
// :101-108, :319-340): the reference keeps `_img` as a W x H float64 array of
// {0., 1.}; here it is 1 bit per pixel in 8x8 tiles (layout in common.cuh).
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void sbp_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *sbp_last_error(void) { return g_err; }
extern "C" int32_t sbp_abi_version(void) { return SBP_ABI_VERSION; }

// ------------------------------------------------------------- context -----
extern "C" int32_t sbp_ctx_create(int32_t device, void *stream_or_null, sbp_ctx **out) {
    SBP_REQUIRE(out != nullptr, "sbp_ctx_create: out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        sbp_set_error("no CUDA device available (%s); libsbpgpu has no CPU path",
                      e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return SBP_ERR_NODEV;
    }
    SBP_REQUIRE(device >= 0 && device < count, "sbp_ctx_create: bad device ordinal");
    SbpDeviceGuard guard(device);
    SBP_CUDA(guard.err);
    cudaDeviceProp prop;
    SBP_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        sbp_set_error("device %d is sm_%d%d; libsbpgpu is built for sm_100a (B200) only", device,
                      prop.major, prop.minor);
        return SBP_ERR_NODEV;
    }
    sbp_ctx *ctx = new sbp_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (stream_or_null) {
        ctx->stream = (cudaStream_t)stream_or_null;
    } else {
        SBP_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    SBP_CUDA(cudaMalloc(&ctx->d_flags, 16 * sizeof(int32_t)));
    SBP_CUDA(cudaMemset(ctx->d_flags, 0, 16 * sizeof(int32_t)));
    SBP_CUDA(cudaMallocHost(&ctx->h_flags, 16 * sizeof(int32_t)));
    *out = ctx;
    return SBP_OK;
}

extern "C" int32_t sbp_ctx_destroy(sbp_ctx *ctx) {
    if (!ctx) return SBP_OK;
    SbpDeviceGuard guard(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < SBP_SCRATCH_SLOTS; ++i)
        if (ctx->d_scratch[i]) cudaFree(ctx->d_scratch[i]);
    if (ctx->d_flags) cudaFree(ctx->d_flags);
    if (ctx->h_flags) cudaFreeHost(ctx->h_flags);
    if (ctx->switch_event) cudaEventDestroy(ctx->switch_event);
    for (cudaEvent_t e : ctx->aux_events) cudaEventDestroy(e);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return SBP_OK;
}

extern "C" int32_t sbp_ctx_set_stream(sbp_ctx *ctx, void *stream) {
    SBP_REQUIRE(ctx != nullptr, "sbp_ctx_set_stream: ctx is NULL");
    if ((cudaStream_t)stream == ctx->stream) return SBP_OK;
    SBP_DEVICE(ctx);
    if (ctx->own_stream) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
        ctx->own_stream = false;
    } else {
        // the context's scratch buffers and flag words are shared by everything launched through
        // it: work already queued on the old stream must be ordered before work on the new one.
        // (Not possible while either stream is being captured into a CUDA graph: a capture runs
        // on one stream from its first to its last launch and the library never switches inside.)
        cudaStreamCaptureStatus st_old = cudaStreamCaptureStatusNone, st_new = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(ctx->stream, &st_old);
        cudaStreamIsCapturing((cudaStream_t)stream, &st_new);
        if (st_old == cudaStreamCaptureStatusNone && st_new == cudaStreamCaptureStatusNone) {
            if (!ctx->switch_event)
                SBP_CUDA(cudaEventCreateWithFlags(&ctx->switch_event, cudaEventDisableTiming));
            SBP_CUDA(cudaEventRecord(ctx->switch_event, ctx->stream));
            SBP_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, ctx->switch_event, 0));
        } else {
            cudaGetLastError();
        }
    }
    ctx->stream = (cudaStream_t)stream;
    return SBP_OK;
}

extern "C" int32_t sbp_ctx_sync(sbp_ctx *ctx) {
    SBP_REQUIRE(ctx != nullptr, "sbp_ctx_sync: ctx is NULL");
    SBP_DEVICE(ctx);
    SBP_CUDA(cudaStreamSynchronize(ctx->stream));
    return SBP_OK;
}

extern "C" int32_t sbp_ctx_info(sbp_ctx *ctx, int32_t *sm_count, int64_t *launches) {
    SBP_REQUIRE(ctx != nullptr, "sbp_ctx_info: ctx is NULL");
    if (sm_count) *sm_count = ctx->sm_count;
    if (launches) *launches = ctx->launches;
    return SBP_OK;
}

extern "C" int32_t sbp_ctx_scratch_generation(sbp_ctx *ctx, int64_t *generation) {
    SBP_REQUIRE(ctx && generation, "sbp_ctx_scratch_generation: NULL argument");
    *generation = ctx->scratch_generation;
    return SBP_OK;
}

int sbp_ctx_scratch(sbp_ctx *ctx, int slot, size_t bytes, void **out) {
    if (bytes == 0) bytes = 16;
    if (ctx->d_scratch_bytes[slot] < bytes) {
        // stream-ordered users of the old buffer must finish before it is freed
        SBP_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_scratch[slot]) SBP_CUDA(cudaFree(ctx->d_scratch[slot]));
        ctx->d_scratch[slot] = nullptr;
        ctx->d_scratch_bytes[slot] = 0;
        size_t want = bytes + bytes / 4;
        want = (want + 255) & ~(size_t)255;
        SBP_CUDA(cudaMalloc(&ctx->d_scratch[slot], want));
        ctx->d_scratch_bytes[slot] = want;
        ctx->scratch_generation++;       // addresses frozen into captured graphs are now stale
    }
    *out = ctx->d_scratch[slot];
    return SBP_OK;
}

// --------------------------------------------------------- K6: bit-pack -----
// One thread builds one 32-bit word = 8 (x) x 4 (y) pixels.  The input is the
// reference's `_img` order, byte (x, y) at x * H + y, so the four y-neighbours
// of a column are contiguous.
__global__ void k_pack_map(const uint8_t *__restrict__ free_xy, int W, int H, int SX,
                           uint32_t n_words, uint32_t *__restrict__ words) {
    for (uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x; wi < n_words;
         wi += gridDim.x * blockDim.x) {
        uint32_t sec = wi >> 3, sub = wi & 7u;
        int sx = (int)(sec % (uint32_t)SX), sy = (int)(sec / (uint32_t)SX);
        int x0 = sx * 16 + (int)((sub >> 1) & 1u) * 8;
        int y0 = sy * 16 + (int)((sub >> 2) & 1u) * 8 + (int)(sub & 1u) * 4;
        uint32_t w = 0;
#pragma unroll
        for (int dx = 0; dx < 8; ++dx) {
            int x = x0 + dx;
#pragma unroll
            for (int dy = 0; dy < 4; ++dy) {
                int y = y0 + dy;
                if (x < W && y < H && free_xy[(size_t)x * (size_t)H + (size_t)y] != 0)
                    w |= 1u << ((dy << 3) | dx);
            }
        }
        words[wi] = w;
    }
}

static int32_t map_create_common(sbp_ctx *ctx, const uint8_t *d_free_xy, int32_t W, int32_t H,
                                 sbp_map **out) {
    sbp_map *m = new sbp_map();
    m->ctx = ctx;
    int SX = (W + 15) / 16, SY = (H + 15) / 16;
    uint64_t n_words64 = (uint64_t)SX * (uint64_t)SY * 8ull;
    if (n_words64 > 0xffffffffull / 4ull) {
        delete m;
        sbp_set_error("map %dx%d too large for 32-bit word indexing", W, H);
        return SBP_ERR_ARG;
    }
    uint32_t n_words = (uint32_t)n_words64;
    SBP_CUDA(cudaMalloc(&m->d_words, (size_t)n_words * 4));
    m->view.words = m->d_words;
    m->view.W = W;
    m->view.H = H;
    m->view.SX = SX;
    m->view.SY = SY;
    m->view.n_words = n_words;
    m->packed_bytes = (int64_t)n_words * 4;
    // shared-memory residency: leave room for the barrier and per-CTA bookkeeping
    m->smem_resident = (size_t)m->packed_bytes + 1024 <= (size_t)ctx->max_smem_optin;
    int threads = 256;
    int blocks = (int)((n_words + threads - 1) / threads);
    if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
    k_pack_map<<<blocks, threads, 0, ctx->stream>>>(d_free_xy, W, H, SX, n_words, m->d_words);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    *out = m;
    return SBP_OK;
}

extern "C" int32_t sbp_map_create(sbp_ctx *ctx, const uint8_t *free_xy, int32_t W, int32_t H,
                                  sbp_map **out) {
    SBP_NVTX("sbp_map_create");
    SBP_REQUIRE(ctx && free_xy && out, "sbp_map_create: NULL argument");
    SBP_REQUIRE(W > 0 && H > 0 && W <= 65536 && H <= 65536, "sbp_map_create: 1 <= W,H <= 65536");
    SBP_DEVICE(ctx);
    uint8_t *d_in = nullptr;
    size_t bytes = (size_t)W * (size_t)H;
    SBP_CUDA(cudaMalloc(&d_in, bytes));
    cudaError_t e = cudaMemcpyAsync(d_in, free_xy, bytes, cudaMemcpyHostToDevice, ctx->stream);
    int32_t rc = SBP_OK;
    if (e != cudaSuccess) {
        sbp_set_error("sbp_map_create: H2D copy failed: %s", cudaGetErrorString(e));
        rc = SBP_ERR_CUDA;
    } else {
        rc = map_create_common(ctx, d_in, W, H, out);
    }
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_in);
    return rc;
}

extern "C" int32_t sbp_map_create_device(sbp_ctx *ctx, const uint8_t *free_xy_device, int32_t W,
                                         int32_t H, sbp_map **out) {
    SBP_NVTX("sbp_map_create_device");
    SBP_REQUIRE(ctx && free_xy_device && out, "sbp_map_create_device: NULL argument");
    SBP_REQUIRE(W > 0 && H > 0 && W <= 65536 && H <= 65536,
                "sbp_map_create_device: 1 <= W,H <= 65536");
    SBP_DEVICE(ctx);
    int32_t rc = map_create_common(ctx, free_xy_device, W, H, out);
    if (rc == SBP_OK) SBP_CUDA(cudaStreamSynchronize(ctx->stream));
    return rc;
}

// Row-major twin of the tiled map (arm link sweeps): one thread per 32-pixel word.
__global__ void k_rows_from_tiles(MapView mv, int WR, uint32_t n_row_words, uint32_t *__restrict__ rows) {
    for (uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x; wi < n_row_words;
         wi += gridDim.x * blockDim.x) {
        const int y = (int)(wi / (uint32_t)WR), x0 = (int)(wi % (uint32_t)WR) * 32;
        uint32_t w = 0;
        for (int b = 0; b < 32 && y < mv.H; ++b)            // y >= H: padding words
            if (x0 + b < mv.W && map_free<true>(mv.words, mv.SX, x0 + b, y)) w |= 1u << b;
        rows[wi] = w;
    }
}

int32_t sbp_map_ensure_rows(sbp_map *m) {
    if (m->d_rows) return SBP_OK;
    sbp_ctx *ctx = m->ctx;
    // an ODD number of words per row: lanes that walk different rows at similar x then fall
    // into different shared-memory banks (bank = (y * WR + (x >> 5)) mod 32; with WR = 32, the
    // 1024-pixel arm map, 78 % of the wavefronts were bank conflicts)
    const int WR = ((m->view.W + 31) / 32) | 1;
    // padded to whole 16-byte units: the bulk copy that stages it moves multiples of 16 bytes
    const uint64_t nw = ((uint64_t)WR * (uint64_t)m->view.H + 3ull) & ~3ull;
    SBP_REQUIRE(nw <= 0xffffffffull / 4ull, "row-major map copy too large for 32-bit word indexing");
    SBP_CUDA(cudaMalloc(&m->d_rows, (size_t)nw * 4));
    m->rows = m->view;
    m->rows.words = m->d_rows;
    m->rows.SX = WR;
    m->rows.n_words = (uint32_t)nw;
    m->rows_smem_resident = (size_t)nw * 4 + 1024 <= (size_t)ctx->max_smem_optin;
    int blocks = (int)((nw + 255) / 256);
    if (blocks > ctx->sm_count * 16) blocks = ctx->sm_count * 16;
    k_rows_from_tiles<<<blocks, 256, 0, ctx->stream>>>(m->view, WR, (uint32_t)nw, m->d_rows);
    ctx->launches++;
    SBP_CUDA(cudaGetLastError());
    return SBP_OK;
}

// Coarse level: a 2-bit code per 8x8 tile (= one aligned 64-bit pair of map words): 1 = all 64
// pixels free, 2 = all blocked, 0 = mixed.  A 32-bit word holds the 4 x 4 tiles of a 32 x 32-pixel
// square, words row major.  Two planes: plane 0 indexed (u, v) = (x tile, y tile) for x-major lines,
// plane 1 its transpose, indexed (u, v) = (y tile, x tile), for y-major lines -- a walk then indexes
// its plane by (major tile, minor tile) whatever its orientation (device_fns.cuh:TileCodes):
//   word = (v >> 2) * WW + (u >> 2),  shift = (((v & 3) << 2) | (u & 3)) * 2.
// One thread per word.
__global__ void k_tile_codes(MapView mv, int WW, int WH, int transposed, uint32_t *__restrict__ af) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)WW * WH) return;
    const int wu = (int)(t % WW), wv = (int)(t / WW);
    uint32_t out = 0;
    for (int j = 0; j < 16; ++j) {
        const int u = wu * 4 + (j & 3), v = wv * 4 + (j >> 2);
        const int tx = transposed ? v : u, ty = transposed ? u : v;
        if (tx >= mv.SX * 2 || ty >= mv.SY * 2) continue;        // (tiles past the last sector: mixed)
        const uint32_t sec = (uint32_t)(ty >> 1) * (uint32_t)mv.SX + (uint32_t)(tx >> 1);
        const uint32_t wi = sec * 8u + (uint32_t)((ty & 1) * 4 + (tx & 1) * 2);
        const uint2 w = *reinterpret_cast<const uint2 *>(mv.words + wi);
        // a tile that sticks out of the map holds padding zeros: never "all free"; "all blocked"
        // is still right for every pixel inside the map
        const uint32_t code = (w.x & w.y) == 0xffffffffu ? 1u : ((w.x | w.y) == 0u ? 2u : 0u);
        out |= code << (2 * j);
    }
    af[t] = out;
}

// Free-run planes over the tile codes: does an axis-aligned rectangle of tiles hold free pixels only?
// Plane 0: bit b of word [k][u] = tile (u, 16 k + b) is all free (0 past the map's last tile row).
__global__ void k_tile_runs0(const uint32_t *__restrict__ codes, int AW, int TX, int TY, int RK, uint32_t *__restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)RK * TX) return;
    const int u = (int)(t % TX), k = (int)(t / TX);
    uint32_t w = 0;
    for (int b = 0; b < 32; ++b) {
        const int v = 16 * k + b;
        if (v >= TY) break;
        const uint32_t c = codes[(size_t)(v >> 2) * AW + (u >> 2)] >> (((v & 3) << 3) | ((u & 3) << 1));
        w |= (c & 1u) << b;
    }
    out[t] = w;
}

// Plane j from plane j - 1: a run of 2^j tiles = two runs of 2^(j-1), the second `half` columns on.
__global__ void k_tile_runs_next(const uint32_t *__restrict__ in, int TX, int RK, int half, uint32_t *__restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)RK * TX) return;
    const int u = (int)(t % TX);
    out[t] = u + half < TX ? (in[t] & in[t + half]) : 0u;
}

int32_t sbp_map_ensure_allfree(sbp_map *m) {
    if (m->af.runs) return SBP_OK;
    sbp_ctx *ctx = m->ctx;
    const int TX = m->view.SX * 2, TY = m->view.SY * 2;         // tiles
    const int WX = (TX + 3) / 4, WY = (TY + 3) / 4;             // words of 4 x 4 tiles
    const int64_t words = (int64_t)WX * WY;
    if (!m->d_allfree) SBP_CUDA(cudaMalloc(&m->d_allfree, (size_t)words * 2 * 4));
    k_tile_codes<<<(unsigned)((words + 255) / 256), 256, 0, ctx->stream>>>(m->view, WX, WY, 0, m->d_allfree);
    k_tile_codes<<<(unsigned)((words + 255) / 256), 256, 0, ctx->stream>>>(m->view, WY, WX, 1, m->d_allfree + words);
    ctx->launches += 2;
    SBP_CUDA(cudaGetLastError());
    const int RK = (TY + 15) / 16;
    const int64_t rwords = (int64_t)RK * TX;
    if (!m->d_runs) SBP_CUDA(cudaMalloc(&m->d_runs, (size_t)rwords * 4 * 4));
    const unsigned rb = (unsigned)((rwords + 255) / 256);
    k_tile_runs0<<<rb, 256, 0, ctx->stream>>>(m->d_allfree, WX, TX, TY, RK, m->d_runs);
    for (int j = 1; j < 4; ++j)
        k_tile_runs_next<<<rb, 256, 0, ctx->stream>>>(m->d_runs + (size_t)(j - 1) * rwords, TX, RK, 1 << (j - 1),
                                                      m->d_runs + (size_t)j * rwords);
    ctx->launches += 4;
    SBP_CUDA(cudaGetLastError());
    m->af.words = m->d_allfree;
    m->af.AW = WX;
    m->af.AH = WY;
    m->af.runs = m->d_runs;
    m->af.RU = TX;
    m->af.RK = RK;
    return SBP_OK;
}

extern "C" int32_t sbp_map_destroy(sbp_map *map) {
    if (!map) return SBP_OK;
    SbpDeviceGuard guard(map->ctx->device);
    cudaStreamSynchronize(map->ctx->stream);
    if (map->d_words) cudaFree(map->d_words);
    if (map->d_rows) cudaFree(map->d_rows);
    if (map->d_allfree) cudaFree(map->d_allfree);
    if (map->d_runs) cudaFree(map->d_runs);
    delete map;
    return SBP_OK;
}

extern "C" int32_t sbp_map_info(sbp_map *map, int32_t *W, int32_t *H, int64_t *packed_bytes,
                                int32_t *smem_resident) {
    SBP_REQUIRE(map != nullptr, "sbp_map_info: map is NULL");
    if (W) *W = map->view.W;
    if (H) *H = map->view.H;
    if (packed_bytes) *packed_bytes = map->packed_bytes;
    if (smem_resident) *smem_resident = map->smem_resident ? 1 : 0;
    return SBP_OK;
}
