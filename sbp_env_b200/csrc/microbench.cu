// microbench.cu -- measured denominators for the gather / fp64 rooflines that
// MEASURED_PEAKS.json does not hold: random 4-byte gathers
// from shared memory, random 32-byte-sector gathers from an L2- or HBM-resident
// buffer, and the unfused fp64 add/mul issue rate.  Bench utilities, not part of the
// declared ABI (like sbp_tune).
#include "common.cuh"

__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s; }

// which = 0: shared-memory random word gather (bank-conflict limited, like a warp whose
// lanes walk 32 unrelated edges)
__global__ void __launch_bounds__(1024) k_mb_smem(int words, int iters, uint32_t *out) {
    extern __shared__ uint32_t sm[];
    for (int i = threadIdx.x; i < words; i += blockDim.x) sm[i] = i * 2654435761u;
    __syncthreads();
    uint32_t s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u, acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 8
        for (int u = 0; u < 8; ++u) acc ^= sm[lcg(s) % (uint32_t)words];
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// which = 1: random 32-B sector gather from global memory (one 4-byte load per sector)
__global__ void __launch_bounds__(256) k_mb_gather(const uint32_t *__restrict__ buf, uint64_t sectors,
                                                    int iters, uint32_t *out) {
    uint32_t s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u, acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 8
        for (int u = 0; u < 8; ++u) {
            uint64_t r = ((uint64_t)lcg(s) << 16) ^ lcg(s);
            acc ^= __ldg(buf + (r % sectors) * 8);
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// which = 2: independent unfused fp64 multiply/add chains (the kNN inner loop's pipe)
__global__ void __launch_bounds__(256) k_mb_fp64(int iters, double *out) {
    double a0 = threadIdx.x * 1e-3 + 1.0, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 4
        for (int u = 0; u < 4; ++u) {
            a0 = __dadd_rn(__dmul_rn(a0, b), c); a1 = __dadd_rn(__dmul_rn(a1, b), c);
            a2 = __dadd_rn(__dmul_rn(a2, b), c); a3 = __dadd_rn(__dmul_rn(a3, b), c);
        }
    }
    if (a0 + a1 + a2 + a3 == 12345.678) out[0] = a0;
}

// Returns in *ms the average duration of one launch and in *units the number of
// gathers (which 0/1) or fp64 operations (which 2) per launch.
extern "C" int32_t sbp_microbench(sbp_ctx *ctx, int32_t which, int64_t bytes, int32_t iters,
                                  double *ms, double *units) {
    SBP_REQUIRE(ctx && ms && units, "sbp_microbench: NULL argument");
    SBP_CUDA(cudaSetDevice(ctx->device));
    cudaEvent_t e0, e1;
    SBP_CUDA(cudaEventCreate(&e0));
    SBP_CUDA(cudaEventCreate(&e1));
    uint32_t *d_out = nullptr, *buf = nullptr;
    SBP_CUDA(cudaMalloc(&d_out, 64));
    const int reps = 5;
    float t = 0;
    if (which == 0) {
        int words = (int)(bytes / 4);
        SBP_CUDA(cudaFuncSetAttribute(k_mb_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        int grid = ctx->sm_count * (bytes <= 100 * 1024 ? 2 : 1);
        k_mb_smem<<<grid, 1024, bytes, ctx->stream>>>(words, iters, d_out);
        SBP_CUDA(cudaEventRecord(e0, ctx->stream));
        for (int r = 0; r < reps; ++r) k_mb_smem<<<grid, 1024, bytes, ctx->stream>>>(words, iters, d_out);
        SBP_CUDA(cudaEventRecord(e1, ctx->stream));
        *units = (double)grid * 1024.0 * iters * 8.0;
    } else if (which == 1) {
        SBP_CUDA(cudaMalloc(&buf, (size_t)bytes));
        SBP_CUDA(cudaMemsetAsync(buf, 1, (size_t)bytes, ctx->stream));
        int grid = ctx->sm_count * 8;
        k_mb_gather<<<grid, 256, 0, ctx->stream>>>(buf, (uint64_t)bytes / 32, iters, d_out);
        SBP_CUDA(cudaEventRecord(e0, ctx->stream));
        for (int r = 0; r < reps; ++r)
            k_mb_gather<<<grid, 256, 0, ctx->stream>>>(buf, (uint64_t)bytes / 32, iters, d_out);
        SBP_CUDA(cudaEventRecord(e1, ctx->stream));
        *units = (double)grid * 256.0 * iters * 8.0;
    } else {
        int grid = ctx->sm_count * 8;
        k_mb_fp64<<<grid, 256, 0, ctx->stream>>>(iters, (double *)d_out);
        SBP_CUDA(cudaEventRecord(e0, ctx->stream));
        for (int r = 0; r < reps; ++r) k_mb_fp64<<<grid, 256, 0, ctx->stream>>>(iters, (double *)d_out);
        SBP_CUDA(cudaEventRecord(e1, ctx->stream));
        *units = (double)grid * 256.0 * iters * 4.0 * 8.0;   // 4 unrolled x 4 chains x (mul+add)
    }
    SBP_CUDA(cudaEventSynchronize(e1));
    SBP_CUDA(cudaEventElapsedTime(&t, e0, e1));
    *ms = t / reps;
    ctx->launches += reps + 1;
    if (buf) cudaFree(buf);
    cudaFree(d_out);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return SBP_OK;
}
