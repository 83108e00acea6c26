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

// which = 3 / 4: independent fp32 FMA chains, packed (fma.rn.f32x2 -> FFMA2, two FMAs per
// lane and instruction, the filter scan's instruction) or scalar (FFMA): the measured
// denominator of the kNN filter's FMA-pipe roofline.  8 independent accumulators per thread,
// the multiplier a scalar broadcast as in the scan.
__global__ void __launch_bounds__(256) k_mb_ffma2(int iters, float *out) {
    uint64_t acc[8];
    float q = 1.0000001f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float v = threadIdx.x * 1e-3f + i;
        asm("mov.b64 %0, {%1, %2};" : "=l"(acc[i]) : "f"(v), "f"(v + 0.5f));
    }
    uint64_t q2, c2;
    asm("mov.b64 %0, {%1, %1};" : "=l"(q2) : "f"(q));
    asm("mov.b64 %0, {%1, %1};" : "=l"(c2) : "f"(1e-9f));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(q2), "l"(c2));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
        s += lo + hi;
    }
    if (s == 12345.678f) out[0] = s;
}

__global__ void __launch_bounds__(256) k_mb_ffma(int iters, float *out) {
    float acc[16];
    const float q = 1.0000001f, c = 1e-9f;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 16; ++i)
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[i]) : "f"(q), "f"(c));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 12345.678f) out[0] = s;
}

// which = 5 / 6: three-input (FMNMX3) and two-input (FMNMX) fp32 minima on independent
// chains: the ALU-pipe denominator of the filter scan's min reduction.
template <bool THREE>
__global__ void __launch_bounds__(256) k_mb_fmin(int iters, float *out) {
    float acc[16];
    const float b = threadIdx.x * 0.5f + 3.f, c = threadIdx.x * 0.25f + 5.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3f + i + 100.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (THREE) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(acc[i]) : "f"(b), "f"(c));
                else asm volatile("min.f32 %0, %0, %1;" : "+f"(acc[i]) : "f"(b));
            }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 12345.678f) out[0] = s;
}

// which = 7: FFMA2 and FMNMX interleaved 1:1 on independent chains -- do the packed FMAs and
// the minima share a datapath (times add) or overlap (the slower one hides the other)?
__global__ void __launch_bounds__(256) k_mb_mix(int iters, float *out) {
    uint64_t acc[8];
    float mn[8];
    const float b = threadIdx.x * 0.5f + 3.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float v = threadIdx.x * 1e-3f + i;
        asm("mov.b64 %0, {%1, %2};" : "=l"(acc[i]) : "f"(v), "f"(v + 0.5f));
        mn[i] = v + 100.f;
    }
    uint64_t q2, c2;
    asm("mov.b64 %0, {%1, %1};" : "=l"(q2) : "f"(1.0000001f));
    asm("mov.b64 %0, {%1, %1};" : "=l"(c2) : "f"(1e-9f));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(q2), "l"(c2));
                asm volatile("min.f32 %0, %0, %1;" : "+f"(mn[i]) : "f"(b));
            }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
        s += lo + hi + mn[i];
    }
    if (s == 12345.678f) out[0] = s;
}

// Returns in *ms the average duration of one launch and in *units the number of
// gathers (which 0/1), fp64 operations (which 2) or fp32 FMAs (which 3/4) per launch.
extern "C" int32_t sbp_microbench(sbp_ctx *ctx, int32_t which, int64_t bytes, int32_t iters,
                                  double *ms, double *units) {
    SBP_REQUIRE(ctx && ms && units, "sbp_microbench: NULL argument");
    SBP_DEVICE(ctx);
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
    } else if (which >= 3 && which <= 7) {
        int grid = ctx->sm_count * 8;
        auto launch = [&]() {
            if (which == 3) k_mb_ffma2<<<grid, 256, 0, ctx->stream>>>(iters, (float *)d_out);
            else if (which == 4) k_mb_ffma<<<grid, 256, 0, ctx->stream>>>(iters, (float *)d_out);
            else if (which == 5) k_mb_fmin<true><<<grid, 256, 0, ctx->stream>>>(iters, (float *)d_out);
            else if (which == 6) k_mb_fmin<false><<<grid, 256, 0, ctx->stream>>>(iters, (float *)d_out);
            else k_mb_mix<<<grid, 256, 0, ctx->stream>>>(iters, (float *)d_out);   // 8 FFMA2 + 8 FMNMX per 16 "units"
        };
        launch();
        SBP_CUDA(cudaEventRecord(e0, ctx->stream));
        for (int r = 0; r < reps; ++r) launch();
        SBP_CUDA(cudaEventRecord(e1, ctx->stream));
        // units = fp32 FMAs (3/4) or min instructions x 32 lanes (5/6) per launch:
        // 4 unrolled x (8 packed x 2 | 16 scalar) per thread and iteration
        *units = (double)grid * 256.0 * iters * 4.0 * 16.0;
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
