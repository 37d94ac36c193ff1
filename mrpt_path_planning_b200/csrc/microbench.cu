// Roofline denominators that MEASURED_PEAKS.json does not carry: CUDA-core FP32 and FP64
// FMA throughput, measured on the box with dependent-chain-free FMA loops.
#include "mppb_internal.cuh"

namespace mppb
{
namespace
{
template <typename T>
__global__ void fma_peak_kernel(T* out, int iters, T a, T b)
{
    T acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = static_cast<T>(threadIdx.x + i);
    for (int it = 0; it < iters; it++)
    {
#pragma unroll
        for (int u = 0; u < 8; u++)
        {
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = acc[i] * a + b;
        }
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i];
    if (s == static_cast<T>(-1)) out[0] = s;  // keep the chain alive
}

template <typename T>
int measure(mppb_ctx* ctx, double* out_tflops)
{
    MPPB_CUDA(ctx, cudaSetDevice(ctx->device));
    MPPB_CUDA(ctx, ctx->bboxScratch.reserve(64));
    const int blocks = ctx->numSMs * 16, threads = 256, iters = sizeof(T) == 4 ? 4096 : 1024;
    double    best   = 0;
    for (int rep = 0; rep < 5; rep++)
    {
        MPPB_CUDA(ctx, cudaEventRecord(ctx->evBegin, ctx->stream));
        fma_peak_kernel<T><<<blocks, threads, 0, ctx->stream>>>(ctx->bboxScratch.as<T>(), iters, T(1.0000001), T(1e-7));
        ctx->launches++;
        MPPB_CUDA(ctx, cudaEventRecord(ctx->evEnd, ctx->stream));
        MPPB_CUDA(ctx, cudaEventSynchronize(ctx->evEnd));
        float ms = 0;
        MPPB_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->evBegin, ctx->evEnd));
        const double flops = 2.0 * 64.0 * iters * static_cast<double>(blocks) * threads;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    MPPB_CUDA(ctx, cudaGetLastError());
    *out_tflops = best;
    return MPPB_OK;
}
}  // namespace
}  // namespace mppb

extern "C" {
int mppb_measure_fp32_tflops(mppb_ctx* ctx, double* out_tflops)
{
    if (!ctx || !out_tflops) return MPPB_ERR_INVALID_ARG;
    return mppb::measure<float>(ctx, out_tflops);
}
int mppb_measure_fp64_tflops(mppb_ctx* ctx, double* out_tflops)
{
    if (!ctx || !out_tflops) return MPPB_ERR_INVALID_ARG;
    return mppb::measure<double>(ctx, out_tflops);
}
}
