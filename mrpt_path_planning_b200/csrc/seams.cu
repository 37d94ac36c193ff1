// Single-call seams of the reference's free functions (include/mppb.h, "free-function seams"): the same arithmetic as
// the batched kernels, for callers that hold ONE pose or ONE trajectory, as the reference's own call sites do
// (TPS_Astar.cpp:828-847 -> transform_pc_square_clipping; :744-745 -> tp_obstacles_single_path).
//
// transform_clip_kernel: mpp::transform_pc_square_clipping (src/algos/transform_pc_square_clipping.cpp:9-43) — points
// inside the +-maxDist square around the pose (global axes, FP64 compare), composed with the inverse pose in FP64,
// stored as float, in INPUT ORDER (the reference appends them as it scans). Order-preserving compaction: flag pass,
// prefix sum (scan.cuh), scatter pass. Compiled with -fmad=false (build.py); cos/sin of the heading from the host.
#include <algorithm>
#include <cmath>

#include "mppb_internal.cuh"
#include "scan.cuh"

namespace mppb
{
namespace
{
__device__ __forceinline__ bool in_square(float gx, float gy, double x0, double y0, double D)
{
    // transform_pc_square_clipping.cpp:30-31: skipped if abs(g - center) > D on either axis (NaN compares false: kept,
    // exactly as the reference's `>` does — callers do not feed non-finite points)
    return !(fabs(static_cast<double>(gx) - x0) > D) && !(fabs(static_cast<double>(gy) - y0) > D);
}
__global__ void flag_kernel(const float* __restrict__ xs, const float* __restrict__ ys, uint32_t n, double x0, double y0, double D,
                            uint32_t* __restrict__ flags)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = in_square(xs[i], ys[i], x0, y0, D) ? 1u : 0u;
}
__global__ void scatter_kernel(const float* __restrict__ xs, const float* __restrict__ ys, uint32_t n, double x0, double y0, double c,
                               double s, double D, const uint32_t* __restrict__ pos, float* __restrict__ ox, float* __restrict__ oy)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gx = xs[i], gy = ys[i];
    if (!in_square(gx, gy, x0, y0, D)) return;
    // CPose2D unary minus, then composePoint: (ix + gx*c) + gy*s ; (iy - gx*s) + gy*c
    const double ix = -x0 * c - y0 * s, iy = x0 * s - y0 * c;
    const double lx = (ix + static_cast<double>(gx) * c) + static_cast<double>(gy) * s;
    const double ly = (iy - static_cast<double>(gx) * s) + static_cast<double>(gy) * c;
    ox[pos[i]]      = static_cast<float>(lx);  // CPointsMap stores float
    oy[pos[i]]      = static_cast<float>(ly);
}
}  // namespace
}  // namespace mppb

using namespace mppb;

extern "C" {

int mppb_transform_pc_square_clipping(mppb_ctx* ctx, const float* xs, const float* ys, size_t n, const double* pose_xyphi,
                                      double max_dist, float* out_x, float* out_y, size_t* out_n)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    if (!pose_xyphi || !out_n || (n && (!xs || !ys || !out_x || !out_y))) return fail(ctx, MPPB_ERR_INVALID_ARG, "null pointers");
    *out_n = 0;
    if (n == 0) return MPPB_OK;
    if (n > 0xfffffff0ull) return fail(ctx, MPPB_ERR_UNSUPPORTED, "cloud larger than 2^32 points");
    MPPB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t   st = ctx->stream;
    const uint32_t N  = static_cast<uint32_t>(n);
    // scratch: raw x, y | flags | positions | out x, y | scan scratch
    MPPB_CUDA(ctx, ctx->dStage.reserve(2 * n * sizeof(float)));
    MPPB_CUDA(ctx, ctx->dStage2.reserve((2 * n + 4) * sizeof(uint32_t)));
    MPPB_CUDA(ctx, ctx->dStage3.reserve(2 * n * sizeof(float)));
    MPPB_CUDA(ctx, ctx->dOut.reserve((4096 + 8) * sizeof(uint32_t)));
    float*    dX    = ctx->dStage.as<float>();
    float*    dY    = dX + n;
    uint32_t* flags = ctx->dStage2.as<uint32_t>();
    uint32_t* pos   = flags + n + 1;
    float*    oX    = ctx->dStage3.as<float>();
    float*    oY    = oX + n;
    MPPB_CUDA(ctx, cudaMemcpyAsync(dX, xs, n * sizeof(float), cudaMemcpyHostToDevice, st));
    MPPB_CUDA(ctx, cudaMemcpyAsync(dY, ys, n * sizeof(float), cudaMemcpyHostToDevice, st));
    const double   x0 = pose_xyphi[0], y0 = pose_xyphi[1], c = std::cos(pose_xyphi[2]), s = std::sin(pose_xyphi[2]);
    const unsigned blocks = (N + 255) / 256;
    flag_kernel<<<blocks, 256, 0, st>>>(dX, dY, N, x0, y0, max_dist, flags);
    uint32_t* grand = pos + n;  // the scan leaves the total behind the prefix sums
    if (device_exclusive_scan(flags, N, pos, ctx->dOut.as<uint32_t>(), nullptr, st) < 0)
        return fail(ctx, MPPB_ERR_UNSUPPORTED, "cloud too large for the scan");
    scatter_kernel<<<blocks, 256, 0, st>>>(dX, dY, N, x0, y0, c, s, max_dist, pos, oX, oY);
    ctx->launches += 5;
    uint32_t total = 0;
    MPPB_CUDA(ctx, cudaMemcpyAsync(&total, grand, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    MPPB_CUDA(ctx, cudaStreamSynchronize(st));
    MPPB_CUDA(ctx, cudaGetLastError());
    if (total)
    {
        MPPB_CUDA(ctx, cudaMemcpyAsync(out_x, oX, total * sizeof(float), cudaMemcpyDeviceToHost, st));
        MPPB_CUDA(ctx, cudaMemcpyAsync(out_y, oY, total * sizeof(float), cudaMemcpyDeviceToHost, st));
        MPPB_CUDA(ctx, cudaStreamSynchronize(st));
    }
    *out_n = total;
    return MPPB_OK;
}

}  // extern "C"
