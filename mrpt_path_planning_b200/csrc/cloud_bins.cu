// Uniform-bin index over the global obstacle cloud (built once per cloud upload).
//
// The reference rescans the whole cloud for every expanded node
// (transform_pc_square_clipping.cpp:26-36, "MRPT_TODO Impl actual cache" at
// TPS_Astar.cpp:835). Here the cloud is counting-sorted by bin once, so a node only
// touches the bin rows overlapping its clipping square. All kernels are HBM streaming
// passes over the cloud (8 B/point read, 8 B/point written).
#include <cfloat>
#include <cmath>
#include <limits>

#include "mppb_internal.cuh"
#include "scan.cuh"

namespace mppb
{
namespace
{
__device__ __forceinline__ int ordered_int(float f)
{
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
inline float from_ordered_int(int i)
{
    const int   j = i >= 0 ? i : i ^ 0x7fffffff;
    float       f;
    memcpy(&f, &j, sizeof(f));
    return f;
}

// scratch: [0]=minx [1]=miny [2]=maxx [3]=maxy (ordered ints)
// a point takes part in the index if it is finite and inside the optional clipping box
// (ObstacleSource.cpp:29-40: float comparisons, inclusive bounds)
struct KeepTest
{
    bool  hasBox;
    float minx, miny, maxx, maxy;
    __device__ __forceinline__ bool operator()(float x, float y) const
    {
        if (!(isfinite(x) && isfinite(y))) return false;
        if (hasBox && (x < minx || x > maxx || y < miny || y > maxy)) return false;
        return true;
    }
};

__global__ void bbox_kernel(const float* __restrict__ xs, const float* __restrict__ ys, size_t n, KeepTest keep,
                            int* scratch)
{
    float mnx = FLT_MAX, mny = FLT_MAX, mxx = -FLT_MAX, mxy = -FLT_MAX;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x)
    {
        const float x = xs[i], y = ys[i];
        if (keep(x, y))
        {
            mnx = fminf(mnx, x);
            mxx = fmaxf(mxx, x);
            mny = fminf(mny, y);
            mxy = fmaxf(mxy, y);
        }
    }
    for (int o = 16; o > 0; o >>= 1)
    {
        mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    if ((threadIdx.x & 31) == 0)
    {
        atomicMin(&scratch[0], ordered_int(mnx));
        atomicMin(&scratch[1], ordered_int(mny));
        atomicMax(&scratch[2], ordered_int(mxx));
        atomicMax(&scratch[3], ordered_int(mxy));
    }
}

__global__ void bin_count_kernel(
    const float* __restrict__ xs, const float* __restrict__ ys, size_t n, KeepTest keep, float minx, float miny,
    float invBin, int nbx, int nby, uint32_t* counts)
{
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x)
    {
        const float x = xs[i], y = ys[i];
        if (keep(x, y))
        {
            const int b = bin_coord(y, miny, invBin, nby) * nbx + bin_coord(x, minx, invBin, nbx);
            atomicAdd(&counts[b], 1u);
        }
    }
}

__global__ void bin_scatter_kernel(
    const float* __restrict__ xs, const float* __restrict__ ys, size_t n, KeepTest keep, float minx, float miny,
    float invBin, int nbx, int nby, const uint32_t* __restrict__ start, uint32_t* cursor, float2* sorted)
{
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x)
    {
        const float x = xs[i], y = ys[i];
        if (keep(x, y))
        {
            const int      b   = bin_coord(y, miny, invBin, nby) * nbx + bin_coord(x, minx, invBin, nbx);
            const uint32_t pos = start[b] + atomicAdd(&cursor[b], 1u);
            sorted[pos]        = make_float2(x, y);
        }
    }
}
}  // namespace

int build_cloud_bins(mppb_ctx* ctx, CloudStore& cs, double binSize)
{
    cudaStream_t st = ctx->stream;
    const size_t n  = cs.rawN;
    cs.bins         = CloudBins();
    const KeepTest keep{cs.hasBox, cs.box[0], cs.box[1], cs.box[2], cs.box[3]};
    if (n == 0) return MPPB_OK;
    if (n > 0xfffffff0ull) return fail(ctx, MPPB_ERR_UNSUPPORTED, "obstacle cloud larger than 2^32 points");
    if (!(binSize > 0)) binSize = 0.5;

    const float* xs = cs.rawX.as<float>();
    const float* ys = cs.rawY.as<float>();

    // 1) bounding box of the finite points
    MPPB_CUDA(ctx, ctx->bboxScratch.reserve(4 * sizeof(int)));
    const int init[4] = {INT32_MAX, INT32_MAX, INT32_MIN, INT32_MIN};
    MPPB_CUDA(ctx, cudaMemcpyAsync(ctx->bboxScratch.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
    const int threads = 256;
    const int blocks  = static_cast<int>(std::min<size_t>((n + threads - 1) / threads, ctx->numSMs * 8));
    bbox_kernel<<<blocks, threads, 0, st>>>(xs, ys, n, keep, ctx->bboxScratch.as<int>());
    ctx->launches++;
    int h[4];
    MPPB_CUDA(ctx, cudaMemcpyAsync(h, ctx->bboxScratch.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    MPPB_CUDA(ctx, cudaStreamSynchronize(st));
    const float minx = from_ordered_int(h[0]), miny = from_ordered_int(h[1]);
    const float maxx = from_ordered_int(h[2]), maxy = from_ordered_int(h[3]);
    // No point kept (all non-finite, or all outside the clipping box): every warp still folded its neutral
    // FLT_MAX / -FLT_MAX into the scratch, so the test is on the decoded box, not on the initial pattern.
    if (h[0] == INT32_MAX || !(minx <= maxx) || !(miny <= maxy)) return MPPB_OK;  // cs.bins stays empty (n == 0)

    // 2) bin grid: at most 2^24 bins
    double s = binSize;
    for (;;)
    {
        const double nbx = std::floor((static_cast<double>(maxx) - minx) / s) + 1;
        const double nby = std::floor((static_cast<double>(maxy) - miny) / s) + 1;
        if (nbx * nby <= 16777216.0 && nbx < 1e6 && nby < 1e6) break;
        s *= 2;
    }
    CloudBins b;
    b.minx        = minx;
    b.miny        = miny;
    b.invBin      = static_cast<float>(1.0 / s);
    b.nbx         = static_cast<int>(std::floor((static_cast<double>(maxx) - minx) / s)) + 1;
    b.nby         = static_cast<int>(std::floor((static_cast<double>(maxy) - miny) / s)) + 1;
    const int nb  = b.nbx * b.nby;
    cs.binSize    = s;

    // 3) histogram, scan, scatter
    MPPB_CUDA(ctx, cs.binStart.reserve((static_cast<size_t>(nb) + 1) * sizeof(uint32_t)));
    MPPB_CUDA(ctx, cs.binCursor.reserve(static_cast<size_t>(nb) * sizeof(uint32_t)));
    // (+2 points: the staged form of tpobs_grid_kernel copies 16-byte-aligned pieces, which may take one point past a
    // run's end; past the last point it finds a sentinel that no clipping square contains)
    MPPB_CUDA(ctx, cs.sortedPts.reserve((n + 2) * sizeof(float2)));
    MPPB_CUDA(ctx, cudaMemsetAsync(cs.binCursor.p, 0, static_cast<size_t>(nb) * sizeof(uint32_t), st));
    bin_count_kernel<<<blocks, threads, 0, st>>>(xs, ys, n, keep, b.minx, b.miny, b.invBin, b.nbx, b.nby,
                                                cs.binCursor.as<uint32_t>());
    // counts -> start (exclusive scan, start[nb] = total); the counts are zeroed: they become the scatter cursor
    MPPB_CUDA(ctx, cs.scanScratch.reserve((4096 + 1) * sizeof(uint32_t)));
    if (device_exclusive_scan(cs.binCursor.as<uint32_t>(), static_cast<uint32_t>(nb), cs.binStart.as<uint32_t>(),
                              cs.scanScratch.as<uint32_t>(), cs.binCursor.as<uint32_t>(), st) < 0)
        return fail(ctx, MPPB_ERR_UNSUPPORTED, "bin table too large");
    ctx->launches += 2;
    bin_scatter_kernel<<<blocks, threads, 0, st>>>(xs, ys, n, keep, b.minx, b.miny, b.invBin, b.nbx, b.nby,
                                                  cs.binStart.as<uint32_t>(), cs.binCursor.as<uint32_t>(),
                                                  cs.sortedPts.as<float2>());
    ctx->launches += 3;
    uint32_t total = 0;
    MPPB_CUDA(ctx, cudaMemcpyAsync(&total, cs.binStart.as<uint32_t>() + nb, sizeof(uint32_t),
                                   cudaMemcpyDeviceToHost, st));
    MPPB_CUDA(ctx, cudaStreamSynchronize(st));
    {
        const float  inf       = std::numeric_limits<float>::infinity();
        const float2 sentinel[2] = {make_float2(inf, inf), make_float2(inf, inf)};
        MPPB_CUDA(ctx, cudaMemcpyAsync(cs.sortedPts.as<float2>() + total, sentinel, sizeof(sentinel), cudaMemcpyHostToDevice, st));
    }
    MPPB_CUDA(ctx, cudaGetLastError());
    b.n        = total;
    b.pts      = cs.sortedPts.as<float2>();
    b.binStart = cs.binStart.as<uint32_t>();
    cs.bins    = b;
    return MPPB_OK;
}

}  // namespace mppb
