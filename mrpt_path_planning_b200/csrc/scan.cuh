// Exclusive prefix sum of a uint32 array on the device, in three small launches (tile scan, scan of the tile totals,
// add-back). Used for the bin table of the cloud index (cloud_bins.cu) and the per-cell entry counts of the
// collision-grid builder (colgrid_build.cu). n up to 2^24 elements (tile totals fit one CTA).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace mppb
{
namespace scan_detail
{
constexpr int kScanThreads = 1024;
constexpr int kScanPer     = 4;                          // elements per thread
constexpr int kScanTile    = kScanThreads * kScanPer;    // elements per CTA

// block-wide exclusive scan of one value per thread; returns the exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_exclusive(uint32_t v, uint32_t* total)
{
    __shared__ uint32_t warpSum[32];
    const int           lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t            x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const uint32_t t = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += t;
    }
    if (lane == 31) warpSum[warp] = x;
    __syncthreads();
    if (warp == 0)
    {
        uint32_t w = warpSum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        warpSum[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const uint32_t before = (warp ? warpSum[warp - 1] : 0u) + (x - v);
    *total                = warpSum[31];
    __syncthreads();  // warpSum may be reused by a following call
    return before;
}

// tile scan: out[i] = exclusive prefix within the tile; tileTotals[tile] = tile sum
static __global__ void __launch_bounds__(kScanThreads)
    scan_tiles_kernel(const uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ out, uint32_t* __restrict__ tileTotals)
{
    const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanPer;
    uint32_t       v[kScanPer], sum = 0;
#pragma unroll
    for (int j = 0; j < kScanPer; j++)
    {
        v[j] = base + j < n ? in[base + j] : 0u;
        sum += v[j];
    }
    uint32_t total;
    uint32_t run = block_exclusive(sum, &total);
#pragma unroll
    for (int j = 0; j < kScanPer; j++)
    {
        if (base + j < n) out[base + j] = run;
        run += v[j];
    }
    if (threadIdx.x == 0) tileTotals[blockIdx.x] = total;
}

// exclusive scan of the tile totals in place (one CTA, numTiles <= 4096); grand total -> *grand
static __global__ void __launch_bounds__(kScanThreads) scan_totals_kernel(uint32_t* tileTotals, uint32_t numTiles, uint32_t* grand)
{
    const uint32_t base = threadIdx.x * kScanPer;
    uint32_t       v[kScanPer], sum = 0;
#pragma unroll
    for (int j = 0; j < kScanPer; j++)
    {
        v[j] = base + j < numTiles ? tileTotals[base + j] : 0u;
        sum += v[j];
    }
    uint32_t total;
    uint32_t run = block_exclusive(sum, &total);
#pragma unroll
    for (int j = 0; j < kScanPer; j++)
    {
        if (base + j < numTiles) tileTotals[base + j] = run;
        run += v[j];
    }
    if (threadIdx.x == 0) *grand = total;
}

// out[i] += tile offset; optionally zero the input (it then serves as a scatter cursor)
static __global__ void __launch_bounds__(kScanThreads)
    scan_addback_kernel(uint32_t* __restrict__ out, uint32_t n, const uint32_t* __restrict__ tileTotals, uint32_t* zeroThis)
{
    const uint32_t off  = tileTotals[blockIdx.x];
    const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanPer;
#pragma unroll
    for (int j = 0; j < kScanPer; j++)
        if (base + j < n)
        {
            out[base + j] += off;
            if (zeroThis) zeroThis[base + j] = 0u;
        }
}
}  // namespace scan_detail

// out[0..n) = exclusive prefix sums of in[0..n), out[n] = total. `scratch` holds >= 4096 + 1 uint32.
// Returns the number of kernels launched (3), or -1 if n is too large.
static inline int device_exclusive_scan(const uint32_t* in, uint32_t n, uint32_t* out, uint32_t* scratch, uint32_t* zeroThis,
                                 cudaStream_t st)
{
    using namespace scan_detail;
    const uint32_t tiles = (n + kScanTile - 1) / kScanTile;
    if (tiles > static_cast<uint32_t>(kScanTile)) return -1;
    scan_tiles_kernel<<<tiles, kScanThreads, 0, st>>>(in, n, out, scratch);
    scan_totals_kernel<<<1, kScanThreads, 0, st>>>(scratch, tiles, out + n);
    scan_addback_kernel<<<tiles, kScanThreads, 0, st>>>(out, n, scratch, zeroThis);
    return 3;
}

}  // namespace mppb
