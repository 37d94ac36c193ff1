// edge_cost_kernel — stage (c) of the hot path: cost evaluators over the interpolated samples
// of candidate edges.
//
// Replaces, for a batch of edges at once, the per-edge calls of Planner::cost_path_segment
// (Planner.cpp:15-29):
//   CostEvaluatorCostMap::operator()           (CostEvaluatorCostMap.cpp:122-155)
//     eval_single_pose: truncating cell lookup, 0 outside the grid          (:157-162)
//   CostEvaluatorPreferredWaypoint::operator() (CostEvaluatorPreferredWaypoint.cpp:59-92)
//     eval_single_pose: radius search over the waypoint list                (:94-120)
// Each sample pose is stateFrom.pose (+) relPose (CostEvaluatorCostMap.cpp:149-150); only its
// (x, y) matter. cos/sin of the start heading come from the host C library; the composition,
// the cell index ((x - x_min) / res, C truncation) and the waypoint arithmetic (float squared
// distances summed as nanoflann's L2_Simple adaptor does, matches visited by ascending distance)
// use the reference's operation order. Compiled with -fmad=false: results are bit-identical to
// the CPU evaluation (no transcendental function is involved, sqrt is IEEE-exact).
//
// One thread per (edge, evaluator slot): the samples of an edge are folded in time order, which
// the "max keeps the last maximum / average divides the running sum" rule requires. The double
// costmap (4.8 MB at config C2) stays L2-resident; nothing here is bandwidth-bound.
#include <math_constants.h>

#include <algorithm>
#include <type_traits>

#include "evaluators.cuh"
#include "mppb_internal.cuh"

namespace mppb
{
namespace
{
__global__ void __launch_bounds__(128)
    edge_cost_kernel(const EvaluatorDev* __restrict__ evals, int nEvals, int nEdges,
                     const double* __restrict__ from /* [n][4] x,y,cos,sin */,
                     const uint32_t* __restrict__ offsets, const double* __restrict__ rel, double* __restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nEdges * nEvals) return;
    const int           e = t / nEvals, slot = t - e * nEvals;
    const EvaluatorDev& E = evals[slot];
    const double        x0 = from[4 * e + 0], y0 = from[4 * e + 1], c = from[4 * e + 2], s = from[4 * e + 3];
    double              cost = .0;
    unsigned            n    = 0;
    for (uint32_t i = offsets[e]; i < offsets[e + 1]; i++)
    {
        const double bx = rel[2 * i], by = rel[2 * i + 1];
        // TPose2D::operator+ : x + b.x*cos - b.y*sin ; y + b.x*sin + b.y*cos
        const double px = x0 + bx * c - by * s;
        const double py = y0 + bx * s + by * c;
        const double v  = E.kind == 1 ? costmap_at(E, px, py) : waypoints_at(E, px, py);
        if (E.useAverage)
        {
            cost += v;
            ++n;
        }
        else if (v >= cost)
        {
            cost = v;
            n    = 1;
        }
    }
    // the reference asserts n > 0 (an edge always has >= 2 samples); an empty list yields NaN here
    out[t] = cost / static_cast<double>(n);
}

// Small batches (the tentative edges of ONE A* expansion: a few dozen edges): one CTA. The general kernel walks an
// edge's samples one by one — for inputs read in place from pinned host memory that is a PCIe read plus a costmap read
// per sample, one after the other (≈12 us for 7 samples). Here the CTA first copies the whole input into shared
// memory (coalesced, one PCIe round trip), then every thread issues the evaluator look-ups of all samples of its edge
// together and folds them in time order afterwards. Same arithmetic, same results.
constexpr int kSmallMaxPairs   = 1024;  // (edge, evaluator) pairs = threads of the CTA
constexpr int kSmallMaxEdges   = 512;
constexpr int kSmallMaxSamples = 1024;
constexpr int kSmallRound      = 8;  // samples whose look-ups are in flight together
__global__ void __launch_bounds__(kSmallMaxPairs)
    edge_cost_small_kernel(const EvaluatorDev* __restrict__ evals, int nEvals, int nEdges, int nSamples,
                           const double* __restrict__ from /* [n][4] x,y,cos,sin */,
                           const uint32_t* __restrict__ offsets, const double* __restrict__ rel, double* __restrict__ out)
{
    extern __shared__ double sm[];
    double*   sFrom = sm;
    double*   sRel  = sm + 4 * nEdges;
    uint32_t* sOff  = reinterpret_cast<uint32_t*>(sRel + 2 * nSamples);
    // four loads per thread in flight per round (a load-store loop would wait for every PCIe read in turn)
    auto stage = [&](auto* dst, const auto* src, int n) {
        for (int base = threadIdx.x; base < n; base += 4 * blockDim.x)
        {
            std::remove_reference_t<decltype(dst[0])> r[4];
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (base + j * static_cast<int>(blockDim.x) < n) r[j] = src[base + j * blockDim.x];
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (base + j * static_cast<int>(blockDim.x) < n) dst[base + j * blockDim.x] = r[j];
        }
    };
    stage(sFrom, from, 4 * nEdges);
    stage(sRel, rel, 2 * nSamples);
    stage(sOff, offsets, nEdges + 1);
    __syncthreads();
    const int t = threadIdx.x;
    if (t >= nEdges * nEvals) return;
    const int           e = t / nEvals, slot = t - e * nEvals;
    const EvaluatorDev& E = evals[slot];
    const double        x0 = sFrom[4 * e + 0], y0 = sFrom[4 * e + 1], c = sFrom[4 * e + 2], s = sFrom[4 * e + 3];
    const uint32_t      first = sOff[e], n = sOff[e + 1] - first;
    const bool          isMap = E.kind == 1, average = E.useAverage != 0;
    double              cost = .0;
    unsigned            cnt  = 0;
    for (uint32_t base = 0; base < n; base += kSmallRound)
    {
        double v[kSmallRound];
#pragma unroll
        for (int j = 0; j < kSmallRound; j++)
        {
            v[j] = .0;
            if (base + j < n)
            {
                const double bx = sRel[2 * (first + base + j)], by = sRel[2 * (first + base + j) + 1];
                // TPose2D::operator+ : x + b.x*cos - b.y*sin ; y + b.x*sin + b.y*cos
                const double px = x0 + bx * c - by * s;
                const double py = y0 + bx * s + by * c;
                v[j]            = isMap ? costmap_at(E, px, py) : waypoints_at(E, px, py);
            }
        }
#pragma unroll
        for (int j = 0; j < kSmallRound; j++)
        {
            if (base + j >= n) break;
            if (average)
            {
                cost += v[j];
                ++cnt;
            }
            else if (v[j] >= cost)
            {
                cost = v[j];
                cnt  = 1;
            }
        }
    }
    out[t] = cost / static_cast<double>(cnt);
}

// costmap builder: nearest float squared distance from each cell centre to the cloud
// (CostEvaluatorCostMap.cpp:91-97 with nanoflann's float L2: dx*dx + dy*dy)
__global__ void __launch_bounds__(256)
    costmap_d2_kernel(CloudBins bins, double xmin, double ymin, double res, int nx, int ny, float clearance,
                      float* __restrict__ out)
{
    const int cx = blockIdx.x * blockDim.x + threadIdx.x;
    const int cy = blockIdx.y * blockDim.y + threadIdx.y;
    if (cx >= nx || cy >= ny) return;
    // CDynamicGrid::idx2x = x_min + (cx + 0.5) * res, narrowed to float by the caller (:91,94)
    const float x    = static_cast<float>(xmin + (cx + 0.5) * res);
    const float y    = static_cast<float>(ymin + (cy + 0.5) * res);
    float       best = CUDART_INF_F;
    if (bins.n > 0)
    {
        // any point with sqrt(d2) < clearance lies inside this (slightly widened) window
        const float w   = clearance * 1.0001f + 1e-6f;
        const int   bx0 = bin_coord(x - w, bins.minx, bins.invBin, bins.nbx);
        const int   bx1 = bin_coord(x + w, bins.minx, bins.invBin, bins.nbx);
        const int   by0 = bin_coord(y - w, bins.miny, bins.invBin, bins.nby);
        const int   by1 = bin_coord(y + w, bins.miny, bins.invBin, bins.nby);
        for (int row = by0; row <= by1; row++)
        {
            const size_t   r   = static_cast<size_t>(row) * bins.nbx;
            const uint32_t beg = bins.binStart[r + bx0], end = bins.binStart[r + bx1 + 1];
            for (uint32_t p = beg; p < end; p++)
            {
                const float2 g  = bins.pts[p];
                const float  dx = x - g.x, dy = y - g.y;
                const float  d2 = dx * dx + dy * dy;
                best            = d2 < best ? d2 : best;
            }
        }
    }
    out[cx + static_cast<size_t>(cy) * nx] = best;
}
}  // namespace

int launch_edge_costs(mppb_ctx* ctx, size_t nEdges, const double* dFrom, const uint32_t* dOffsets, const double* dRel,
                      double* dOut, size_t nSamples)
{
    if (nEdges == 0 || ctx->numEvals == 0) return MPPB_OK;
    const size_t total = nEdges * static_cast<size_t>(ctx->numEvals);
    if (total > 0x7fffffffull) return fail(ctx, MPPB_ERR_INVALID_ARG, "too many edges in one batch");
    if (nSamples > 0 && total <= kSmallMaxPairs && nEdges <= kSmallMaxEdges && nSamples <= kSmallMaxSamples)
    {
        // one A* expansion's worth of edges: the single-CTA kernel (sample count known to the caller)
        const size_t   smem    = (4 * nEdges + 2 * nSamples) * sizeof(double) + (nEdges + 1) * sizeof(uint32_t);
        const unsigned threads = std::max(256u, static_cast<unsigned>((total + 31) / 32 * 32));  // >= 256 for the staging
        edge_cost_small_kernel<<<1, threads, smem, ctx->stream>>>(ctx->evalDescs.as<EvaluatorDev>(), ctx->numEvals,
                                                                  static_cast<int>(nEdges), static_cast<int>(nSamples),
                                                                  dFrom, dOffsets, dRel, dOut);
        ctx->launches++;
        MPPB_CUDA(ctx, cudaGetLastError());
        return MPPB_OK;
    }
    const unsigned blocks = static_cast<unsigned>((total + 127) / 128);
    edge_cost_kernel<<<blocks, 128, 0, ctx->stream>>>(ctx->evalDescs.as<EvaluatorDev>(), ctx->numEvals,
                                                      static_cast<int>(nEdges), dFrom, dOffsets, dRel, dOut);
    ctx->launches++;
    MPPB_CUDA(ctx, cudaGetLastError());
    return MPPB_OK;
}

int launch_costmap_d2(mppb_ctx* ctx, const CloudStore& cs, double xmin, double ymin, double res, int nx, int ny,
                      float clearance, float* dOut)
{
    if (nx <= 0 || ny <= 0) return MPPB_OK;
    const dim3 block(32, 8);
    const dim3 grid((nx + block.x - 1) / block.x, (ny + block.y - 1) / block.y);
    costmap_d2_kernel<<<grid, block, 0, ctx->stream>>>(cs.bins, xmin, ymin, res, nx, ny, clearance, dOut);
    ctx->launches++;
    MPPB_CUDA(ctx, cudaGetLastError());
    return MPPB_OK;
}

}  // namespace mppb
