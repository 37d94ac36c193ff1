// Device-side arithmetic of the two built-in cost evaluators at one sample pose (shared by edge_cost.cu and expand.cu).
//   CostEvaluatorCostMap::eval_single_pose            (CostEvaluatorCostMap.cpp:157-162): truncating cell look-up, 0 outside
//   CostEvaluatorPreferredWaypoint::eval_single_pose  (CostEvaluatorPreferredWaypoint.cpp:94-120): radius search over the list
// Reference operation order; every translation unit that includes this is compiled with -fmad=false (build.py), so the
// results are bit-identical to the CPU evaluation (no transcendental function is involved, sqrt is IEEE-exact).
#pragma once
#include "mppb_internal.cuh"

namespace mppb
{
__device__ __forceinline__ double costmap_at(const EvaluatorDev& E, double x, double y)
{
    // CDynamicGrid::cellByPos -> x2idx = int((x - x_min) / res), null outside [0,size)
    const int cx = static_cast<int>((x - E.xmin) / E.res);
    const int cy = static_cast<int>((y - E.ymin) / E.res);
    if (cx < 0 || cx >= E.nx || cy < 0 || cy >= E.ny) return .0;
    return E.cells[cx + static_cast<size_t>(cy) * E.nx];
}

__device__ __forceinline__ double waypoints_at(const EvaluatorDev& E, double x, double y)
{
    const float qx = static_cast<float>(x), qy = static_cast<float>(y);
    double      cost    = E.scale;
    float       last    = -1.f;  // squared distance of the previous match
    int         lastIdx = -1;
    for (;;)
    {
        // next match in ascending (distance, index) order
        float best    = 0.f;
        int   bestIdx = -1;
        for (int i = 0; i < E.nWps; i++)
        {
            const float2 w  = E.wps[i];
            const float  dx = qx - w.x, dy = qy - w.y;
            const float  d2 = dx * dx + dy * dy;
            if (!(d2 < E.radiusSqrF)) continue;
            if (!(d2 > last || (d2 == last && i > lastIdx))) continue;
            if (bestIdx < 0 || d2 < best)
            {
                best    = d2;
                bestIdx = i;
            }
        }
        if (bestIdx < 0) break;
        last             = best;
        lastIdx          = bestIdx;
        const float dist = sqrtf(best);
        if (dist > E.radius || dist < 0) continue;
        cost -= E.scale * (1.0 - sqrt(static_cast<double>(dist) / E.radius));
    }
    return cost > .0 ? cost : .0;
}

// value of evaluator E at the global point (px, py)
__device__ __forceinline__ double evaluator_at(const EvaluatorDev& E, double px, double py)
{
    return E.kind == 1 ? costmap_at(E, px, py) : waypoints_at(E, px, py);
}
}  // namespace mppb
