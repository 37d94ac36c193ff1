// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product.
//
// CPU restatement of the free functions of the hot path (stage a/b drivers), following
// /root/reference/mrpt_path_planning/src/algos/*.cpp line by line. Pinned against
// the reference's own sources compiled as oracle/_ref (tests/test_oracle_vs_reference.py); the MRPT
// layer underneath is assumed.
#pragma once
#include "orc_ptg.hpp"

namespace orc
{
// src/algos/transform_pc_square_clipping.cpp:9-43
inline void transform_pc_square_clipping(
    const PointCloud& inMap, const CPose2D& asSeenFrom, const double MAX_DIST_XY, PointCloud& outMap,
    bool appendToOutMap = true)
{
    const size_t nObs = inMap.size();
    if (!appendToOutMap) outMap.clear();
    const CPose2D invPose = asSeenFrom.inverse();
    for (size_t obs = 0; obs < nObs; obs++)
    {
        const double gx = inMap.xs[obs], gy = inMap.ys[obs];
        if (std::abs(gx - asSeenFrom.x) > MAX_DIST_XY || std::abs(gy - asSeenFrom.y) > MAX_DIST_XY) continue;
        double ox, oy;
        invPose.composePoint(gx, gy, ox, oy);
        outMap.insertPointFast(static_cast<float>(ox), static_cast<float>(oy));
    }
}

// src/algos/tp_obstacles_single_path.cpp:9-40
inline double tp_obstacles_single_path(int tp_space_k_direction, const PointCloud& localObstacles, const PTG& ptg)
{
    const size_t nObs = localObstacles.size();
    double       out  = 0;
    ptg.initTPObstacleSingle(tp_space_k_direction, out);
    for (size_t obs = 0; obs < nObs; obs++)
    {
        const float ox = localObstacles.xs[obs], oy = localObstacles.ys[obs];
        // the PTG virtual takes uint16_t k (tp_obstacles_single_path.cpp:31-32)
        ptg.updateTPObstacleSingle(ox, oy, static_cast<uint16_t>(tp_space_k_direction), out);
    }
    return out;
}

// TPS_Astar.cpp:828-847 (no caching in the reference)
inline PointCloud cached_local_obstacles(
    const TPose2D& queryPose, const std::vector<const PointCloud*>& globalObstacles, double MAX_PTG_XY_DIST)
{
    PointCloud out;
    for (const PointCloud* obs : globalObstacles)
        transform_pc_square_clipping(*obs, CPose2D(queryPose), MAX_PTG_XY_DIST, out);
    return out;
}

}  // namespace orc
