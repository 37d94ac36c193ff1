// Cost evaluator plugins of the mpp:: mirror (product code). The arithmetic over obstacle points
// and edge samples runs on the GPU through the C ABI; the host keeps the plugin objects, their
// parameters and the one transcendental (pow) whose rounding must match the reference's libm.
// Reference: src/algos/CostEvaluatorCostMap.cpp:54-162, src/algos/CostEvaluatorPreferredWaypoint.cpp:49-120,
// src/algos/Planner.cpp:15-29.
#include <cmath>
#include <cstring>

#include "mpp.h"

namespace mpp
{
namespace
{
void ck(mppb_ctx* ctx, int rc)
{
    if (rc != MPPB_OK) throw std::runtime_error(std::string("libmpp_b200: ") + mppb_last_error(ctx));
}

// flatten MoveEdgeSE2_TPS::interpolatedPath into the (start pose, relative sample list) the C ABI takes
struct EdgeSamples
{
    double                from[3];
    std::vector<uint32_t> offsets{0};
    std::vector<double>   rel;
    explicit EdgeSamples(const MoveEdgeSE2_TPS& e)
    {
        MPP_ASSERT(!e.interpolatedPath.empty());
        from[0] = e.stateFrom.pose.x;
        from[1] = e.stateFrom.pose.y;
        from[2] = e.stateFrom.pose.phi;
        for (const auto& kv : e.interpolatedPath)
        {
            rel.push_back(kv.second.x);
            rel.push_back(kv.second.y);
        }
        offsets.push_back(static_cast<uint32_t>(e.interpolatedPath.size()));
    }
};
// the evaluator's private context, created on first use on the device of `like`
mppb_ctx* private_ctx(detail::EvaluatorGpu& g, mppb_ctx* like)
{
    if (!g.ctx)
    {
        if (mppb_ctx_create(mppb_ctx_device(like), nullptr, &g.ctx) != MPPB_OK)
            throw std::runtime_error(std::string("cost evaluator: no CUDA context: ") + mppb_last_global_error());
        g.uploaded = false;
    }
    return g.ctx;
}
// true if the resident copy was made from these parameters; records them otherwise
bool same_key(detail::EvaluatorGpu& g, const double (&k)[6])
{
    if (g.uploaded && std::memcmp(g.key, k, sizeof(k)) == 0) return true;
    std::memcpy(g.key, k, sizeof(k));
    return false;
}
}  // namespace

// ---- CostEvaluatorCostMap ----------------------------------------------------------------------------
CostEvaluatorCostMap::Parameters CostEvaluatorCostMap::Parameters::FromYAML(const Yaml& c)
{
    MPP_ASSERT(c.isMap());
    Parameters p;
    p.resolution                 = c.getDouble("resolution");
    p.preferredClearanceDistance = c.getDouble("preferredClearanceDistance");
    p.maxCost                    = c.getDouble("maxCost");
    p.useAverageOfPath           = c.getBool("useAverageOfPath");
    p.maxRadiusFromRobot         = c.getDouble("maxRadiusFromRobot");
    return p;
}

CostEvaluatorCostMap::Ptr CostEvaluatorCostMap::FromStaticPointObstacles(mppb_ctx* ctx, const PointCloud& obsPts,
                                                                         const Parameters& p,
                                                                         const std::optional<TPose2D>& curRobotPose)
{
    if (!ctx) throw std::runtime_error("CostEvaluatorCostMap needs a libmpp_b200 context (no CPU fallback)");
    MPP_ASSERT(!obsPts.empty());
    auto cm     = std::make_shared<CostEvaluatorCostMap>();
    cm->params_ = p;
    cm->ctx_    = ctx;
    const float D = static_cast<float>(p.preferredClearanceDistance);
    // required area: obstacle bounding box grown by D, or a square around the robot (float arithmetic
    // of TBoundingBoxf / TPoint3Df in the reference)
    float minx, miny, maxx, maxy;
    obsPts.boundingBox(minx, miny, maxx, maxy);
    minx -= D;
    miny -= D;
    maxx += D;
    maxy += D;
    if (p.maxRadiusFromRobot > 0)
    {
        MPP_ASSERT(curRobotPose.has_value());
        const double R = p.maxRadiusFromRobot + D;
        minx           = static_cast<float>(curRobotPose->x - R);
        miny           = static_cast<float>(curRobotPose->y - R);
        maxx           = static_cast<float>(curRobotPose->x + R);
        maxy           = static_cast<float>(curRobotPose->y + R);
    }
    // CDynamicGrid::setSize
    const double res = p.resolution;
    cm->x_min        = res * round_i(minx / res);
    cm->y_min        = res * round_i(miny / res);
    cm->x_max        = res * round_i(maxx / res);
    cm->y_max        = res * round_i(maxy / res);
    cm->size_x       = round_i((cm->x_max - cm->x_min) / res);
    cm->size_y       = round_i((cm->y_max - cm->y_min) / res);
    MPP_ASSERT(cm->size_x > 0 && cm->size_y > 0);
    const size_t ncells = static_cast<size_t>(cm->size_x) * cm->size_y;
    cm->cells.assign(ncells, .0);

    // nearest obstacle per cell centre: GPU (float squared distances, exact minimum below D)
    std::vector<float> d2(ncells);
    ck(ctx, mppb_costmap_nearest_d2(ctx, obsPts.xs.data(), obsPts.ys.data(), obsPts.size(), cm->x_min, cm->y_min, res,
                                    cm->size_x, cm->size_y, D, d2.data()));
    // cost shaping: the only cells that matter are the few with d < D; pow() stays on the host so
    // that it is the same libm call as in the reference (CostEvaluatorCostMap.cpp:99-101)
    for (size_t i = 0; i < ncells; i++)
    {
        const float d = std::sqrt(d2[i]);
        if (d < D)
        {
            const double cost = p.maxCost * std::pow(-0.99999 + 1. / (d / D), 0.4);
            MPP_ASSERT(cost >= .0);
            cm->cells[i] = cost;
        }
    }
    return cm;
}

double CostEvaluatorCostMap::operator()(const MoveEdgeSE2_TPS& edge) const
{
    if (!ctx_) throw std::runtime_error("CostEvaluatorCostMap is not bound to a libmpp_b200 context");
    const EdgeSamples           s(edge);
    std::lock_guard<std::mutex> lk(gpu_->m);
    mppb_ctx*                   c = private_ctx(*gpu_, ctx_);
    const double key[6] = {x_min, y_min, params_.resolution, static_cast<double>(size_x) * size_y,
                           params_.useAverageOfPath ? 1.0 : 0.0, static_cast<double>(reinterpret_cast<uintptr_t>(cells.data()))};
    if (!same_key(*gpu_, key))  // the grid travels once (it is immutable once built by FromStaticPointObstacles)
    {
        mppb_costmap_desc d{x_min, y_min, params_.resolution, size_x, size_y, cells.data(), params_.useAverageOfPath ? 1 : 0};
        ck(c, mppb_clear_evaluators(c));
        ck(c, mppb_set_costmap(c, 0, &d));
        gpu_->uploaded = true;
    }
    double out = 0;
    ck(c, mppb_eval_edge_costs(c, 1, s.from, s.offsets.data(), s.rel.data(), &out));
    return out;
}

// ---- CostEvaluatorPreferredWaypoint --------------------------------------------------------------------
CostEvaluatorPreferredWaypoint::Parameters CostEvaluatorPreferredWaypoint::Parameters::FromYAML(const Yaml& c)
{
    MPP_ASSERT(c.isMap());
    Parameters p;
    p.waypointInfluenceRadius = c.getDouble("waypointInfluenceRadius");
    p.costScale               = c.getDouble("costScale");
    p.useAverageOfPath        = c.getBool("useAverageOfPath");
    return p;
}

double CostEvaluatorPreferredWaypoint::operator()(const MoveEdgeSE2_TPS& edge) const
{
    if (!ctx_) throw std::runtime_error("CostEvaluatorPreferredWaypoint is not bound to a libmpp_b200 context");
    const EdgeSamples           s(edge);
    std::lock_guard<std::mutex> lk(gpu_->m);
    mppb_ctx*                   c = private_ctx(*gpu_, ctx_);
    double key[6] = {params_.waypointInfluenceRadius, params_.costScale, params_.useAverageOfPath ? 1.0 : 0.0,
                     static_cast<double>(waypoints_.size()), 0, 0};
    for (const auto& w : waypoints_) key[4] += w.x, key[5] += w.y;
    if (!same_key(*gpu_, key))
    {
        std::vector<double> xs, ys;
        for (const auto& w : waypoints_)
        {
            xs.push_back(w.x);
            ys.push_back(w.y);
        }
        ck(c, mppb_clear_evaluators(c));
        ck(c, mppb_set_waypoints(c, 0, xs.data(), ys.data(), xs.size(), params_.waypointInfluenceRadius,
                                 params_.costScale, params_.useAverageOfPath ? 1 : 0));
        gpu_->uploaded = true;
    }
    double out = 0;
    ck(c, mppb_eval_edge_costs(c, 1, s.from, s.offsets.data(), s.rel.data(), &out));
    return out;
}

void CostEvaluatorPreferredWaypoint::setPreferredWaypoints(const std::vector<TPoint2D>& pts)
{
    std::lock_guard<std::mutex> lk(gpu_->m);
    waypoints_     = pts;
    gpu_->uploaded = false;
}

std::vector<TPoint2D> waypoint_targets_from_yaml(const Yaml& y)
{
    std::vector<TPoint2D> pts;
    for (const Yaml& wp : y.sequence())
    {
        const auto t = wp.getDoubleVector("target");
        if (t.size() < 2) throw std::runtime_error("Waypoint `target` must have at least 2 coordinates");
        pts.push_back({t[0], t[1]});
    }
    return pts;
}

// ---- Planner::cost_path_segment ---------------------------------------------------------------------------
cost_t Planner::cost_path_segment(const MoveEdgeSE2_TPS& edge) const
{
    cost_t c = edge.estimatedExecTime;
    for (const auto& ce : costEvaluators_)
    {
        MPP_ASSERT(ce);
        c += (*ce)(edge);
    }
    return c;
}

}  // namespace mpp
