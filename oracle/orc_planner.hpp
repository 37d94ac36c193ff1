// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product.
//
// CPU restatement of the planner-level code of the hot path and of its callers
// (/root/reference/mrpt_path_planning/...): the data model (include/mpp/data/*.h), the cost
// evaluators (src/algos/CostEvaluator*.cpp), edge_interpolated_path, Planner::cost_path_segment,
// TPS_Astar (whole file), refine_trajectory, plan_to_trajectory / save_to_txt and
// ObstacleSource clipping. Each function cites the reference lines it follows; MRPT behaviour
// that is not under /root/reference is marked [MRPT]. Pinned against
// the reference's own sources compiled as oracle/_ref (tests/test_oracle_vs_reference.py).
#pragma once
#include <chrono>
#include <deque>
#include <functional>
#include <list>
#include <map>
#include <optional>
#include <set>
#include <sstream>
#include <unordered_map>
#include <unordered_set>

#include "orc_algos.hpp"
#include "orc_holo.hpp"

namespace orc
{
using TNodeID = uint64_t;

// include/mpp/data/SE2_KinState.h:65-87, src/data/SE2_KinState.cpp:36-57
struct SE2_KinState
{
    TPose2D  pose{0, 0, 0};
    TTwist2D vel;
};
struct SE2orR2_KinState
{
    bool     isPoint = false;  // PoseOrPoint variant: pose (x,y,phi) or point (x,y)
    TPose2D  pose;             // phi unused for a point
    TTwist2D vel;
    SE2_KinState asSE2KinState() const
    {
        SE2_KinState s;
        s.vel  = vel;
        s.pose = isPoint ? TPose2D(pose.x, pose.y, .0) : pose;
        return s;
    }
};

// include/mpp/data/MoveEdgeSE2_TPS.h:20-68
struct MoveEdge
{
    TNodeID      parentId = static_cast<TNodeID>(-1);
    SE2_KinState stateFrom, stateTo;
    double       cost                 = std::numeric_limits<double>::max();
    int8_t       ptgIndex             = -1;
    int16_t      ptgPathIndex         = -1;
    double       ptgDist              = std::numeric_limits<double>::max();
    double       ptgTrimmableSpeed    = 1.0;
    TPose2D      ptgFinalRelativeGoal;
    double       ptgFinalGoalRelSpeed = .0;
    double       estimatedExecTime    = .0;
    std::map<double, TPose2D> interpolatedPath;
    uint32_t     ptgStep = 0;  // bookkeeping for tests only (relTrgStep of the accepted neighbour)

    // src/data/MoveEdgeSE2_TPS.cpp:14-31
    NavDynState getPTGDynState() const
    {
        NavDynState d;
        d.curVelLocal = stateFrom.vel;
        d.curVelLocal.rotate(-stateFrom.pose.phi);
        d.relTarget      = ptgFinalRelativeGoal;
        d.targetRelSpeed = ptgFinalGoalRelSpeed;
        return d;
    }
};

// include/mpp/data/MotionPrimitivesTree.h:72-300 (+ [MRPT] CDirectedTree storage)
struct MotionTree
{
    struct Node : SE2_KinState
    {
        TNodeID                nodeID = static_cast<TNodeID>(-1);
        std::optional<TNodeID> parentID;
        double                 cost = 0;
    };
    struct EdgeInfo
    {
        TNodeID  id;
        MoveEdge data;
    };
    TNodeID                                root = 0;
    std::map<TNodeID, std::list<EdgeInfo>> edges_to_children;
    std::map<TNodeID, Node>                nodes;

    void insert_root_node(TNodeID id, const SE2_KinState& s)
    {
        ORC_ASSERT(nodes.empty());
        Node n;
        static_cast<SE2_KinState&>(n) = s;
        n.nodeID                      = id;
        n.cost                        = 0;
        nodes[id]                     = n;
    }
    void insert_node_and_edge(TNodeID parentId, TNodeID childId, const SE2_KinState& childState, const MoveEdge& e)
    {
        edges_to_children[parentId].push_back(EdgeInfo{childId, e});
        const double newCost = nodes.at(parentId).cost + e.cost;
        Node         n;
        static_cast<SE2_KinState&>(n) = childState;
        n.nodeID                      = childId;
        n.parentID                    = parentId;
        n.cost                        = newCost;
        nodes[childId]                = n;
    }
    void rewire_node_parent(TNodeID nodeId, const MoveEdge& newEdgeFromParent)
    {
        Node&      node        = nodes.at(nodeId);
        const auto oldParentId = *node.parentID;
        auto&      oldEdges    = edges_to_children[oldParentId];
        bool       done        = false;
        for (auto it = oldEdges.begin(); it != oldEdges.end(); ++it)
            if (it->id == nodeId)
            {
                oldEdges.erase(it);
                done = true;
                break;
            }
        if (!done) throw std::runtime_error("[rewire_node_parent] could not find edge from former parent");
        const TNodeID parentId = newEdgeFromParent.parentId;
        edges_to_children[parentId].push_back(EdgeInfo{nodeId, newEdgeFromParent});
        const double newCost = nodes.at(parentId).cost + newEdgeFromParent.cost;
        ORC_ASSERT(newCost <= node.cost);
        node.parentID = parentId;
        node.cost     = newCost;
    }
    const MoveEdge& edge_to_parent(TNodeID nodeId) const
    {
        const Node& node = nodes.at(nodeId);
        for (const auto& e : edges_to_children.at(*node.parentID))
            if (e.id == nodeId) return e.data;
        throw std::runtime_error("Could not find edge to parent");
    }
    // :259-294
    void backtrack_path(TNodeID target, std::list<Node>& outPath, std::list<MoveEdge*>& edgeList) const
    {
        auto it = nodes.find(target);
        if (it == nodes.end()) throw std::runtime_error("backtrackPath: target_node not found in tree!");
        const Node* node = &it->second;
        for (;;)
        {
            outPath.push_front(*node);
            if (!node->parentID.has_value()) break;
            edgeList.push_front(const_cast<MoveEdge*>(&edge_to_parent(node->nodeID)));
            node = &nodes.at(*node->parentID);
        }
    }
};

using PtgList = std::vector<std::shared_ptr<PTG>>;

// src/algos/edge_interpolated_path.cpp:9-73
inline void edge_interpolated_path(
    MoveEdge& edge, const PtgList& ptgs, const TPose2D& reconstrRelPose, size_t ptg_step,
    const std::optional<size_t>& numSegments = std::nullopt)
{
    size_t nSeg = 0;
    if (numSegments.has_value())
        nSeg = *numSegments;
    else
    {
        ORC_ASSERT(edge.interpolatedPath.size() > 1);
        nSeg = edge.interpolatedPath.size();
    }
    auto&        ptg = ptgs.at(edge.ptgIndex);
    const double dt  = ptg->getPathStepDuration();
    auto&        ip  = edge.interpolatedPath;
    ip.clear();
    ip[0 * dt] = TPose2D(0, 0, 0);
    for (size_t i = 0; i < nSeg; i++)
    {
        const auto   iStep = ((i + 1) * ptg_step) / (nSeg + 2);
        const double t     = iStep * dt;
        ip[t]              = ptg->getPathPose(edge.ptgPathIndex, iStep);
    }
    ip[ptg_step * dt]      = reconstrRelPose;
    edge.estimatedExecTime = ptg_step * dt;
}

// include/mpp/algos/CostEvaluator.h:22-43
struct CostEvaluator
{
    virtual ~CostEvaluator()                             = default;
    virtual double operator()(const MoveEdge& edge) const = 0;
};

// shared "max or average over the interpolated samples" reducer
// (CostEvaluatorCostMap.cpp:122-155 == CostEvaluatorPreferredWaypoint.cpp:59-92)
template <class F>
inline double reduce_over_path(const MoveEdge& edge, bool useAverageOfPath, F evalSinglePose)
{
    double cost = .0;
    size_t n    = 0;
    ORC_ASSERT(!edge.interpolatedPath.empty());
    for (const auto& kv : edge.interpolatedPath)
    {
        const double c = evalSinglePose(edge.stateFrom.pose + kv.second);
        ORC_ASSERT(c >= .0);
        if (useAverageOfPath)
        {
            cost += c;
            ++n;
        }
        else if (c >= cost)
        {
            cost = c;
            n    = 1;
        }
    }
    ORC_ASSERT(n);
    return cost / n;
}

// src/algos/CostEvaluatorCostMap.cpp
struct CostEvaluatorCostMap : CostEvaluator
{
    struct Parameters
    {
        double resolution = 0.05, preferredClearanceDistance = 0.4, maxCost = 2.0, maxRadiusFromRobot = .0;
        bool   useAverageOfPath = false;
    } params;
    DynGrid<double> costmap;

    // :54-120
    static std::shared_ptr<CostEvaluatorCostMap> FromStaticPointObstacles(
        const PointCloud& obsPts, const Parameters& p, const std::optional<TPose2D>& curRobotPose)
    {
        auto cm    = std::make_shared<CostEvaluatorCostMap>();
        cm->params = p;
        ORC_ASSERT(obsPts.size() > 0);
        const float D = p.preferredClearanceDistance;
        // [MRPT] CPointsMap::boundingBox() -> TBoundingBoxf
        float minx = obsPts.xs[0], maxx = minx, miny = obsPts.ys[0], maxy = miny;
        for (size_t i = 1; i < obsPts.size(); i++)
        {
            minx = std::min(minx, obsPts.xs[i]);
            maxx = std::max(maxx, obsPts.xs[i]);
            miny = std::min(miny, obsPts.ys[i]);
            maxy = std::max(maxy, obsPts.ys[i]);
        }
        minx -= D;
        miny -= D;
        maxx += D;
        maxy += D;
        if (p.maxRadiusFromRobot > 0)
        {
            ORC_ASSERT(curRobotPose.has_value());
            const double R = p.maxRadiusFromRobot + D;
            minx           = static_cast<float>(curRobotPose->x - R);
            miny           = static_cast<float>(curRobotPose->y - R);
            maxx           = static_cast<float>(curRobotPose->x + R);
            maxy           = static_cast<float>(curRobotPose->y + R);
        }
        cm->costmap.setSize(minx, maxx, miny, maxy, p.resolution, .0);
        auto& g = cm->costmap;
        for (int cy = 0; cy < g.size_y; cy++)
        {
            const float y = g.idx2y(cy);
            for (int cx = 0; cx < g.size_x; cx++)
            {
                const float x = g.idx2x(cx);
                const float d = std::sqrt(obsPts.closestPointSqrDist(x, y));
                if (d < D)
                {
                    const double cost = p.maxCost * std::pow(-0.99999 + 1. / (d / D), 0.4);
                    ORC_ASSERT(cost >= .0);
                    *g.cellByIndex(cx, cy) = cost;
                }
            }
        }
        return cm;
    }
    // :157-162
    double eval_single_pose(const TPose2D& p) const
    {
        const double* cell = costmap.cellByPos(p.x, p.y);
        return cell ? *cell : .0;
    }
    // :122-155
    double operator()(const MoveEdge& edge) const override
    {
        return reduce_over_path(edge, params.useAverageOfPath, [this](const TPose2D& p) { return eval_single_pose(p); });
    }
};

// src/algos/CostEvaluatorPreferredWaypoint.cpp
struct CostEvaluatorPreferredWaypoint : CostEvaluator
{
    struct Parameters
    {
        double waypointInfluenceRadius = 1.5, costScale = 10.0;
        bool   useAverageOfPath = true;
    } params;
    PointCloud waypoints;  // float storage (CSimplePointsMap)

    // :49-57
    void setPreferredWaypoints(const std::vector<TPoint2D>& pts)
    {
        waypoints.clear();
        for (const auto& pt : pts) waypoints.insertPointFast(static_cast<float>(pt.x), static_cast<float>(pt.y));
    }
    // :94-120. [MRPT/nanoflann] kdTreeRadiusSearch2D: float query, float squared distances
    // (d0*d0 + d1*d1), matches with dist^2 < radius^2, sorted by ascending distance.
    double eval_single_pose(const TPose2D& p) const
    {
        const double       inflRadius = params.waypointInfluenceRadius;
        const float        qx = p.x, qy = p.y;
        const float        maxRadiusSqr = square(inflRadius);
        std::vector<float> near;
        for (size_t i = 0; i < waypoints.size(); i++)
        {
            const float dx = qx - waypoints.xs[i], dy = qy - waypoints.ys[i];
            const float d2 = dx * dx + dy * dy;
            if (d2 < maxRadiusSqr) near.push_back(d2);
        }
        std::sort(near.begin(), near.end());
        double cost = params.costScale;
        for (const float d2 : near)
        {
            const float dist = std::sqrt(d2);
            if (dist > inflRadius || dist < 0) continue;
            cost -= params.costScale * (1.0f - std::sqrt(dist / inflRadius));
        }
        return std::max(.0, cost);
    }
    // :59-92
    double operator()(const MoveEdge& edge) const override
    {
        return reduce_over_path(edge, params.useAverageOfPath, [this](const TPose2D& p) { return eval_single_pose(p); });
    }
};

// src/interfaces/ObstacleSource.cpp:19-42 (static point cloud source)
inline PointCloud apply_clipping_box(const PointCloud& in, double minx, double miny, double maxx, double maxy)
{
    const float fminx = minx, fminy = miny, fmaxx = maxx, fmaxy = maxy;
    PointCloud  out;
    for (size_t i = 0; i < in.size(); i++)
    {
        if (in.xs[i] < fminx || in.xs[i] > fmaxx || in.ys[i] < fminy || in.ys[i] > fmaxy) continue;
        out.insertPointFast(in.xs[i], in.ys[i]);
    }
    return out;
}

// include/mpp/algos/within_bbox.h
inline bool within_bbox(const TPose2D& p, const TPose2D& max, const TPose2D& min)
{
    return p.x < max.x && p.y < max.y && p.phi < max.phi + 1e-6 && p.x > min.x && p.y > min.y && p.phi > min.phi - 1e-6;
}
inline bool within_bbox_pt(const TPose2D& p, const TPose2D& max, const TPose2D& min)
{
    return p.x < max.x && p.y < max.y && p.x > min.x && p.y > min.y;
}

struct PlannerInput
{
    SE2_KinState            stateStart;
    SE2orR2_KinState        stateGoal;
    TPose2D                 worldBboxMin, worldBboxMax;
    std::vector<PointCloud> obstacles;  // one static cloud per ObstacleSource
    PtgList                 ptgs;
};
struct PlannerOutput
{
    bool                   success         = false;
    double                 computationTime = 0;
    double                 pathCost        = std::numeric_limits<double>::max();
    std::optional<TNodeID> goalNodeId, bestNodeId;
    double                 bestNodeIdCostToGoal = std::numeric_limits<double>::max();
    MotionTree             motionTree;
    // statistics for the benchmark (not in the reference)
    size_t nExpanded = 0, nTpsPointsEvaluated = 0, nEdgesCosted = 0;
};

// include/mpp/algos/TPS_Astar.h:24-57
struct TPS_Astar_Parameters
{
    double              grid_resolution_xy       = 0.20;
    double              grid_resolution_yaw      = 5.0 * M_PI / 180.0;
    double              SE2_metricAngleWeight    = 1.0;
    double              heuristic_heading_weight = 0.1;
    uint32_t            max_ptg_trajectories_to_explore = 20;
    std::vector<double> ptg_sample_timestamps           = {1.0, 3.0, 5.0};
    uint32_t            max_ptg_speeds_to_explore       = 3;
    size_t              pathInterpolatedSegments        = 5;
    double              maximumComputationTime          = std::numeric_limits<double>::max();
    // determinism for tests/benchmarks (not in the reference): stop after this many expansions
    size_t maxExpansions = std::numeric_limits<size_t>::max();
};

// src/algos/TPS_Astar.cpp + include/mpp/algos/TPS_Astar.h
class TPS_Astar
{
   public:
    TPS_Astar_Parameters                        params_;
    std::vector<std::shared_ptr<CostEvaluator>> costEvaluators_;

    // src/algos/Planner.cpp:15-29
    double cost_path_segment(const MoveEdge& edge) const
    {
        double c = edge.estimatedExecTime;
        for (const auto& ce : costEvaluators_)
        {
            ORC_ASSERT(ce);
            c += (*ce)(edge);
        }
        return c;
    }

   private:
    // TPS_Astar.h:90-170
    struct NodeCoords
    {
        int32_t                idxX = 0, idxY = 0;
        std::optional<int32_t> idxYaw;
        NodeCoords() = default;
        NodeCoords(int32_t ix, int32_t iy) : idxX(ix), idxY(iy) {}
        NodeCoords(int32_t ix, int32_t iy, int32_t iphi) : idxX(ix), idxY(iy), idxYaw(iphi) {}
        bool operator==(const NodeCoords& o) const
        {
            if (idxX != o.idxX || idxY != o.idxY) return false;
            if (idxYaw.has_value() != o.idxYaw.has_value()) return false;
            if (idxYaw.has_value() && idxYaw.value() != o.idxYaw.value()) return false;
            return true;
        }
        bool sameLocation(const NodeCoords& o) const
        {
            if (idxX != o.idxX || idxY != o.idxY) return false;
            if (idxYaw.has_value() && o.idxYaw.has_value()) return idxYaw.value() == o.idxYaw.value();
            return true;
        }
    };
    struct NodeCoordsHash
    {
        size_t operator()(const NodeCoords& x) const
        {
            size_t res = 17;
            res        = res * 31 + std::hash<int32_t>()(x.idxX);
            res        = res * 31 + std::hash<int32_t>()(x.idxY);
            if (x.idxYaw) res = res * 31 + std::hash<int32_t>()(x.idxYaw.value());
            return res;
        }
    };
    struct Node
    {
        std::optional<TNodeID>     id;
        SE2_KinState               state;
        double                     gScore = std::numeric_limits<double>::max();
        double                     fScore = std::numeric_limits<double>::max();
        std::optional<const Node*> cameFrom;
        bool                       pendingInOpenSet = false;
        bool                       visited          = false;
    };
    std::unordered_map<NodeCoords, Node, NodeCoordsHash> grid_;

    // TPS_Astar.h:224-230 (float arguments; [MRPT] wrapToPi<float>)
    int32_t x2idx(float x) const { return x / params_.grid_resolution_xy; }
    int32_t y2idx(float y) const { return y / params_.grid_resolution_xy; }
    int32_t phi2idx(float yaw) const
    {
        float      a       = yaw + static_cast<float>(M_PI);
        const bool was_neg = a < 0;
        a                  = std::fmod(a, static_cast<float>(2.0 * M_PI));
        if (was_neg) a += static_cast<float>(2.0 * M_PI);
        const float phi = a - static_cast<float>(M_PI);
        return phi / params_.grid_resolution_yaw;
    }
    NodeCoords nodeGridCoords(const TPose2D& p) const { return NodeCoords(x2idx(p.x), y2idx(p.y), phi2idx(p.phi)); }
    NodeCoords nodeGridCoordsPt(const TPose2D& p) const { return NodeCoords(x2idx(p.x), y2idx(p.y)); }

    // TPS_Astar.cpp:513-525
    Node& getOrCreateNodeByPose(const SE2_KinState& p, TNodeID& nextFreeId)
    {
        Node& n = grid_[nodeGridCoords(p.pose)];
        if (!n.id.has_value())
        {
            n.id    = nextFreeId++;
            n.state = p;
        }
        return n;
    }

    // TPS_Astar.cpp:477-511, 527-537. For a point goal `goal - from.pose` is read as the plain
    // vector difference TPoint2D - TPoint2D(from.pose) ([MRPT] implicit conversion; ambiguous
    // without the MRPT headers); for a pose goal it is the SE(2) "ominus".
    double heuristic(const SE2_KinState& from, const SE2orR2_KinState& goal) const
    {
        double dist, relX, relY;
        if (goal.isPoint)
        {
            relX = goal.pose.x - from.pose.x;
            relY = goal.pose.y - from.pose.y;
            dist = std::sqrt(square(from.pose.x - goal.pose.x) + square(from.pose.y - goal.pose.y));
        }
        else
        {
            const TPose2D relPose = goal.pose - from.pose;
            relX                  = relPose.x;
            relY                  = relPose.y;
            dist                  = relPose.norm() + params_.SE2_metricAngleWeight * std::abs(relPose.phi);
        }
        const double relNorm = std::sqrt(relX * relX + relY * relY);
        const double distHeading =
            (relNorm < 0.1) ? 0.0 : std::abs(angDistance(std::atan2(relY, relX), from.pose.phi));
        return dist + params_.heuristic_heading_weight * distHeading;
    }

    // TPS_Astar.h:278-290
    struct path_to_neighbor_t
    {
        std::optional<size_t>      ptgIndex;
        std::optional<int>         ptgTrajIndex;
        std::optional<uint32_t>    relTrgStep;
        std::optional<NavDynState> ptgDynState;
        double                     ptgTrimmableSpeed = 1.0;
        double                     ptgDist           = std::numeric_limits<double>::max();
        TPose2D                    relReconstrPose;
        NodeCoords                 neighborNodeCoords;
    };
    struct TPS_point
    {
        int      k;
        uint32_t step;
        double   speed;
    };
    using nodes_with_desired_speed_t = std::unordered_map<NodeCoords, double, NodeCoordsHash>;

    // TPS_Astar.cpp:538-826
    std::vector<path_to_neighbor_t> find_feasible_paths_to_neighbors(
        const Node& from, const PtgList& trs, const SE2orR2_KinState& goalState,
        const std::vector<const PointCloud*>& globalObstacles, double MAX_XY_OBSTACLES_CLIPPING_DIST,
        const nodes_with_desired_speed_t& nodesWithSpeed, const TPose2D& worldBboxMin, const TPose2D& worldBboxMax,
        PlannerOutput& stats)
    {
        const NodeCoords iGoalCoords =
            goalState.isPoint ? nodeGridCoordsPt(goalState.pose) : nodeGridCoords(goalState.pose);
        const TPose2D relGoal  = goalState.asSE2KinState().pose - from.state.pose;
        const double  halfCell = params_.grid_resolution_xy * 0.5;
        const PointCloud localObstacles =
            cached_local_obstacles(from.state.pose, globalObstacles, MAX_XY_OBSTACLES_CLIPPING_DIST);
        std::unordered_map<NodeCoords, path_to_neighbor_t, NodeCoordsHash> bestPaths;

        for (size_t ptgIdx = 0; ptgIdx < trs.size(); ptgIdx++)
        {
            auto&        ptg          = trs.at(ptgIdx);
            const bool   ptgTrimmable = ptg->isSpeedTrimmable();
            const double ptg_dt       = ptg->getPathStepDuration();
            {
                NavDynState ds;
                (ds.curVelLocal = from.state.vel).rotate(-from.state.pose.phi);
                ds.relTarget = relGoal;
                if (const auto it = nodesWithSpeed.find(iGoalCoords); it != nodesWithSpeed.end())
                    ds.targetRelSpeed = it->second;
                else
                    ds.targetRelSpeed = 0;
                ptg->updateNavDynamicState(ds);
            }
            std::set<int>          trajIdxsToConsider;
            std::vector<TPS_point> tpsPointsToConsider;
            std::set<size_t>       targetTpsPointIndx;
            ORC_ASSERT(params_.max_ptg_trajectories_to_explore >= 2);
            for (size_t i = 0; i < params_.max_ptg_trajectories_to_explore; i++)
            {
                // integer division inside mrpt::round (TPS_Astar.cpp:619-623)
                const int trjIdx =
                    mrpt_round(i * (ptg->numPaths - 1) / (params_.max_ptg_trajectories_to_explore - 1));
                trajIdxsToConsider.insert(trjIdx);
            }
            std::vector<double> speedsToConsider;
            if (ptgTrimmable)
            {
                ORC_ASSERT(params_.max_ptg_speeds_to_explore >= 1);
                const double speedStep = 1.0 / params_.max_ptg_speeds_to_explore;
                for (double s = speedStep; s < 1.001; s += speedStep) speedsToConsider.push_back(s);
            }
            else
                speedsToConsider.push_back(1.0);
            {
                int          relTrg_k       = 0;
                double       relTrg_d       = 0;
                const double queryTolerance = params_.grid_resolution_xy;
                if (ptg->inverseMap_WS2TP(relGoal.x, relGoal.y, relTrg_k, relTrg_d, queryTolerance))
                {
                    uint32_t relTrg_step = 0;
                    // NOTE: the reference passes the *normalised* distance here (TPS_Astar.cpp:658)
                    if (ptg->getPathStepForDist(relTrg_k, relTrg_d, relTrg_step))
                    {
                        for (auto speed : speedsToConsider)
                        {
                            tpsPointsToConsider.push_back(TPS_point{relTrg_k, relTrg_step, speed});
                            targetTpsPointIndx.insert(tpsPointsToConsider.size() - 1);
                        }
                    }
                    trajIdxsToConsider.insert(relTrg_k);
                }
            }
            for (const auto speed : speedsToConsider)
                for (const auto trjIdx : trajIdxsToConsider)
                    for (double t : params_.ptg_sample_timestamps)
                    {
                        const uint32_t trjStep  = mrpt_round(t / ptg_dt);
                        const auto     maxSteps = ptg->getPathStepCount(trjIdx);
                        ORC_ASSERT(maxSteps >= 1);
                        if (trjStep >= maxSteps) continue;
                        tpsPointsToConsider.push_back(TPS_point{trjIdx, trjStep, speed});
                    }

            std::unordered_set<NodeCoords, NodeCoordsHash> goalNodeCoords;
            for (size_t tpsPtIdx = 0; tpsPtIdx < tpsPointsToConsider.size(); tpsPtIdx++)
            {
                const auto& tpsPt = tpsPointsToConsider[tpsPtIdx];
                if (tpsPt.step == 0) continue;
                if (ptgTrimmable && tpsPt.speed != ptg->trimmableSpeed) ptg->trimmableSpeed = tpsPt.speed;
                const double  relTrgDist      = ptg->getPathDist(tpsPt.k, tpsPt.step);
                const TPose2D relReconstrPose = ptg->getPathPose(tpsPt.k, tpsPt.step);
                const TPose2D absPose         = from.state.pose + relReconstrPose;
                if (absPose.x < worldBboxMin.x || absPose.y < worldBboxMin.y || absPose.phi < worldBboxMin.phi)
                    continue;
                if (absPose.x > worldBboxMax.x - halfCell || absPose.y > worldBboxMax.y - halfCell ||
                    absPose.phi > worldBboxMax.phi)
                    continue;
                const NodeCoords nc = nodeGridCoords(absPose);
                stats.nTpsPointsEvaluated++;
                const double freeDistance = tp_obstacles_single_path(tpsPt.k, localObstacles, *ptg);
                if (relTrgDist >= freeDistance) continue;
                if (targetTpsPointIndx.count(tpsPtIdx) != 0)
                    goalNodeCoords.insert(nc);
                else if (goalNodeCoords.count(nc) != 0)
                    continue;
                auto& path = bestPaths[nc];
                if (relTrgDist < path.ptgDist)
                {
                    path.ptgDist            = relTrgDist;
                    path.ptgIndex           = ptgIdx;
                    path.ptgTrajIndex       = tpsPt.k;
                    path.relReconstrPose    = relReconstrPose;
                    path.relTrgStep         = tpsPt.step;
                    path.neighborNodeCoords = nc;
                    path.ptgDynState        = ptg->dynState;
                }
            }
        }
        std::vector<path_to_neighbor_t> neighbors;
        for (const auto& kv : bestPaths)
        {
            if (!kv.second.ptgIndex.has_value()) continue;
            neighbors.emplace_back(kv.second);
        }
        return neighbors;
    }

   public:
    // TPS_Astar.cpp:86-475
    PlannerOutput plan(const PlannerInput& in)
    {
        const auto t0 = std::chrono::steady_clock::now();
        ORC_ASSERT(!in.ptgs.empty());
        ORC_ASSERT(in.worldBboxMin != in.worldBboxMax);
        ORC_ASSERT(within_bbox(in.stateStart.pose, in.worldBboxMax, in.worldBboxMin));
        if (in.stateGoal.isPoint)
            ORC_ASSERT(within_bbox_pt(in.stateGoal.pose, in.worldBboxMax, in.worldBboxMin));
        else
            ORC_ASSERT(within_bbox(in.stateGoal.pose, in.worldBboxMax, in.worldBboxMin));

        PlannerOutput po;
        auto&         tree        = po.motionTree;
        double        MAX_XY_DIST = 0;
        for (const auto& ptg : in.ptgs) MAX_XY_DIST = std::max(MAX_XY_DIST, ptg->refDistance);
        ORC_ASSERT(MAX_XY_DIST > 0);

        std::vector<PointCloud> clipped;
        clipped.reserve(in.obstacles.size());
        for (const auto& os : in.obstacles)
            clipped.push_back(
                apply_clipping_box(os, in.worldBboxMin.x, in.worldBboxMin.y, in.worldBboxMax.x, in.worldBboxMax.y));
        std::vector<const PointCloud*> obstaclePoints;
        for (const auto& c : clipped) obstaclePoints.push_back(&c);

        tree.edges_to_children.clear();
        grid_.clear();
        std::multimap<double, Node*> openSet;
        TNodeID                      nextFreeId = 0;
        {
            Node& n = getOrCreateNodeByPose(in.stateStart, nextFreeId);
            n.state = in.stateStart;
            tree.root = n.id.value();
            tree.insert_root_node(tree.root, n.state);
            n.gScore           = 0;
            n.fScore           = heuristic(n.state, in.stateGoal);
            n.pendingInOpenSet = true;
            openSet.insert({n.fScore, &n});
        }
        Node* nodeGoal = &getOrCreateNodeByPose(in.stateGoal.asSE2KinState(), nextFreeId);
        po.goalNodeId  = nodeGoal->id.value();
        const NodeCoords goalCellIndices =
            in.stateGoal.isPoint ? nodeGridCoordsPt(in.stateGoal.pose) : nodeGridCoords(in.stateGoal.pose);
        nodes_with_desired_speed_t nodesWithDesiredSpeed;
        nodesWithDesiredSpeed[goalCellIndices] = 0;

        while (!openSet.empty())
        {
            Node& current = *openSet.begin()->second;
            if (const auto curNodeGridIdx = nodeGridCoords(current.state.pose);
                curNodeGridIdx.sameLocation(goalCellIndices))
            {
                if (in.stateGoal.isPoint)
                {
                    current.state.pose.x    = in.stateGoal.pose.x;
                    current.state.pose.y    = in.stateGoal.pose.y;
                    nodeGoal                = &current;
                    po.goalNodeId           = nodeGoal->id.value();
                    po.bestNodeId           = po.goalNodeId;
                    po.bestNodeIdCostToGoal = 0;
                }
                else
                {
                    current.state.pose      = in.stateGoal.pose;
                    po.bestNodeIdCostToGoal = 0;
                }
                static_cast<SE2_KinState&>(tree.nodes.at(*current.id)).pose = current.state.pose;
                break;
            }
            current.pendingInOpenSet = false;
            current.visited          = true;
            openSet.erase(openSet.begin());
            po.nExpanded++;

            const auto neighbors = find_feasible_paths_to_neighbors(
                current, in.ptgs, in.stateGoal, obstaclePoints, MAX_XY_DIST, nodesWithDesiredSpeed, in.worldBboxMin,
                in.worldBboxMax, po);

            for (const auto& edge : neighbors)
            {
                auto& ptg = *in.ptgs.at(edge.ptgIndex.value());
                ptg.updateNavDynamicState(edge.ptgDynState.value());
                if (ptg.isSpeedTrimmable()) ptg.trimmableSpeed = edge.ptgTrimmableSpeed;
                const uint32_t ptg_step        = edge.relTrgStep.value();
                const TPose2D& reconstrRelPose = edge.relReconstrPose;
                const TPose2D  q_i             = current.state.pose + reconstrRelPose;
                const TTwist2D relTwist        = ptg.getPathTwist(edge.ptgTrajIndex.value(), ptg_step);
                SE2_KinState   x_i;
                x_i.pose = q_i;
                (x_i.vel = relTwist).rotate(current.state.pose.phi);
                if (q_i.x < in.worldBboxMin.x || q_i.y < in.worldBboxMin.y || q_i.phi < in.worldBboxMin.phi) continue;
                if (q_i.x > in.worldBboxMax.x || q_i.y > in.worldBboxMax.y || q_i.phi > in.worldBboxMax.phi) continue;
                Node& neighborNode = getOrCreateNodeByPose(x_i, nextFreeId);
                if (neighborNode.visited) continue;

                MoveEdge newEdge;
                newEdge.parentId             = current.id.value();
                newEdge.ptgDist              = edge.ptgDist;
                newEdge.ptgIndex             = static_cast<int8_t>(edge.ptgIndex.value());
                newEdge.ptgPathIndex         = static_cast<int16_t>(edge.ptgTrajIndex.value());
                newEdge.ptgTrimmableSpeed    = edge.ptgTrimmableSpeed;
                newEdge.ptgFinalGoalRelSpeed = 0;
                newEdge.ptgFinalRelativeGoal = in.stateGoal.asSE2KinState().pose - current.state.pose;
                newEdge.stateFrom            = current.state;
                newEdge.stateTo              = x_i;
                newEdge.ptgStep              = ptg_step;
                edge_interpolated_path(newEdge, in.ptgs, reconstrRelPose, ptg_step, params_.pathInterpolatedSegments);
                newEdge.cost = cost_path_segment(newEdge);
                po.nEdgesCosted++;
                ORC_ASSERT(newEdge.cost > .0);

                const double tentative_gScore = current.gScore + newEdge.cost;
                if (tentative_gScore >= neighborNode.gScore) continue;
                const bool hasToRewire = neighborNode.cameFrom.has_value();
                neighborNode.cameFrom  = &current;
                neighborNode.gScore    = tentative_gScore;
                const double costToGoal = heuristic(neighborNode.state, in.stateGoal);
                neighborNode.fScore     = tentative_gScore + costToGoal;
                if (!neighborNode.pendingInOpenSet)
                {
                    neighborNode.pendingInOpenSet = true;
                    openSet.insert({neighborNode.fScore, &neighborNode});
                }
                neighborNode.state = x_i;
                if (hasToRewire)
                    tree.rewire_node_parent(neighborNode.id.value(), newEdge);
                else
                    tree.insert_node_and_edge(newEdge.parentId, neighborNode.id.value(), neighborNode.state, newEdge);
                if (costToGoal < po.bestNodeIdCostToGoal)
                {
                    po.bestNodeIdCostToGoal = costToGoal;
                    po.bestNodeId           = neighborNode.id.value();
                }
            }
            const double tNow = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (tNow > params_.maximumComputationTime) break;
            if (po.nExpanded >= params_.maxExpansions) break;
        }
        po.success = po.goalNodeId == po.bestNodeId;
        if (po.bestNodeId) po.pathCost = tree.nodes.at(*po.bestNodeId).cost;
        po.computationTime = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return po;
    }
};

// src/algos/refine_trajectory.cpp:14-89
inline void refine_trajectory(
    const std::list<MotionTree::Node>& inPath, std::list<MoveEdge*>& edgesToRefine, const PtgList& ptgInfo,
    const double ptg_tolerance_dist = 0.10)
{
    const size_t nEdges = edgesToRefine.size();
    ORC_ASSERT(inPath.size() == nEdges + 1);
    auto itPath = inPath.begin();
    auto itEdge = edgesToRefine.begin();
    for (size_t i = 0; i < nEdges; i++, itPath++, itEdge++)
    {
        MoveEdge& edge = **itEdge;
        auto&     ptg  = ptgInfo.at(edge.ptgIndex);
        ptg->updateNavDynamicState(edge.getPTGDynState());
        if (ptg->isSpeedTrimmable()) ptg->trimmableSpeed = edge.ptgTrimmableSpeed;
        auto itNextPath = itPath;
        itNextPath++;
        const TPose2D deltaNodes = itNextPath->pose - itPath->pose;
        if (deltaNodes.x == 0 && deltaNodes.y == 0) continue;
        int        newK        = -1;
        double     newNormDist = 0;
        const bool ok = ptg->inverseMap_WS2TP(deltaNodes.x, deltaNodes.y, newK, newNormDist, ptg_tolerance_dist);
        if (!ok) continue;  // the reference only prints a warning here
        const double newDist    = newNormDist * ptg->refDistance;
        uint32_t     newPtgStep = 0;
        ptg->getPathStepForDist(newK, newDist, newPtgStep);
        edge.ptgPathIndex = newK;
        edge.ptgDist      = newDist;
        edge.ptgStep      = newPtgStep;
        edge_interpolated_path(edge, ptgInfo, deltaNodes, newPtgStep);
    }
}

// src/algos/trajectories.cpp:15-69
struct trajectory_state_t
{
    SE2_KinState state;
    size_t       ptgIndex = 0;
    int          ptgPathIndex = 0;
    uint32_t     ptgStep = 0;
};
using trajectory_t = std::map<double, trajectory_state_t>;

inline trajectory_t plan_to_trajectory(
    const std::list<MoveEdge*>& planEdges, const PtgList& ptgInfo, const double samplePeriod = 50e-3)
{
    ORC_ASSERT(samplePeriod > 0.);
    trajectory_t out;
    double       currentEdgeStartTime = .0;
    TPose2D      endOfLastEdge(0, 0, 0);
    for (const auto& edge : planEdges)
    {
        auto&        ptg    = ptgInfo.at(edge->ptgIndex);
        const double ptg_dt = ptg->getPathStepDuration();
        ORC_ASSERT(ptg_dt > 0.);
        ptg->updateNavDynamicState(edge->getPTGDynState());
        if (ptg->isSpeedTrimmable()) ptg->trimmableSpeed = edge->ptgTrimmableSpeed;
        uint32_t   ptgFinalStep = 0;
        const bool ok           = ptg->getPathStepForDist(edge->ptgPathIndex, edge->ptgDist, ptgFinalStep);
        ORC_ASSERT(ok);
        const uint32_t stepIncr = std::max<uint32_t>(1, mrpt_round(samplePeriod / ptg_dt));
        for (uint32_t step = 0;; step += stepIncr)
        {
            if (step > ptgFinalStep) step = ptgFinalStep;
            trajectory_state_t& ts = out[currentEdgeStartTime + step * ptg_dt];
            ts.state.pose          = endOfLastEdge + ptg->getPathPose(edge->ptgPathIndex, step);
            ts.state.vel           = ptg->getPathTwist(edge->ptgPathIndex, step);
            ts.ptgIndex            = edge->ptgIndex;
            ts.ptgPathIndex        = edge->ptgPathIndex;
            ts.ptgStep             = step;
            if (step == ptgFinalStep) break;
        }
        currentEdgeStartTime += ptgFinalStep * ptg_dt;
        endOfLastEdge = edge->stateTo.pose - planEdges.front()->stateFrom.pose;
    }
    return out;
}

// src/algos/trajectories.cpp:71-106
inline std::string trajectory_to_txt(const trajectory_t& traj)
{
    std::string s;
    char        buf[512];
    snprintf(
        buf, sizeof(buf), "%% %15s  %15s %15s %15s  %15s %15s %15s %15s %15s %15s\n", "Time [s]", "x_global [m]",
        "y_global [m]", "phi [rad]", "vx_local [m]", "vy_local[m]", "omega [rad/s]", "PTG_index", "PTG_traj_index",
        "PTG_step");
    s += buf;
    for (const auto& kv : traj)
    {
        const auto& ts = kv.second;
        snprintf(
            buf, sizeof(buf), "%15.03f %15.03f %15.03f %15.03f  %15.03f %15.03f %15.03f   %15u %15u %15u\n", kv.first,
            ts.state.pose.x, ts.state.pose.y, ts.state.pose.phi, ts.state.vel.vx, ts.state.vel.vy, ts.state.vel.omega,
            static_cast<unsigned int>(ts.ptgIndex), static_cast<unsigned int>(ts.ptgPathIndex),
            static_cast<unsigned int>(ts.ptgStep));
        s += buf;
    }
    return s;
}

}  // namespace orc
