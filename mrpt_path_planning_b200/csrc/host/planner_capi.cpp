// C ABI over the mpp::TPS_Astar mirror (see include/mppb.h, section "planner").
#include <thread>
#include <cstring>
#include <memory>
#include <string>

#include "../../../include/mppb.h"
#include "host_internal.h"

struct mppb_planner
{
    mppb_ctx*                                    ctx = nullptr;
    std::unique_ptr<mpp::TPS_Astar>              planner;
    mpp::TrajectoriesAndRobotShape               trs;
    std::vector<mpp::ObstacleSource::Ptr>        obstacles;
    std::vector<mpp::PointCloud::Ptr>            clouds;  // the clouds behind `obstacles`, same order
    std::vector<mpp::PlannerInput>               inputs;
    std::vector<mpp::PlannerOutput>              outputs;
    std::vector<mpp::MotionPrimitivesTreeSE2::path_t>          paths;
    std::vector<mpp::MotionPrimitivesTreeSE2::edge_sequence_t> pathEdges;
    std::vector<char>                            havePath;  // paths[i] / pathEdges[i] are built (on first use)
    std::string                                  lastError, text;
};

namespace
{
thread_local std::string g_plannerErr;
template <class F>
int guarded(mppb_planner* p, F&& fn)
{
    try
    {
        fn();
        return MPPB_OK;
    }
    catch (const std::exception& e)
    {
        if (p)
            p->lastError = e.what();
        else
            g_plannerErr = e.what();
        return MPPB_ERR_STATE;
    }
}
void require(bool cond, const char* msg)
{
    if (!cond) throw std::runtime_error(msg);
}
// motion tree and backtracked path of one query, built on first use
void ensure_path(mppb_planner* p, size_t query)
{
    require(query < p->outputs.size() && query < p->paths.size() && query < p->havePath.size(), "query index out of range");
    if (p->havePath[query]) return;
    mpp::PlannerOutput& o = p->outputs[query];
    o.materializeTree();
    if (o.bestNodeId) std::tie(p->paths[query], p->pathEdges[query]) = o.motionTree.backtrack_path(*o.bestNodeId);
    p->havePath[query] = 1;
}
}  // namespace

namespace mpp
{
int selftest_worker_pool(int n_threads, size_t count, int rounds);  // planner.cpp
int selftest_open_set(uint64_t seed, size_t ops, int distinct_keys);  // planner.cpp
int selftest_lattice_index(uint64_t seed, size_t ops, int range);     // planner.cpp
int selftest_cell_order(uint64_t seed, size_t trials, int max_keys);  // planner.cpp
}

extern "C" {

int mppb_selftest_worker_pool(int n_threads, size_t count, int rounds)
{
    try
    {
        return mpp::selftest_worker_pool(n_threads, count, rounds);
    }
    catch (const std::exception&)
    {
        return -2;
    }
}

int mppb_selftest_open_set(uint64_t seed, size_t ops, int distinct_keys)
{
    try
    {
        return mpp::selftest_open_set(seed, ops, distinct_keys);
    }
    catch (const std::exception&)
    {
        return -2;
    }
}

int mppb_selftest_lattice_index(uint64_t seed, size_t ops, int range)
{
    try
    {
        return mpp::selftest_lattice_index(seed, ops, range);
    }
    catch (const std::exception&)
    {
        return -2;
    }
}

int mppb_selftest_cell_order(uint64_t seed, size_t trials, int max_keys)
{
    try
    {
        return mpp::selftest_cell_order(seed, trials, max_keys);
    }
    catch (const std::exception&)
    {
        return -2;
    }
}

int mppb_planner_create(mppb_ctx* ctx, mppb_planner** out)
{
    if (!ctx || !out) return MPPB_ERR_INVALID_ARG;
    *out = nullptr;
    return guarded(nullptr, [&]() {
        auto* p    = new mppb_planner();
        p->ctx     = ctx;
        p->planner = std::make_unique<mpp::TPS_Astar>(ctx);
        *out       = p;
    });
}
void        mppb_planner_destroy(mppb_planner* p) { delete p; }
const char* mppb_planner_last_error(const mppb_planner* p) { return p ? p->lastError.c_str() : g_plannerErr.c_str(); }

int mppb_planner_set_params(mppb_planner* p, const mppb_astar_params* a)
{
    if (!p || !a) return MPPB_ERR_INVALID_ARG;
    return guarded(p, [&]() {
        require(a->num_timestamps > 0 && a->ptg_sample_timestamps, "ptg_sample_timestamps is required");
        auto& q                           = p->planner->params_;
        q.grid_resolution_xy              = a->grid_resolution_xy;
        q.grid_resolution_yaw             = a->grid_resolution_yaw;
        q.SE2_metricAngleWeight           = a->SE2_metricAngleWeight;
        q.heuristic_heading_weight        = a->heuristic_heading_weight;
        q.max_ptg_trajectories_to_explore = a->max_ptg_trajectories_to_explore;
        q.max_ptg_speeds_to_explore       = a->max_ptg_speeds_to_explore;
        q.pathInterpolatedSegments        = a->path_interpolated_segments;
        q.maximumComputationTime =
            a->maximum_computation_time > 0 ? a->maximum_computation_time : std::numeric_limits<double>::max();
        q.maxExpansions = a->max_expansions ? a->max_expansions : std::numeric_limits<size_t>::max();
        q.ptg_sample_timestamps.assign(a->ptg_sample_timestamps, a->ptg_sample_timestamps + a->num_timestamps);
    });
}

int mppb_planner_add_ptg(mppb_planner* p, mppb_ptg* ptg)
{
    if (!p || !ptg) return MPPB_ERR_INVALID_ARG;
    return guarded(p, [&]() {
        auto list = p->trs.ptgs;
        list.emplace_back(ptg->obj.get(), [](mpp::ptg_t*) {});  // borrowed: the PTG object outlives the planner
        p->trs.initFromPtgs(std::move(list));
    });
}

int mppb_planner_add_obstacles(mppb_planner* p, const float* xs, const float* ys, size_t n)
{
    if (!p || (n && (!xs || !ys))) return MPPB_ERR_INVALID_ARG;
    return guarded(p, [&]() {
        auto pc = std::make_shared<mpp::PointCloud>();
        pc->xs.assign(xs, xs + n);
        pc->ys.assign(ys, ys + n);
        p->obstacles.push_back(mpp::ObstacleSource::FromStaticPointcloud(pc));
        p->clouds.push_back(pc);
    });
}

int mppb_planner_update_obstacles(mppb_planner* p, size_t source, const float* xs, const float* ys, size_t n)
{
    if (!p || (n && (!xs || !ys))) return MPPB_ERR_INVALID_ARG;
    return guarded(p, [&]() {
        require(source < p->clouds.size(), "obstacle source index out of range");
        // in place: with an unchanged point count the buffers keep their addresses, as a fixed-size scan does
        mpp::PointCloud& pc = *p->clouds[source];
        pc.xs.assign(xs, xs + n);
        pc.ys.assign(ys, ys + n);
    });
}

int mppb_planner_add_costmap(mppb_planner* p, const float* xs, const float* ys, size_t n, double resolution,
                             double preferred_clearance, double max_cost, int use_average, double max_radius_from_robot,
                             const double* robot_pose_xyphi)
{
    if (!p || !xs || !ys) return MPPB_ERR_INVALID_ARG;
    return guarded(p, [&]() {
        mpp::PointCloud pc;
        pc.xs.assign(xs, xs + n);
        pc.ys.assign(ys, ys + n);
        mpp::CostEvaluatorCostMap::Parameters q;
        q.resolution                 = resolution;
        q.preferredClearanceDistance = preferred_clearance;
        q.maxCost                    = max_cost;
        q.useAverageOfPath           = use_average != 0;
        q.maxRadiusFromRobot         = max_radius_from_robot;
        std::optional<mpp::TPose2D> pose;
        if (robot_pose_xyphi) pose = mpp::TPose2D(robot_pose_xyphi[0], robot_pose_xyphi[1], robot_pose_xyphi[2]);
        p->planner->costEvaluators_.push_back(mpp::CostEvaluatorCostMap::FromStaticPointObstacles(p->ctx, pc, q, pose));
    });
}

int mppb_planner_costmap_cells(mppb_planner* p, int evaluator, double* info5, double* cells)
{
    if (!p) return MPPB_ERR_INVALID_ARG;
    return guarded(p, [&]() {
        require(evaluator >= 0 && evaluator < static_cast<int>(p->planner->costEvaluators_.size()), "bad evaluator index");
        auto* cm = dynamic_cast<mpp::CostEvaluatorCostMap*>(p->planner->costEvaluators_[evaluator].get());
        require(cm != nullptr, "evaluator is not a costmap");
        if (info5)
        {
            info5[0] = cm->x_min;
            info5[1] = cm->y_min;
            info5[2] = cm->params_.resolution;
            info5[3] = cm->size_x;
            info5[4] = cm->size_y;
        }
        if (cells) std::memcpy(cells, cm->cells.data(), cm->cells.size() * sizeof(double));
    });
}

int mppb_planner_add_waypoints(mppb_planner* p, const double* xs, const double* ys, size_t n, double influence_radius,
                               double cost_scale, int use_average)
{
    if (!p || (n && (!xs || !ys))) return MPPB_ERR_INVALID_ARG;
    return guarded(p, [&]() {
        auto ev                             = std::make_shared<mpp::CostEvaluatorPreferredWaypoint>(p->ctx);
        ev->params_.waypointInfluenceRadius = influence_radius;
        ev->params_.costScale               = cost_scale;
        ev->params_.useAverageOfPath        = use_average != 0;
        std::vector<mpp::TPoint2D> pts(n);
        for (size_t i = 0; i < n; i++) pts[i] = {xs[i], ys[i]};
        ev->setPreferredWaypoints(pts);
        p->planner->costEvaluators_.push_back(ev);
    });
}

int mppb_planner_plan(mppb_planner* p, size_t n_queries, const mppb_plan_query* qs, int n_threads, mppb_plan_result* res)
{
    if (!p || !qs || !res || n_queries == 0) return MPPB_ERR_INVALID_ARG;
    return guarded(p, [&]() {
        p->inputs.assign(n_queries, mpp::PlannerInput());
        for (size_t i = 0; i < n_queries; i++)
        {
            const mppb_plan_query& q  = qs[i];
            mpp::PlannerInput&     in = p->inputs[i];
            in.stateStart.pose        = mpp::TPose2D(q.start[0], q.start[1], q.start[2]);
            in.stateStart.vel         = mpp::TTwist2D(q.start[3], q.start[4], q.start[5]);
            if (q.goal_is_point)
                in.stateGoal.state = mpp::PoseOrPoint(mpp::TPoint2D{q.goal[0], q.goal[1]});
            else
                in.stateGoal.state = mpp::PoseOrPoint(mpp::TPose2D(q.goal[0], q.goal[1], q.goal[2]));
            in.stateGoal.vel = mpp::TTwist2D(q.goal[3], q.goal[4], q.goal[5]);
            in.worldBboxMin  = mpp::TPose2D(q.bbox_min[0], q.bbox_min[1], q.bbox_min[2]);
            in.worldBboxMax  = mpp::TPose2D(q.bbox_max[0], q.bbox_max[1], q.bbox_max[2]);
            in.obstacles     = p->obstacles;
            in.ptgs          = p->trs;
        }
        // results of the previous call go first: paths/pathEdges hold raw pointers into the trees of `outputs`, and
        // a plan that throws must not leave them dangling behind a non-empty `outputs`
        p->paths.clear();
        p->pathEdges.clear();
        p->havePath.clear();
        if (p->outputs.size() > 16)
        {
            // the previous batch's motion trees are hundreds of thousands of small nodes: drop them on several
            // threads instead of one by one when the vector is reassigned (0.3 s for 512 queries otherwise)
            const size_t             nT = std::min<size_t>(16, std::max(1u, std::thread::hardware_concurrency()));
            std::vector<std::thread> ts;
            for (size_t t = 0; t < nT; t++)
                ts.emplace_back([p, t, nT]() {
                    for (size_t i = t; i < p->outputs.size(); i += nT) p->outputs[i] = mpp::PlannerOutput();
                });
            for (auto& t : ts) t.join();
        }
        p->outputs.clear();
        if (n_queries == 1)
            p->outputs.push_back(p->planner->plan(p->inputs[0]));
        else  // a batch: motion trees and paths are built per query when somebody asks (ensure_path below)
            p->outputs = p->planner->plan_batch(p->inputs, n_threads, false);
        p->paths.assign(n_queries, {});
        p->pathEdges.assign(n_queries, {});
        p->havePath.assign(n_queries, 0);
        for (size_t i = 0; i < n_queries; i++)
        {
            const mpp::PlannerOutput& o = p->outputs[i];
            mppb_plan_result& r = res[i];
            r.success           = o.success ? 1 : 0;
            r.path_cost         = o.pathCost;
            r.tree_nodes        = o.treeNodes;
            r.tree_parents      = o.treeParents;
            r.nodes_expanded    = o.nodesExpanded;
            r.tps_points        = o.tpsPointsEvaluated;
            r.edges_costed      = o.edgesCosted;
            r.best_node         = o.bestNodeId ? static_cast<int64_t>(*o.bestNodeId) : -1;
            r.goal_node         = o.goalNodeId ? static_cast<int64_t>(*o.goalNodeId) : -1;
            r.computation_time  = o.computationTime;
            r.best_cost_to_goal = o.bestNodeIdCostToGoal;
            r.path_len          = static_cast<uint32_t>(o.pathLength);
        }
    });
}

int mppb_planner_plan_multi(mppb_planner** planners, int n_planners, size_t n_queries, const mppb_plan_query* qs,
                            int n_threads_per_planner, mppb_plan_result* res)
{
    if (!planners || n_planners < 1 || !qs || !res || n_queries == 0) return MPPB_ERR_INVALID_ARG;
    for (int i = 0; i < n_planners; i++)
        if (!planners[i]) return MPPB_ERR_INVALID_ARG;
    // query q goes to planner q % n_planners as its local query q / n_planners; every planner plans its share on a
    // thread of its own (its context's device, its own worker pool), all at once
    const size_t                               P = static_cast<size_t>(n_planners);
    std::vector<std::vector<mppb_plan_query>>  share(P);
    std::vector<std::vector<mppb_plan_result>> out(P);
    for (size_t q = 0; q < n_queries; q++) share[q % P].push_back(qs[q]);
    std::vector<int>         rc(P, MPPB_OK);
    std::vector<std::thread> ts;
    for (size_t i = 0; i < P; i++)
    {
        out[i].resize(share[i].size());
        if (share[i].empty()) continue;
        ts.emplace_back([&, i]() {
            rc[i] = mppb_planner_plan(planners[i], share[i].size(), share[i].data(), n_threads_per_planner, out[i].data());
        });
    }
    for (auto& t : ts) t.join();
    for (size_t i = 0; i < P; i++)
        if (rc[i] != MPPB_OK) return rc[i];  // (the failing planner holds the message: mppb_planner_last_error)
    for (size_t q = 0; q < n_queries; q++) res[q] = out[q % P][q / P];
    return MPPB_OK;
}

int mppb_planner_refine(mppb_planner* p, size_t query)
{
    if (!p) return MPPB_ERR_INVALID_ARG;
    return guarded(p, [&]() {
        ensure_path(p, query);
        mpp::refine_trajectory(p->paths[query], p->pathEdges[query], p->inputs[query].ptgs);
    });
}

int mppb_planner_path(mppb_planner* p, size_t query, double* nodes, double* edges)
{
    if (!p) return MPPB_ERR_INVALID_ARG;
    return guarded(p, [&]() {
        ensure_path(p, query);
        size_t i = 0;
        if (nodes)
            for (const auto& n : p->paths[query])
            {
                double* o = nodes + 8 * i++;
                o[0]      = static_cast<double>(n.nodeID_);
                o[1]      = n.pose.x;
                o[2]      = n.pose.y;
                o[3]      = n.pose.phi;
                o[4]      = n.vel.vx;
                o[5]      = n.vel.vy;
                o[6]      = n.vel.omega;
                o[7]      = n.cost_;
            }
        i = 0;
        if (edges)
            for (const mpp::MoveEdgeSE2_TPS* e : p->pathEdges[query])
            {
                double* o = edges + 8 * i++;
                o[0]      = e->ptgIndex;
                o[1]      = e->ptgPathIndex;
                o[2]      = e->ptgStep;
                o[3]      = e->ptgDist;
                o[4]      = e->cost;
                o[5]      = e->estimatedExecTime;
                o[6]      = e->ptgTrimmableSpeed;
                o[7]      = static_cast<double>(e->interpolatedPath.size());
            }
    });
}

int64_t mppb_planner_tree(mppb_planner* p, size_t query, double* out, int64_t cap)
{
    if (!p || query >= p->outputs.size()) return -1;
    int64_t n  = 0;
    const int rc = guarded(p, [&]() {
        ensure_path(p, query);
        const auto& tree = p->outputs[query].motionTree;
        for (const auto& kv : tree.nodes())
        {
            if (out && n < cap)
            {
                double* o = out + 7 * n;
                o[0]      = static_cast<double>(kv.first);
                o[1]      = kv.second.parentID_ ? static_cast<double>(*kv.second.parentID_) : -1;
                o[2]      = kv.second.pose.x;
                o[3]      = kv.second.pose.y;
                o[4]      = kv.second.pose.phi;
                o[5]      = kv.second.cost_;
                o[6]      = kv.second.parentID_ ? tree.edge_to_parent(kv.first).cost : 0;
            }
            n++;
        }
    });
    return rc == MPPB_OK ? n : -1;
}

int64_t mppb_planner_trajectory_txt(mppb_planner* p, size_t query, double sample_period, char* buf, int64_t cap)
{
    if (!p || query >= p->outputs.size() || query >= p->pathEdges.size()) return -1;
    const int rc = guarded(p, [&]() {
        ensure_path(p, query);
        auto traj = mpp::plan_to_trajectory(p->pathEdges[query], p->inputs[query].ptgs, sample_period);
        // the trajectory is relative to the start pose; path-planner-cli re-bases it (path-planner-cli.cpp:396-399)
        const mpp::TPose2D start = p->outputs[query].originalInput.stateStart.pose;
        for (auto& kv : traj) kv.second.state.pose = start + kv.second.state.pose;
        p->text = mpp::trajectory_as_text(traj);
    });
    if (rc != MPPB_OK) return -1;
    if (buf && cap > 0)
    {
        const int64_t m = std::min<int64_t>(cap - 1, static_cast<int64_t>(p->text.size()));
        std::memcpy(buf, p->text.data(), m);
        buf[m] = 0;
    }
    return static_cast<int64_t>(p->text.size());
}

int mppb_planner_gpu_stats(const mppb_planner* p, double* out6)
{
    if (!p || !out6) return MPPB_ERR_INVALID_ARG;
    for (int i = 0; i < 7; i++) out6[6 + i] = p->planner->lastGpuStats.phaseSeconds[i];
    const auto& s = p->planner->lastGpuStats;
    out6[0]       = static_cast<double>(s.collisionBatches);
    out6[1]       = static_cast<double>(s.costBatches);
    out6[2]       = static_cast<double>(s.nodesEvaluated);
    out6[3]       = static_cast<double>(s.holoCandidates);
    out6[4]       = static_cast<double>(s.edgesCosted);
    out6[5]       = s.gpuSeconds;
    out6[13]      = static_cast<double>(s.holoNearTies);
    return MPPB_OK;
}

}  // extern "C"
