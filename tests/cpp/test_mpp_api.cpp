// Compiled test of the C++ surface of the mpp:: mirror (csrc/host/mpp.h) — the interface a user of the reference
// programs against: a user CostEvaluator subclass (evaluated on the host per edge, Planner.cpp:22-26), the progress
// callback (TPS_Astar.cpp:432-451), maximumComputationTime returning the best node so far (:424-429, :468), the class
// registry (registerClasses.cpp:16-37) and the single-call seams of the free functions. Needs a GPU.
//   test_mpp_api <ptgs.ini> <obstacles.txt>
#include <cmath>
#include <cstdio>
#include <iostream>

#include "mpp.h"

#define CHECK(cond)                                                                 \
    do {                                                                            \
        if (!(cond))                                                                \
        {                                                                           \
            std::fprintf(stderr, "CHECK failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
            return 1;                                                               \
        }                                                                           \
    } while (0)

// a user plug-in: constant cost per edge; must see complete edges (interpolated path, states, PTG fields)
struct ConstantCost : mpp::CostEvaluator
{
    mutable size_t calls = 0, badEdges = 0;
    double         operator()(const mpp::MoveEdgeSE2_TPS& e) const override
    {
        calls++;
        if (e.interpolatedPath.size() < 2 || e.ptgIndex < 0 || e.ptgPathIndex < 0 || !(e.estimatedExecTime > 0)) badEdges++;
        return 0.25;
    }
};

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    mppb_ctx* ctx = nullptr;
    if (mppb_ctx_create(0, nullptr, &ctx) != MPPB_OK)
    {
        std::fprintf(stderr, "no context: %s\n", mppb_last_global_error());
        return 3;
    }
    auto cloud = std::make_shared<mpp::PointCloud>();
    CHECK(cloud->load2D_from_text_file(argv[2]) && cloud->size() > 10);
    mpp::PlannerInput pi;
    pi.stateStart.pose = mpp::TPose2D(0.5, 0, 0);
    pi.stateGoal.state = mpp::PoseOrPoint(mpp::TPose2D(4, 2.5, 45.0 * M_PI / 180.0));
    pi.worldBboxMin    = mpp::TPose2D(-1.5, -2.0, -M_PI);
    pi.worldBboxMax    = mpp::TPose2D(6.0, 4.5, M_PI);
    pi.obstacles.push_back(mpp::ObstacleSource::FromStaticPointcloud(cloud));
    pi.ptgs.initFromConfigFile(mpp::ConfigFile(argv[1]), "SelfDriving", ctx);
    CHECK(pi.ptgs.initialized() && pi.ptgs.ptgs.size() == 2);

    // --- class registry: the library's planner by name, unknown names answer null ---
    auto& reg = mpp::ClassRegistry::instance();
    CHECK(reg.createPlanner("no::SuchPlanner", ctx) == nullptr);
    std::shared_ptr<mpp::Planner> byName = reg.createPlanner("mpp::TPS_Astar", ctx);
    CHECK(byName && dynamic_cast<mpp::TPS_Astar*>(byName.get()));

    mpp::TPS_Astar planner(ctx);
    planner.params_.grid_resolution_xy              = 0.40;
    planner.params_.grid_resolution_yaw             = 25.0 * M_PI / 180.0;
    planner.params_.SE2_metricAngleWeight           = 0.1;
    planner.params_.max_ptg_trajectories_to_explore = 13;
    planner.params_.ptg_sample_timestamps           = {0.5, 1.5, 4.0};

    // --- (1) no evaluators: the device-side expansion path ---
    const mpp::PlannerOutput plain = planner.plan(pi);
    CHECK(plain.success && plain.nodesExpanded > 3);

    // the reference's profiler sections exist under their names (TPS_Astar.cpp:89-833)
    {
        const auto& sec = planner.profiler_().sections();
        for (const char* name : {"plan", "plan.iter", "find_feasible", "find_feasible.loop1", "find_feasible.ptgUpdateDyn",
                                 "find_feasible.loop2", "find_feasible.tp_obstacles_single", "find_feasible.finalFill",
                                 "cached_local_obstacles"})
            CHECK(sec.count(name) == 1);
        CHECK(sec.at("plan").calls == 1 && sec.at("plan.iter").calls == plain.nodesExpanded && sec.at("plan").seconds > 0);
        CHECK(sec.at("find_feasible.tp_obstacles_single").calls == plain.tpsPointsEvaluated);
        CHECK(!planner.profiler_().getStatsAsText().empty());
    }

    // --- (2) a user evaluator: called on the host for every tentative edge, its value enters every edge cost ---
    auto user = std::make_shared<ConstantCost>();
    planner.costEvaluators_.push_back(user);
    const mpp::PlannerOutput withUser = planner.plan(pi);
    CHECK(withUser.success);
    CHECK(user->calls == withUser.edgesCosted && user->calls > 10 && user->badEdges == 0);
    {
        auto [path, edges] = withUser.motionTree.backtrack_path(*withUser.bestNodeId);
        double cost = 0;
        for (const mpp::MoveEdgeSE2_TPS* e : edges)
        {
            CHECK(e->cost == e->estimatedExecTime + 0.25);  // cost_path_segment: execution time + evaluators
            cost = cost + e->cost;
        }
        CHECK(cost == withUser.pathCost && edges.size() + 1 == path.size());
    }

    // --- (3) progress callback: period 0 -> after every expansion that has a best node ---
    size_t callbacks = 0, badCallbacks = 0;
    planner.progressCallbackCallPeriod_ = 0.0;
    planner.progressCallback_           = [&](const mpp::ProgressCallbackData& pcd) {
        callbacks++;
        if (!pcd.tree || !pcd.bestFinalNode || pcd.bestPath.empty() || !pcd.originalPlanInput || !pcd.costEvaluators ||
            pcd.tree->nodes().at(*pcd.bestFinalNode).cost_ != pcd.bestCostFromStart || !(pcd.bestCostToGoal >= 0))
            badCallbacks++;
    };
    const mpp::PlannerOutput withCb = planner.plan(pi);
    CHECK(withCb.success && callbacks >= 3 && badCallbacks == 0);
    CHECK(withCb.pathCost == withUser.pathCost);
    // the same on the device-expansion path (no host-plugin evaluator)
    planner.costEvaluators_.clear();
    callbacks = 0;
    const mpp::PlannerOutput plainCb = planner.plan(pi);
    CHECK(plainCb.success && callbacks >= 3 && badCallbacks == 0 && plainCb.pathCost == plain.pathCost);
    CHECK(plainCb.motionTree.nodes().size() == plain.motionTree.nodes().size());
    planner.progressCallback_ = nullptr;

    // --- (4) time budget exhausted: not an error, the best node so far is the answer ---
    planner.params_.maximumComputationTime = 1e-9;
    const mpp::PlannerOutput timedOut      = planner.plan(pi);
    CHECK(!timedOut.success && timedOut.nodesExpanded == 1 && timedOut.bestNodeId.has_value());
    CHECK(timedOut.motionTree.nodes().count(*timedOut.bestNodeId) == 1);
    planner.params_.maximumComputationTime = 1e9;

    // --- (5) single-call seams: clip + transform in input order, then one path's free distance over the local points,
    //         against the planner-independent batched entry point on the same pose ---
    const mpp::TPose2D pose(1.2, 0.7, 0.6);
    mpp::PointCloud    local;
    mpp::transform_pc_square_clipping(ctx, *cloud, pose, 5.0, local, false);
    {
        size_t m = 0;
        const double c = std::cos(pose.phi), s = std::sin(pose.phi);
        for (size_t i = 0; i < cloud->size(); i++)
        {
            const double gx = cloud->xs[i], gy = cloud->ys[i];
            if (std::abs(gx - pose.x) > 5.0 || std::abs(gy - pose.y) > 5.0) continue;
            const double ix = -pose.x * c - pose.y * s, iy = pose.x * s - pose.y * c;
            CHECK(m < local.size());
            CHECK(local.xs[m] == static_cast<float>((ix + gx * c) + gy * s));
            CHECK(local.ys[m] == static_cast<float>((iy - gx * s) + gy * c));
            m++;
        }
        CHECK(m == local.size() && m > 10);
    }
    {
        mppb_ctx* ctx2 = nullptr;
        CHECK(mppb_ctx_create(0, nullptr, &ctx2) == MPPB_OK);
        const mpp::ptg_t& ptg = *pi.ptgs.ptgs[0];
        auto*             dd  = dynamic_cast<const mpp::DiffDrive_C*>(&ptg);
        CHECK(dd);
        const mppb_ptg_grid_desc d = dd->gridDesc();
        CHECK(mppb_set_ptg_grid(ctx2, 0, &d) == MPPB_OK);
        CHECK(mppb_set_obstacles(ctx2, cloud->xs.data(), cloud->ys.data(), cloud->size(), 0.0, 0) == MPPB_OK);
        std::vector<float> row(static_cast<size_t>(d.num_paths));
        const double       p3[3] = {pose.x, pose.y, pose.phi};
        CHECK(mppb_eval_nodes(ctx2, 1, p3, 5.0, row.data()) == MPPB_OK);
        size_t finite = 0;
        for (int k = 0; k < d.num_paths; k += 7)
        {
            const double viaSeam  = mpp::tp_obstacles_single_path(ctx, k, local, ptg);
            const double viaBatch = std::min(ptg.initTPObstacleSingle(static_cast<uint16_t>(k)), static_cast<double>(row[static_cast<size_t>(k)]));
            CHECK(viaSeam == viaBatch);
            finite += viaSeam < ptg.getRefDistance();
        }
        CHECK(finite > 0);
        mppb_ctx_destroy(ctx2);
    }
    std::printf("OK user_evaluator_calls=%zu callbacks=%zu expansions=%zu\n", user->calls, callbacks, plain.nodesExpanded);
    mppb_ctx_destroy(ctx);
    return 0;
}
