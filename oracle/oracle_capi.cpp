// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product.
//
// extern "C" surface of the CPU oracle so that tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs can call it through ctypes.
// Pinned against oracle/_ref (the reference's own sources behind the same interface, oracle/ref/ref_capi.cpp);
// the MRPT layer is assumed.
#include <chrono>
#include <cstring>

#include <atomic>
#include <mutex>
#include <thread>

#include "orc_planner.hpp"

using namespace orc;

static thread_local std::string g_err;

#define ORC_TRY try {
#define ORC_CATCH(ret)              \
    }                               \
    catch (const std::exception& e) \
    {                               \
        g_err = e.what();           \
        return ret;                 \
    }

static int hw_threads()
{
    const unsigned n = std::thread::hardware_concurrency();
    return n ? static_cast<int>(n) : 1;
}

// run fn(b) for b in [0,B) on nthreads host threads (dynamic schedule); rethrows the first error
template <class F>
static void parallel_for(long B, int nthreads, F fn)
{
    if (nthreads <= 0) nthreads = hw_threads();
    if (nthreads > B) nthreads = B > 0 ? static_cast<int>(B) : 1;
    std::atomic<long> next{0};
    std::mutex        mtx;
    std::string       err;
    auto              worker = [&]() {
        for (;;)
        {
            const long b = next.fetch_add(1);
            if (b >= B) break;
            try
            {
                fn(b);
            }
            catch (const std::exception& e)
            {
                std::lock_guard<std::mutex> lk(mtx);
                err = e.what();
            }
        }
    };
    if (nthreads == 1)
        worker();
    else
    {
        std::vector<std::thread> th;
        for (int i = 0; i < nthreads; i++) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
    if (!err.empty()) throw std::runtime_error(err);
}

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

int orc_num_threads() { return hw_threads(); }

// ---- DiffDrive_C PTG ----------------------------------------------------------------
void* orc_diffdrive_new(
    int numPaths, double refDist, double res, double vmax, double wmax_rad, double K, double turnRadiusRef,
    const double* shape_x, const double* shape_y, int nverts)
{
    ORC_TRY
    auto* p                   = new DiffDriveC();
    p->numPaths               = numPaths;
    p->refDistance            = refDist;
    p->resolution             = res;
    p->V_MAX                  = vmax;
    p->W_MAX                  = wmax_rad;
    p->K                      = K;
    p->turningRadiusReference = turnRadiusRef;
    p->setRobotShape(std::vector<double>(shape_x, shape_x + nverts), std::vector<double>(shape_y, shape_y + nverts));
    p->initialize();
    return static_cast<PTG*>(p);
    ORC_CATCH(nullptr)
}
void orc_ptg_free(void* h) { delete static_cast<PTG*>(h); }
int  orc_ptg_num_paths(void* h) { return static_cast<PTG*>(h)->numPaths; }
double orc_ptg_ref_distance(void* h) { return static_cast<PTG*>(h)->refDistance; }
double orc_ptg_max_radius(void* h) { return static_cast<PTG*>(h)->getMaxRobotRadius(); }
double orc_ptg_step_duration(void* h) { return static_cast<PTG*>(h)->getPathStepDuration(); }
long   orc_ptg_step_count(void* h, int k)
{
    ORC_TRY
    return static_cast<long>(static_cast<PTG*>(h)->getPathStepCount(k));
    ORC_CATCH(-1)
}
double orc_ptg_init_dist(void* h, int k)
{
    double v = 0;
    static_cast<PTG*>(h)->initTPObstacleSingle(k, v);
    return v;
}
int orc_ptg_path_pose(void* h, int k, unsigned step, double* xyphi_dist)
{
    ORC_TRY
    auto*         p  = static_cast<PTG*>(h);
    const TPose2D q  = p->getPathPose(k, step);
    xyphi_dist[0]    = q.x;
    xyphi_dist[1]    = q.y;
    xyphi_dist[2]    = q.phi;
    xyphi_dist[3]    = p->getPathDist(k, step);
    return 0;
    ORC_CATCH(-1)
}
int orc_ptg_inverse_map(void* h, double x, double y, double tol, int* k, double* d)
{
    ORC_TRY
    return static_cast<PTG*>(h)->inverseMap_WS2TP(x, y, *k, *d, tol) ? 1 : 0;
    ORC_CATCH(-1)
}
int orc_ptg_step_for_dist(void* h, int k, double dist, unsigned* step)
{
    ORC_TRY
    uint32_t s = 0;
    const bool ok = static_cast<PTG*>(h)->getPathStepForDist(k, dist, s);
    *step         = s;
    return ok ? 1 : 0;
    ORC_CATCH(-1)
}
// trajectory dump: out is [steps][7] floats (x,y,phi,t,dist,v,w)
int orc_diffdrive_path(void* h, int k, float* out)
{
    ORC_TRY
    auto* p = dynamic_cast<DiffDriveC*>(static_cast<PTG*>(h));
    ORC_ASSERT(p);
    const auto& tr = p->trajectory.at(k);
    for (size_t i = 0; i < tr.size(); i++)
    {
        const float v[7] = {tr[i].x, tr[i].y, tr[i].phi, tr[i].t, tr[i].dist, tr[i].v, tr[i].w};
        std::memcpy(out + 7 * i, v, sizeof(v));
    }
    return 0;
    ORC_CATCH(-1)
}
int orc_diffdrive_grid_info(void* h, double* xmin, double* ymin, double* res, int* nx, int* ny, long* nentries)
{
    ORC_TRY
    auto* p = dynamic_cast<DiffDriveC*>(static_cast<PTG*>(h));
    ORC_ASSERT(p);
    const auto& g = p->collisionGrid;
    *xmin         = g.x_min;
    *ymin         = g.y_min;
    *res          = g.res;
    *nx           = g.size_x;
    *ny           = g.size_y;
    long n        = 0;
    for (const auto& c : g.cells) n += static_cast<long>(c.size());
    *nentries = n;
    return 0;
    ORC_CATCH(-1)
}
// CSR export in cell-major order, entries in the reference's list (insertion) order.
int orc_diffdrive_grid_csr(void* h, uint32_t* offsets, uint16_t* ks, float* ds)
{
    ORC_TRY
    auto* p = dynamic_cast<DiffDriveC*>(static_cast<PTG*>(h));
    ORC_ASSERT(p);
    const auto& g = p->collisionGrid;
    uint32_t    n = 0;
    for (size_t c = 0; c < g.cells.size(); c++)
    {
        offsets[c] = n;
        for (const auto& e : g.cells[c])
        {
            ks[n] = e.first;
            ds[n] = e.second;
            n++;
        }
    }
    offsets[g.cells.size()] = n;
    return 0;
    ORC_CATCH(-1)
}
int orc_ptg_point_inside_shape(void* h, double x, double y)
{
    return static_cast<PTG*>(h)->isPointInsideRobotShape(x, y) ? 1 : 0;
}

// ---- stage (a): clip + transform ------------------------------------------------------
// returns number of local points written (out_x/out_y sized n)
long orc_local_obstacles(
    const float* xs, const float* ys, size_t n, const double* pose_xyphi, double clip, float* out_x, float* out_y)
{
    ORC_TRY
    PointCloud in;
    in.xs.assign(xs, xs + n);
    in.ys.assign(ys, ys + n);
    PointCloud out;
    transform_pc_square_clipping(in, CPose2D(TPose2D(pose_xyphi[0], pose_xyphi[1], pose_xyphi[2])), clip, out);
    std::memcpy(out_x, out.xs.data(), out.size() * sizeof(float));
    std::memcpy(out_y, out.ys.data(), out.size() * sizeof(float));
    return static_cast<long>(out.size());
    ORC_CATCH(-1)
}

// ---- stage (a)+(b) for a batch of node poses ------------------------------------------
// out[B][sum_p K_p] doubles = free TP-distance per (node, ptg, k) exactly as
// find_feasible_paths_to_neighbors obtains it (TPS_Astar.cpp:561-562, 744-745).
// mode 0: reference call pattern (tp_obstacles_single_path once per k);
// mode 1: the PTG's all-k virtual (DiffDriveCollisionGridBased.cpp:813-824), same values.
// Returns elapsed wall seconds of the evaluation loop (excluding input copies), <0 on error.
double orc_eval_nodes(
    void** ptgs, int nptg, const float* xs, const float* ys, size_t n, const double* poses_xyphi, size_t B,
    double clip, double* out, int mode, int nthreads)
{
    ORC_TRY
    PointCloud cloud;
    cloud.xs.assign(xs, xs + n);
    cloud.ys.assign(ys, ys + n);
    std::vector<const PointCloud*> clouds{&cloud};
    size_t                         sumK = 0;
    for (int p = 0; p < nptg; p++) sumK += static_cast<PTG*>(ptgs[p])->numPaths;
    const auto t0 = std::chrono::steady_clock::now();
    parallel_for(static_cast<long>(B), nthreads, [&](long b) {
        {
            const TPose2D    pose(poses_xyphi[3 * b], poses_xyphi[3 * b + 1], poses_xyphi[3 * b + 2]);
            const PointCloud local = cached_local_obstacles(pose, clouds, clip);
            double*          o     = out + b * sumK;
            for (int p = 0; p < nptg; p++)
            {
                const PTG* ptg = static_cast<PTG*>(ptgs[p]);
                const int  K   = ptg->numPaths;
                auto*      dd  = dynamic_cast<const DiffDriveC*>(ptg);
                if (mode == 1 && dd)
                {
                    for (int k = 0; k < K; k++) ptg->initTPObstacleSingle(k, o[k]);
                    for (size_t i = 0; i < local.size(); i++)
                        dd->updateTPObstacleAll(local.xs[i], local.ys[i], o);
                }
                else
                {
                    for (int k = 0; k < K; k++) o[k] = tp_obstacles_single_path(k, local, *ptg);
                }
                o += K;
            }
        }
    });
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
    ORC_CATCH(-1.0)
}


// ---- solve_poly* (second-opinion tests against numpy.roots) ----------------------------
int orc_solve_poly4(double a, double b, double c, double d, double* x) { return poly::solve_poly4(x, a, b, c, d); }
int orc_solve_poly3(double a, double b, double c, double* x) { return poly::solve_poly3(x, a, b, c); }
int orc_solve_poly2(double a, double b, double c, double* r)
{
    return poly::solve_poly2(a, b, c, r[0], r[1]);
}

// ---- HolonomicBlend PTG ------------------------------------------------------------------
void* orc_holo_new(
    int numPaths, double refDist, double T_ramp_max, double vmax, double wmax_rad, double robotRadius,
    const char* expr_V, const char* expr_W)
{
    ORC_TRY
    auto* p        = new HolonomicBlend();
    p->numPaths    = numPaths;
    p->refDistance = refDist;
    p->T_ramp_max  = T_ramp_max;
    p->V_MAX       = vmax;
    p->W_MAX       = wmax_rad;
    p->robotRadius = robotRadius;
    if (expr_V && *expr_V) p->expr_V = expr_V;
    if (expr_W && *expr_W) p->expr_W = expr_W;
    p->initialize();
    return static_cast<PTG*>(p);
    ORC_CATCH(nullptr)
}
// dyn = vxLocal, vyLocal, omega, relTargetX, relTargetY, relTargetPhi, targetRelSpeed
int orc_ptg_update_dyn_state(void* h, const double* dyn, int force)
{
    ORC_TRY
    NavDynState ds;
    ds.curVelLocal.vx    = dyn[0];
    ds.curVelLocal.vy    = dyn[1];
    ds.curVelLocal.omega = dyn[2];
    ds.relTarget         = TPose2D(dyn[3], dyn[4], dyn[5]);
    ds.targetRelSpeed    = dyn[6];
    static_cast<PTG*>(h)->updateNavDynamicState(ds, force != 0);
    return 0;
    ORC_CATCH(-1)
}
void orc_ptg_set_trim_speed(void* h, double s) { static_cast<PTG*>(h)->trimmableSpeed = s; }
int  orc_ptg_path_twist(void* h, int k, unsigned step, double* out3)
{
    ORC_TRY
    const TTwist2D t = static_cast<PTG*>(h)->getPathTwist(k, step);
    out3[0]          = t.vx;
    out3[1]          = t.vy;
    out3[2]          = t.omega;
    return 0;
    ORC_CATCH(-1)
}
// one updateTPObstacleSingle call: returns the updated tp_obstacle_k
double orc_ptg_update_tpobs_single(void* h, double ox, double oy, int k, double tp_obstacle_k)
{
    static_cast<PTG*>(h)->updateTPObstacleSingle(ox, oy, k, tp_obstacle_k);
    return tp_obstacle_k;
}

// Stage (a)+(b) for explicit candidates (node, k, speed) of ONE trimmable PTG, as
// find_feasible_paths_to_neighbors evaluates them (TPS_Astar.cpp:587-611, 721-745).
// nodes: [B][6] = x, y, phi, vx, vy, omega (global velocity); relTargets: [B][3].
// out_raw[m]  = min over in-window points of the collision distance (+inf if none)
// out_free[m] = tp_obstacles_single_path() with the init value taken at the candidate's speed
//               on a fresh step-count cache.
double orc_eval_candidates(
    void* h, const float* xs, const float* ys, size_t n, const double* nodes, const double* relTargets, size_t B,
    const int* cand_node, const int* cand_k, const double* cand_speed, size_t M, double clip, double* out_raw,
    double* out_free)
{
    ORC_TRY
    PTG*       ptg = static_cast<PTG*>(h);
    PointCloud cloud;
    cloud.xs.assign(xs, xs + n);
    cloud.ys.assign(ys, ys + n);
    std::vector<const PointCloud*> clouds{&cloud};
    std::vector<PointCloud>        locals(B);
    const auto                     t0 = std::chrono::steady_clock::now();
    for (size_t b = 0; b < B; b++)
        locals[b] = cached_local_obstacles(TPose2D(nodes[6 * b], nodes[6 * b + 1], nodes[6 * b + 2]), clouds, clip);
    for (size_t m = 0; m < M; m++)
    {
        const size_t b = cand_node[m];
        NavDynState  ds;
        ds.curVelLocal.vx    = nodes[6 * b + 3];
        ds.curVelLocal.vy    = nodes[6 * b + 4];
        ds.curVelLocal.omega = nodes[6 * b + 5];
        ds.curVelLocal.rotate(-nodes[6 * b + 2]);
        ds.relTarget      = TPose2D(relTargets[3 * b], relTargets[3 * b + 1], relTargets[3 * b + 2]);
        ds.targetRelSpeed = 0;
        ptg->trimmableSpeed = cand_speed[m];
        ptg->updateNavDynamicState(ds, true /*fresh caches*/);
        ptg->trimmableSpeed = cand_speed[m];
        double raw = std::numeric_limits<double>::infinity();
        const PointCloud& L = locals[b];
        for (size_t i = 0; i < L.size(); i++) ptg->updateTPObstacleSingle(L.xs[i], L.ys[i], cand_k[m], raw);
        out_raw[m] = raw;
        if (out_free) out_free[m] = tp_obstacles_single_path(cand_k[m], L, *ptg);
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    ORC_CATCH(-1.0)
}

// ---- cost evaluators ---------------------------------------------------------------------
struct OrcCostmap
{
    std::shared_ptr<CostEvaluatorCostMap> cm;
};
void* orc_costmap_new(
    const float* xs, const float* ys, size_t n, double resolution, double clearance, double maxCost, int useAverage,
    double maxRadiusFromRobot, const double* robot_xyphi)
{
    ORC_TRY
    PointCloud c;
    c.xs.assign(xs, xs + n);
    c.ys.assign(ys, ys + n);
    CostEvaluatorCostMap::Parameters p;
    p.resolution                 = resolution;
    p.preferredClearanceDistance = clearance;
    p.maxCost                    = maxCost;
    p.useAverageOfPath           = useAverage != 0;
    p.maxRadiusFromRobot         = maxRadiusFromRobot;
    std::optional<TPose2D> pose;
    if (robot_xyphi) pose = TPose2D(robot_xyphi[0], robot_xyphi[1], robot_xyphi[2]);
    auto* h = new OrcCostmap();
    h->cm   = CostEvaluatorCostMap::FromStaticPointObstacles(c, p, pose);
    return h;
    ORC_CATCH(nullptr)
}
void orc_costmap_free(void* h) { delete static_cast<OrcCostmap*>(h); }
// info: xmin, ymin, res, nx, ny
void orc_costmap_info(void* h, double* info)
{
    const auto& g = static_cast<OrcCostmap*>(h)->cm->costmap;
    info[0]       = g.x_min;
    info[1]       = g.y_min;
    info[2]       = g.res;
    info[3]       = g.size_x;
    info[4]       = g.size_y;
}
void orc_costmap_cells(void* h, double* out)
{
    const auto& g = static_cast<OrcCostmap*>(h)->cm->costmap;
    std::memcpy(out, g.cells.data(), g.cells.size() * sizeof(double));
}
double orc_costmap_eval_pose(void* h, double x, double y)
{
    return static_cast<OrcCostmap*>(h)->cm->eval_single_pose(TPose2D(x, y, 0));
}

struct OrcWaypoints
{
    std::shared_ptr<CostEvaluatorPreferredWaypoint> ev;
};
void* orc_waypoints_new(const double* xs, const double* ys, size_t n, double radius, double scale, int useAverage)
{
    ORC_TRY
    auto* h                             = new OrcWaypoints();
    h->ev                               = std::make_shared<CostEvaluatorPreferredWaypoint>();
    h->ev->params.waypointInfluenceRadius = radius;
    h->ev->params.costScale               = scale;
    h->ev->params.useAverageOfPath        = useAverage != 0;
    std::vector<TPoint2D> pts(n);
    for (size_t i = 0; i < n; i++)
    {
        pts[i].x = xs[i];
        pts[i].y = ys[i];
    }
    h->ev->setPreferredWaypoints(pts);
    return h;
    ORC_CATCH(nullptr)
}
void   orc_waypoints_free(void* h) { delete static_cast<OrcWaypoints*>(h); }
double orc_waypoints_eval_pose(void* h, double x, double y)
{
    return static_cast<OrcWaypoints*>(h)->ev->eval_single_pose(TPose2D(x, y, 0));
}

// Per-edge evaluator costs for explicit sample lists (stage c): edge e has samples
// [sample_offsets[e], sample_offsets[e+1]) of relative poses (x,y) composed with from_xyphi[e].
// kind 0 = costmap handle, 1 = waypoint handle. Mirrors reduce_over_path.
int orc_eval_edge_costs(
    void* evaluator, int kind, size_t nEdges, const double* from_xyphi, const uint32_t* sample_offsets,
    const double* rel_xy, double* out_cost)
{
    ORC_TRY
    for (size_t e = 0; e < nEdges; e++)
    {
        MoveEdge edge;
        edge.stateFrom.pose = TPose2D(from_xyphi[3 * e], from_xyphi[3 * e + 1], from_xyphi[3 * e + 2]);
        for (uint32_t s = sample_offsets[e]; s < sample_offsets[e + 1]; s++)
            edge.interpolatedPath[static_cast<double>(s)] = TPose2D(rel_xy[2 * s], rel_xy[2 * s + 1], 0);
        if (kind == 0)
            out_cost[e] = (*static_cast<OrcCostmap*>(evaluator)->cm)(edge);
        else
            out_cost[e] = (*static_cast<OrcWaypoints*>(evaluator)->ev)(edge);
    }
    return 0;
    ORC_CATCH(-1)
}

// ---- TPS_Astar ----------------------------------------------------------------------------
struct OrcPlanner
{
    TPS_Astar     planner;
    PlannerInput  in;
    PlannerOutput out;
    bool          planned = false;
    std::list<MotionTree::Node> path;
    std::list<MoveEdge*>        pathEdges;
    std::string                 text;
};
void* orc_planner_new() { return new OrcPlanner(); }
void  orc_planner_free(void* h) { delete static_cast<OrcPlanner*>(h); }
// p: grid_xy, grid_yaw_rad, SE2_metricAngleWeight, heuristic_heading_weight, max_traj, max_speeds,
//    pathInterpolatedSegments, maximumComputationTime (<=0: unlimited), maxExpansions (<=0: unlimited)
void orc_planner_set_params(void* h, const double* p, const double* timestamps, int nTimestamps)
{
    auto& q                           = static_cast<OrcPlanner*>(h)->planner.params_;
    q.grid_resolution_xy              = p[0];
    q.grid_resolution_yaw             = p[1];
    q.SE2_metricAngleWeight           = p[2];
    q.heuristic_heading_weight        = p[3];
    q.max_ptg_trajectories_to_explore = static_cast<uint32_t>(p[4]);
    q.max_ptg_speeds_to_explore       = static_cast<uint32_t>(p[5]);
    q.pathInterpolatedSegments        = static_cast<size_t>(p[6]);
    q.maximumComputationTime          = p[7] > 0 ? p[7] : std::numeric_limits<double>::max();
    q.maxExpansions = p[8] > 0 ? static_cast<size_t>(p[8]) : std::numeric_limits<size_t>::max();
    q.ptg_sample_timestamps.assign(timestamps, timestamps + nTimestamps);
}
void orc_planner_add_ptg(void* h, void* ptg)
{
    static_cast<OrcPlanner*>(h)->in.ptgs.emplace_back(static_cast<PTG*>(ptg), [](PTG*) {});
}
void orc_planner_add_obstacles(void* h, const float* xs, const float* ys, size_t n)
{
    PointCloud c;
    c.xs.assign(xs, xs + n);
    c.ys.assign(ys, ys + n);
    static_cast<OrcPlanner*>(h)->in.obstacles.push_back(std::move(c));
}
void orc_planner_add_costmap(void* h, void* costmap)
{
    static_cast<OrcPlanner*>(h)->planner.costEvaluators_.push_back(static_cast<OrcCostmap*>(costmap)->cm);
}
void orc_planner_add_waypoints(void* h, void* wps)
{
    static_cast<OrcPlanner*>(h)->planner.costEvaluators_.push_back(static_cast<OrcWaypoints*>(wps)->ev);
}
// start: x y phi vx vy w ; goal: x y phi vx vy w (+ goalIsPoint) ; bbox min/max: x y phi
// result: success, pathCost, nTreeNodes, nTreeParents, nExpanded, nTpsPoints, nEdgesCosted, bestNodeId,
//         goalNodeId, computationTime, bestNodeIdCostToGoal
int orc_planner_plan(
    void* h, const double* start, const double* goal, int goalIsPoint, const double* bboxMin, const double* bboxMax,
    double* result)
{
    ORC_TRY
    auto& P                 = *static_cast<OrcPlanner*>(h);
    P.in.stateStart.pose    = TPose2D(start[0], start[1], start[2]);
    P.in.stateStart.vel.vx  = start[3];
    P.in.stateStart.vel.vy  = start[4];
    P.in.stateStart.vel.omega = start[5];
    P.in.stateGoal.isPoint  = goalIsPoint != 0;
    P.in.stateGoal.pose     = TPose2D(goal[0], goal[1], goalIsPoint ? 0.0 : goal[2]);
    P.in.stateGoal.vel.vx   = goal[3];
    P.in.stateGoal.vel.vy   = goal[4];
    P.in.stateGoal.vel.omega = goal[5];
    P.in.worldBboxMin       = TPose2D(bboxMin[0], bboxMin[1], bboxMin[2]);
    P.in.worldBboxMax       = TPose2D(bboxMax[0], bboxMax[1], bboxMax[2]);
    P.out                   = P.planner.plan(P.in);
    P.planned               = true;
    P.path.clear();
    P.pathEdges.clear();
    if (P.out.bestNodeId) P.out.motionTree.backtrack_path(*P.out.bestNodeId, P.path, P.pathEdges);
    result[0]  = P.out.success ? 1 : 0;
    result[1]  = P.out.pathCost;
    result[2]  = static_cast<double>(P.out.motionTree.nodes.size());
    result[3]  = static_cast<double>(P.out.motionTree.edges_to_children.size());
    result[4]  = static_cast<double>(P.out.nExpanded);
    result[5]  = static_cast<double>(P.out.nTpsPointsEvaluated);
    result[6]  = static_cast<double>(P.out.nEdgesCosted);
    result[7]  = P.out.bestNodeId ? static_cast<double>(*P.out.bestNodeId) : -1;
    result[8]  = P.out.goalNodeId ? static_cast<double>(*P.out.goalNodeId) : -1;
    result[9]  = P.out.computationTime;
    result[10] = P.out.bestNodeIdCostToGoal;
    return 0;
    ORC_CATCH(-1)
}
int orc_planner_refine(void* h)
{
    ORC_TRY
    auto& P = *static_cast<OrcPlanner*>(h);
    refine_trajectory(P.path, P.pathEdges, P.in.ptgs);
    return 0;
    ORC_CATCH(-1)
}
long orc_planner_path_len(void* h) { return static_cast<long>(static_cast<OrcPlanner*>(h)->path.size()); }
// nodes: [n][8] = id, x, y, phi, vx, vy, omega, cost ; edges: [n-1][8] = ptgIndex, k, step, ptgDist, cost,
// estimatedExecTime, trimSpeed, nInterpolated
int orc_planner_path(void* h, double* nodes, double* edges)
{
    ORC_TRY
    auto&  P = *static_cast<OrcPlanner*>(h);
    size_t i = 0;
    for (const auto& n : P.path)
    {
        double* o = nodes + 8 * i++;
        o[0]      = static_cast<double>(n.nodeID);
        o[1]      = n.pose.x;
        o[2]      = n.pose.y;
        o[3]      = n.pose.phi;
        o[4]      = n.vel.vx;
        o[5]      = n.vel.vy;
        o[6]      = n.vel.omega;
        o[7]      = n.cost;
    }
    i = 0;
    for (const MoveEdge* e : P.pathEdges)
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
    return 0;
    ORC_CATCH(-1)
}
// whole motion tree: [n][7] = id, parent(-1 root), x, y, phi, cost, edgeCost(to parent)
long orc_planner_tree(void* h, double* out, long cap)
{
    ORC_TRY
    auto& P = *static_cast<OrcPlanner*>(h);
    long  i = 0;
    for (const auto& kv : P.out.motionTree.nodes)
    {
        if (out && i < cap)
        {
            double* o = out + 7 * i;
            o[0]      = static_cast<double>(kv.first);
            o[1]      = kv.second.parentID ? static_cast<double>(*kv.second.parentID) : -1;
            o[2]      = kv.second.pose.x;
            o[3]      = kv.second.pose.y;
            o[4]      = kv.second.pose.phi;
            o[5]      = kv.second.cost;
            o[6]      = kv.second.parentID ? P.out.motionTree.edge_to_parent(kv.first).cost : 0;
        }
        i++;
    }
    return i;
    ORC_CATCH(-1)
}
// plan_to_trajectory + save_to_txt text; returns needed length
long orc_planner_trajectory_txt(void* h, double samplePeriod, char* buf, long cap)
{
    ORC_TRY
    auto& P = *static_cast<OrcPlanner*>(h);
    auto  traj = plan_to_trajectory(P.pathEdges, P.in.ptgs, samplePeriod);
    // trajectories are relative to the start pose; the CLI re-bases them (path-planner-cli.cpp:396-399)
    for (auto& kv : traj) kv.second.state.pose = P.in.stateStart.pose + kv.second.state.pose;
    P.text = trajectory_to_txt(traj);
    if (buf && cap > 0)
    {
        const long m = std::min<long>(cap - 1, static_cast<long>(P.text.size()));
        std::memcpy(buf, P.text.data(), m);
        buf[m] = 0;
    }
    return static_cast<long>(P.text.size());
    ORC_CATCH(-1)
}

}  // extern "C"
