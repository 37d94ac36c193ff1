// ORACLE / REFERENCE BUILD — TEST INFRASTRUCTURE ONLY. Not part of the shipped product.
//
// The same extern "C" surface as oracle/oracle_capi.cpp (orc_* symbols, same argument meaning),
// implemented on top of the REFERENCE'S OWN classes and functions: mpp::TPS_Astar,
// mpp::ptg::DiffDrive_C, mpp::ptg::HolonomicBlend, mpp::CostEvaluatorCostMap,
// mpp::CostEvaluatorPreferredWaypoint, mpp::transform_pc_square_clipping,
// mpp::tp_obstacles_single_path, mpp::edge_interpolated_path, mpp::refine_trajectory,
// mpp::plan_to_trajectory, mpp::save_to_txt — whose translation units are compiled unmodified from
// /root/reference (oracle/Makefile, target `ref`) against the MRPT stand-in oracle/ref/mrpt_shim.h.
// Nothing here restates reference logic; it only marshals arguments. This file is compiled with
// -fno-access-control so that the tests can read private tables (collision grid, trajectories,
// costmap cells) and call private members (cached_local_obstacles, eval_single_pose).
#include <mpp/algos/CostEvaluatorCostMap.h>
#include <mpp/algos/CostEvaluatorPreferredWaypoint.h>
#include <mpp/algos/TPS_Astar.h>
#include <mpp/algos/edge_interpolated_path.h>
#include <mpp/algos/refine_trajectory.h>
#include <mpp/algos/render_tree.h>
#include <mpp/algos/tp_obstacles_single_path.h>
#include <mpp/algos/trajectories.h>
#include <mpp/algos/transform_pc_square_clipping.h>
#include <mpp/ptgs/DiffDrive_C.h>
#include <mpp/ptgs/HolonomicBlend.h>
#include <unistd.h>

#include <atomic>
#include <thread>

using mrpt::math::TPoint2D;
using mrpt::math::TPose2D;
using ptg_t = mpp::ptg_t;

// TPS_Astar.cpp names render_tree() for its optional debug dump (saveDebugVisualizationDecimation,
// off by default); render_tree.cpp is visualisation and is not compiled.
auto mpp::render_tree(const MotionPrimitivesTreeSE2&, const PlannerInput&, const RenderOptions&)
    -> std::shared_ptr<mrpt::opengl::CSetOfObjects>
{
    return mrpt::opengl::CSetOfObjects::Create();
}

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

struct RefPTG
{
    std::shared_ptr<ptg_t> p;
};
static ptg_t* P(void* h) { return static_cast<RefPTG*>(h)->p.get(); }

static mrpt::maps::CSimplePointsMap::Ptr make_cloud(const float* xs, const float* ys, size_t n)
{
    auto pc = mrpt::maps::CSimplePointsMap::Create();
    pc->reserve(n);
    for (size_t i = 0; i < n; i++) pc->insertPointFast(xs[i], ys[i], 0);
    return pc;
}

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }
int         orc_num_threads() { return hw_threads(); }
int         orc_is_reference_build() { return 1; }

// ---- PTGs -------------------------------------------------------------------------------
void* orc_diffdrive_new(
    int numPaths, double refDist, double res, double vmax, double wmax_rad, double K, double turnRadiusRef,
    const double* shape_x, const double* shape_y, int nverts)
{
    ORC_TRY
    auto p                    = std::make_shared<mpp::ptg::DiffDrive_C>();
    p->m_alphaValuesCount     = static_cast<uint16_t>(numPaths);
    p->refDistance            = refDist;
    p->m_resolution           = res;
    p->V_MAX                  = vmax;
    p->W_MAX                  = wmax_rad;
    p->K                      = K;
    p->turningRadiusReference = turnRadiusRef;
    mrpt::math::CPolygon poly;
    for (int i = 0; i < nverts; i++) poly.AddVertex(shape_x[i], shape_y[i]);
    p->setRobotShape(poly);
    p->initialize("", false);
    return new RefPTG{p};
    ORC_CATCH(nullptr)
}

void* orc_holo_new(
    int numPaths, double refDist, double T_ramp_max, double vmax, double wmax_rad, double robotRadius,
    const char* expr_V, const char* expr_W)
{
    ORC_TRY
    auto p                = std::make_shared<mpp::ptg::HolonomicBlend>();
    p->m_alphaValuesCount = static_cast<uint16_t>(numPaths);
    p->refDistance        = refDist;
    p->T_ramp_max         = T_ramp_max;
    p->V_MAX              = vmax;
    p->W_MAX              = wmax_rad;
    if (expr_V && *expr_V) p->expr_V = expr_V;
    if (expr_W && *expr_W) p->expr_W = expr_W;
    p->setRobotShapeRadius(robotRadius);
    p->initialize("", false);
    return new RefPTG{p};
    ORC_CATCH(nullptr)
}

// TrajectoriesAndRobotShape::initFromConfigFile on an INI file (the reference's own loader);
// writes up to `cap` PTG handles, returns how many the file defines.
int orc_ptgs_from_ini(const char* iniPath, const char* section, void** out, int cap)
{
    ORC_TRY
    // share/ptgs_ackermann_vehicle.ini:22,30 names MRPT's upstream class `CPTG_DiffDrive_C`, whose
    // source is not available; it is aliased to the in-repo fork mpp::ptg::DiffDrive_C (SURVEY.md §7
    // hard part 7).
    mrpt::rtti::registerClassCustomName("CPTG_DiffDrive_C", CLASS_ID(mpp::ptg::DiffDrive_C));
    mrpt::config::CConfigFile      cfg(iniPath);
    mpp::TrajectoriesAndRobotShape trs;
    // initFromConfigFile writes its collision-grid cache to "./ptg_grid_NNN.dat.gz": the stream
    // stand-ins never open, so nothing is read or written.
    trs.initFromConfigFile(cfg, section);
    int n = 0;
    for (auto& p : trs.ptgs)
    {
        if (n < cap) out[n] = new RefPTG{p};
        n++;
    }
    return n;
    ORC_CATCH(-1)
}

void   orc_ptg_free(void* h) { delete static_cast<RefPTG*>(h); }
int    orc_ptg_num_paths(void* h) { return P(h)->getPathCount(); }
double orc_ptg_ref_distance(void* h) { return P(h)->getRefDistance(); }
double orc_ptg_max_radius(void* h) { return P(h)->getMaxRobotRadius(); }
double orc_ptg_step_duration(void* h) { return P(h)->getPathStepDuration(); }
long   orc_ptg_step_count(void* h, int k)
{
    ORC_TRY
    return static_cast<long>(P(h)->getPathStepCount(k));
    ORC_CATCH(-1)
}
double orc_ptg_init_dist(void* h, int k)
{
    double v = 0;
    P(h)->initTPObstacleSingle(k, v);
    return v;
}
int orc_ptg_path_pose(void* h, int k, unsigned step, double* xyphi_dist)
{
    ORC_TRY
    const TPose2D q = P(h)->getPathPose(k, step);
    xyphi_dist[0]   = q.x;
    xyphi_dist[1]   = q.y;
    xyphi_dist[2]   = q.phi;
    xyphi_dist[3]   = P(h)->getPathDist(k, step);
    return 0;
    ORC_CATCH(-1)
}
int orc_ptg_inverse_map(void* h, double x, double y, double tol, int* k, double* d)
{
    ORC_TRY
    return P(h)->inverseMap_WS2TP(x, y, *k, *d, tol) ? 1 : 0;
    ORC_CATCH(-1)
}
int orc_ptg_step_for_dist(void* h, int k, double dist, unsigned* step)
{
    ORC_TRY
    uint32_t   s  = 0;
    const bool ok = P(h)->getPathStepForDist(k, dist, s);
    *step         = s;
    return ok ? 1 : 0;
    ORC_CATCH(-1)
}
int orc_diffdrive_path(void* h, int k, float* out)
{
    ORC_TRY
    auto* p = dynamic_cast<mpp::ptg::DiffDriveCollisionGridBased*>(P(h));
    ASSERT_(p);
    const auto& tr = p->m_trajectory.at(k);
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
    auto* p = dynamic_cast<mpp::ptg::DiffDriveCollisionGridBased*>(P(h));
    ASSERT_(p);
    const auto& g = p->m_collisionGrid;
    *xmin         = g.m_x_min;
    *ymin         = g.m_y_min;
    *res          = g.m_resolution;
    *nx           = static_cast<int>(g.m_size_x);
    *ny           = static_cast<int>(g.m_size_y);
    long n        = 0;
    for (const auto& c : g.m_map) n += static_cast<long>(c.size());
    *nentries = n;
    return 0;
    ORC_CATCH(-1)
}
int orc_diffdrive_grid_csr(void* h, uint32_t* offsets, uint16_t* ks, float* ds)
{
    ORC_TRY
    auto* p = dynamic_cast<mpp::ptg::DiffDriveCollisionGridBased*>(P(h));
    ASSERT_(p);
    const auto& g = p->m_collisionGrid;
    uint32_t    n = 0;
    for (size_t c = 0; c < g.m_map.size(); c++)
    {
        offsets[c] = n;
        for (const auto& e : g.m_map[c])
        {
            ks[n] = e.first;
            ds[n] = e.second;
            n++;
        }
    }
    offsets[g.m_map.size()] = n;
    return 0;
    ORC_CATCH(-1)
}
int orc_ptg_point_inside_shape(void* h, double x, double y) { return P(h)->isPointInsideRobotShape(x, y) ? 1 : 0; }

// ---- stage (a) ----------------------------------------------------------------------------
long orc_local_obstacles(
    const float* xs, const float* ys, size_t n, const double* pose_xyphi, double clip, float* out_x, float* out_y)
{
    ORC_TRY
    auto                         in = make_cloud(xs, ys, n);
    mrpt::maps::CSimplePointsMap out;
    mpp::transform_pc_square_clipping(
        *in, mrpt::poses::CPose2D(TPose2D(pose_xyphi[0], pose_xyphi[1], pose_xyphi[2])), clip, out);
    std::memcpy(out_x, out.getPointsBufferRef_x().data(), out.size() * sizeof(float));
    std::memcpy(out_y, out.getPointsBufferRef_y().data(), out.size() * sizeof(float));
    return static_cast<long>(out.size());
    ORC_CATCH(-1)
}

// ---- stage (a)+(b) for a batch of node poses: TPS_Astar::cached_local_obstacles (TPS_Astar.cpp:828)
// followed by tp_obstacles_single_path per (PTG, k) (the call made at TPS_Astar.cpp:744-745).
double orc_eval_nodes(
    void** ptgs, int nptg, const float* xs, const float* ys, size_t n, const double* poses_xyphi, size_t B,
    double clip, double* out, int mode, int nthreads)
{
    ORC_TRY
    mrpt::maps::CPointsMap::Ptr              cloud = make_cloud(xs, ys, n);
    std::vector<mrpt::maps::CPointsMap::Ptr> clouds{cloud};
    size_t                                   sumK = 0;
    for (int p = 0; p < nptg; p++) sumK += P(ptgs[p])->getPathCount();
    const auto t0 = std::chrono::steady_clock::now();
    parallel_for(static_cast<long>(B), nthreads, [&](long b) {
        thread_local mpp::TPS_Astar planner;  // only for its cached_local_obstacles member
        const TPose2D               pose(poses_xyphi[3 * b], poses_xyphi[3 * b + 1], poses_xyphi[3 * b + 2]);
        const auto                  local = planner.cached_local_obstacles(pose, clouds, clip);
        double*                     o     = out + b * sumK;
        for (int p = 0; p < nptg; p++)
        {
            const ptg_t* ptg = P(ptgs[p]);
            const int    K   = ptg->getPathCount();
            if (mode == 1)
            {
                std::vector<double> tp;
                ptg->initTPObstacles(tp);
                size_t       nObs;
                const float *oxs, *oys, *ozs;
                local->getPointsBuffer(nObs, oxs, oys, ozs);
                for (size_t i = 0; i < nObs; i++) ptg->updateTPObstacle(oxs[i], oys[i], tp);
                for (int k = 0; k < K; k++) o[k] = tp[k];
            }
            else
            {
                for (int k = 0; k < K; k++) o[k] = mpp::tp_obstacles_single_path(k, *local, *ptg);
            }
            o += K;
        }
    });
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
    ORC_CATCH(-1.0)
}

int orc_solve_poly4(double a, double b, double c, double d, double* x) { return mrpt::math::solve_poly4(x, a, b, c, d); }
int orc_solve_poly3(double a, double b, double c, double* x) { return mrpt::math::solve_poly3(x, a, b, c); }
int orc_solve_poly2(double a, double b, double c, double* r) { return mrpt::math::solve_poly2(a, b, c, r[0], r[1]); }

static ptg_t::TNavDynamicState dyn_from(const double* dyn)
{
    ptg_t::TNavDynamicState ds;
    ds.curVelLocal.vx    = dyn[0];
    ds.curVelLocal.vy    = dyn[1];
    ds.curVelLocal.omega = dyn[2];
    ds.relTarget         = TPose2D(dyn[3], dyn[4], dyn[5]);
    ds.targetRelSpeed    = dyn[6];
    return ds;
}
int orc_ptg_update_dyn_state(void* h, const double* dyn, int force)
{
    ORC_TRY
    P(h)->updateNavDynamicState(dyn_from(dyn), force != 0);
    return 0;
    ORC_CATCH(-1)
}
void orc_ptg_set_trim_speed(void* h, double s)
{
    if (auto* t = dynamic_cast<mpp::ptg::SpeedTrimmablePTG*>(P(h)); t) t->trimmableSpeed_ = s;
}
int orc_ptg_path_twist(void* h, int k, unsigned step, double* out3)
{
    ORC_TRY
    const auto t = P(h)->getPathTwist(k, step);
    out3[0]      = t.vx;
    out3[1]      = t.vy;
    out3[2]      = t.omega;
    return 0;
    ORC_CATCH(-1)
}
double orc_ptg_update_tpobs_single(void* h, double ox, double oy, int k, double tp_obstacle_k)
{
    P(h)->updateTPObstacleSingle(ox, oy, k, tp_obstacle_k);
    return tp_obstacle_k;
}

double orc_eval_candidates(
    void* h, const float* xs, const float* ys, size_t n, const double* nodes, const double* relTargets, size_t B,
    const int* cand_node, const int* cand_k, const double* cand_speed, size_t M, double clip, double* out_raw,
    double* out_free)
{
    ORC_TRY
    ptg_t* ptg  = P(h);
    auto*  trim = dynamic_cast<mpp::ptg::SpeedTrimmablePTG*>(ptg);
    ASSERT_(trim);
    mrpt::maps::CPointsMap::Ptr              cloud = make_cloud(xs, ys, n);
    std::vector<mrpt::maps::CPointsMap::Ptr> clouds{cloud};
    std::vector<mrpt::maps::CPointsMap::Ptr> locals(B);
    mpp::TPS_Astar                           planner;
    const auto                               t0 = std::chrono::steady_clock::now();
    for (size_t b = 0; b < B; b++)
        locals[b] = planner.cached_local_obstacles(TPose2D(nodes[6 * b], nodes[6 * b + 1], nodes[6 * b + 2]), clouds, clip);
    for (size_t m = 0; m < M; m++)
    {
        const size_t            b = cand_node[m];
        ptg_t::TNavDynamicState ds;
        ds.curVelLocal = mrpt::math::TTwist2D(nodes[6 * b + 3], nodes[6 * b + 4], nodes[6 * b + 5]);
        ds.curVelLocal.rotate(-nodes[6 * b + 2]);
        ds.relTarget         = TPose2D(relTargets[3 * b], relTargets[3 * b + 1], relTargets[3 * b + 2]);
        ds.targetRelSpeed    = 0;
        trim->trimmableSpeed_ = cand_speed[m];
        ptg->updateNavDynamicState(ds, true);
        trim->trimmableSpeed_ = cand_speed[m];
        double       raw      = std::numeric_limits<double>::infinity();
        size_t       nObs;
        const float *oxs, *oys, *ozs;
        locals[b]->getPointsBuffer(nObs, oxs, oys, ozs);
        for (size_t i = 0; i < nObs; i++) ptg->updateTPObstacleSingle(oxs[i], oys[i], cand_k[m], raw);
        out_raw[m] = raw;
        if (out_free) out_free[m] = mpp::tp_obstacles_single_path(cand_k[m], *locals[b], *ptg);
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    ORC_CATCH(-1.0)
}

// ---- cost evaluators ----------------------------------------------------------------------
struct RefCostmap
{
    mpp::CostEvaluatorCostMap::Ptr cm;
};
void* orc_costmap_new(
    const float* xs, const float* ys, size_t n, double resolution, double clearance, double maxCost, int useAverage,
    double maxRadiusFromRobot, const double* robot_xyphi)
{
    ORC_TRY
    auto                                  c = make_cloud(xs, ys, n);
    mpp::CostEvaluatorCostMap::Parameters p;
    p.resolution                 = resolution;
    p.preferredClearanceDistance = clearance;
    p.maxCost                    = maxCost;
    p.useAverageOfPath           = useAverage != 0;
    p.maxRadiusFromRobot         = maxRadiusFromRobot;
    std::optional<TPose2D> pose;
    if (robot_xyphi) pose = TPose2D(robot_xyphi[0], robot_xyphi[1], robot_xyphi[2]);
    auto* h = new RefCostmap();
    h->cm   = mpp::CostEvaluatorCostMap::FromStaticPointObstacles(*c, p, pose);
    return h;
    ORC_CATCH(nullptr)
}
void orc_costmap_free(void* h) { delete static_cast<RefCostmap*>(h); }
void orc_costmap_info(void* h, double* info)
{
    const auto& g = static_cast<RefCostmap*>(h)->cm->costmap_;
    info[0]       = g.getXMin();
    info[1]       = g.getYMin();
    info[2]       = g.getResolution();
    info[3]       = static_cast<double>(g.getSizeX());
    info[4]       = static_cast<double>(g.getSizeY());
}
void orc_costmap_cells(void* h, double* out)
{
    const auto& g = static_cast<RefCostmap*>(h)->cm->costmap_;
    std::memcpy(out, g.m_map.data(), g.m_map.size() * sizeof(double));
}
double orc_costmap_eval_pose(void* h, double x, double y)
{
    return static_cast<RefCostmap*>(h)->cm->eval_single_pose(TPose2D(x, y, 0));
}

struct RefWaypoints
{
    mpp::CostEvaluatorPreferredWaypoint::Ptr ev;
};
void* orc_waypoints_new(const double* xs, const double* ys, size_t n, double radius, double scale, int useAverage)
{
    ORC_TRY
    auto* h                                = new RefWaypoints();
    h->ev                                  = mpp::CostEvaluatorPreferredWaypoint::Create();
    h->ev->params_.waypointInfluenceRadius = radius;
    h->ev->params_.costScale               = scale;
    h->ev->params_.useAverageOfPath        = useAverage != 0;
    std::vector<TPoint2D> pts(n);
    for (size_t i = 0; i < n; i++) pts[i] = TPoint2D(xs[i], ys[i]);
    h->ev->setPreferredWaypoints(pts);
    return h;
    ORC_CATCH(nullptr)
}
void   orc_waypoints_free(void* h) { delete static_cast<RefWaypoints*>(h); }
double orc_waypoints_eval_pose(void* h, double x, double y)
{
    return static_cast<RefWaypoints*>(h)->ev->eval_single_pose(TPose2D(x, y, 0));
}

int orc_eval_edge_costs(
    void* evaluator, int kind, size_t nEdges, const double* from_xyphi, const uint32_t* sample_offsets,
    const double* rel_xy, double* out_cost)
{
    ORC_TRY
    for (size_t e = 0; e < nEdges; e++)
    {
        mpp::MoveEdgeSE2_TPS edge;
        edge.stateFrom.pose = TPose2D(from_xyphi[3 * e], from_xyphi[3 * e + 1], from_xyphi[3 * e + 2]);
        for (uint32_t s = sample_offsets[e]; s < sample_offsets[e + 1]; s++)
            edge.interpolatedPath[static_cast<double>(s)] = TPose2D(rel_xy[2 * s], rel_xy[2 * s + 1], 0);
        if (kind == 0)
            out_cost[e] = (*static_cast<RefCostmap*>(evaluator)->cm)(edge);
        else
            out_cost[e] = (*static_cast<RefWaypoints*>(evaluator)->ev)(edge);
    }
    return 0;
    ORC_CATCH(-1)
}

// ---- TPS_Astar ------------------------------------------------------------------------------
namespace
{
// Appended after the user's evaluators: adds +0.0 to the edge cost (exact) and counts calls, i.e.
// how many tentative edges Planner::cost_path_segment scored.
class CountingEvaluator : public mpp::CostEvaluator
{
   public:
    double         operator()(const mpp::MoveEdgeSE2_TPS&) const override
    {
        calls++;
        return 0.0;
    }
    mutable size_t calls = 0;
};

struct ExpansionCapReached : public std::exception
{
};

struct RefPlanner
{
    mpp::TPS_Astar                               planner;
    mpp::PlannerInput                            in;
    mpp::PlannerOutput                           out;
    std::vector<mpp::CostEvaluator::Ptr>         evaluators;
    std::shared_ptr<CountingEvaluator>           counter = std::make_shared<CountingEvaluator>();
    size_t                                       maxExpansions = 0;  // 0 = unlimited
    mpp::MotionPrimitivesTreeSE2::path_t          path;
    mpp::MotionPrimitivesTreeSE2::edge_sequence_t pathEdges;
    std::string                                   text;
};
}  // namespace

void* orc_planner_new() { return new RefPlanner(); }
void  orc_planner_free(void* h) { delete static_cast<RefPlanner*>(h); }
void  orc_planner_set_params(void* h, const double* p, const double* timestamps, int nTimestamps)
{
    auto& R                           = *static_cast<RefPlanner*>(h);
    auto& q                           = R.planner.params_;
    q.grid_resolution_xy              = p[0];
    q.grid_resolution_yaw             = p[1];
    q.SE2_metricAngleWeight           = p[2];
    q.heuristic_heading_weight        = p[3];
    q.max_ptg_trajectories_to_explore = static_cast<uint32_t>(p[4]);
    q.max_ptg_speeds_to_explore       = static_cast<uint32_t>(p[5]);
    q.pathInterpolatedSegments        = static_cast<size_t>(p[6]);
    q.maximumComputationTime          = p[7] > 0 ? p[7] : std::numeric_limits<double>::max();
    R.maxExpansions                   = p[8] > 0 ? static_cast<size_t>(p[8]) : 0;
    q.ptg_sample_timestamps.assign(timestamps, timestamps + nTimestamps);
}
void orc_planner_add_ptg(void* h, void* ptg)
{
    auto& R = *static_cast<RefPlanner*>(h);
    R.in.ptgs.ptgs.push_back(static_cast<RefPTG*>(ptg)->p);
    R.in.ptgs.initialized_ = true;  // what initFromConfigFile sets (TrajectoriesAndRobotShape.cpp:94)
}
void orc_planner_add_obstacles(void* h, const float* xs, const float* ys, size_t n)
{
    static_cast<RefPlanner*>(h)->in.obstacles.push_back(mpp::ObstacleSource::FromStaticPointcloud(make_cloud(xs, ys, n)));
}
void orc_planner_add_costmap(void* h, void* costmap)
{
    static_cast<RefPlanner*>(h)->evaluators.push_back(static_cast<RefCostmap*>(costmap)->cm);
}
void orc_planner_add_waypoints(void* h, void* wps)
{
    static_cast<RefPlanner*>(h)->evaluators.push_back(static_cast<RefWaypoints*>(wps)->ev);
}

int orc_planner_plan(
    void* h, const double* start, const double* goal, int goalIsPoint, const double* bboxMin, const double* bboxMax,
    double* result)
{
    ORC_TRY
    auto& R               = *static_cast<RefPlanner*>(h);
    R.in.stateStart.pose  = TPose2D(start[0], start[1], start[2]);
    R.in.stateStart.vel   = mrpt::math::TTwist2D(start[3], start[4], start[5]);
    if (goalIsPoint)
        R.in.stateGoal.state = mpp::PoseOrPoint(TPoint2D(goal[0], goal[1]));
    else
        R.in.stateGoal.state = mpp::PoseOrPoint(TPose2D(goal[0], goal[1], goal[2]));
    R.in.stateGoal.vel  = mrpt::math::TTwist2D(goal[3], goal[4], goal[5]);
    R.in.worldBboxMin   = TPose2D(bboxMin[0], bboxMin[1], bboxMin[2]);
    R.in.worldBboxMax   = TPose2D(bboxMax[0], bboxMax[1], bboxMax[2]);

    R.planner.costEvaluators_ = R.evaluators;
    R.planner.costEvaluators_.push_back(R.counter);
    R.counter->calls = 0;
    R.planner.profiler_().clear();

    // Expansion cap (a determinism aid of the tests, not a reference feature): the planner's
    // progress callback fires at the end of every A* iteration when its period is negative
    // (TPS_Astar.cpp:438-455); at the N-th call the tree is copied out and the search is aborted
    // by an exception. The reference code itself is untouched.
    mpp::MotionPrimitivesTreeSE2          cappedTree;
    std::optional<mrpt::graphs::TNodeID>  cappedBest;
    double                                cappedBestCostToGoal = 0;
    size_t                                iters                = 0;
    if (R.maxExpansions)
    {
        R.planner.progressCallbackCallPeriod_ = -1.0;
        R.planner.progressCallback_           = [&](const mpp::ProgressCallbackData& pcd) {
            if (++iters >= R.maxExpansions)
            {
                cappedTree           = *pcd.tree;
                cappedBest           = pcd.bestFinalNode;
                cappedBestCostToGoal = pcd.bestCostToGoal;
                throw ExpansionCapReached();
            }
        };
    }
    else
        R.planner.progressCallback_ = nullptr;

    const double t0 = mrpt::Clock::nowDouble();
    try
    {
        R.out = R.planner.plan(R.in);
    }
    catch (const ExpansionCapReached&)
    {
        R.out                       = mpp::PlannerOutput();
        R.out.originalInput         = R.in;
        R.out.motionTree            = cappedTree;
        R.out.bestNodeId            = cappedBest;
        R.out.bestNodeIdCostToGoal  = cappedBestCostToGoal;
        const auto startCell        = R.planner.nodeGridCoords(R.in.stateStart.pose);
        const auto goalCell         = R.planner.nodeGridCoords(R.in.stateGoal.asSE2KinState().pose);
        R.out.goalNodeId            = (startCell == goalCell) ? 0 : 1;  // creation order, TPS_Astar.cpp:156-176
        R.out.success               = R.out.goalNodeId == R.out.bestNodeId;
        if (R.out.bestNodeId) R.out.pathCost = R.out.motionTree.nodes().at(*R.out.bestNodeId).cost_;
        R.out.computationTime = mrpt::Clock::nowDouble() - t0;
    }
    R.path.clear();
    R.pathEdges.clear();
    if (R.out.bestNodeId)
    {
        auto [p, e] = R.out.motionTree.backtrack_path(*R.out.bestNodeId);
        R.path      = std::move(p);
        R.pathEdges = std::move(e);
    }
    const auto& cnt = R.planner.profiler_().counts_;
    auto        get = [&](const char* k) -> double {
        auto it = cnt.find(k);
        return it == cnt.end() ? 0.0 : static_cast<double>(it->second);
    };
    result[0]  = R.out.success ? 1 : 0;
    result[1]  = R.out.pathCost;
    result[2]  = static_cast<double>(R.out.motionTree.nodes().size());
    result[3]  = static_cast<double>(R.out.motionTree.edges_to_children.size());
    result[4]  = get("find_feasible");
    result[5]  = get("find_feasible.tp_obstacles_single");
    result[6]  = static_cast<double>(R.counter->calls);
    result[7]  = R.out.bestNodeId ? static_cast<double>(*R.out.bestNodeId) : -1;
    result[8]  = R.out.goalNodeId ? static_cast<double>(*R.out.goalNodeId) : -1;
    result[9]  = R.out.computationTime;
    result[10] = R.out.bestNodeIdCostToGoal;
    return 0;
    ORC_CATCH(-1)
}
int orc_planner_refine(void* h)
{
    ORC_TRY
    auto& R = *static_cast<RefPlanner*>(h);
    mpp::refine_trajectory(R.path, R.pathEdges, R.in.ptgs);
    return 0;
    ORC_CATCH(-1)
}
long orc_planner_path_len(void* h) { return static_cast<long>(static_cast<RefPlanner*>(h)->path.size()); }
int  orc_planner_path(void* h, double* nodes, double* edges)
{
    ORC_TRY
    auto&  R = *static_cast<RefPlanner*>(h);
    size_t i = 0;
    for (const auto& n : R.path)
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
    for (const mpp::MoveEdgeSE2_TPS* e : R.pathEdges)
    {
        double*      o  = edges + 8 * i++;
        const double dt = R.in.ptgs.ptgs.at(e->ptgIndex)->getPathStepDuration();
        o[0]            = e->ptgIndex;
        o[1]            = e->ptgPathIndex;
        o[2]            = static_cast<double>(std::lround(e->estimatedExecTime / dt));  // estimatedExecTime = step*dt
        o[3]            = e->ptgDist;
        o[4]            = e->cost;
        o[5]            = e->estimatedExecTime;
        o[6]            = e->ptgTrimmableSpeed;
        o[7]            = static_cast<double>(e->interpolatedPath.size());
    }
    return 0;
    ORC_CATCH(-1)
}
long orc_planner_tree(void* h, double* out, long cap)
{
    ORC_TRY
    auto& R = *static_cast<RefPlanner*>(h);
    long  i = 0;
    for (const auto& kv : R.out.motionTree.nodes())
    {
        if (out && i < cap)
        {
            double* o = out + 7 * i;
            o[0]      = static_cast<double>(kv.first);
            o[1]      = kv.second.parentID_ ? static_cast<double>(*kv.second.parentID_) : -1;
            o[2]      = kv.second.pose.x;
            o[3]      = kv.second.pose.y;
            o[4]      = kv.second.pose.phi;
            o[5]      = kv.second.cost_;
            o[6]      = kv.second.parentID_ ? R.out.motionTree.edge_to_parent(kv.first).cost : 0;
        }
        i++;
    }
    return i;
    ORC_CATCH(-1)
}
long orc_planner_trajectory_txt(void* h, double samplePeriod, char* buf, long cap)
{
    ORC_TRY
    auto& R    = *static_cast<RefPlanner*>(h);
    auto  traj = mpp::plan_to_trajectory(R.pathEdges, R.in.ptgs, samplePeriod);
    // trajectories are relative to the start pose; the CLI re-bases them (path-planner-cli.cpp:396-399)
    for (auto& kv : traj) kv.second.state.pose = R.in.stateStart.pose + kv.second.state.pose;
    char tmpl[] = "/tmp/mpp_ref_traj_XXXXXX";
    const int fd = mkstemp(tmpl);
    ASSERT_(fd >= 0);
    close(fd);
    const bool ok = mpp::save_to_txt(traj, tmpl);
    std::ifstream     f(tmpl);
    std::stringstream ss;
    ss << f.rdbuf();
    unlink(tmpl);
    ASSERT_(ok);
    R.text = ss.str();
    if (buf && cap > 0)
    {
        const long m = std::min<long>(cap - 1, static_cast<long>(R.text.size()));
        std::memcpy(buf, R.text.data(), m);
        buf[m] = 0;
    }
    return static_cast<long>(R.text.size());
    ORC_CATCH(-1)
}

}  // extern "C"
