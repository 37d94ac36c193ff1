// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product.
//
// extern "C" surface of the CPU oracle so that tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs can call it through ctypes.
// Parity UNPINNED (no reference golden vectors exist).
#include <chrono>
#include <cstring>

#include <atomic>
#include <mutex>
#include <thread>

#include "orc_algos.hpp"

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

}  // extern "C"
