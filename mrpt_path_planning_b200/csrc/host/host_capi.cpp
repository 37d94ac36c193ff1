// C ABI over the host-side mpp:: mirror objects (PTGs, later: cost evaluators, planner).
#include <cstring>
#include <memory>
#include <string>

#include "../../../include/mppb.h"
#include "host_internal.h"

namespace
{
thread_local std::string g_ptgErr;
}

#define HOST_TRY try {
#define HOST_CATCH(ret)             \
    }                               \
    catch (const std::exception& e) \
    {                               \
        g_ptgErr = e.what();        \
        return ret;                 \
    }

extern "C" {

const char* mppb_ptg_last_error(void) { return g_ptgErr.c_str(); }

int mppb_ptg_create_diffdrive_c(const mppb_diffdrive_c_params* p, mppb_ptg** out)
{
    return mppb_ptg_create_diffdrive_c_on(nullptr, p, out);
}

int mppb_ptg_create_diffdrive_c_on(mppb_ctx* ctx, const mppb_diffdrive_c_params* p, mppb_ptg** out)
{
    HOST_TRY
    if (!p || !out) throw std::runtime_error("null argument");
    *out = nullptr;
    if (p->shape_nverts < 3 || !p->shape_x || !p->shape_y) throw std::runtime_error("robot shape needs >= 3 vertices");
    mpp::DiffDriveCParams q;
    q.num_paths              = p->num_paths;
    q.refDistance            = p->ref_distance;
    q.resolution             = p->resolution;
    q.v_max_mps              = p->v_max_mps;
    q.w_max_rps              = p->w_max_rps;
    q.K                      = p->K;
    q.turningRadiusReference = p->turning_radius_ref;
    q.shape_x.assign(p->shape_x, p->shape_x + p->shape_nverts);
    q.shape_y.assign(p->shape_y, p->shape_y + p->shape_nverts);
    auto* h = new mppb_ptg();
    try
    {
        h->obj = std::make_unique<mpp::DiffDrive_C>(q, ctx);
    }
    catch (...)
    {
        delete h;
        throw;
    }
    *out = h;
    return MPPB_OK;
    HOST_CATCH(MPPB_ERR_INVALID_ARG)
}

void mppb_ptg_destroy(mppb_ptg* ptg) { delete ptg; }

int    mppb_ptg_num_paths(const mppb_ptg* p) { return p ? p->obj->getPathCount() : 0; }
double mppb_ptg_ref_distance(const mppb_ptg* p) { return p ? p->obj->getRefDistance() : 0; }
double mppb_ptg_step_duration(const mppb_ptg* p) { return p ? p->obj->getPathStepDuration() : 0; }
double mppb_ptg_max_radius(const mppb_ptg* p) { return p ? p->obj->getMaxRobotRadius() : 0; }

int64_t mppb_ptg_step_count(const mppb_ptg* p, int k)
{
    HOST_TRY
    if (!p || k < 0 || k >= p->obj->getPathCount()) throw std::runtime_error("bad path index");
    return static_cast<int64_t>(p->obj->getPathStepCount(static_cast<uint16_t>(k)));
    HOST_CATCH(-1)
}
int mppb_ptg_path_pose(const mppb_ptg* p, int k, uint32_t step, double* out4)
{
    HOST_TRY
    if (!p || !out4 || k < 0 || k >= p->obj->getPathCount()) throw std::runtime_error("bad argument");
    const mpp::TPose2D q = p->obj->getPathPose(static_cast<uint16_t>(k), step);
    out4[0]              = q.x;
    out4[1]              = q.y;
    out4[2]              = q.phi;
    out4[3]              = p->obj->getPathDist(static_cast<uint16_t>(k), step);
    return MPPB_OK;
    HOST_CATCH(MPPB_ERR_INVALID_ARG)
}
int mppb_ptg_step_for_dist(const mppb_ptg* p, int k, double dist, uint32_t* out_step)
{
    HOST_TRY
    if (!p || !out_step || k < 0 || k >= p->obj->getPathCount()) throw std::runtime_error("bad argument");
    return p->obj->getPathStepForDist(static_cast<uint16_t>(k), dist, *out_step) ? 1 : 0;
    HOST_CATCH(MPPB_ERR_INVALID_ARG)
}
int mppb_ptg_inverse_map(const mppb_ptg* p, double x, double y, double tol, int* out_k, double* out_d)
{
    HOST_TRY
    if (!p || !out_k || !out_d) throw std::runtime_error("bad argument");
    return p->obj->inverseMap_WS2TP(x, y, *out_k, *out_d, tol) ? 1 : 0;
    HOST_CATCH(MPPB_ERR_INVALID_ARG)
}
double mppb_ptg_init_dist(const mppb_ptg* p, int k)
{
    HOST_TRY
    if (!p || k < 0 || k >= p->obj->getPathCount()) throw std::runtime_error("bad path index");
    return p->obj->initTPObstacleSingle(static_cast<uint16_t>(k));
    HOST_CATCH(-1.0)
}
int mppb_ptg_grid_export(const mppb_ptg* p, mppb_ptg_grid_desc* out)
{
    HOST_TRY
    if (!p || !out) throw std::runtime_error("bad argument");
    auto* dd = dynamic_cast<const mpp::DiffDrive_C*>(p->obj.get());
    if (!dd) throw std::runtime_error("PTG is not of the collision-grid family");
    *out = dd->gridDesc();
    return MPPB_OK;
    HOST_CATCH(MPPB_ERR_UNSUPPORTED)
}
int mppb_ptg_paths_export(const mppb_ptg* p, mppb_ptg_paths_desc* out)
{
    HOST_TRY
    if (!p || !out) throw std::runtime_error("bad argument");
    auto* dd = dynamic_cast<const mpp::DiffDrive_C*>(p->obj.get());
    if (!dd) throw std::runtime_error("PTG is not of the collision-grid family");
    *out = dd->pathsDesc();
    return MPPB_OK;
    HOST_CATCH(MPPB_ERR_UNSUPPORTED)
}
int mppb_ctx_add_ptg(mppb_ctx* ctx, const mppb_ptg* p)
{
    if (!ctx || !p) return MPPB_ERR_INVALID_ARG;
    if (auto* holo = dynamic_cast<const mpp::HolonomicBlend*>(p->obj.get()))
    {
        const mppb_ptg_holo_desc d = holo->holoDesc();
        return mppb_set_ptg_holo(ctx, mppb_num_ptgs(ctx), &d);
    }
    mppb_ptg_grid_desc d;
    int                rc = mppb_ptg_grid_export(p, &d);
    if (rc != MPPB_OK) return rc;
    const int idx = mppb_num_ptgs(ctx);
    rc            = mppb_set_ptg_grid(ctx, idx, &d);
    if (rc != MPPB_OK) return rc;
    // the sampled paths too: the device-side expansion (mppb_expand_nodes) looks end poses and distances up in them
    const mppb_ptg_paths_desc pd = static_cast<const mpp::DiffDrive_C*>(p->obj.get())->pathsDesc();
    return mppb_set_ptg_paths(ctx, idx, &pd);
}

int mppb_ptg_create_holo(const mppb_holo_params* p, mppb_ptg** out)
{
    HOST_TRY
    if (!p || !out) throw std::runtime_error("null argument");
    *out = nullptr;
    mpp::HolonomicBlendParams q;
    q.num_paths    = p->num_paths;
    q.refDistance  = p->ref_distance;
    q.T_ramp_max   = p->T_ramp_max;
    q.v_max_mps    = p->v_max_mps;
    q.w_max_rps    = p->w_max_rps;
    q.robot_radius = p->robot_radius;
    if (p->expr_V && *p->expr_V) q.expr_V = p->expr_V;
    if (p->expr_W && *p->expr_W) q.expr_W = p->expr_W;
    auto* h = new mppb_ptg();
    try
    {
        h->obj = std::make_unique<mpp::HolonomicBlend>(q);
    }
    catch (...)
    {
        delete h;
        throw;
    }
    *out = h;
    return MPPB_OK;
    HOST_CATCH(MPPB_ERR_INVALID_ARG)
}
int mppb_ptg_update_dyn_state(mppb_ptg* p, const double* dyn, int force)
{
    HOST_TRY
    if (!p || !dyn) throw std::runtime_error("bad argument");
    mpp::TNavDynamicState ds;
    ds.curVelLocal    = mpp::TTwist2D(dyn[0], dyn[1], dyn[2]);
    ds.relTarget      = mpp::TPose2D(dyn[3], dyn[4], dyn[5]);
    ds.targetRelSpeed = dyn[6];
    p->obj->updateNavDynamicState(ds, force != 0);
    return MPPB_OK;
    HOST_CATCH(MPPB_ERR_INVALID_ARG)
}
int mppb_ptg_set_trim_speed(mppb_ptg* p, double speed)
{
    if (!p) return MPPB_ERR_INVALID_ARG;
    p->obj->trimmableSpeed_ = speed;
    return MPPB_OK;
}
int mppb_ptg_path_twist(const mppb_ptg* p, int k, uint32_t step, double* out3)
{
    HOST_TRY
    if (!p || !out3 || k < 0 || k >= p->obj->getPathCount()) throw std::runtime_error("bad argument");
    const mpp::TTwist2D t = p->obj->getPathTwist(static_cast<uint16_t>(k), step);
    out3[0]               = t.vx;
    out3[1]               = t.vy;
    out3[2]               = t.omega;
    return MPPB_OK;
    HOST_CATCH(MPPB_ERR_INVALID_ARG)
}
int mppb_ptg_holo_params(const mppb_ptg* p, int k, mppb_holo_cand* out)
{
    HOST_TRY
    if (!p || !out || k < 0 || k >= p->obj->getPathCount()) throw std::runtime_error("bad argument");
    auto* holo = dynamic_cast<const mpp::HolonomicBlend*>(p->obj.get());
    if (!holo) throw std::runtime_error("PTG is not a HolonomicBlend");
    holo->holoParams(static_cast<uint16_t>(k), *out);
    return MPPB_OK;
    HOST_CATCH(MPPB_ERR_UNSUPPORTED)
}

}  // extern "C"
