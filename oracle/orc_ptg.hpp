// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product.
//
// CPU restatement of the PTG ("parameterized trajectory generator") arithmetic on the
// hot path. Each function cites the reference lines it follows (paths relative to
// /root/reference/mrpt_path_planning/). Behaviour of the un-vendored MRPT base class
// (mrpt::nav::CParameterizedTrajectoryGenerator) is restated from SURVEY.md §8c and is
// marked [MRPT]. The mpp:: parts are pinned against the reference's own sources
// compiled as oracle/_ref (tests/test_oracle_vs_reference.py); the [MRPT] parts stay assumed.
#pragma once
#include "orc_math.hpp"

#include <map>
#include <memory>
#include <utility>

namespace orc
{
// [MRPT] TNavDynamicState (TPS_Astar.cpp:588-603)
struct NavDynState
{
    TTwist2D curVelLocal;
    TPose2D  relTarget{20.0, 0, 0};
    double   targetRelSpeed = 0;
    bool     operator==(const NavDynState& o) const
    {
        return curVelLocal == o.curVelLocal && relTarget == o.relTarget && targetRelSpeed == o.targetRelSpeed;
    }
};

// [MRPT] CParameterizedTrajectoryGenerator — only what the hot path calls.
class PTG
{
   public:
    virtual ~PTG() = default;

    static constexpr int INVALID_PATH_INDEX = -1;

    unsigned numPaths    = 0;
    double   refDistance = 0;

    // [MRPT] index2alpha / alpha2index (DiffDriveCollisionGridBased.cpp:123, HolonomicBlend.cpp:723)
    double index2alpha(unsigned k) const { return M_PI * (-1.0 + 2.0 * (k + 0.5) / numPaths); }
    int    alpha2index(double alpha) const
    {
        // upstream calls wrapToPi(alpha) and discards the result; restated as a no-op.
        int k = mrpt_round(0.5 * (numPaths * (1.0 + alpha / M_PI) - 1.0));
        if (k < 0) k = 0;
        if (k >= static_cast<int>(numPaths)) k = numPaths - 1;
        return k;
    }

    virtual bool     isSpeedTrimmable() const { return false; }
    double           trimmableSpeed = 1.0;  // SpeedTrimmablePTG.h:17-24

    virtual size_t   getPathStepCount(unsigned k) const               = 0;
    virtual TPose2D  getPathPose(unsigned k, uint32_t step) const     = 0;
    virtual double   getPathDist(unsigned k, uint32_t step) const     = 0;
    virtual bool     getPathStepForDist(unsigned k, double dist, uint32_t& step) const = 0;
    virtual double   getPathStepDuration() const                      = 0;
    virtual TTwist2D getPathTwist(unsigned k, uint32_t step) const
    {
        // [MRPT] default implementation: finite differences between consecutive steps.
        if (step > 0)
        {
            const TPose2D p1 = getPathPose(k, step), p0 = getPathPose(k, step - 1);
            const double  dt = getPathStepDuration();
            TTwist2D      tw;
            tw.vx    = (p1.x - p0.x) / dt;
            tw.vy    = (p1.y - p0.y) / dt;
            tw.omega = (p1.phi - p0.phi) / dt;
            return tw;
        }
        return dynState.curVelLocal;
    }
    virtual bool inverseMap_WS2TP(double x, double y, int& k, double& d, double tol) const = 0;
    virtual bool supportSpeedAtTarget() const { return false; }
    virtual void onNewNavDynamicState() {}

    // [MRPT] updateNavDynamicState (TPS_Astar.cpp:281,608)
    void updateNavDynamicState(const NavDynState& ds, bool force = false)
    {
        if (!force && ds == dynState) return;
        dynState = ds;
        target_k = INVALID_PATH_INDEX;
        onNewNavDynamicState();
        if (supportSpeedAtTarget())
        {
            int    k = 0;
            double d = 0;
            if (inverseMap_WS2TP(ds.relTarget.x, ds.relTarget.y, k, d, 1.0) && d > 0.01 && d < 0.99)
            {
                target_k = k;
                onNewNavDynamicState();
            }
        }
    }
    NavDynState dynState;
    int         target_k = INVALID_PATH_INDEX;

    // [MRPT] robot shape
    virtual bool   isPointInsideRobotShape(double x, double y) const = 0;
    virtual double getMaxRobotRadius() const                        = 0;

    // [MRPT] initTPObstacleSingle (tp_obstacles_single_path.cpp:24)
    void initTPObstacleSingle(unsigned k, double& out) const
    {
        out = std::min(refDistance, getPathDist(k, static_cast<uint32_t>(getPathStepCount(k) - 1)));
    }
    // [MRPT] internal_TPObsDistancePostprocess with COLL_BEH_BACK_AWAY
    // (DiffDriveCollisionGridBased.cpp:836, HolonomicBlend.cpp:839)
    void tpObsPostprocess(double ox, double oy, double newDist, double& inout) const
    {
        if (!isPointInsideRobotShape(ox, oy))
        {
            if (newDist < inout) inout = newDist;
            return;
        }
        if (newDist < getMaxRobotRadius()) return;  // moving away from the obstacle: ignore
        inout = 0;
    }
    virtual void updateTPObstacleSingle(double ox, double oy, unsigned k, double& tp_obstacle_k) const = 0;
};

// ---------------------------------------------------------------------------------------
// DiffDriveCollisionGridBased + DiffDrive_C  (src/ptgs/DiffDriveCollisionGridBased.cpp,
// src/ptgs/DiffDrive_C.cpp)
// ---------------------------------------------------------------------------------------
struct TCPoint
{
    float x = 0, y = 0, phi = 0, t = 0, dist = 0, v = 0, w = 0;
};

class DiffDriveC : public PTG
{
   public:
    // parameters (ini keys resolution, v_max_mps, w_max_dps, K; .h:140-144)
    double V_MAX = 0, W_MAX = 0, turningRadiusReference = 0.10, resolution = 0.05, K = 1.0;
    double stepTimeDuration = 0.01;
    Polygon2D robotShape;
    double    robotMaxRadius = 0;

    using CollisionCell = std::vector<std::pair<uint16_t, float>>;
    std::vector<std::vector<TCPoint>> trajectory;
    DynGrid<CollisionCell>            collisionGrid;

    struct LambdaCell
    {
        uint16_t k_min = 0xffff, k_max = 0;
        uint32_t n_min = 0xffffffffu, n_max = 0;
        bool     isEmpty() const { return k_min == 0xffff; }
    };
    DynGrid<LambdaCell> lambdaGrid;

    bool isSpeedTrimmable() const override { return true; }  // DiffDrive_C.h:41

    // [MRPT] CPTG_RobotShape_Polygonal::setRobotShape
    void setRobotShape(const std::vector<double>& xs, const std::vector<double>& ys)
    {
        ORC_ASSERT(xs.size() == ys.size() && xs.size() >= 3);
        robotShape.vx  = xs;
        robotShape.vy  = ys;
        robotMaxRadius = 0;
        for (size_t i = 0; i < xs.size(); i++)
            robotMaxRadius = std::max(robotMaxRadius, std::sqrt(xs[i] * xs[i] + ys[i] * ys[i]));
    }
    bool   isPointInsideRobotShape(double x, double y) const override { return robotShape.contains(x, y); }
    double getMaxRobotRadius() const override { return robotMaxRadius; }

    // DiffDrive_C.cpp:70-80
    void steering(float alpha, float& v, float& w) const
    {
        v = V_MAX * sign_nz(K);
        w = (alpha / M_PI) * W_MAX * sign_nz(K);
    }

    // DiffDriveCollisionGridBased.cpp:103-201. State is float; cos()/sin() resolve to the
    // double C functions when only <cmath> is visible (g++/libstdc++), so the products are
    // formed in double and rounded to float on the "+=" store. Define
    // ORC_SIM_FLOAT_TRIG to restate the other reading (cosf/sinf, all-float products).
    void simulateTrajectories(float max_time, float max_dist, float dt)
    {
        trajectory.clear();
        stepTimeDuration = dt;
        trajectory.resize(numPaths);
        float x_min = 1e3f, x_max = -1e3f, y_min = 1e3f, y_max = -1e3f;
        for (unsigned k = 0; k < numPaths; k++)
        {
            const float          alpha = index2alpha(k);
            std::vector<TCPoint> pts;
            float                t = 0, dist = 0, turned = 0, x = 0, y = 0, phi = 0, v = 0, w = 0;
            pts.push_back(TCPoint{x, y, phi, t, dist, v, w});
            while (t < max_time && dist < max_dist && std::fabs(turned) < 1.95 * M_PI)
            {
                steering(alpha, v, w);
#ifdef ORC_SIM_FLOAT_TRIG
                x += cosf(phi) * v * dt;
                y += sinf(phi) * v * dt;
#else
                x += ::cos(static_cast<double>(phi)) * v * dt;
                y += ::sin(static_cast<double>(phi)) * v * dt;
#endif
                phi += w * dt;
                turned += w * dt;
                // mrpt::square<float>(v) is a float product; the second term is double.
                float v_inTPSpace = ::sqrt((v * v) + square(w * turningRadiusReference));
                dist += v_inTPSpace * dt;
                t += dt;
                pts.back().v = v;
                pts.back().w = w;
                pts.push_back(TCPoint{x, y, phi, t, dist, v, w});
                x_min = std::min(x_min, x);
                x_max = std::max(x_max, x);
                y_min = std::min(y_min, y);
                y_max = std::max(y_max, y);
            }
            pts.back().v = v;
            pts.back().w = w;
            pts.push_back(TCPoint{x, y, phi, t, dist, v, w});
            trajectory[k] = std::move(pts);
        }
        // :181-200 "lambda" grid used by the look-up inverse map
        lambdaGrid.setSize(x_min - 0.5f, x_max + 0.5f, y_min - 0.5f, y_max + 0.5f, 0.25f, LambdaCell());
        for (unsigned k = 0; k < numPaths; k++)
        {
            const uint32_t M = static_cast<uint32_t>(trajectory[k].size());
            for (uint32_t n = 0; n < M; n++)
            {
                LambdaCell* c = lambdaGrid.cellByPos(trajectory[k][n].x, trajectory[k][n].y);
                ORC_ASSERT(c);
                c->k_min = std::min<uint16_t>(c->k_min, k);
                c->k_max = std::max<uint16_t>(c->k_max, k);
                c->n_min = std::min(c->n_min, n);
                c->n_max = std::max(c->n_max, n);
            }
        }
    }

    // :233-258
    void updateCellInfo(int icx, int icy, uint16_t k, float dist)
    {
        CollisionCell* cell = collisionGrid.cellByIndex(icx, icy);
        if (!cell) return;
        for (auto& e : *cell)
            if (e.first == k)
            {
                if (dist < e.second) e.second = dist;
                return;
            }
        cell->emplace_back(k, dist);
    }

    // :615-764 (cache file I/O skipped: the tables are always rebuilt)
    void initialize()
    {
        ORC_ASSERT(robotShape.size() >= 3 && refDistance > 0 && V_MAX > 0 && W_MAX > 0 && resolution > 0);
        simulateTrajectories(100.0f, refDistance, 1.0e-3f);
        collisionGrid.setSize(-refDistance, refDistance, -refDistance, refDistance, resolution);
        const int    grid_cx_max = collisionGrid.size_x - 1, grid_cy_max = collisionGrid.size_y - 1;
        const double half_cell = collisionGrid.res * 0.5;
        const size_t nVerts    = robotShape.size();
        Polygon2D    poly;
        poly.vx.resize(nVerts);
        poly.vy.resize(nVerts);
        for (unsigned k = 0; k < numPaths; k++)
        {
            const size_t nPoints = getPathStepCount(k);
            ORC_ASSERT(nPoints > 1);
            for (size_t n = 0; n < nPoints - 1; n++)
            {
                const TPose2D p = getPathPose(k, n);
                double bbminx = std::numeric_limits<double>::max(), bbminy = bbminx;
                double bbmaxx = -std::numeric_limits<double>::max(), bbmaxy = bbmaxx;
                for (size_t m = 0; m < nVerts; m++)
                {
                    poly.vx[m] = p.x + std::cos(p.phi) * robotShape.vx[m] - std::sin(p.phi) * robotShape.vy[m];
                    poly.vy[m] = p.y + std::sin(p.phi) * robotShape.vx[m] + std::cos(p.phi) * robotShape.vy[m];
                    bbmaxx     = std::max(bbmaxx, poly.vx[m]);
                    bbmaxy     = std::max(bbmaxy, poly.vy[m]);
                    bbminx     = std::min(bbminx, poly.vx[m]);
                    bbminy     = std::min(bbminy, poly.vy[m]);
                }
                const int ix_min = std::max(0, collisionGrid.x2idx(bbminx) - 1);
                const int iy_min = std::max(0, collisionGrid.y2idx(bbminy) - 1);
                const int ix_max = std::min(collisionGrid.x2idx(bbmaxx) + 1, grid_cx_max);
                const int iy_max = std::min(collisionGrid.y2idx(bbmaxy) + 1, grid_cy_max);
                for (int ix = ix_min; ix < ix_max; ix++)
                {
                    const double cx = collisionGrid.idx2x(ix) - half_cell;
                    for (int iy = iy_min; iy < iy_max; iy++)
                    {
                        const double cy = collisionGrid.idx2y(iy) - half_cell;
                        if (poly.contains(cx, cy))
                        {
                            const float d = getPathDist(k, n);
                            updateCellInfo(ix, iy, k, d);
                            updateCellInfo(ix - 1, iy, k, d);
                            updateCellInfo(ix, iy - 1, k, d);
                            updateCellInfo(ix - 1, iy - 1, k, d);
                        }
                    }
                }
            }
        }
    }

    // :766-811, 880-896
    size_t  getPathStepCount(unsigned k) const override { return trajectory.at(k).size(); }
    TPose2D getPathPose(unsigned k, uint32_t step) const override
    {
        const TCPoint& p = trajectory.at(k).at(step);
        return TPose2D(p.x, p.y, p.phi);
    }
    double getPathDist(unsigned k, uint32_t step) const override { return trajectory.at(k).at(step).dist; }
    bool   getPathStepForDist(unsigned k, double dist, uint32_t& out_step) const override
    {
        const auto&  tr        = trajectory.at(k);
        const size_t numPoints = tr.size();
        for (size_t n = 0; n < numPoints - 1; n++)
            if (tr[n + 1].dist >= dist)
            {
                out_step = n;
                return true;
            }
        out_step = numPoints - 1;
        return false;
    }
    double   getPathStepDuration() const override { return stepTimeDuration; }
    TTwist2D getPathTwist(unsigned k, uint32_t step) const override
    {
        const TCPoint& p = trajectory.at(k).at(step);
        TTwist2D       tw{p.v, 0, p.w};
        tw.rotate(p.phi);
        return tw;
    }

    // DiffDrive_C.cpp:88-154 (closed-form inverse map)
    bool inverseMap_WS2TP(double x, double y, int& k_out, double& d_out, double /*tol*/) const override
    {
        bool is_exact = true;
        if (y != 0)
        {
            double       R    = (x * x + y * y) / (2 * y);
            const double Rmin = std::abs(V_MAX / W_MAX);
            double       theta;
            if (K > 0)
                theta = (y > 0) ? std::atan2(x, std::fabs(R) - y) : std::atan2(x, y + std::fabs(R));
            else
                theta = (y > 0) ? std::atan2(-x, std::fabs(R) - y) : std::atan2(-x, y + std::fabs(R));
            theta = wrapTo2Pi(theta);
            d_out = static_cast<float>(theta * (std::fabs(R) + turningRadiusReference));
            if (std::abs(R) < Rmin)
            {
                is_exact = false;
                R        = Rmin * sign_nz(R);
            }
            const double a = M_PI * V_MAX / (W_MAX * R);
            k_out          = alpha2index(static_cast<float>(a));
        }
        else
        {
            if (sign_nz(x) == sign_nz(K))
            {
                k_out    = alpha2index(0);
                d_out    = x;
                is_exact = true;
            }
            else
            {
                k_out    = numPaths - 1;
                d_out    = 1e+3;
                is_exact = false;
            }
        }
        d_out = d_out / refDistance;
        ORC_ASSERT(k_out >= 0 && k_out < static_cast<int>(numPaths));
        return is_exact;
    }

    // :826-838 with CCollisionGrid::getTPObstacle :220-227 (float narrowing of ox, oy)
    void updateTPObstacleSingle(double ox, double oy, unsigned k, double& tp_obstacle_k) const override
    {
        const float          fx = ox, fy = oy;
        const CollisionCell* cell = collisionGrid.cellByPos(fx, fy);
        if (!cell) return;
        for (const auto& e : *cell)
            if (e.first == static_cast<uint16_t>(k))
            {
                const double dist = e.second;
                tpObsPostprocess(ox, oy, dist, tp_obstacle_k);
            }
    }
    // :813-824 (all k at once; identical arithmetic, used to speed up large test cases)
    void updateTPObstacleAll(double ox, double oy, double* tp_obstacles) const
    {
        const float          fx = ox, fy = oy;
        const CollisionCell* cell = collisionGrid.cellByPos(fx, fy);
        if (!cell) return;
        for (const auto& e : *cell) tpObsPostprocess(ox, oy, static_cast<double>(e.second), tp_obstacles[e.first]);
    }
};

}  // namespace orc
