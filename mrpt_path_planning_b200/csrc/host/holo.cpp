// Host side of the HolonomicBlend PTG (product code). See ptg.h.
//
// Only the path geometry the planner queries lives here (reference: src/ptgs/HolonomicBlend.cpp
// getPathPose :535-590, getPathDist :592-623, getPathStepForDist :625-717, getPathStepCount
// :514-533, inverseMap_WS2TP_with_Tramp :309-484, internal_params_from_dir_and_dynstate :942-973).
// The arithmetic keeps the reference's operation order: these values decide lattice cells and
// the free/blocked comparison, so they must round identically.
#include <cstring>
#include <algorithm>
#include <limits>

#include "ptg.h"

namespace mpp
{
namespace
{
// real roots of a x^2 + b x + c (mrpt::math::solve_poly2): 0, 1 or 2 roots, ascending
int quadratic_roots(double a, double b, double c, double& lo, double& hi)
{
    constexpr double tiny = 1e-14;
    if (std::fabs(a) < tiny)
    {
        if (std::fabs(b) < tiny) return 0;
        lo = -c / b;
        hi = 1e99;
        return 1;
    }
    double disc = b * b - 4 * a * c;
    if (disc < 0) return 0;
    disc = std::sqrt(disc);
    lo   = (-b + disc) / (2 * a);
    hi   = (-b - disc) / (2 * a);
    if (hi < lo) std::swap(lo, hi);
    return 2;
}

// solve M x = rhs for a 4x4 system (partial pivoting); M and rhs are overwritten
void solve4(double M[4][4], double rhs[4], double x[4])
{
    for (int c = 0; c < 4; c++)
    {
        int best = c;
        for (int r = c + 1; r < 4; r++)
            if (std::fabs(M[r][c]) > std::fabs(M[best][c])) best = r;
        if (best != c)
        {
            for (int j = 0; j < 4; j++) std::swap(M[c][j], M[best][j]);
            std::swap(rhs[c], rhs[best]);
        }
        for (int r = c + 1; r < 4; r++)
        {
            const double f = M[r][c] / M[c][c];
            for (int j = c; j < 4; j++) M[r][j] -= f * M[c][j];
            rhs[r] -= f * rhs[c];
        }
    }
    for (int r = 3; r >= 0; r--)
    {
        double acc = rhs[r];
        for (int j = r + 1; j < 4; j++) acc -= M[r][j] * x[j];
        x[r] = acc / M[r][r];
    }
}
}  // namespace

HolonomicBlend::HolonomicBlend(const HolonomicBlendParams& p) : p_(p)
{
    MPP_ASSERT(p_.T_ramp_max > 0 && p_.v_max_mps > 0 && p_.w_max_rps > 0);
    MPP_ASSERT(p_.num_paths > 0 && p_.num_paths < 65536 && p_.robot_radius > 0 && p_.refDistance > 0);
    numPaths_    = p_.num_paths;
    refDistance_ = p_.refDistance;
    for (RuntimeExpr* e : {&exprV_, &exprW_})
    {
        e->bind("trimmable_speed", &trimmableSpeed_);
        e->bind("dir", &exprDir_);
        e->bind("target_dir", &targetDir_);
        e->bind("target_dist", &targetDist_);
        e->bind("V_MAX", &p_.v_max_mps);
        e->bind("W_MAX", &p_.w_max_rps);
        e->bind("T_ramp_max", &p_.T_ramp_max);
        e->bind("target_x", &dynState_.relTarget.x);
        e->bind("target_y", &dynState_.relTarget.y);
        e->bind("target_phi", &dynState_.relTarget.phi);
        e->bind("vxi", &dynState_.curVelLocal.vx);
        e->bind("vyi", &dynState_.curVelLocal.vy);
        e->bind("wi", &dynState_.curVelLocal.omega);
        e->bind("target_rel_speed", &dynState_.targetRelSpeed);
    }
    exprV_.compile(p_.expr_V.empty() ? "V_MAX" : p_.expr_V);
    exprW_.compile(p_.expr_W.empty() ? "W_MAX" : p_.expr_W);
    // do |V| and |omega| read the dynamic state? If not (the shipped configurations: functions of dir and the trimmed
    // speed only), memoised results stay valid when the state changes
    exprsReadDynState_ = false;
    for (const RuntimeExpr* e : {&exprV_, &exprW_})
        for (const double* v : {&targetDir_, &targetDist_, &dynState_.relTarget.x, &dynState_.relTarget.y,
                                &dynState_.relTarget.phi, &dynState_.curVelLocal.vx, &dynState_.curVelLocal.vy,
                                &dynState_.curVelLocal.omega, &dynState_.targetRelSpeed})
            if (e->references(v)) exprsReadDynState_ = true;
    memoGen_++;
    initialized_ = true;
}

void HolonomicBlend::onNewNavDynamicState()
{
    stepCountCache_.assign(numPaths_, -1);
    if (exprsReadDynState_) memoGen_++;  // drops every memoised paramsForDir() result
    targetDir_  = target_k_ >= 0 ? index2alpha(static_cast<unsigned>(target_k_)) : .0;
    targetDist_ = dynState_.relTarget.norm();
}

HolonomicBlend::Internal HolonomicBlend::paramsForDir(double dir) const
{
    uint64_t db, sb;
    std::memcpy(&db, &dir, sizeof(db));
    std::memcpy(&sb, &trimmableSpeed_, sizeof(sb));
    Memo& m = memo_[((db * 0x9E3779B97F4A7C15ull) ^ (sb * 0xC2B2AE3D27D4EB4Full)) >> 55];
    if (m.gen == memoGen_ && dir == m.dir && trimmableSpeed_ == m.speed)
    {
        Internal q = m.q;  // vf, wf, vxf, vyf, T_ramp; the initial velocity is the current state's
        q.vxi      = dynState_.curVelLocal.vx;
        q.vyi      = dynState_.curVelLocal.vy;
        return q;
    }
    Internal q;
    q.T_ramp = p_.T_ramp_max;
    exprDir_ = dir;
    q.vf     = std::abs(exprV_.eval());
    q.wf     = sign_or_zero(dir) * std::abs(exprW_.eval());
    q.vxi    = dynState_.curVelLocal.vx;
    q.vyi    = dynState_.curVelLocal.vy;
    q.vxf    = q.vf * std::cos(dir);
    q.vyf    = q.vf * std::sin(dir);
    m.q      = q;
    m.dir    = dir;
    m.speed  = trimmableSpeed_;
    m.gen    = memoGen_;
    return q;
}

void HolonomicBlend::holoParams(uint16_t k, mppb_holo_cand& out) const
{
    const Internal q = paramsForDir(index2alpha(k));
    out.T_ramp       = q.T_ramp;
    out.vxi          = q.vxi;
    out.vyi          = q.vyi;
    out.vxf          = q.vxf;
    out.vyf          = q.vyf;
    out.vf           = q.vf;
}

// trapezoid rule, 20 panels, of sqrt(a t^2 + b t + c) on [0,T]
double HolonomicBlend::rampIntegral(double T, double a, double b, double c)
{
    MPP_ASSERT(a >= .0 && c >= .0);
    constexpr unsigned panels = 20;
    const double       h      = T / (panels);
    double             prev = std::sqrt(c), t = .0, sum = .0;
    for (unsigned i = 0; i < panels; i++)
    {
        t += h;
        double radicand = a * t * t + b * t + c;
        MPP_ASSERT(radicand > -1e-5);
        if (radicand < 0) radicand = .0;
        const double cur = std::sqrt(radicand);
        sum += h * (prev + cur) * 0.5;
        prev = cur;
    }
    return sum;
}

// path length travelled during the velocity ramp, up to time t
double HolonomicBlend::rampDistance(double k2, double k4, double vxi, double vyi, double t)
{
    const double c = (vxi * vxi + vyi * vyi);
    if (!(std::abs(k2) > eps || std::abs(k4) > eps)) return std::sqrt(c) * t;
    const double a = ((k2 * k2) * 4.0 + (k4 * k4) * 4.0);
    const double b = (k2 * vxi * 4.0 + k4 * vyi * 4.0);
    if (std::abs(b) < eps && std::abs(c) < eps) return std::sqrt(a) * (t * t) * 0.5;
    return rampIntegral(t, a, b, c);
}

double HolonomicBlend::pathDistAt(uint32_t step, double T_ramp, double vxf, double vyf) const
{
    const double t    = PATH_TIME_STEP * step;
    const double TR2_ = 1.0 / (2 * T_ramp);
    const double vxi = dynState_.curVelLocal.vx, vyi = dynState_.curVelLocal.vy;
    const double k2 = (vxf - vxi) * TR2_, k4 = (vyf - vyi) * TR2_;
    if (t < T_ramp) return rampDistance(k2, k4, vxi, vyi, t);
    return (t - T_ramp) * p_.v_max_mps + rampDistance(k2, k4, vxi, vyi, T_ramp);
}

double HolonomicBlend::getPathDist(uint16_t k, uint32_t step) const
{
    const Internal q = paramsForDir(index2alpha(k));
    return pathDistAt(step, q.T_ramp, q.vxf, q.vyf);
}

// The step count is cached per k and only dropped when the dynamic state changes — NOT when the
// trimmed speed changes (reference behaviour, HolonomicBlend.cpp:182-185,514-533): callers see
// the count computed at whatever speed was current on the first query.
size_t HolonomicBlend::getPathStepCount(uint16_t k) const
{
    if (stepCountCache_.size() > k && stepCountCache_[k] > 0) return stepCountCache_[k];
    uint32_t step = 0;
    if (!getPathStepForDist(k, refDistance_, step))
        throw std::runtime_error("Could not solve closed-form distance for k=" + std::to_string(k));
    MPP_ASSERT(step > 0);
    if (stepCountCache_.size() != numPaths_) stepCountCache_.assign(numPaths_, -1);
    stepCountCache_[k] = static_cast<int>(step);
    return step;
}

TPose2D HolonomicBlend::getPathPose(uint16_t k, uint32_t step) const
{
    const double   t    = PATH_TIME_STEP * step;
    const double   dir  = index2alpha(k);
    const Internal q    = paramsForDir(dir);
    const double   TR2_ = 1.0 / (2 * q.T_ramp);
    const double   wi   = dynState_.curVelLocal.omega;
    TPose2D        p;
    if (t < q.T_ramp)
    {
        p.x = q.vxi * t + t * t * TR2_ * (q.vxf - q.vxi);
        p.y = q.vyi * t + t * t * TR2_ * (q.vyf - q.vyi);
        // heading: accelerate towards wf until aligned with the motion direction
        double    r1 = 0, r2 = 0;
        const int n = quadratic_roots(TR2_ * (q.wf - wi), wi, -dir, r1, r2);
        if (n != 2)
            p.phi = .0;
        else if (t > std::max(r1, r2))
            p.phi = dir;
        else
            p.phi = wi * t + t * t * TR2_ * (q.wf - wi);
    }
    else
    {
        p.x = q.T_ramp * 0.5 * (q.vxi + q.vxf) + (t - q.T_ramp) * q.vxf;
        p.y = q.T_ramp * 0.5 * (q.vyi + q.vyf) + (t - q.T_ramp) * q.vyf;
        const double tAligned = (dir - q.T_ramp * 0.5 * (wi + q.wf)) / q.wf + q.T_ramp;
        if (t > tAligned)
            p.phi = dir;
        else
            p.phi = q.T_ramp * 0.5 * (wi + q.wf) + (t - q.T_ramp) * q.wf;
    }
    return p;
}

bool HolonomicBlend::getPathStepForDist(uint16_t k, double dist, uint32_t& out_step) const
{
    const Internal q    = paramsForDir(index2alpha(k));
    const double   TR2_ = 1.0 / (2 * q.T_ramp);
    const double   k2 = (q.vxf - q.vxi) * TR2_, k4 = (q.vyf - q.vyi) * TR2_;
    const double   distAtRampEnd = rampDistance(k2, k4, q.vxi, q.vyi, q.T_ramp);
    double         t             = -1;
    if (dist >= distAtRampEnd)
        t = q.T_ramp + (dist - distAtRampEnd) / p_.v_max_mps;
    else if (std::abs(k2) < eps && std::abs(k4) < eps)
        t = (dist) / p_.v_max_mps;
    else
    {
        const double a = ((k2 * k2) * 4.0 + (k4 * k4) * 4.0);
        const double b = (k2 * q.vxi * 4.0 + k4 * q.vyi * 4.0);
        const double c = (q.vxi * q.vxi + q.vyi * q.vyi);
        if (std::abs(b) < eps && std::abs(c) < eps)
            t = std::sqrt(2.0) * 1.0 / std::pow(a, 1.0 / 4.0) * std::sqrt(dist);
        else
        {
            // Newton on f(t) = rampIntegral(t) - dist, f'(t) = sqrt(a t^2 + b t + c)
            t = q.T_ramp * 0.6;
            for (int it = 0; it < 10; it++)
            {
                const double err   = rampIntegral(t, a, b, c) - dist;
                const double slope = std::sqrt(a * t * t + b * t + c);
                MPP_ASSERT(std::abs(slope) > 1e-40);
                t -= (err) / slope;
                if (t < 0) t = .0;
                if (std::abs(err) < 1e-3) break;
            }
        }
    }
    if (t >= 0)
    {
        out_step = static_cast<uint32_t>(round_i(t / PATH_TIME_STEP));
        return true;
    }
    return false;
}

// Newton iterations on q = [t, vxf, vyf, T_ramp] for the trajectory that passes through (x,y)
bool HolonomicBlend::inverseMap_WS2TP(double x, double y, int& out_k, double& out_d, double /*tol*/) const
{
    MPP_ASSERT(x != 0 || y != 0);
    constexpr double kStopSpeed = 0.10 * 1.05;  // relative speed below which "reach and stop" applies
    constexpr double kErrMax    = 1e-3;
    const double     vxi = dynState_.curVelLocal.vx, vyi = dynState_.curVelLocal.vy;
    const double     Tmax = p_.T_ramp_max, Vmax = p_.v_max_mps;
    double           q[4] = {Tmax * 1.1, Vmax * x / std::sqrt(x * x + y * y), Vmax * y / std::sqrt(x * x + y * y), Tmax};
    bool             found = false;
    for (int it = 0; !found && it < 25; it++)
    {
        const double t = q[0], vxf = q[1], vyf = q[2];
        exprDir_                = std::atan2(vyf, vxf);
        const double vAbs       = std::abs(exprV_.eval());
        const double vSq        = vAbs * vAbs;
        const bool   stopAtGoal = vSq < kStopSpeed * kStopSpeed;
        const double Tr         = q[3];
        const double TR_ = 1.0 / (Tr), TR2_ = 1.0 / (2 * Tr);
        double       r[4];
        double       J[4][4] = {{0}};
        const bool   cruise  = t >= Tr;
        if (cruise)
        {
            r[0]    = 0.5 * Tr * (vxi + vxf) + (t - Tr) * vxf - x;
            r[1]    = 0.5 * Tr * (vyi + vyf) + (t - Tr) * vyf - y;
            J[0][0] = vxf;
            J[0][1] = 0.5 * Tr + t;
            J[1][0] = vyf;
            J[1][2] = 0.5 * Tr + t;
        }
        else
        {
            r[0]    = vxi * t + t * t * TR2_ * (vxf - vxi) - x;
            r[1]    = vyi * t + t * t * TR2_ * (vyf - vyi) - y;
            J[0][0] = vxi + t * TR_ * (vxf - vxi);
            J[0][1] = TR2_ * t * t;
            J[1][0] = vyi + t * TR_ * (vyf - vyi);
            J[1][2] = TR2_ * t * t;
        }
        r[2] = vxf * vxf + vyf * vyf - vSq;
        r[3] = stopAtGoal ? Tr - t : 0;
        if (stopAtGoal)
        {
            if (cruise)
            {
                J[0][3] = 0.5 * (vxi - vxf);
                J[1][3] = 0.5 * (vyi - vyf);
            }
            else
            {
                J[0][3] = -t * t * TR2_ * (vxf - vxi);
                J[1][3] = -t * t * TR2_ * (vyf - vyi);
            }
            J[3][0] = -1;
            J[3][3] = 1;
        }
        else
        {
            q[3]    = Tmax;  // keep the prescribed ramp time
            J[3][3] = 1;
        }
        J[2][1]            = 2 * vxf;
        J[2][2]            = 2 * vyf;
        const double rnorm = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
        double       dq[4];
        solve4(J, r, dq);
        for (int i = 0; i < 4; i++) q[i] -= dq[i];
        found = rnorm < kErrMax;
    }
    if (!(found && q[0] >= .0)) return false;
    out_k               = alpha2index(std::atan2(q[2], q[1]));
    const unsigned step = static_cast<unsigned>(q[0] / PATH_TIME_STEP);
    out_d               = pathDistAt(step, q[3], q[1], q[2]) / refDistance_;
    return true;
}

}  // namespace mpp
