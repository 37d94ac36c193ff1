// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product.
//
// CPU restatement of mpp::ptg::HolonomicBlend
// (/root/reference/mrpt_path_planning/src/ptgs/HolonomicBlend.cpp; each method cites the lines
// it follows). MRPT pieces it leans on (solve_poly*, the expression compiler, the circular
// robot shape) are restated in orc_poly.hpp / orc_expr.hpp / below and marked [MRPT].
// Pinned against HolonomicBlend.cpp itself, compiled as oracle/_ref (tests/test_oracle_vs_reference.py);
// the MRPT layer underneath is assumed.
#pragma once
#include "orc_expr.hpp"
#include "orc_poly.hpp"
#include "orc_ptg.hpp"

namespace orc
{
class HolonomicBlend : public PTG
{
   public:
    // HolonomicBlend.cpp:57-58
    static constexpr double PATH_TIME_STEP = 10e-3;
    static constexpr double eps            = 1e-4;

    double      T_ramp_max = -1.0, V_MAX = -1.0, W_MAX = -1.0;
    double      turningRadiusReference = 0.30;
    double      robotRadius            = 0;
    std::string expr_V = "V_MAX", expr_W = "W_MAX", expr_T_ramp = "T_ramp_max";  // :908-911

    struct InternalParams
    {
        double vxi = 0, vyi = 0, vxf = 0, vyf = 0, vf = 0, wf = 0, T_ramp = 1;
    };

    HolonomicBlend()
    {
        // :889-906 symbol table (pointers into this object)
        const std::map<std::string, double*> vars = {
            {"trimmable_speed", &trimmableSpeed},
            {"dir", &m_expr_dir},
            {"target_dir", &m_expr_target_dir},
            {"target_dist", &m_expr_target_dist},
            {"V_MAX", &V_MAX},
            {"W_MAX", &W_MAX},
            {"T_ramp_max", &T_ramp_max},
            {"target_x", &dynState.relTarget.x},
            {"target_y", &dynState.relTarget.y},
            {"target_phi", &dynState.relTarget.phi},
            {"vxi", &dynState.curVelLocal.vx},
            {"vyi", &dynState.curVelLocal.vy},
            {"wi", &dynState.curVelLocal.omega},
            {"target_rel_speed", &dynState.targetRelSpeed}};
        m_expr_v.register_symbol_table(vars);
        m_expr_w.register_symbol_table(vars);
    }
    HolonomicBlend(const HolonomicBlend&)            = delete;
    HolonomicBlend& operator=(const HolonomicBlend&) = delete;

    // :915-940
    void initialize()
    {
        ORC_ASSERT(T_ramp_max > 0 && V_MAX > 0 && W_MAX > 0 && numPaths > 0 && robotRadius > 0);
        m_expr_v.compile(expr_V);
        m_expr_w.compile(expr_W);
        m_pathStepCountCache.clear();
    }

    bool isSpeedTrimmable() const override { return true; }      // HolonomicBlend.h:33
    bool supportSpeedAtTarget() const override { return true; }  // HolonomicBlend.h:62

    // [MRPT] CPTG_RobotShape_Circular
    bool   isPointInsideRobotShape(double x, double y) const override { return ::hypot(x, y) < robotRadius; }
    double getMaxRobotRadius() const override { return robotRadius; }

    // :182-192
    void onNewNavDynamicState() override
    {
        m_pathStepCountCache.assign(numPaths, -1);
        m_expr_target_dir = .0;
        if (target_k != INVALID_PATH_INDEX) m_expr_target_dir = index2alpha(target_k);
        m_expr_target_dist = dynState.relTarget.norm();
    }

    // :942-973
    InternalParams internal_params_from_dir_and_dynstate(const double dir) const
    {
        InternalParams p;
        p.T_ramp   = T_ramp_max;
        m_expr_dir = dir;
        p.vf       = std::abs(m_expr_v.eval());
        p.wf       = sign_with_zero(dir) * std::abs(m_expr_w.eval());
        p.vxi      = dynState.curVelLocal.vx;
        p.vyi      = dynState.curVelLocal.vy;
        p.vxf      = p.vf * std::cos(dir);
        p.vyf      = p.vf * std::sin(dir);
        return p;
    }

    // :96-131
    static double calc_trans_distance_t_below_Tramp_abc(double T, double a, double b, double c)
    {
        double             d         = .0;
        const unsigned int NUM_STEPS = 20;
        ORC_ASSERT(a >= .0);
        ORC_ASSERT(c >= .0);
        double       feval_t = std::sqrt(c);
        double       feval_tp1;
        const double At = T / (NUM_STEPS);
        double       t  = .0;
        for (unsigned int i = 0; i < NUM_STEPS; i++)
        {
            t += At;
            double dd = a * t * t + b * t + c;
            ORC_ASSERT(dd > -1e-5);
            if (dd < 0) dd = .0;
            feval_tp1 = std::sqrt(dd);
            d += At * (feval_t + feval_tp1) * 0.5;
            feval_t = feval_tp1;
        }
        return d;
    }
    // :150-180
    static double calc_trans_distance_t_below_Tramp(double k2, double k4, double vxi, double vyi, double t)
    {
        const double c = (vxi * vxi + vyi * vyi);
        if (std::abs(k2) > eps || std::abs(k4) > eps)
        {
            const double a = ((k2 * k2) * 4.0 + (k4 * k4) * 4.0);
            const double b = (k2 * vxi * 4.0 + k4 * vyi * 4.0);
            if (std::abs(b) < eps && std::abs(c) < eps) return std::sqrt(a) * (t * t) * 0.5;
            return calc_trans_distance_t_below_Tramp_abc(t, a, b, c);
        }
        return std::sqrt(c) * t;
    }

    // :514-533 (the cache is NOT invalidated when trimmableSpeed changes — reproduced on purpose)
    size_t getPathStepCount(unsigned k) const override
    {
        if (m_pathStepCountCache.size() > k && m_pathStepCountCache[k] > 0) return m_pathStepCountCache[k];
        uint32_t step = 0;
        if (!getPathStepForDist(k, refDistance, step))
            throw std::runtime_error("Could not solve closed-form distance for k=" + std::to_string(k));
        ORC_ASSERT(step > 0);
        if (m_pathStepCountCache.size() != numPaths) m_pathStepCountCache.assign(numPaths, -1);
        m_pathStepCountCache[k] = step;
        return step;
    }

    // :535-590
    TPose2D getPathPose(unsigned k, uint32_t step) const override
    {
        const double t    = PATH_TIME_STEP * step;
        const double dir  = index2alpha(k);
        const auto   _    = internal_params_from_dir_and_dynstate(dir);
        const double TR2_ = 1.0 / (2 * _.T_ramp);
        TPose2D      p;
        if (t < _.T_ramp)
        {
            p.x = _.vxi * t + t * t * TR2_ * (_.vxf - _.vxi);
            p.y = _.vyi * t + t * t * TR2_ * (_.vyf - _.vyi);
        }
        else
        {
            p.x = _.T_ramp * 0.5 * (_.vxi + _.vxf) + (t - _.T_ramp) * _.vxf;
            p.y = _.T_ramp * 0.5 * (_.vyi + _.vyf) + (t - _.T_ramp) * _.vyf;
        }
        const double wi = dynState.curVelLocal.omega;
        if (t < _.T_ramp)
        {
            const double a = TR2_ * (_.wf - wi), b = (wi), c = -dir;
            double       r1 = 0, r2 = 0;
            const int    nroots = poly::solve_poly2(a, b, c, r1, r2);
            if (nroots != 2)
                p.phi = .0;
            else
            {
                const double t_solve = std::max(r1, r2);
                if (t > t_solve)
                    p.phi = dir;
                else
                    p.phi = wi * t + t * t * TR2_ * (_.wf - wi);
            }
        }
        else
        {
            const double t_solve = (dir - _.T_ramp * 0.5 * (wi + _.wf)) / _.wf + _.T_ramp;
            if (t > t_solve)
                p.phi = dir;
            else
                p.phi = _.T_ramp * 0.5 * (wi + _.wf) + (t - _.T_ramp) * _.wf;
        }
        return p;
    }

    // :592-623
    double getPathDist(unsigned k, uint32_t step) const override
    {
        const auto pp = internal_params_from_dir_and_dynstate(index2alpha(k));
        return internal_getPathDist(step, pp.T_ramp, pp.vxf, pp.vyf);
    }
    double internal_getPathDist(uint32_t step, double T_ramp, double vxf, double vyf) const
    {
        const double t    = PATH_TIME_STEP * step;
        const double TR2_ = 1.0 / (2 * T_ramp);
        const double vxi = dynState.curVelLocal.vx, vyi = dynState.curVelLocal.vy;
        const double k2 = (vxf - vxi) * TR2_, k4 = (vyf - vyi) * TR2_;
        if (t < T_ramp) return calc_trans_distance_t_below_Tramp(k2, k4, vxi, vyi, t);
        return (t - T_ramp) * V_MAX + calc_trans_distance_t_below_Tramp(k2, k4, vxi, vyi, T_ramp);
    }

    // :625-717
    bool getPathStepForDist(unsigned k, double dist, uint32_t& out_step) const override
    {
        const double dir  = index2alpha(k);
        const auto   _    = internal_params_from_dir_and_dynstate(dir);
        const double TR2_ = 1.0 / (2 * _.T_ramp);
        const double k2 = (_.vxf - _.vxi) * TR2_, k4 = (_.vyf - _.vyi) * TR2_;
        const double dist_trans_T_ramp = calc_trans_distance_t_below_Tramp(k2, k4, _.vxi, _.vyi, _.T_ramp);
        double       t_solved          = -1;
        if (dist >= dist_trans_T_ramp)
            t_solved = _.T_ramp + (dist - dist_trans_T_ramp) / V_MAX;
        else
        {
            if (std::abs(k2) < eps && std::abs(k4) < eps)
                t_solved = (dist) / V_MAX;
            else
            {
                const double a = ((k2 * k2) * 4.0 + (k4 * k4) * 4.0);
                const double b = (k2 * _.vxi * 4.0 + k4 * _.vyi * 4.0);
                const double c = (_.vxi * _.vxi + _.vyi * _.vyi);
                if (std::abs(b) < eps && std::abs(c) < eps)
                    t_solved = std::sqrt(2.0) * 1.0 / std::pow(a, 1.0 / 4.0) * std::sqrt(dist);
                else
                {
                    t_solved = _.T_ramp * 0.6;
                    for (int iters = 0; iters < 10; iters++)
                    {
                        double       err  = calc_trans_distance_t_below_Tramp_abc(t_solved, a, b, c) - dist;
                        const double diff = std::sqrt(a * t_solved * t_solved + b * t_solved + c);
                        ORC_ASSERT(std::abs(diff) > 1e-40);
                        t_solved -= (err) / diff;
                        if (t_solved < 0) t_solved = .0;
                        if (std::abs(err) < 1e-3) break;
                    }
                }
            }
        }
        if (t_solved >= 0)
        {
            out_step = mrpt_round(t_solved / PATH_TIME_STEP);
            return true;
        }
        return false;
    }
    double getPathStepDuration() const override { return PATH_TIME_STEP; }

    // :301-484. The 4x4 linear solve is mrpt CMatrixFixed::lu_solve (Eigen partial-pivot LU) [MRPT]:
    // restated as Gaussian elimination with partial pivoting.
    bool inverseMap_WS2TP(double x, double y, int& out_k, double& out_d, double /*tol*/) const override
    {
        double dummy;
        return inverseMap_WS2TP_with_Tramp(x, y, out_k, out_d, dummy);
    }
    bool inverseMap_WS2TP_with_Tramp(double x, double y, int& out_k, double& out_d, double& out_T_ramp) const
    {
        ORC_ASSERT(x != 0 || y != 0);
        const double REL_SPEED_TO_CONSIDER_REACH_AND_STOP = 0.10 * 1.05;
        const double err_threshold                        = 1e-3;
        const double vxi = dynState.curVelLocal.vx, vyi = dynState.curVelLocal.vy;
        double       q[4];
        q[0]             = T_ramp_max * 1.1;
        q[1]             = V_MAX * x / std::sqrt(x * x + y * y);
        q[2]             = V_MAX * y / std::sqrt(x * x + y * y);
        q[3]             = T_ramp_max;
        double err_mod   = std::numeric_limits<double>::max();
        bool   sol_found = false;
        for (int iters = 0; !sol_found && iters < 25; iters++)
        {
            const double t   = q[0];
            const double vxf = q[1], vyf = q[2];
            const double alpha = std::atan2(vyf, vxf);
            m_expr_dir         = alpha;
            const double V_MAXsq      = square(std::abs(m_expr_v.eval()));
            const bool   stopAtTarget = V_MAXsq < square(REL_SPEED_TO_CONSIDER_REACH_AND_STOP);
            const double T_ramp       = q[3];
            const double TR_ = 1.0 / (T_ramp), TR2_ = 1.0 / (2 * T_ramp);
            double       r[4];
            if (t >= T_ramp)
            {
                r[0] = 0.5 * T_ramp * (vxi + vxf) + (t - T_ramp) * vxf - x;
                r[1] = 0.5 * T_ramp * (vyi + vyf) + (t - T_ramp) * vyf - y;
            }
            else
            {
                r[0] = vxi * t + t * t * TR2_ * (vxf - vxi) - x;
                r[1] = vyi * t + t * t * TR2_ * (vyf - vyi) - y;
            }
            r[2] = vxf * vxf + vyf * vyf - V_MAXsq;
            r[3] = stopAtTarget ? T_ramp - t : 0;
            double J[4][4] = {{0}};
            if (t >= T_ramp)
            {
                J[0][0] = vxf;
                J[0][1] = 0.5 * T_ramp + t;
                J[1][0] = vyf;
                J[1][2] = 0.5 * T_ramp + t;
                if (stopAtTarget)
                {
                    J[0][3] = 0.5 * (vxi - vxf);
                    J[1][3] = 0.5 * (vyi - vyf);
                }
                else
                {
                    q[3]    = T_ramp_max;
                    J[3][3] = 1;
                }
            }
            else
            {
                J[0][0] = vxi + t * TR_ * (vxf - vxi);
                J[0][1] = TR2_ * t * t;
                J[1][0] = vyi + t * TR_ * (vyf - vyi);
                J[1][2] = TR2_ * t * t;
                if (stopAtTarget)
                {
                    J[0][3] = -t * t * TR2_ * (vxf - vxi);
                    J[1][3] = -t * t * TR2_ * (vyf - vyi);
                }
                else
                {
                    q[3]    = T_ramp_max;
                    J[3][3] = 1;
                }
            }
            if (stopAtTarget)
            {
                J[3][0] = -1;
                J[3][3] = 1;
            }
            J[2][1] = 2 * vxf;
            J[2][2] = 2 * vyf;
            double q_incr[4];
            lu_solve4(J, r, q_incr);
            for (int i = 0; i < 4; i++) q[i] -= q_incr[i];
            err_mod   = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
            sol_found = (err_mod < err_threshold);
        }
        if (sol_found && q[0] >= .0)
        {
            const double alpha = std::atan2(q[2], q[1]);
            out_k              = alpha2index(alpha);
            out_T_ramp         = q[3];
            const double       t = q[0], vxf = q[1], vyf = q[2];
            const unsigned int solved_step = t / PATH_TIME_STEP;
            const double       found_dist  = internal_getPathDist(solved_step, out_T_ramp, vxf, vyf);
            out_d                          = found_dist / refDistance;
            return true;
        }
        return false;
    }

    // :719-840
    void updateTPObstacleSingle(double ox, double oy, unsigned k, double& tp_obstacle_k) const override
    {
        const double R               = robotRadius;
        const double dir             = index2alpha(k);
        const auto   _               = internal_params_from_dir_and_dynstate(dir);
        const double TR2_            = 1.0 / (2 * _.T_ramp);
        const double TR_2            = _.T_ramp * 0.5;
        const double T_ramp_thres099 = _.T_ramp * 0.99;
        const double T_ramp_thres101 = _.T_ramp * 1.01;
        double       sol_t           = -1.0;
        const double k2 = (_.vxf - _.vxi) * TR2_, k4 = (_.vyf - _.vyi) * TR2_;
        const double a = (k2 * k2 + k4 * k4);
        const double b = (k2 * _.vxi * 2.0 + k4 * _.vyi * 2.0);
        const double c = -(k2 * ox * 2.0 + k4 * oy * 2.0 - _.vxi * _.vxi - _.vyi * _.vyi);
        const double d = -(ox * _.vxi * 2.0 + oy * _.vyi * 2.0);
        const double e = -R * R + ox * ox + oy * oy;
        double       roots[4];
        int          num_real_sols = 0;
        if (std::abs(a) > eps)
            num_real_sols = poly::solve_poly4(roots, b / a, c / a, d / a, e / a);
        else if (std::abs(b) > eps)
            num_real_sols = poly::solve_poly3(roots, c / b, d / b, e / b);
        else
        {
            const double discr = d * d - 4 * c * e;
            if (discr >= 0)
            {
                num_real_sols = 2;
                roots[0]      = (-d + std::sqrt(discr)) / (2 * c);
                roots[1]      = (-d - std::sqrt(discr)) / (2 * c);
            }
            else
                num_real_sols = 0;
        }
        for (int i = 0; i < num_real_sols; i++)
        {
            if (roots[i] == roots[i] && std::isfinite(roots[i]) && roots[i] >= .0 && roots[i] <= _.T_ramp * 1.01)
            {
                if (sol_t < 0)
                    sol_t = roots[i];
                else if (roots[i] < sol_t)
                    sol_t = roots[i];
            }
        }
        if (sol_t < 0 || sol_t > T_ramp_thres101)
        {
            sol_t            = -1.0;
            const double c1  = TR_2 * (_.vxi - _.vxf) - ox;
            const double c2  = TR_2 * (_.vyi - _.vyf) - oy;
            const double xa  = _.vf * _.vf;
            const double xb  = 2 * (c1 * _.vxf + c2 * _.vyf);
            const double xc  = c1 * c1 + c2 * c2 - R * R;
            const double dsc = xb * xb - 4 * xa * xc;
            if (dsc >= 0)
            {
                const double sol_t0 = (-xb + std::sqrt(dsc)) / (2 * xa);
                const double sol_t1 = (-xb - std::sqrt(dsc)) / (2 * xa);
                if (sol_t0 < _.T_ramp && sol_t1 < _.T_ramp)
                    sol_t = -1.0;
                else if (sol_t0 < _.T_ramp && sol_t1 >= T_ramp_thres099)
                    sol_t = sol_t1;
                else if (sol_t1 < _.T_ramp && sol_t0 >= T_ramp_thres099)
                    sol_t = sol_t0;
                else if (sol_t1 >= T_ramp_thres099 && sol_t0 >= T_ramp_thres099)
                    sol_t = std::min(sol_t0, sol_t1);
            }
        }
        if (sol_t < 0) return;
        double dist;
        if (sol_t < _.T_ramp)
            dist = calc_trans_distance_t_below_Tramp(k2, k4, _.vxi, _.vyi, sol_t);
        else
            dist = (sol_t - _.T_ramp) * V_MAX + calc_trans_distance_t_below_Tramp(k2, k4, _.vxi, _.vyi, _.T_ramp);
        tpObsPostprocess(ox, oy, dist, tp_obstacle_k);
    }

   private:
    // x = J^-1 r by Gaussian elimination with partial (row) pivoting
    static void lu_solve4(double J[4][4], const double r[4], double x[4])
    {
        double A[4][5];
        for (int i = 0; i < 4; i++)
        {
            for (int j = 0; j < 4; j++) A[i][j] = J[i][j];
            A[i][4] = r[i];
        }
        for (int col = 0; col < 4; col++)
        {
            int piv = col;
            for (int i = col + 1; i < 4; i++)
                if (std::fabs(A[i][col]) > std::fabs(A[piv][col])) piv = i;
            if (piv != col)
                for (int j = 0; j < 5; j++) std::swap(A[col][j], A[piv][j]);
            for (int i = col + 1; i < 4; i++)
            {
                const double f = A[i][col] / A[col][col];
                for (int j = col; j < 5; j++) A[i][j] -= f * A[col][j];
            }
        }
        for (int i = 3; i >= 0; i--)
        {
            double s = A[i][4];
            for (int j = i + 1; j < 4; j++) s -= A[i][j] * x[j];
            x[i] = s / A[i][i];
        }
    }

    mutable std::vector<int> m_pathStepCountCache;
    Expression               m_expr_v, m_expr_w;
    mutable double           m_expr_dir         = 0;
    double                   m_expr_target_dir  = 0;
    double                   m_expr_target_dist = 0;
};

}  // namespace orc
