// tpobs_holo_kernel — stage (a)+(b) of the hot path for HolonomicBlend PTGs, fused.
//
// Replaces, for a batch of candidates (lattice node, trajectory k, trimmed speed), the reference's
//   cached_local_obstacles -> transform_pc_square_clipping         (TPS_Astar.cpp:561-562,
//                                                                   transform_pc_square_clipping.cpp:26-42)
//   tp_obstacles_single_path -> HolonomicBlend::updateTPObstacleSingle per obstacle point
//                                                                  (tp_obstacles_single_path.cpp:26-34,
//                                                                   HolonomicBlend.cpp:719-840)
// out[cand] = min over the node's in-window points of the TP-space collision distance of the
// candidate's trajectory (+INF when no point constrains it); the caller applies
// initTPObstacleSingle's value. The per-candidate velocity parameters (vxi, vyi, vxf, vyf, vf,
// T_ramp) are evaluated on the host exactly as internal_params_from_dir_and_dynstate
// (HolonomicBlend.cpp:942-973) does, because they involve user expressions and libm cos/sin.
//
// Geometric pre-filter: a point can only change the result if it comes within the robot radius of the robot
// centre's path, i.e. of C = {ramp parabola, t in [0, 1.01 T]} U {cruise line, t >= 0.99 T} (the time ranges the
// reference accepts roots in, HolonomicBlend.cpp:783-825). The parabola is a quadratic Bezier curve and lies in the
// triangle of its control points; the cruise part is a ray. Each in-window point is first transformed in FP32
// (hi/lo split of the node position) and tested against that triangle and ray with a 1 cm safety margin over the
// radius (FP32 error of the test ~1e-5 m; root-finding slop of the exact path ~1e-12 m); only the few per cent of
// points that pass are queued per warp and run, 32 at a time, through the exact FP64 chain below. A point the
// filter drops would have left the minimum untouched, so results are bit-identical to the unfiltered kernel
// (tests/test_tpobs_holo_gpu.py::test_filter_is_conservative compares the two).
//
// Work decomposition: one warp per candidate. The warp walks the bin rows of the cloud index that
// overlap the node's clipping square; lanes take consecutive points (coalesced 8-byte loads),
// clip and transform them with the reference's FP64 operation order, round to float as
// CPointsMap stores them, and solve the time-to-collision quartic in FP64 (poly34.cuh). The
// per-candidate minimum is an order-independent min, reduced with warp shuffles.
// This file is compiled with -fmad=false (see build.py): all arithmetic rounds as on the
// reference's x86-64 build; acos/cos in the cubic's trigonometric branch are CUDA libm.
#include <algorithm>
#include <cfloat>

#include <math_constants.h>

#include "mppb_internal.cuh"
#include "poly34.cuh"

namespace mppb
{
namespace
{
constexpr int      kThreads = 128;
constexpr int      kWarps   = kThreads / 32;
constexpr unsigned kFull    = 0xffffffffu;
constexpr double   kEps     = 1e-4;  // HolonomicBlend::eps (HolonomicBlend.cpp:58)

// HolonomicBlend.cpp:96-131: 20-step trapezoid of sqrt(a t^2 + b t + c) over [0,T]
__device__ __forceinline__ double ramp_integral_numeric(double T, double a, double b, double c)
{
    double       d  = .0;
    double       f0 = sqrt(c);
    const double At = T / 20;
    double       t  = .0;
#pragma unroll 4
    for (int i = 0; i < 20; i++)
    {
        t += At;
        double dd = a * t * t + b * t + c;
        if (dd < 0) dd = .0;
        const double f1 = sqrt(dd);
        d += At * (f0 + f1) * 0.5;
        f0 = f1;
    }
    return d;
}
// The same sum, bit for bit, by a whole warp whose lanes hold the same arguments (the candidate prologue): the 21
// square roots — the expensive part of the scalar loop, 1.8 us per candidate — are taken by 21 lanes at once; the
// abscissae and the sum are still accumulated in the reference's order.
__device__ __forceinline__ double ramp_integral_numeric_warp(double T, double a, double b, double c, int lane)
{
    const double At = T / 20;
    double       t = .0, tMine = .0;
#pragma unroll
    for (int i = 0; i < 20; i++)
    {
        t += At;
        if (i + 1 == lane) tMine = t;  // lane i+1 takes the i-th step's f1; lane 0 takes f(0) = sqrt(c)
    }
    double dd = lane == 0 ? c : a * tMine * tMine + b * tMine + c;
    if (lane != 0 && dd < 0) dd = .0;
    const double f     = sqrt(dd);
    const double fPrev = __shfl_up_sync(kFull, f, 1);
    const double term  = At * (fPrev + f) * 0.5;  // lanes 1..20: the i-th step's increment, i = lane - 1
    double       d     = .0;
#pragma unroll
    for (int i = 1; i <= 20; i++) d += __shfl_sync(kFull, term, i);
    return d;
}
// HolonomicBlend.cpp:150-180
__device__ __forceinline__ double ramp_distance(double k2, double k4, double vxi, double vyi, double t)
{
    const double c = (vxi * vxi + vyi * vyi);
    if (fabs(k2) > kEps || fabs(k4) > kEps)
    {
        const double a = ((k2 * k2) * 4.0 + (k4 * k4) * 4.0);
        const double b = (k2 * vxi * 4.0 + k4 * vyi * 4.0);
        if (fabs(b) < kEps && fabs(c) < kEps) return sqrt(a) * (t * t) * 0.5;
        return ramp_integral_numeric(t, a, b, c);
    }
    return sqrt(c) * t;
}

// ramp_distance() evaluated by a warp with uniform arguments (same branches, same value in every lane)
__device__ __forceinline__ double ramp_distance_warp(double k2, double k4, double vxi, double vyi, double t, int lane)
{
    const double c = (vxi * vxi + vyi * vyi);
    if (fabs(k2) > kEps || fabs(k4) > kEps)
    {
        const double a = ((k2 * k2) * 4.0 + (k4 * k4) * 4.0);
        const double b = (k2 * vxi * 4.0 + k4 * vyi * 4.0);
        if (fabs(b) < kEps && fabs(c) < kEps) return sqrt(a) * (t * t) * 0.5;
        return ramp_integral_numeric_warp(t, a, b, c, lane);
    }
    return sqrt(c) * t;
}

struct CandConsts
{
    double T, vxi, vyi, vxf, vyf, vf;
    double k2, k4, qa, qb, TR_2, T099, T101, R, Vmax, distAtT;
};

// the per-candidate constants of HolonomicBlend::updateTPObstacleSingle (HolonomicBlend.cpp:722-749), by a whole warp
// whose lanes hold the same candidate (the 20-panel distance integral is shared out among the lanes)
__device__ __forceinline__ CandConsts make_consts(const mppb_holo_cand& cd, const HoloPtgDev& P, int lane)
{
    CandConsts q;
    q.T               = cd.T_ramp;
    q.vxi             = cd.vxi;
    q.vyi             = cd.vyi;
    q.vxf             = cd.vxf;
    q.vyf             = cd.vyf;
    q.vf              = cd.vf;
    q.R               = P.robotRadius;
    q.Vmax            = P.vMax;
    const double TR2_ = 1.0 / (2 * q.T);
    q.TR_2            = q.T * 0.5;
    q.T099            = q.T * 0.99;
    q.T101            = q.T * 1.01;
    q.k2              = (q.vxf - q.vxi) * TR2_;
    q.k4              = (q.vyf - q.vyi) * TR2_;
    q.qa              = (q.k2 * q.k2 + q.k4 * q.k4);
    q.qb              = (q.k2 * q.vxi * 2.0 + q.k4 * q.vyi * 2.0);
    q.distAtT         = ramp_distance_warp(q.k2, q.k4, q.vxi, q.vyi, q.T, lane);
    return q;
}

// HolonomicBlend.cpp:719-840 for one local obstacle point; returns the candidate's updated minimum
__device__ __forceinline__ double update_tp_obstacle(const CandConsts& q, double ox, double oy, double cur)
{
    double       sol_t = -1.0;
    const double a = q.qa, b = q.qb;
    const double c = -(q.k2 * ox * 2.0 + q.k4 * oy * 2.0 - q.vxi * q.vxi - q.vyi * q.vyi);
    const double d = -(ox * q.vxi * 2.0 + oy * q.vyi * 2.0);
    const double e = -q.R * q.R + ox * ox + oy * oy;
    double       roots[4];
    int          nreal = 0;
    if (fabs(a) > kEps)
        nreal = p34::quartic(roots, b / a, c / a, d / a, e / a);
    else if (fabs(b) > kEps)
        nreal = p34::cubic(roots, c / b, d / b, e / b);
    else
    {
        const double discr = d * d - 4 * c * e;
        if (discr >= 0)
        {
            nreal    = 2;
            roots[0] = (-d + sqrt(discr)) / (2 * c);
            roots[1] = (-d - sqrt(discr)) / (2 * c);
        }
    }
    for (int i = 0; i < nreal; i++)
    {
        const double r = roots[i];
        if (r == r && isfinite(r) && r >= .0 && r <= q.T * 1.01)
        {
            if (sol_t < 0)
                sol_t = r;
            else if (r < sol_t)
                sol_t = r;
        }
    }
    if (sol_t < 0 || sol_t > q.T101)
    {
        sol_t              = -1.0;
        const double c1    = q.TR_2 * (q.vxi - q.vxf) - ox;
        const double c2    = q.TR_2 * (q.vyi - q.vyf) - oy;
        const double xa    = q.vf * q.vf;
        const double xb    = 2 * (c1 * q.vxf + c2 * q.vyf);
        const double xc    = c1 * c1 + c2 * c2 - q.R * q.R;
        const double discr = xb * xb - 4 * xa * xc;
        if (discr >= 0)
        {
            const double t0 = (-xb + sqrt(discr)) / (2 * xa);
            const double t1 = (-xb - sqrt(discr)) / (2 * xa);
            if (t0 < q.T && t1 < q.T)
                sol_t = -1.0;
            else if (t0 < q.T && t1 >= q.T099)
                sol_t = t1;
            else if (t1 < q.T && t0 >= q.T099)
                sol_t = t0;
            else if (t1 >= q.T099 && t0 >= q.T099)
                sol_t = fmin(t0, t1);
        }
    }
    if (sol_t < 0) return cur;
    double dist;
    if (sol_t < q.T)
        dist = ramp_distance(q.k2, q.k4, q.vxi, q.vyi, sol_t);
    else
        dist = (sol_t - q.T) * q.Vmax + q.distAtT;
    // internal_TPObsDistancePostprocess, circular robot, COLL_BEH_BACK_AWAY
    if (!(hypot(ox, oy) < q.R)) return dist < cur ? dist : cur;
    if (dist < q.R) return cur;
    return 0.0;
}

// conservative FP32 description of the region a colliding obstacle point must lie in (robot frame)
struct PathHull
{
    float ax[3], ay[3], ex[3], ey[3], inv[3];  // triangle edges: start A, vector E, 1/|E|^2 (0 if degenerate)
    float qx, qy, ux, uy;                      // cruise ray: origin and unit direction
    float thr2;                                // (R + margin)^2
    bool  enabled;
};
__device__ __forceinline__ PathHull make_hull(const CandConsts& q)
{
    PathHull     h;
    const double tau = q.T101;
    const double px[3] = {0.0, q.vxi * tau * 0.5, q.vxi * tau + q.k2 * tau * tau};
    const double py[3] = {0.0, q.vyi * tau * 0.5, q.vyi * tau + q.k4 * tau * tau};
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
        const int j = (i + 1) % 3;
        h.ax[i]     = static_cast<float>(px[i]);
        h.ay[i]     = static_cast<float>(py[i]);
        h.ex[i]     = static_cast<float>(px[j] - px[i]);
        h.ey[i]     = static_cast<float>(py[j] - py[i]);
        const float l2 = h.ex[i] * h.ex[i] + h.ey[i] * h.ey[i];
        h.inv[i]       = l2 > 1e-12f ? 1.0f / l2 : 0.0f;
    }
    // cruise line p(t) = T/2 (vi + vf) + vf (t - T), from t = 0.99 T on
    h.qx = static_cast<float>(q.TR_2 * (q.vxi + q.vxf) - 0.01 * q.T * q.vxf);
    h.qy = static_cast<float>(q.TR_2 * (q.vyi + q.vyf) - 0.01 * q.T * q.vyf);
    const double vn = sqrt(q.vxf * q.vxf + q.vyf * q.vyf);
    h.ux            = static_cast<float>(vn > 0 ? q.vxf / vn : 0.0);
    h.uy            = static_cast<float>(vn > 0 ? q.vyf / vn : 0.0);
    const float thr = static_cast<float>(q.R) + 0.01f;
    h.thr2          = thr * thr;
    // degenerate or out-of-range parameters: no filtering (every in-window point takes the exact path)
    const double big = 1e3;
    h.enabled = vn > 1e-6 && q.vf > 1e-6 && isfinite(q.T) && q.T > 0 && q.T < big && fabs(q.vxi) < big && fabs(q.vyi) < big &&
                fabs(q.vxf) < big && fabs(q.vyf) < big && q.R > 0 && q.R < big &&
                fabs(vn - q.vf) <= 1e-6 * q.vf;  // (vxf, vyf) must be vf * (cos, sin), as the host computes it
    return h;
}
// true if the point (lx, ly) may be within thr of the path (squared distances, FP32)
__device__ __forceinline__ bool hull_may_hit(const PathHull& h, float lx, float ly)
{
    float d2   = FLT_MAX;
    int   npos = 0, nneg = 0;
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
        const float rx = lx - h.ax[i], ry = ly - h.ay[i];
        const float t  = fminf(fmaxf((rx * h.ex[i] + ry * h.ey[i]) * h.inv[i], 0.f), 1.f);
        const float sx = rx - t * h.ex[i], sy = ry - t * h.ey[i];
        d2             = fminf(d2, sx * sx + sy * sy);
        const float cr = h.ex[i] * ry - h.ey[i] * rx;
        npos += cr > 0.f;
        nneg += cr < 0.f;
    }
    if (npos == 3 || nneg == 3) return true;  // inside the control triangle
    {
        const float rx = lx - h.qx, ry = ly - h.qy;
        const float t  = fmaxf(rx * h.ux + ry * h.uy, 0.f);
        const float sx = rx - t * h.ux, sy = ry - t * h.uy;
        d2             = fminf(d2, sx * sx + sy * sy);
    }
    return d2 <= h.thr2;
}

// ---- bin culling ------------------------------------------------------------------------------------------------
// The hull above is thin (a triangle and a ray, inflated by the robot radius), the node's clipping square is not:
// of the ~440 cloud bins of a 10 m window a candidate's hull touches a few dozen. Per bin row the warp therefore
// narrows the row's run [bx0, bx1] to the bins the inflated hull can reach inside that row's strip of y, in FP64 and
// with a margin (thr = R + 1 cm, against FP32 rounding of ~1e-6 m in the exact chain's inputs), so only a tenth of
// the window's points are loaded and filtered at all. The hull is taken in the node's "d frame" (global axes, origin
// at the node), where a bin row is a horizontal strip. Conservative by construction: a point that can collide lies
// within R of the path, the path lies in (triangle U ray), and bin_coord() is monotone.
// widens [xl, xh] by the x extent of segment A-B inside the strip ya <= y <= yb
__device__ __forceinline__ void strip_extent(double ax, double ay, double bx, double by, double ya, double yb, double& xl,
                                             double& xh)
{
    if (fmax(ay, by) < ya || fmin(ay, by) > yb) return;
    const double dy = by - ay;
    double       t0 = 0.0, t1 = 1.0;
    if (fabs(dy) > 1e-9)
    {
        const double ta = (ya - ay) / dy, tb = (yb - ay) / dy;
        t0              = fmax(0.0, fmin(ta, tb));
        t1              = fmin(1.0, fmax(ta, tb));
    }
    const double xa = ax + t0 * (bx - ax), xb = ax + t1 * (bx - ax);
    xl = fmin(xl, fmin(xa, xb));
    xh = fmax(xh, fmax(xa, xb));
}
// x extent, in the d frame, of the candidate's hull (control triangle of the ramp + cruise ray cut where it leaves
// `reach`) inside the strip yLo <= y <= yHi; empty: .x > .y. Out of line and fed by value: it runs once per bin row and
// must not cost the point loop registers. (c, s): the node's heading.
__device__ __noinline__ double2 hull_strip_extent(double T, double vxi, double vyi, double vxf, double vyf, double c, double s,
                                                  double reach, double yLo, double yHi)
{
    const double tau = T * 1.01, TR2_ = 1.0 / (2 * T);
    const double k2 = (vxf - vxi) * TR2_, k4 = (vyf - vyi) * TR2_;
    double       lx[5] = {0.0, vxi * tau * 0.5, vxi * tau + k2 * tau * tau, T * 0.5 * (vxi + vxf) - 0.01 * T * vxf, 0.0};
    double       ly[5] = {0.0, vyi * tau * 0.5, vyi * tau + k4 * tau * tau, T * 0.5 * (vyi + vyf) - 0.01 * T * vyf, 0.0};
    const double vn  = sqrt(vxf * vxf + vyf * vyf);
    // a ray point at parameter t is at least t - |Q| from the node; window points are within `reach` of it
    const double len = hypot(lx[3], ly[3]) + reach;
    lx[4]            = lx[3] + (vxf / vn) * len;
    ly[4]            = ly[3] + (vyf / vn) * len;
    double vx[5], vy[5];
#pragma unroll
    for (int i = 0; i < 5; i++)
    {
        vx[i] = lx[i] * c - ly[i] * s;  // robot frame -> d frame
        vy[i] = lx[i] * s + ly[i] * c;
    }
    double xl = CUDART_INF, xh = -CUDART_INF;
    strip_extent(vx[0], vy[0], vx[1], vy[1], yLo, yHi, xl, xh);
    strip_extent(vx[1], vy[1], vx[2], vy[2], yLo, yHi, xl, xh);
    strip_extent(vx[2], vy[2], vx[0], vy[0], yLo, yHi, xl, xh);
    strip_extent(vx[3], vy[3], vx[4], vy[4], yLo, yHi, xl, xh);
    return make_double2(xl, xh);
}

#ifndef MPPB_HOLO_MIN_BLOCKS
#define MPPB_HOLO_MIN_BLOCKS 6  // 80 registers: the FP32 filter loop dominates, the rare exact path may spill
#endif
__global__ void __launch_bounds__(kThreads, MPPB_HOLO_MIN_BLOCKS)
    tpobs_holo_kernel(CloudBins bins, const HoloPtgDev* __restrict__ ptgs, const mppb_holo_cand* __restrict__ cands,
                      int nCands, int nNodes, const double* __restrict__ nodes /* [n][4] x,y,cos,sin */,
                      const float4* __restrict__ boxes /* NULL or [n] minx,miny,maxx,maxy */, double clip,
                      int useFilter, double* __restrict__ out)
{
    __shared__ float2 queue[kWarps][64];  // per warp: points that passed the filter and wait for the exact path
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ci   = blockIdx.x * kWarps + warp;
    if (ci >= nCands) return;
    // one node (a single query's expansion): its row does not have to wait for the candidate record, which, like the
    // row, may sit in pinned host memory (two PCIe round trips in a row otherwise)
    double row0[4] = {0, 0, 0, 0};
    if (nNodes == 1)
        for (int j = 0; j < 4; j++) row0[j] = nodes[j];
    const mppb_holo_cand cd = cands[ci];
    const HoloPtgDev     P  = ptgs[cd.ptg_index];

    const CandConsts q = make_consts(cd, P, lane);

    const double x0 = nNodes == 1 ? row0[0] : nodes[4 * cd.node + 0], y0 = nNodes == 1 ? row0[1] : nodes[4 * cd.node + 1];
    const double c = nNodes == 1 ? row0[2] : nodes[4 * cd.node + 2], s = nNodes == 1 ? row0[3] : nodes[4 * cd.node + 3];
    const double ix = -x0 * c - y0 * s;  // CPose2D unary minus
    const double iy = x0 * s - y0 * c;

    // optional per-node world box (ObstacleSource.cpp:29-40: float compare, inclusive bounds)
    const bool   hasBox = boxes != nullptr;
    const float4 box    = hasBox ? boxes[cd.node] : make_float4(0.f, 0.f, 0.f, 0.f);

    // the exact per-point chain of the reference (clip, transform, float rounding, time to collision)
    auto exact = [&](float2 g, double cur) -> double {
        const double gx = g.x, gy = g.y;
        // transform_pc_square_clipping.cpp:30-31
        if (fabs(gx - x0) > clip || fabs(gy - y0) > clip) return cur;
        const double ox = (ix + gx * c) + gy * s;
        const double oy = (iy - gx * s) + gy * c;
        const float  fx = static_cast<float>(ox), fy = static_cast<float>(oy);  // CPointsMap stores float
        return update_tp_obstacle(q, static_cast<double>(fx), static_cast<double>(fy), cur);
    };
    PathHull    hull   = make_hull(q);
    const float clipF  = static_cast<float>(clip);
    const float x0h = static_cast<float>(x0), x0l = static_cast<float>(x0 - static_cast<double>(x0h));
    const float y0h = static_cast<float>(y0), y0l = static_cast<float>(y0 - static_cast<double>(y0h));
    const float cf = static_cast<float>(c), sf = static_cast<float>(s);
    const float clipHi = clipF + 1e-4f * fmaxf(clipF, 1.f);
    const bool  filter = useFilter && hull.enabled && isfinite(clipF) && clipF < 1e3f;
    int         qn     = 0;  // queued points of this warp (same value in all lanes)

    double best = CUDART_INF;
    if (bins.n > 0 && isfinite(x0) && isfinite(y0) && clip >= 0)
    {
        // conservative float window -> bin range (the exact test is per point, in double)
        const float lox = float_below(x0 - clip);
        const float hix = float_above(x0 + clip);
        const float loy = float_below(y0 - clip);
        const float hiy = float_above(y0 + clip);
        const int   bx0 = bin_coord(lox, bins.minx, bins.invBin, bins.nbx);
        const int   bx1 = bin_coord(hix, bins.minx, bins.invBin, bins.nbx);
        const int   by0 = bin_coord(loy, bins.miny, bins.invBin, bins.nby);
        const int   by1 = bin_coord(hiy, bins.miny, bins.invBin, bins.nby);
        // bin culling (see HullD): margin = robot radius + 1 cm + what the float row boundaries can be off by
        const bool   cull       = filter && useFilter == 1;
        const double binSize    = 1.0 / static_cast<double>(bins.invBin);
        const double cullMargin = q.R + 0.01 + 1e-3 +
                                  1e-6 * (fabs(static_cast<double>(bins.miny)) + fabs(y0) + fabs(x0) +
                                          fabs(static_cast<double>(bins.minx)) + clip + binSize * (bins.nby + bins.nbx));
        const double cullReach = 1.5 * clip + cullMargin + 1.0;  // > sqrt(2) clip: no window point is farther from the node
        // The window's points, 32 at a time across bin rows: the runs of up to 32 rows are fetched together (one
        // row per lane, so their binStart reads are ONE round trip instead of one per row — on a sparse map the
        // rows hold a dozen points each and a row-by-row loop is a chain of dependent loads), a warp scan of the
        // run lengths gives every flat index its (row, offset), and lanes load consecutive flat indices. Points
        // that pass the filter are queued, and whenever 32 are waiting (or the window is exhausted) the warp runs
        // them through the exact chain — a single call site, so the FP64 code exists once.
        int      rowBase = by0;
        uint32_t f0 = 0, total = 0, runBeg = 0, excl = 0;
        bool     groupOpen = false;
        for (;;)
        {
            bool haveChunk = false;
            while (rowBase <= by1)
            {
                if (!groupOpen)
                {
                    const int nRows  = min(32, by1 - rowBase + 1);
                    uint32_t  runLen = 0;
                    runBeg           = 0;
                    if (lane < nRows)
                    {
                        int cx0 = bx0, cx1 = bx1;
                        if (cull)
                        {
                            // this row's strip of y in the d frame (first/last rows also hold what lies beyond the grid)
                            const int    row = rowBase + lane;
                            const double yLo = row == 0 ? -CUDART_INF : (bins.miny + row * binSize - y0) - cullMargin;
                            const double yHi =
                                row == bins.nby - 1 ? CUDART_INF : (bins.miny + (row + 1) * binSize - y0) + cullMargin;
                            const double2 ext = hull_strip_extent(q.T, q.vxi, q.vyi, q.vxf, q.vyf, c, s, cullReach, yLo, yHi);
                            const double  xl = ext.x, xh = ext.y;
                            if (xl <= xh)
                            {
                                cx0 = max(bx0, bin_coord(__double2float_rd(x0 + xl - cullMargin), bins.minx, bins.invBin, bins.nbx));
                                cx1 = min(bx1, bin_coord(__double2float_ru(x0 + xh + cullMargin), bins.minx, bins.invBin, bins.nbx));
                            }
                            else
                                cx1 = cx0 - 1;
                        }
                        if (cx0 <= cx1)
                        {
                            const size_t r = static_cast<size_t>(rowBase + lane) * bins.nbx;
                            runBeg         = bins.binStart[r + cx0];
                            runLen         = bins.binStart[r + cx1 + 1] - runBeg;
                        }
                    }
                    uint32_t incl = runLen;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1)
                    {
                        const uint32_t t = __shfl_up_sync(kFull, incl, o);
                        if (lane >= o) incl += t;
                    }
                    total     = __shfl_sync(kFull, incl, 31);
                    excl      = incl - runLen;
                    f0        = 0;
                    groupOpen = true;
                }
                if (f0 < total)
                {
                    haveChunk = true;
                    break;
                }
                groupOpen = false;
                rowBase += 32;
            }
            if (haveChunk)
            {
                const uint32_t f = f0 + lane;
                // the last row whose first flat index is <= f (empty rows share their successor's index and lose)
                int row = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1)
                {
                    const uint32_t v = __shfl_sync(kFull, excl, row + step);
                    if (v <= f) row += step;
                }
                const uint32_t p = __shfl_sync(kFull, runBeg, row) + (f - __shfl_sync(kFull, excl, row));
                bool           pass = false;
                float2         g    = make_float2(0.f, 0.f);
                if (f < total)
                {
                    g    = bins.pts[p];
                    pass = !(hasBox && (g.x < box.x || g.x > box.z || g.y < box.y || g.y > box.w));
                    if (filter)
                    {
                        // FP32 transform (|error| < 3e-6 * max(clip,1) m) and the conservative tests
                        const float dx = (g.x - x0h) - x0l, dy = (g.y - y0h) - y0l;
                        pass           = pass && fmaxf(fabsf(dx), fabsf(dy)) <= clipHi;
                        pass           = pass && hull_may_hit(hull, dx * cf + dy * sf, dy * cf - dx * sf);
                    }
                }
                f0 += 32;
                const unsigned m = __ballot_sync(kFull, pass);
                if (pass) queue[warp][qn + __popc(m & ((1u << lane) - 1u))] = g;
                qn += __popc(m);
            }
            if (qn >= 32 || (!haveChunk && qn > 0))
            {
                __syncwarp();
                const int take = min(qn, 32);
                if (lane < take) best = exact(queue[warp][qn - take + lane], best);
                qn -= take;
                __syncwarp();
            }
            if (!haveChunk && qn == 0) break;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fmin(best, __shfl_xor_sync(kFull, best, o));
    if (lane == 0) out[ci] = best;
}
// ---- one CTA per node (diagnostic alternative, mppb_set_holo_by_node) ------------------------------------------------
// The warp-per-candidate kernel above makes every candidate of a node walk the node's window on its own: the bin-row
// runs are located, loaded, box- and clip-tested once per candidate — 129 times per node for the reference's demo
// parameters. Here a CTA owns a node: (1) all threads walk the window ONCE and keep the points that may lie inside the
// clipping square (conservative FP32 test; the exact FP64 test still decides) in shared memory, (2) each warp then
// takes the node's candidates in turn over that dense array: FP32 transform, hull filter, queue, exact chain — the
// same per-point arithmetic as above on the same set of points, so the results are the same minima (tested bit for
// bit). Measured on B200 (21 504 candidates of 512 nodes, 2 621 in-window points each): 0.90 ms against 0.85 ms of the
// kernel above — the time goes into the per-(candidate, point) filter and the exact chain, not into the window walk,
// which the candidates of one node, sitting in consecutive warps, share through L1 anyway. Kept as an opt-in.
// candBegin[b] .. candBegin[b+1]: the candidates of node b (the caller's array is grouped by node).
constexpr int kNodeThreads = 256;
constexpr int kNodeWarps   = kNodeThreads / 32;
constexpr int kNodeMaxPts  = 6144;  // window points held per pass (48 KB)
constexpr int kNodeRows    = 64;    // bin rows located together
__global__ void __launch_bounds__(kNodeThreads, 3)
    tpobs_holo_node_kernel(CloudBins bins, const HoloPtgDev* __restrict__ ptgs, const mppb_holo_cand* __restrict__ cands,
                           const uint32_t* __restrict__ candBegin, const double* __restrict__ nodes /* [n][4] x,y,cos,sin */,
                           const float4* __restrict__ boxes, double clip, int useFilter, double* __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char dynSmem[];
    float2* const sPts  = reinterpret_cast<float2*>(dynSmem);                                   // [kNodeMaxPts]
    double* const sBest = reinterpret_cast<double*>(dynSmem + sizeof(float2) * kNodeMaxPts);  // [candidates of the node]
    __shared__ float2   queue[kNodeWarps][64];
    __shared__ uint32_t rowBeg[kNodeRows], rowPrefix[kNodeRows + 1];
    __shared__ unsigned sCount;

    const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int      node = blockIdx.x;
    const uint32_t cBeg = candBegin[node], cEnd = candBegin[node + 1];
    if (cBeg == cEnd) return;
    const double x0 = nodes[4 * node + 0], y0 = nodes[4 * node + 1], c = nodes[4 * node + 2], s = nodes[4 * node + 3];
    const double ix = -x0 * c - y0 * s;  // CPose2D unary minus
    const double iy = x0 * s - y0 * c;
    const bool   hasBox = boxes != nullptr;
    const float4 box    = hasBox ? boxes[node] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float  clipF  = static_cast<float>(clip);
    const float  x0h = static_cast<float>(x0), x0l = static_cast<float>(x0 - static_cast<double>(x0h));
    const float  y0h = static_cast<float>(y0), y0l = static_cast<float>(y0 - static_cast<double>(y0h));
    const float  cf = static_cast<float>(c), sf = static_cast<float>(s);
    const float  clipHi     = clipF + 1e-4f * fmaxf(clipF, 1.f);
    const bool   clipFinite = isfinite(clipF) && clipF < 1e3f;

    for (uint32_t i = threadIdx.x; i < cEnd - cBeg; i += kNodeThreads) sBest[i] = CUDART_INF;

    if (bins.n > 0 && isfinite(x0) && isfinite(y0) && clip >= 0)
    {
        const float lox = float_below(x0 - clip);
        const float hix = float_above(x0 + clip);
        const float loy = float_below(y0 - clip);
        const float hiy = float_above(y0 + clip);
        const int   bx0 = bin_coord(lox, bins.minx, bins.invBin, bins.nbx);
        const int   bx1 = bin_coord(hix, bins.minx, bins.invBin, bins.nbx);
        const int   by0 = bin_coord(loy, bins.miny, bins.invBin, bins.nby);
        const int   by1 = bin_coord(hiy, bins.miny, bins.invBin, bins.nby);
        for (int rowBase = by0; rowBase <= by1; rowBase += kNodeRows)
        {
            const int nRows = min(kNodeRows, by1 - rowBase + 1);
            __syncthreads();  // previous pass done with the row tables
            if (threadIdx.x < nRows)
            {
                const size_t   r = static_cast<size_t>(rowBase + threadIdx.x) * bins.nbx;
                const uint32_t b = bins.binStart[r + bx0], e = bins.binStart[r + bx1 + 1];
                rowBeg[threadIdx.x]        = b;
                rowPrefix[threadIdx.x + 1] = e - b;
            }
            __syncthreads();
            if (warp == 0)
            {
                // inclusive scan of up to 64 run lengths, two per lane
                uint32_t a0 = (2 * lane < nRows) ? rowPrefix[2 * lane + 1] : 0u;
                uint32_t a1 = (2 * lane + 1 < nRows) ? rowPrefix[2 * lane + 2] : 0u;
                uint32_t v  = a0 + a1;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1)
                {
                    const uint32_t t = __shfl_up_sync(kFull, v, o);
                    if (lane >= o) v += t;
                }
                if (lane == 0) rowPrefix[0] = 0;
                if (2 * lane < nRows) rowPrefix[2 * lane + 1] = v - a1;
                if (2 * lane + 1 < nRows) rowPrefix[2 * lane + 2] = v;
            }
            __syncthreads();
            const uint32_t total = rowPrefix[nRows];
            for (uint32_t f0 = 0; f0 < total; f0 += kNodeMaxPts)
            {
                // ---- (1) the window's points [f0, f1) of this row group -> shared memory, once for all candidates ----
                const uint32_t f1 = min(total, f0 + static_cast<uint32_t>(kNodeMaxPts));
                __syncthreads();  // the candidates of the previous pass are done with sPts
                if (threadIdx.x == 0) sCount = 0;
                __syncthreads();
                for (uint32_t fb = f0; fb < f1; fb += kNodeThreads)
                {
                    const uint32_t f    = fb + threadIdx.x;
                    bool           keep = false;
                    float2         g    = make_float2(0.f, 0.f);
                    if (f < f1)
                    {
                        int lo = 0, hi = nRows;  // the row whose flat range holds f
                        while (hi - lo > 1)
                        {
                            const int mid = (lo + hi) >> 1;
                            if (rowPrefix[mid] <= f)
                                lo = mid;
                            else
                                hi = mid;
                        }
                        g    = bins.pts[rowBeg[lo] + (f - rowPrefix[lo])];
                        keep = !(hasBox && (g.x < box.x || g.x > box.z || g.y < box.y || g.y > box.w));
                        if (clipFinite)
                        {
                            const float dx = (g.x - x0h) - x0l, dy = (g.y - y0h) - y0l;
                            keep           = keep && fmaxf(fabsf(dx), fabsf(dy)) <= clipHi;  // conservative; exact() decides
                        }
                    }
                    const unsigned m    = __ballot_sync(kFull, keep);
                    unsigned       base = 0;
                    if (lane == 0 && m) base = atomicAdd(&sCount, __popc(m));
                    base = __shfl_sync(kFull, base, 0);
                    if (keep) sPts[base + __popc(m & ((1u << lane) - 1u))] = g;
                }
                __syncthreads();
                const uint32_t nPts = sCount;
                // ---- (2) every candidate of the node over the dense array, one warp per candidate at a time ----
                for (uint32_t ci = cBeg + warp; ci < cEnd; ci += kNodeWarps)
                {
                    const mppb_holo_cand cd = cands[ci];
                    const HoloPtgDev     P  = ptgs[cd.ptg_index];
                    const CandConsts     q  = make_consts(cd, P, lane);
                    const PathHull       hull   = make_hull(q);
                    const bool           filter = useFilter && hull.enabled && clipFinite;
                    auto exact = [&](float2 g, double cur) -> double {
                        const double gx = g.x, gy = g.y;
                        // transform_pc_square_clipping.cpp:30-31
                        if (fabs(gx - x0) > clip || fabs(gy - y0) > clip) return cur;
                        const double ox = (ix + gx * c) + gy * s;
                        const double oy = (iy - gx * s) + gy * c;
                        const float  fx = static_cast<float>(ox), fy = static_cast<float>(oy);  // CPointsMap stores float
                        return update_tp_obstacle(q, static_cast<double>(fx), static_cast<double>(fy), cur);
                    };
                    double best = CUDART_INF;
                    int    qn   = 0;
                    for (uint32_t pb = 0; pb < nPts || qn > 0; pb += 32)
                    {
                        const bool haveChunk = pb < nPts;
                        if (haveChunk)
                        {
                            const uint32_t i    = pb + lane;
                            bool           pass = i < nPts;
                            float2         g    = make_float2(0.f, 0.f);
                            if (pass)
                            {
                                g = sPts[i];
                                if (filter)
                                {
                                    const float dx = (g.x - x0h) - x0l, dy = (g.y - y0h) - y0l;
                                    pass           = hull_may_hit(hull, dx * cf + dy * sf, dy * cf - dx * sf);
                                }
                            }
                            const unsigned m = __ballot_sync(kFull, pass);
                            if (pass) queue[warp][qn + __popc(m & ((1u << lane) - 1u))] = g;
                            qn += __popc(m);
                        }
                        if (qn >= 32 || (!haveChunk && qn > 0))
                        {
                            __syncwarp();
                            const int take = min(qn, 32);
                            if (lane < take) best = exact(queue[warp][qn - take + lane], best);
                            qn -= take;
                            __syncwarp();
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) best = fmin(best, __shfl_xor_sync(kFull, best, o));
                    if (lane == 0) sBest[ci - cBeg] = fmin(sBest[ci - cBeg], best);
                }
            }
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < cEnd - cBeg; i += kNodeThreads) out[cBeg + i] = sBest[i];
}
}  // namespace

// One CTA per node (candidates grouped by node, offsets in dCandBegin[nNodes+1])
int launch_tpobs_holo_by_node(mppb_ctx* ctx, size_t nNodes, const uint32_t* dCandBegin, size_t maxCandsPerNode,
                              const mppb_holo_cand* dCands, const double* dNodes, const float* dBoxes, double clip, double* dOut)
{
    if (nNodes == 0) return MPPB_OK;
    if (ctx->holos.empty()) return fail(ctx, MPPB_ERR_STATE, "no HolonomicBlend PTG set (mppb_set_ptg_holo)");
    const size_t smem = sizeof(float2) * kNodeMaxPts + sizeof(double) * std::max<size_t>(maxCandsPerNode, 1);
    if (smem > 200 * 1024) return fail(ctx, MPPB_ERR_UNSUPPORTED, "too many candidates per node");
    MPPB_CUDA(ctx, cudaFuncSetAttribute(tpobs_holo_node_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    tpobs_holo_node_kernel<<<static_cast<unsigned>(nNodes), kNodeThreads, smem, ctx->stream>>>(
        ctx->bins, ctx->holoDescs.as<HoloPtgDev>(), dCands, dCandBegin, dNodes, reinterpret_cast<const float4*>(dBoxes), clip,
        ctx->holoFilter, dOut);
    ctx->launches++;
    MPPB_CUDA(ctx, cudaGetLastError());
    return MPPB_OK;
}

int launch_tpobs_holo(mppb_ctx* ctx, size_t nCands, const mppb_holo_cand* dCands, const double* dNodes,
                      const float* dBoxes, double clip, double* dOut, size_t nNodes)
{
    const int useFilter = ctx->holoFilter;  // 0: the unfiltered kernel, 2: filter without bin culling (mppb_set_holo_filter)
    if (nCands == 0) return MPPB_OK;
    if (ctx->holos.empty()) return fail(ctx, MPPB_ERR_STATE, "no HolonomicBlend PTG set (mppb_set_ptg_holo)");
    if (nCands > 0x3fffffffull) return fail(ctx, MPPB_ERR_INVALID_ARG, "too many candidates in one batch");
    const unsigned blocks = static_cast<unsigned>((nCands + kWarps - 1) / kWarps);
    tpobs_holo_kernel<<<blocks, kThreads, 0, ctx->stream>>>(ctx->bins, ctx->holoDescs.as<HoloPtgDev>(), dCands,
                                                            static_cast<int>(nCands), nNodes == 1 ? 1 : 0, dNodes,
                                                            reinterpret_cast<const float4*>(dBoxes), clip, useFilter, dOut);
    ctx->launches++;
    MPPB_CUDA(ctx, cudaGetLastError());
    return MPPB_OK;
}

}  // namespace mppb
