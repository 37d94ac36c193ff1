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
// Work decomposition: one warp per candidate. The warp walks the bin rows of the cloud index that
// overlap the node's clipping square; lanes take consecutive points (coalesced 8-byte loads),
// clip and transform them with the reference's FP64 operation order, round to float as
// CPointsMap stores them, and solve the time-to-collision quartic in FP64 (poly34.cuh). The
// per-candidate minimum is an order-independent min, reduced with warp shuffles.
// This file is compiled with -fmad=false (see build.py): all arithmetic rounds as on the
// reference's x86-64 build; acos/cos in the cubic's trigonometric branch are CUDA libm.
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

struct CandConsts
{
    double T, vxi, vyi, vxf, vyf, vf;
    double k2, k4, qa, qb, TR_2, T099, T101, R, Vmax, distAtT;
};

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

__global__ void __launch_bounds__(kThreads)
    tpobs_holo_kernel(CloudBins bins, const HoloPtgDev* __restrict__ ptgs, const mppb_holo_cand* __restrict__ cands,
                      int nCands, const double* __restrict__ nodes /* [n][4] x,y,cos,sin */,
                      const float4* __restrict__ boxes /* NULL or [n] minx,miny,maxx,maxy */, double clip,
                      double* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int ci   = blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (ci >= nCands) return;
    const mppb_holo_cand cd = cands[ci];
    const HoloPtgDev     P  = ptgs[cd.ptg_index];

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
    q.distAtT         = ramp_distance(q.k2, q.k4, q.vxi, q.vyi, q.T);

    const double x0 = nodes[4 * cd.node + 0], y0 = nodes[4 * cd.node + 1];
    const double c = nodes[4 * cd.node + 2], s = nodes[4 * cd.node + 3];
    const double ix = -x0 * c - y0 * s;  // CPose2D unary minus
    const double iy = x0 * s - y0 * c;

    // optional per-node world box (ObstacleSource.cpp:29-40: float compare, inclusive bounds)
    const bool   hasBox = boxes != nullptr;
    const float4 box    = hasBox ? boxes[cd.node] : make_float4(0.f, 0.f, 0.f, 0.f);

    double best = CUDART_INF;
    if (bins.n > 0 && isfinite(x0) && isfinite(y0) && clip >= 0)
    {
        // conservative float window -> bin range (the exact test is per point, in double)
        const float lox = nextafterf(__double2float_rd(x0 - clip), -FLT_MAX);
        const float hix = nextafterf(__double2float_ru(x0 + clip), FLT_MAX);
        const float loy = nextafterf(__double2float_rd(y0 - clip), -FLT_MAX);
        const float hiy = nextafterf(__double2float_ru(y0 + clip), FLT_MAX);
        const int   bx0 = bin_coord(lox, bins.minx, bins.invBin, bins.nbx);
        const int   bx1 = bin_coord(hix, bins.minx, bins.invBin, bins.nbx);
        const int   by0 = bin_coord(loy, bins.miny, bins.invBin, bins.nby);
        const int   by1 = bin_coord(hiy, bins.miny, bins.invBin, bins.nby);
        for (int row = by0; row <= by1; row++)
        {
            const size_t   r   = static_cast<size_t>(row) * bins.nbx;
            const uint32_t beg = bins.binStart[r + bx0], end = bins.binStart[r + bx1 + 1];
            for (uint32_t p = beg + lane; p < end; p += 32)
            {
                const float2 g  = bins.pts[p];
                if (hasBox && (g.x < box.x || g.x > box.z || g.y < box.y || g.y > box.w)) continue;
                const double gx = g.x, gy = g.y;
                // transform_pc_square_clipping.cpp:30-31
                if (fabs(gx - x0) > clip || fabs(gy - y0) > clip) continue;
                const double ox = (ix + gx * c) + gy * s;
                const double oy = (iy - gx * s) + gy * c;
                const float  fx = static_cast<float>(ox), fy = static_cast<float>(oy);  // CPointsMap stores float
                best            = update_tp_obstacle(q, static_cast<double>(fx), static_cast<double>(fy), best);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fmin(best, __shfl_xor_sync(kFull, best, o));
    if (lane == 0) out[ci] = best;
}
}  // namespace

int launch_tpobs_holo(mppb_ctx* ctx, size_t nCands, const mppb_holo_cand* dCands, const double* dNodes,
                      const float* dBoxes, double clip, double* dOut)
{
    if (nCands == 0) return MPPB_OK;
    if (ctx->holos.empty()) return fail(ctx, MPPB_ERR_STATE, "no HolonomicBlend PTG set (mppb_set_ptg_holo)");
    if (nCands > 0x3fffffffull) return fail(ctx, MPPB_ERR_INVALID_ARG, "too many candidates in one batch");
    const unsigned blocks = static_cast<unsigned>((nCands + kWarps - 1) / kWarps);
    tpobs_holo_kernel<<<blocks, kThreads, 0, ctx->stream>>>(ctx->bins, ctx->holoDescs.as<HoloPtgDev>(), dCands,
                                                            static_cast<int>(nCands), dNodes,
                                                            reinterpret_cast<const float4*>(dBoxes), clip, dOut);
    ctx->launches++;
    MPPB_CUDA(ctx, cudaGetLastError());
    return MPPB_OK;
}

}  // namespace mppb
