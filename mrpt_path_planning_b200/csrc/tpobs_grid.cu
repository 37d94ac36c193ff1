// tpobs_grid_kernel — stage (a)+(b) of the hot path for collision-grid PTGs, fused.
//
// Replaces, for a whole batch of lattice nodes at once, the reference's per-node
//   cached_local_obstacles -> transform_pc_square_clipping        (TPS_Astar.cpp:561-562,
//                                                                  transform_pc_square_clipping.cpp:26-42)
//   tp_obstacles_single_path -> updateTPObstacleSingle per (k,pt)  (tp_obstacles_single_path.cpp:26-34,
//                                                                  DiffDriveCollisionGridBased.cpp:826-838)
// out[node][ptgOffset+k] = min over in-window points of the TP-obstacle distance of path k,
// +INF if no point constrains k (the caller applies initTPObstacleSingle's value).
//
// Exactness (bit parity of the free/blocked decision): the local point is formed with the
// reference's operation order in IEEE double without FMA contraction (__dmul_rn/__dadd_rn),
// rounded to float as CPointsMap stores it, promoted again, and mapped to a collision-grid
// cell with a true double division and truncation (CDynamicGrid::x2idx). cos/sin of the node
// heading come from the host C library. The per-path result is an order-independent min,
// so it is reduced with atomicMin on the (non-negative) float bit pattern in shared memory.
//
// Work decomposition: one CTA per node. The node's clipping square selects a range of bin
// rows of the cloud index; each row contributes one contiguous run of points. Threads walk
// the concatenated runs with coalesced 8-byte loads.
#include <cfloat>

#include "mppb_internal.cuh"

namespace mppb
{
namespace
{
constexpr int kThreads     = 128;
constexpr int kRowsPerPass = 64;
constexpr unsigned kInfBits = 0x7f800000u;

struct NodeWindow
{
    double x0, y0, c, s, ix, iy, D;
};

// reference predicate transform_pc_square_clipping.cpp:30-31 (true => point skipped)
__device__ __forceinline__ bool clipped_out(double g, double center, double D)
{
    return fabs(__dsub_rn(g, center)) > D;
}

// mrpt::math::TPolygon2D::contains (PNPOLY) in the reference's operation order.
__device__ bool polygon_contains(const PtgGridDev& g, double px, double py)
{
    bool      c = false;
    const int n = g.nverts;
    for (int i = 0, j = n - 1; i < n; j = i++)
    {
        const double yi = g.shapeY[i], yj = g.shapeY[j];
        if ((yi > py) != (yj > py))
        {
            const double xi = g.shapeX[i], xj = g.shapeX[j];
            const double t  = __dadd_rn(__ddiv_rn(__dmul_rn(__dsub_rn(xj, xi), __dsub_rn(py, yi)), __dsub_rn(yj, yi)), xi);
            if (px < t) c = !c;
        }
    }
    return c;
}

__global__ void __launch_bounds__(kThreads)
    tpobs_grid_kernel(CloudBins bins, const PtgGridDev* __restrict__ ptgs, int numPtgs, int totalPaths,
                      const double* __restrict__ nodes /* [n][4] x,y,cos,sin */, double clip,
                      float* __restrict__ out)
{
    extern __shared__ unsigned char smemRaw[];
    unsigned*                       smin = reinterpret_cast<unsigned*>(smemRaw);  // [totalPaths]
    __shared__ uint32_t             rowBeg[kRowsPerPass];
    __shared__ uint32_t             rowPrefix[kRowsPerPass + 1];

    const int    node = blockIdx.x;
    const double x0 = nodes[4 * node + 0], y0 = nodes[4 * node + 1];
    const double c = nodes[4 * node + 2], s = nodes[4 * node + 3];

    for (int i = threadIdx.x; i < totalPaths; i += kThreads) smin[i] = kInfBits;

    // inverse pose (CPose2D unary minus): ix = -x*c - y*s ; iy = x*s - y*c
    const double ix = __dsub_rn(__dmul_rn(-x0, c), __dmul_rn(y0, s));
    const double iy = __dsub_rn(__dmul_rn(x0, s), __dmul_rn(y0, c));

    // conservative float window for the bin range (the exact test is per point, in double)
    int  bx0 = 0, bx1 = -1, by0 = 0, by1 = -1;
    bool any = bins.n > 0 && isfinite(x0) && isfinite(y0) && clip >= 0;
    if (any)
    {
        const float lox = nextafterf(__double2float_rd(x0 - clip), -FLT_MAX);
        const float hix = nextafterf(__double2float_ru(x0 + clip), FLT_MAX);
        const float loy = nextafterf(__double2float_rd(y0 - clip), -FLT_MAX);
        const float hiy = nextafterf(__double2float_ru(y0 + clip), FLT_MAX);
        bx0             = bin_coord(lox, bins.minx, bins.invBin, bins.nbx);
        bx1             = bin_coord(hix, bins.minx, bins.invBin, bins.nbx);
        by0             = bin_coord(loy, bins.miny, bins.invBin, bins.nby);
        by1             = bin_coord(hiy, bins.miny, bins.invBin, bins.nby);
    }
    __syncthreads();

    for (int rowBase = by0; any && rowBase <= by1; rowBase += kRowsPerPass)
    {
        const int nRows = min(kRowsPerPass, by1 - rowBase + 1);
        if (threadIdx.x < nRows)
        {
            const size_t   r = static_cast<size_t>(rowBase + threadIdx.x) * bins.nbx;
            const uint32_t b = bins.binStart[r + bx0], e = bins.binStart[r + bx1 + 1];
            rowBeg[threadIdx.x]        = b;
            rowPrefix[threadIdx.x + 1] = e - b;
        }
        __syncthreads();
        if (threadIdx.x == 0)
        {
            uint32_t run = 0;
            rowPrefix[0] = 0;
            for (int r = 0; r < nRows; r++)
            {
                run += rowPrefix[r + 1];
                rowPrefix[r + 1] = run;
            }
        }
        __syncthreads();
        const uint32_t total = rowPrefix[nRows];

        for (uint32_t f = threadIdx.x; f < total; f += kThreads)
        {
            // locate the row of flat index f (binary search over <= 64 prefix sums)
            int lo = 0, hi = nRows;
            while (hi - lo > 1)
            {
                const int mid = (lo + hi) >> 1;
                if (rowPrefix[mid] <= f)
                    lo = mid;
                else
                    hi = mid;
            }
            const float2 g  = bins.pts[rowBeg[lo] + (f - rowPrefix[lo])];
            const double gx = g.x, gy = g.y;
            if (clipped_out(gx, x0, clip) || clipped_out(gy, y0, clip)) continue;

            // composePoint with the inverse pose, reference operation order, no FMA
            const double ox = __dadd_rn(__dadd_rn(ix, __dmul_rn(gx, c)), __dmul_rn(gy, s));
            const double oy = __dadd_rn(__dsub_rn(iy, __dmul_rn(gx, s)), __dmul_rn(gy, c));
            const float  fx = __double2float_rn(ox), fy = __double2float_rn(oy);
            const double lx = fx, ly = fy;

            for (int p = 0; p < numPtgs; p++)
            {
                const PtgGridDev& G  = ptgs[p];
                const int         cx = __double2int_rz(__ddiv_rn(__dsub_rn(lx, G.xmin), G.res));
                const int         cy = __double2int_rz(__ddiv_rn(__dsub_rn(ly, G.ymin), G.res));
                if (cx < 0 || cx >= G.nx || cy < 0 || cy >= G.ny) continue;
                const uint32_t cell = cx + cy * G.nx;
                const uint32_t eb = G.offsets[cell], ee = G.offsets[cell + 1];
                if (eb == ee) continue;
                // isPointInsideRobotShape: only points within the shape's bounding radius can be inside
                bool inside = false;
                if (fx * fx + fy * fy <= G.insideGuardR2) inside = polygon_contains(G, lx, ly);
                unsigned* sm = smin + G.outOffset;
                for (uint32_t e = eb; e < ee; e++)
                {
                    const uint2 kd = G.entries[e];
                    unsigned    v  = kd.y;
                    if (inside)
                    {
                        // COLL_BEH_BACK_AWAY: d < maxRobotRadius -> ignore, else block the path
                        if (static_cast<double>(__uint_as_float(kd.y)) < G.maxRadius) continue;
                        v = 0u;
                    }
                    if (v < sm[kd.x]) atomicMin(&sm[kd.x], v);
                }
            }
        }
        __syncthreads();
    }
    __syncthreads();
    float* o = out + static_cast<size_t>(node) * totalPaths;
    for (int i = threadIdx.x; i < totalPaths; i += kThreads) o[i] = __uint_as_float(smin[i]);
}
}  // namespace

int launch_tpobs_grid(mppb_ctx* ctx, size_t nNodes, const double* dNodes, double clip, float* dOut)
{
    if (nNodes == 0) return MPPB_OK;
    if (ctx->grids.empty()) return fail(ctx, MPPB_ERR_STATE, "no PTG tables set (mppb_set_ptg_grid)");
    if (nNodes > 0x7fffffffull) return fail(ctx, MPPB_ERR_INVALID_ARG, "too many nodes in one batch");
    const size_t smem = static_cast<size_t>(ctx->totalPaths) * sizeof(unsigned);
    if (smem > 200 * 1024) return fail(ctx, MPPB_ERR_UNSUPPORTED, "sum of PTG path counts too large");
    if (smem > 48 * 1024)
        MPPB_CUDA(ctx, cudaFuncSetAttribute(tpobs_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            static_cast<int>(smem)));
    tpobs_grid_kernel<<<static_cast<unsigned>(nNodes), kThreads, smem, ctx->stream>>>(
        ctx->bins, ctx->gridDescs.as<PtgGridDev>(), static_cast<int>(ctx->grids.size()), ctx->totalPaths, dNodes,
        clip, dOut);
    ctx->launches++;
    MPPB_CUDA(ctx, cudaGetLastError());
    return MPPB_OK;
}

}  // namespace mppb
