// expand_kernel — one A* expansion per node on the device, for PTG sets of the collision-grid family.
//
// Replaces, for a whole batch of nodes at once, what TPS_Astar::find_feasible_paths_to_neighbors
// (src/algos/TPS_Astar.cpp:538-826) and the per-neighbour block of TPS_Astar::plan (:269-345) compute from the node's
// frozen state alone (see include/mppb.h, "one A* expansion per node"):
//   candidate TPS points in the reference's order (:614-697)  ->  end pose, world-box test, lattice cell (:725-736,
//   TPS_Astar.h:224-230)  ->  blocked iff relTrgDist >= freeDistance (:749) against the collision row that
//   tpobs_grid_kernel wrote just before on the same stream  ->  goal-cell rule (:765-778) and best candidate per reached
//   cell, strict < (:783-796)  ->  neighbour pose and twist (:290-297), interpolated samples
//   (edge_interpolated_path.cpp:51-72) and cost_path_segment (Planner.cpp:15-29; evaluators: evaluators.cuh).
// The sampled paths (x, y, phi, dist, v, w per step: DiffDriveCollisionGridBased.cpp:766-811) are resident in HBM
// together with the host-libm cos/sin of every step's heading; the node's own cos/sin come from the host as well, so
// every value below is formed by IEEE +,-,*,/ and fmod in the reference's operation order (this file is compiled with
// -fmad=false): poses, cells, decisions and costs are bit-identical to the CPU evaluation.
//
// Work decomposition: one CTA per node, one thread per candidate slot (<= 512). The per-cell selection is an all-pairs
// scan over the node's <= 512 candidates in shared memory (a few thousand compares); the evaluator look-ups of the
// surviving edges' samples are issued eight at a time per (edge, evaluator) thread. Nothing here is bandwidth-bound:
// the kernel reads ~100 table entries and writes <= 80 B per reached cell; its job is to spare the host the ~45 us
// per expansion that phases A, C and D cost there.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "evaluators.cuh"
#include "mppb_internal.cuh"

namespace mppb
{
namespace
{
constexpr int kMaxCand = 512;  // candidate slots per node (threads of the CTA)

__device__ __forceinline__ double wrap_to_2pi_d(double a)
{
    const bool neg = a < 0;
    a              = fmod(a, 2.0 * M_PI);
    return neg ? a + 2.0 * M_PI : a;
}
// [MRPT] wrapToPi: result in [-pi, pi)
__device__ __forceinline__ double wrap_to_pi_d(double a) { return wrap_to_2pi_d(a + M_PI) - M_PI; }

// TPS_Astar.h:224-230: lattice indices take their argument as float
__device__ __forceinline__ int x2idx_f(float x, double res) { return static_cast<int>(static_cast<double>(x) / res); }
__device__ __forceinline__ int phi2idx_f(float yaw, double resYaw)
{
    const float pi = static_cast<float>(M_PI), twoPi = static_cast<float>(2.0 * M_PI);
    float       a   = yaw + pi;
    const bool  neg = a < 0;
    a               = fmodf(a, twoPi);
    if (neg) a += twoPi;
    return static_cast<int>(static_cast<double>(a - pi) / resYaw);
}

__global__ void __launch_bounds__(kMaxCand)
    expand_kernel(const __grid_constant__ ExpandParamsDev P, const EvaluatorDev* __restrict__ evalsGlobal, int nEvals,
                  const mppb_expand_node* __restrict__ nodes, const double* __restrict__ nodes4 /* [n][4] x,y,cos,sin */,
                  const float* __restrict__ rows /* [n][totalPaths] */, int totalPaths, mppb_neighbor* __restrict__ out,
                  uint32_t* __restrict__ counts)
{
    // per-candidate arrays in dynamic shared memory, sized by the candidate slots of THIS parameter set (86 for the
    // reference's demo parameters, not the 512 the kernel admits): ~140 bytes per slot, so that a dozen and more CTAs
    // share an SM — the kernel is a chain of dependent look-ups and barriers, and only resident CTAs hide that
    extern __shared__ uint4 sRecRaw[];  // [mc] neighbour records (80 bytes each) come first: see below
    const int        mc       = (P.maxCand + 31) & ~31;  // = the CTA's thread count: every thread owns a slot
    double* const    sVal     = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(sRecRaw) + static_cast<size_t>(mc) * sizeof(mppb_neighbor));
    int* const       sCellAll = reinterpret_cast<int*>(sVal + static_cast<size_t>(mc) * MPPB_MAX_EVALUATORS);
    int* const       sCell[3] = {sCellAll, sCellAll + mc, sCellAll + 2 * mc};
    unsigned* const  sDist    = reinterpret_cast<unsigned*>(sCellAll + 3 * mc);  // relTrgDist as float bits (non-negative table float: unsigned order)
    uint32_t* const  sWinStep = sDist + mc;  // per output position: the winning candidate's step, path and PTG
    uint16_t* const  sWinK    = reinterpret_cast<uint16_t*>(sWinStep + mc);
    uint8_t* const   sState   = reinterpret_cast<uint8_t*>(sWinK + mc);  // bit0 passed phase A, bit1 unblocked, bit2 valid after the goal-cell rule
    uint8_t* const   sLeader  = sState + mc;   // first candidate (in enumeration order) of its cell among the valid ones
    uint8_t* const   sWinPtg  = sLeader + mc;
    __shared__ int   sTps, sOut;
    __shared__ mppb_expand_node sN;  // the node's record (it may live in pinned host memory: fetched once)
    __shared__ double           sCS[2];

    // evaluator descriptors and (short) waypoint lists in shared memory: the waypoint evaluator walks its list once
    // per match and sample
    constexpr int           kMaxWpsShared = 64;
    __shared__ EvaluatorDev sEvals[MPPB_MAX_EVALUATORS];
    __shared__ float2       sWps[MPPB_MAX_EVALUATORS][kMaxWpsShared];
    const int node = blockIdx.x;
    const int c    = threadIdx.x;
    {
        constexpr int kEvalWords = sizeof(EvaluatorDev) / 8;
        static_assert(sizeof(EvaluatorDev) % 8 == 0, "descriptor is copied in 8-byte words");
        for (int i = c; i < nEvals * kEvalWords; i += blockDim.x)
            reinterpret_cast<uint64_t*>(sEvals)[i] = reinterpret_cast<const uint64_t*>(evalsGlobal)[i];
    }
    {
        static_assert(sizeof(mppb_expand_node) % 8 == 0, "record is copied in 8-byte words");
        constexpr int   kWords = sizeof(mppb_expand_node) / 8;
        const uint64_t* src    = reinterpret_cast<const uint64_t*>(nodes + node);
        if (c < kWords) reinterpret_cast<uint64_t*>(&sN)[c] = src[c];
        if (c >= 32 && c < 34) sCS[c - 32] = nodes4[4 * node + 2 + (c - 32)];
        if (c == 0)
        {
            sTps = 0;
            sOut = 0;
        }
    }
    __syncthreads();
    for (int e = 0; e < nEvals; e++)
        if (sEvals[e].kind == 2 && sEvals[e].nWps <= kMaxWpsShared)
        {
            const float2* src = sEvals[e].wps;
            for (int i = c; i < sEvals[e].nWps; i += blockDim.x) sWps[e][i] = src[i];
        }
    __syncthreads();
    if (c == 0)
        for (int e = 0; e < nEvals; e++)
            if (sEvals[e].kind == 2 && sEvals[e].nWps <= kMaxWpsShared) sEvals[e].wps = sWps[e];
    const EvaluatorDev* const evals = sEvals;
    const mppb_expand_node&   N     = sN;
    const double x0 = N.pose[0], y0 = N.pose[1], phi0 = N.pose[2];
    const double cs0 = sCS[0], sn0 = sCS[1];

    // ---- the candidate of this thread (enumeration order = slot order) ----
    int      p = -1, k = -1;
    uint32_t step = 0;
    bool     isTarget = false;
    size_t   idx = 0;  // index of (k, step) in the PTG's path arrays
    if (c < P.candBase[P.nPtgs])
    {
        p = 0;
        while (c >= P.candBase[p + 1]) p++;
        const int slot  = c - P.candBase[p];
        const int goalK = N.goal_k[p];
        if (slot == 0)
        {
            // direct-to-goal TPS point (TPS_Astar.cpp:650-672)
            isTarget = true;
            k        = goalK;
            step     = N.goal_step[p];
            if (goalK < 0 || step == 0xffffffffu) k = -1;
        }
        else
        {
            // merged ascending list: k subset + goal k (std::set of the reference, :619-625,:677)
            const int       m = (slot - 1) / P.nT, t = (slot - 1) - m * P.nT;
            const int       nK = P.nK[p];
            const uint16_t* ks = P.kSubset[p];
            int             pos = nK, len = nK;
            if (goalK >= 0)
            {
                pos = 0;
                while (pos < nK && static_cast<int>(ks[pos]) < goalK) pos++;
                if (!(pos < nK && static_cast<int>(ks[pos]) == goalK)) len = nK + 1;
            }
            if (m < len)
            {
                if (len == nK)
                    k = ks[m];
                else
                    k = m < pos ? ks[m] : (m == pos ? goalK : ks[m - 1]);
                step = P.tsStep[p][t];
                const PathTableDev& T = P.paths[p];
                if (step >= T.begin[k + 1] - T.begin[k]) k = -1;  // trjStep >= getPathStepCount(k) (:689-692)
            }
        }
        if (k >= 0 && step == 0) k = -1;  // no motion (:718)
    }

    // ---- phase A: end pose, world-box test, lattice cell ----
    uint8_t state = 0;
    double  ax = 0, ay = 0, aphi = 0, rx = 0, ry = 0;
    float   distF = 0;
    bool    inBox = false;
    if (k >= 0)
    {
        const PathTableDev& T = P.paths[p];
        idx                   = static_cast<size_t>(T.begin[k]) + step;
        rx                    = T.x[idx];
        ry                    = T.y[idx];
        const double rphi     = T.phi[idx];
        distF                 = T.dist[idx];
        // from.pose (+) relPose with the heading's cos/sin from the host
        ax   = x0 + rx * cs0 - ry * sn0;
        ay   = y0 + rx * sn0 + ry * cs0;
        aphi = wrap_to_pi_d(phi0 + rphi);
        const bool outLo = ax < N.bbox_min[0] || ay < N.bbox_min[1] || aphi < N.bbox_min[2];
        const bool outHi = ax > N.bbox_max[0] - P.halfCell || ay > N.bbox_max[1] - P.halfCell || aphi > N.bbox_max[2];
        inBox            = !outLo && !outHi;
    }
    // Programmatic dependent launch: this kernel may start while tpobs_grid_kernel is still running; everything above
    // reads tables and the node record only. The collision rows must wait for that kernel's completion.
    cudaGridDependencySynchronize();
    if (inBox)
    {
        const PathTableDev& T = P.paths[p];
        state       = 1;
        sCell[0][c] = x2idx_f(static_cast<float>(ax), P.resXY);
        sCell[1][c] = x2idx_f(static_cast<float>(ay), P.resXY);
        sCell[2][c] = phi2idx_f(static_cast<float>(aphi), P.resYaw);
        sDist[c]    = __float_as_uint(distF);
        // blocked iff relTrgDist >= freeDistance = min(initTPObstacleSingle(k), row[k]) (:744-757)
        const double raw      = rows[static_cast<size_t>(node) * totalPaths + T.colOffset + k];
        const double init     = T.initDist[k];
        const double freeDist = raw < init ? raw : init;
        if (!(static_cast<double>(distF) >= freeDist)) state |= 2;
    }
    sState[c] = state;
    if (state & 1) atomicAdd(&sTps, P.copies[p]);
    __syncthreads();

    // ---- goal-cell rule: a cell reached by the PTG's own (unblocked) direct-to-goal point is its alone (:765-778) ----
    if (state & 2)
    {
        bool valid = true;
        if (!isTarget)
        {
            const int t = P.candBase[p];
            if ((sState[t] & 2) && sCell[0][t] == sCell[0][c] && sCell[1][t] == sCell[1][c] && sCell[2][t] == sCell[2][c])
                valid = false;
        }
        if (valid) state |= 4;
        sState[c] = state;
    }
    __syncthreads();

    // ---- best candidate per reached cell: smallest relTrgDist, the earlier one on ties (:783-796) ----
    const int nCand = P.candBase[P.nPtgs];
    bool      winner = false;
    int       leader = c;
    if (state & 4)
    {
        const int cx = sCell[0][c], cy = sCell[1][c], cw = sCell[2][c];
        unsigned  best = sDist[c];
        int       bestIdx = c;
        for (int j = 0; j < nCand; j++)
        {
            if (!(sState[j] & 4) || sCell[0][j] != cx || sCell[1][j] != cy || sCell[2][j] != cw) continue;
            if (j < leader) leader = j;
            if (sDist[j] < best || (sDist[j] == best && j < bestIdx))
            {
                best    = sDist[j];
                bestIdx = j;
            }
        }
        winner = bestIdx == c;
    }
    sLeader[c] = (state & 4) && leader == c;
    __syncthreads();

    // ---- the winners write their cell's record at the position of the cell's first appearance ----
    // (records are assembled in shared memory and leave with 16-byte stores of consecutive threads: the destination
    // may be pinned host memory, where field-by-field stores of single threads would each be a PCIe transaction)
    mppb_neighbor* const o = reinterpret_cast<mppb_neighbor*>(sRecRaw);
    if (winner)
    {
        int pos = 0;
        for (int j = 0; j < leader; j++) pos += sLeader[j];
        sWinStep[pos] = step;
        sWinK[pos]    = static_cast<uint16_t>(k);
        sWinPtg[pos]  = static_cast<uint8_t>(p);
        const PathTableDev& T = P.paths[p];
        mppb_neighbor       r;
        r.cell[0]   = sCell[0][c];
        r.cell[1]   = sCell[1][c];
        r.cell[2]   = sCell[2][c];
        r.ptg_index = static_cast<uint8_t>(p);
        r.k         = static_cast<uint16_t>(k);
        r.step      = step;
        r.ptg_dist  = distF;
        r.pose[0]   = ax;
        r.pose[1]   = ay;
        r.pose[2]   = aphi;
        // getPathTwist: (v, 0, w) rotated by the step's heading (DiffDriveCollisionGridBased.cpp:880-896), then into
        // the global frame by the node's heading (TPS_Astar.cpp:291-297)
        const double v = T.v[idx], w = T.w[idx], tc = T.cs[idx], ts = T.sn[idx];
        const double vy0 = 0.0;
        const double lx = v * tc - vy0 * ts, ly = v * ts + vy0 * tc;
        r.vel[0] = lx * cs0 - ly * sn0;
        r.vel[1] = lx * sn0 + ly * cs0;
        r.vel[2] = w;
        r.in_bbox = !(ax < N.bbox_min[0] || ay < N.bbox_min[1] || aphi < N.bbox_min[2] || ax > N.bbox_max[0] ||
                      ay > N.bbox_max[1] || aphi > N.bbox_max[2]);
        r.cost    = 0;  // filled in below
        o[pos]    = r;
    }
    if (sLeader[c]) atomicAdd(&sOut, 1);
    __syncthreads();
    const int nOut = sOut;

    // ---- evaluator values of every reached cell's edge: thread per (edge, evaluator), look-ups eight at a time ----
    // samples: relative pose (0,0) at t=0, then getPathPose(k, ((i+1)*step)/(nSeg+2)) for i < nSeg, then the end pose;
    // equal time keys collapse (std::map), i.e. a repeated iStep contributes once (edge_interpolated_path.cpp:56-69)
    const int nSamp = P.nSeg + 2;
    for (int item = c; item < nOut * nEvals; item += blockDim.x)
    {
        const int           e = item / nEvals, slot = item - e * nEvals;
        const EvaluatorDev& E = evals[slot];
        const PathTableDev& T  = P.paths[sWinPtg[e]];
        const size_t        b  = T.begin[sWinK[e]];
        const uint32_t      st = sWinStep[e];
        const bool average = E.useAverage != 0;
        double     cost = .0;
        unsigned   cnt  = 0;
        uint32_t   prev = 0xffffffffu;
        for (int base = 0; base < nSamp; base += 8)
        {
            double   v[8];
            uint32_t is[8];
#pragma unroll
            for (int j = 0; j < 8; j++)
            {
                const int i = base + j;
                v[j]        = .0;
                is[j]       = 0xffffffffu;
                if (i >= nSamp) continue;
                // sample i: i == 0 -> step 0; 1..nSeg -> (i*step)/(nSeg+2); nSeg+1 -> the end step
                const uint32_t iStep = i == 0 ? 0u : (i == nSamp - 1 ? st : static_cast<uint32_t>((static_cast<size_t>(i) * st) / static_cast<size_t>(nSamp)));
                is[j]                = iStep;
                const double bx = T.x[b + iStep], by = T.y[b + iStep];
                // TPose2D::operator+ : x + b.x*cos - b.y*sin ; y + b.x*sin + b.y*cos (CostEvaluatorCostMap.cpp:149-150)
                const double px = x0 + bx * cs0 - by * sn0;
                const double py = y0 + bx * sn0 + by * cs0;
                v[j]            = evaluator_at(E, px, py);
            }
#pragma unroll
            for (int j = 0; j < 8; j++)
            {
                if (base + j >= nSamp) break;
                if (is[j] == prev) continue;  // same time key as the previous sample: one map entry
                prev = is[j];
                if (average)
                {
                    cost += v[j];
                    ++cnt;
                }
                else if (v[j] >= cost)
                {
                    cost = v[j];
                    cnt  = 1;
                }
            }
        }
        sVal[item] = cost / static_cast<double>(cnt);
    }
    __syncthreads();
    // cost_path_segment: estimatedExecTime + evaluators in slot order (Planner.cpp:22-26)
    for (int e = c; e < nOut; e += blockDim.x)
    {
        double cc = static_cast<double>(sWinStep[e]) * P.paths[sWinPtg[e]].dt;
        for (int s = 0; s < nEvals; s++) cc += sVal[e * nEvals + s];
        o[e].cost = cc;
    }
    __syncthreads();
    {
        static_assert(sizeof(mppb_neighbor) % sizeof(uint4) == 0, "records leave in 16-byte words");
        constexpr int kWordsPerRec = sizeof(mppb_neighbor) / sizeof(uint4);
        uint4*        dst          = reinterpret_cast<uint4*>(out + static_cast<size_t>(node) * P.maxCand);
        for (int i = c; i < nOut * kWordsPerRec; i += blockDim.x) dst[i] = sRecRaw[i];
    }
    if (c == 0)
    {
        counts[2 * node + 0] = static_cast<uint32_t>(nOut);
        counts[2 * node + 1] = static_cast<uint32_t>(sTps);
    }
}

// mrpt::round: nearest integer, ties to even (lrint in the default rounding mode)
int round_i(double v) { return static_cast<int>(std::lrint(v)); }

// (Re)build the device parameter block from the PTG set and the planner parameters; expandReady says whether the
// expansion is available.
int expand_prepare(mppb_ctx* ctx)
{
    if (ctx->expandReady) return MPPB_OK;
    ctx->expandHost = ExpandParamsDev{};
    if (!ctx->expandParamsSet) return fail(ctx, MPPB_ERR_STATE, "expansion parameters not set (mppb_set_expand_params)");
    const int nP = static_cast<int>(ctx->ptgKinds.size());
    if (nP == 0) return fail(ctx, MPPB_ERR_STATE, "no PTGs set");
    const mppb_astar_params& a = ctx->expandParams;
    ExpandParamsDev&         E = ctx->expandHost;
    E.resXY    = a.grid_resolution_xy;
    E.resYaw   = a.grid_resolution_yaw;
    E.halfCell = a.grid_resolution_xy * 0.5;
    E.nSeg     = static_cast<int>(a.path_interpolated_segments);
    E.nPtgs    = nP;
    E.nT       = static_cast<int>(a.num_timestamps);
    if (E.nT < 1 || E.nT > MPPB_MAX_TIMESTAMPS) return fail(ctx, MPPB_ERR_UNSUPPORTED, "too many ptg_sample_timestamps for the device expansion");
    if (a.max_ptg_trajectories_to_explore < 2) return fail(ctx, MPPB_ERR_INVALID_ARG, "max_ptg_trajectories_to_explore must be >= 2");
    if (a.max_ptg_speeds_to_explore < 1) return fail(ctx, MPPB_ERR_INVALID_ARG, "max_ptg_speeds_to_explore must be >= 1");
    if (E.nSeg < 0 || E.nSeg + 2 > 64) return fail(ctx, MPPB_ERR_UNSUPPORTED, "too many interpolated segments for the device expansion");
    // speeds {1/S, 2/S, ...} < 1.001 (TPS_Astar.cpp:636-640): DiffDrive_C paths ignore the trimmed speed, so every TPS
    // point stands for that many identical ones
    int          nSpeeds = 0;
    const double inc     = 1.0 / a.max_ptg_speeds_to_explore;
    for (double s = inc; s < 1.001; s += inc) nSpeeds++;
    E.candBase[0] = 0;
    int g = 0;
    for (int p = 0; p < nP; p++)
    {
        if (ctx->ptgKinds[p] != 0) return fail(ctx, MPPB_ERR_UNSUPPORTED, "device expansion needs collision-grid PTGs only");
        const PathTableHost& T = ctx->pathTables[p];
        if (!T.set) return fail(ctx, MPPB_ERR_STATE, "sampled paths of a PTG are missing (mppb_set_ptg_paths)");
        const PtgGridHost& G = ctx->grids[g++];
        if (G.K != T.K) return fail(ctx, MPPB_ERR_STATE, "path table and collision grid disagree on the number of paths");
        // k subset: round(i*(K-1)/(M-1)) with the integer division first (TPS_Astar.cpp:619-623), ascending, no duplicates
        std::vector<int> ks;
        const size_t     M = a.max_ptg_trajectories_to_explore;
        for (size_t i = 0; i < M; i++)
        {
            const size_t q = i * static_cast<size_t>(T.K - 1) / (M - 1);
            ks.push_back(round_i(static_cast<double>(q)));
        }
        std::sort(ks.begin(), ks.end());
        ks.erase(std::unique(ks.begin(), ks.end()), ks.end());
        if (ks.size() > MPPB_MAX_EXPLORE_PATHS) return fail(ctx, MPPB_ERR_UNSUPPORTED, "too many trajectories to explore for the device expansion");
        E.nK[p] = static_cast<int>(ks.size());
        for (size_t i = 0; i < ks.size(); i++) E.kSubset[p][i] = static_cast<uint16_t>(ks[i]);
        for (int t = 0; t < E.nT; t++) E.tsStep[p][t] = static_cast<uint32_t>(round_i(ctx->expandTimestamps[t] / T.dt));
        E.copies[p]       = nSpeeds;
        E.candBase[p + 1] = E.candBase[p] + 1 + (E.nK[p] + 1) * E.nT;
        PathTableDev& D   = E.paths[p];
        D.begin = T.begin.as<uint32_t>();
        D.x = T.x.as<float>(), D.y = T.y.as<float>(), D.phi = T.phi.as<float>(), D.dist = T.dist.as<float>();
        D.v = T.v.as<float>(), D.w = T.w.as<float>();
        D.cs = T.cs.as<double>(), D.sn = T.sn.as<double>();
        D.initDist  = T.initDist.as<double>();
        D.K         = T.K;
        D.colOffset = G.outOffset;
        D.dt        = T.dt;
    }
    if (E.candBase[nP] > kMaxCand) return fail(ctx, MPPB_ERR_UNSUPPORTED, "too many candidate TPS points per node for the device expansion");
    E.maxCand = E.candBase[nP];
    ctx->expandReady = true;  // (the block travels as a kernel parameter)
    return MPPB_OK;
}
}  // namespace

void expand_invalidate(mppb_ctx* ctx) { ctx->expandReady = false; }

int launch_expand(mppb_ctx* ctx, size_t nNodes, const mppb_expand_node* dIn, const double* dNodes4, const float* dRows,
                  mppb_neighbor* dOut, uint32_t* dCounts)
{
    if (nNodes == 0) return MPPB_OK;
    const int    threads = std::min(kMaxCand, (ctx->expandHost.maxCand + 31) / 32 * 32);
    // per candidate slot: record 80 + evaluator values 32 + cell 12 + dist 4 + step 4 + k 2 + three flags (see the kernel)
    const size_t mc   = (static_cast<size_t>(ctx->expandHost.maxCand) + 31) & ~static_cast<size_t>(31);
    const size_t smem = mc * (sizeof(mppb_neighbor) + MPPB_MAX_EVALUATORS * sizeof(double) + 3 * 4 + 4 + 4 + 2 + 3);
    if (smem > 40 * 1024)
        MPPB_CUDA(ctx, cudaFuncSetAttribute(expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    // launched with programmatic stream serialization: it may begin (parameters, tables, node records) while the
    // collision kernel before it on the stream is still draining, and waits for its rows inside the kernel
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim            = dim3(static_cast<unsigned>(nNodes));
    cfg.blockDim           = dim3(static_cast<unsigned>(threads));
    cfg.dynamicSmemBytes   = smem;
    cfg.stream             = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs                                          = attr;
    cfg.numAttrs                                       = 1;
    MPPB_CUDA(ctx, cudaLaunchKernelEx(&cfg, expand_kernel, ctx->expandHost, static_cast<const EvaluatorDev*>(ctx->evalDescs.as<EvaluatorDev>()),
                                      ctx->numEvals, dIn, dNodes4, dRows, ctx->totalPaths, dOut, dCounts));
    ctx->launches++;
    return MPPB_OK;
}
}  // namespace mppb

using namespace mppb;

namespace
{
template <class T>
T* mapped_alias_of(const T* hostPtr)
{
    if (!hostPtr) return nullptr;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, hostPtr) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
        return static_cast<T*>(attr.devicePointer);
    cudaGetLastError();
    return nullptr;
}

int expand_nodes_impl(mppb_ctx* ctx, size_t n, const mppb_expand_node* nodes, int useBoxes, double clip, mppb_neighbor* out,
                      uint32_t* outCounts, bool wait)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    if (n == 0) return MPPB_OK;
    if (!nodes || !out || !outCounts) return fail(ctx, MPPB_ERR_INVALID_ARG, "null host pointers");
    const int rc0 = expand_prepare(ctx);
    if (rc0 != MPPB_OK) return rc0;
    MPPB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t maxCand = static_cast<size_t>(ctx->expandHost.maxCand);
    const size_t cols    = static_cast<size_t>(ctx->totalPaths);
    // node rows (x, y, cos, sin) and obstacle boxes for the collision kernel, from the expansion records
    MPPB_CUDA(ctx, ctx->hExpandNodes.reserve(n * 4 * sizeof(double)));
    MPPB_CUDA(ctx, ctx->hExpandBoxes.reserve(n * 4 * sizeof(float)));
    double* hN = ctx->hExpandNodes.as<double>();
    float*  hB = ctx->hExpandBoxes.as<float>();
    for (size_t i = 0; i < n; i++)
    {
        const mppb_expand_node& N = nodes[i];
        hN[4 * i + 0]             = N.pose[0];
        hN[4 * i + 1]             = N.pose[1];
        if (N.cos_phi == 0.0 && N.sin_phi == 0.0)
        {
            hN[4 * i + 2] = std::cos(N.pose[2]);
            hN[4 * i + 3] = std::sin(N.pose[2]);
        }
        else  // taken by the caller (the planner's workers have them at hand: one thread would otherwise take them all)
        {
            hN[4 * i + 2] = N.cos_phi;
            hN[4 * i + 3] = N.sin_phi;
        }
        if (useBoxes)
        {
            // world box as apply_clipping_box sees it: double -> float (ObstacleSource.cpp:29-30)
            hB[4 * i + 0] = static_cast<float>(N.bbox_min[0]);
            hB[4 * i + 1] = static_cast<float>(N.bbox_min[1]);
            hB[4 * i + 2] = static_cast<float>(N.bbox_max[0]);
            hB[4 * i + 3] = static_cast<float>(N.bbox_max[1]);
        }
    }
    // small batches read their inputs in place from pinned memory (one A* expansion of one query); large ones are staged
    const bool              small = n <= 64;
    const double*           dN    = small ? mapped_alias_of(hN) : nullptr;
    const float*            dB    = useBoxes ? (small ? mapped_alias_of(hB) : nullptr) : nullptr;
    const mppb_expand_node* dIn   = small ? mapped_alias_of(nodes) : nullptr;
    if (!dN)
    {
        MPPB_CUDA(ctx, ctx->dExpandNodes.reserve(n * 4 * sizeof(double)));
        MPPB_CUDA(ctx, cudaMemcpyAsync(ctx->dExpandNodes.p, hN, n * 4 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        dN = ctx->dExpandNodes.as<double>();
    }
    if (useBoxes && !dB)
    {
        MPPB_CUDA(ctx, ctx->dExpandBoxes.reserve(n * 4 * sizeof(float)));
        MPPB_CUDA(ctx, cudaMemcpyAsync(ctx->dExpandBoxes.p, hB, n * 4 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        dB = ctx->dExpandBoxes.as<float>();
    }
    if (!dIn)
    {
        MPPB_CUDA(ctx, ctx->dExpandIn.reserve(n * sizeof(mppb_expand_node)));
        MPPB_CUDA(ctx, cudaMemcpyAsync(ctx->dExpandIn.p, nodes, n * sizeof(mppb_expand_node), cudaMemcpyHostToDevice, ctx->stream));
        dIn = ctx->dExpandIn.as<mppb_expand_node>();
    }
    MPPB_CUDA(ctx, ctx->dExpandRows.reserve(n * cols * sizeof(float)));
    ctx->expandRowsNodes = n;
    mppb_neighbor* dOut    = mapped_alias_of(out);
    uint32_t*      dCounts = mapped_alias_of(outCounts);
    const bool     outInPlace = dOut != nullptr && dCounts != nullptr;
    if (!outInPlace)
    {
        MPPB_CUDA(ctx, ctx->dExpandOut.reserve(n * maxCand * sizeof(mppb_neighbor)));
        MPPB_CUDA(ctx, ctx->dExpandCounts.reserve(n * 2 * sizeof(uint32_t)));
        dOut    = ctx->dExpandOut.as<mppb_neighbor>();
        dCounts = ctx->dExpandCounts.as<uint32_t>();
    }
    // (plain form: the expansion kernel is launched behind it with programmatic stream serialization)
    int rc = launch_tpobs_grid(ctx, n, dN, dB, clip, ctx->dExpandRows.as<float>(), /*allowNear=*/false);
    if (rc != MPPB_OK) return rc;
    rc = launch_expand(ctx, n, dIn, dN, ctx->dExpandRows.as<float>(), dOut, dCounts);
    if (rc != MPPB_OK) return rc;
    if (!outInPlace)
    {
        MPPB_CUDA(ctx, cudaMemcpyAsync(out, dOut, n * maxCand * sizeof(mppb_neighbor), cudaMemcpyDeviceToHost, ctx->stream));
        MPPB_CUDA(ctx, cudaMemcpyAsync(outCounts, dCounts, n * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (wait) MPPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MPPB_OK;
}
}  // namespace

extern "C" {

int mppb_set_ptg_paths(mppb_ctx* ctx, int ptg_index, const mppb_ptg_paths_desc* d)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    if (!d) return fail(ctx, MPPB_ERR_INVALID_ARG, "null path descriptor");
    if (ptg_index < 0 || ptg_index >= static_cast<int>(ctx->ptgKinds.size()) || ctx->ptgKinds[ptg_index] != 0)
        return fail(ctx, MPPB_ERR_INVALID_ARG, "ptg_index is not a collision-grid PTG of this context");
    if (d->num_paths <= 0 || !d->path_begin || !d->x || !d->y || !d->phi || !d->dist || !d->v || !d->w || !d->cos_phi ||
        !d->sin_phi || !(d->step_duration > 0) || !(d->ref_distance > 0))
        return fail(ctx, MPPB_ERR_INVALID_ARG, "bad path descriptor");
    for (int k = 0; k < d->num_paths; k++)
        if (d->path_begin[k + 1] <= d->path_begin[k]) return fail(ctx, MPPB_ERR_INVALID_ARG, "every path needs at least one step");
    MPPB_CUDA(ctx, cudaSetDevice(ctx->device));
    MPPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    expand_invalidate(ctx);
    PathTableHost& T = ctx->pathTables[ptg_index];
    const size_t   K = static_cast<size_t>(d->num_paths), n = d->path_begin[K];
    T.K              = d->num_paths;
    T.dt             = d->step_duration;
    T.refDistance    = d->ref_distance;
    T.hostBegin.assign(d->path_begin, d->path_begin + K + 1);
    // initTPObstacleSingle(k) = min(refDistance, getPathDist(k, last step)) (tp_obstacles_single_path.cpp:24)
    std::vector<double> init(K);
    for (size_t k = 0; k < K; k++) init[k] = std::min(d->ref_distance, static_cast<double>(d->dist[d->path_begin[k + 1] - 1]));
    auto up = [&](DevBuf& b, const void* src, size_t bytes) -> cudaError_t {
        cudaError_t e = b.reserve(std::max<size_t>(bytes, 8));
        if (e != cudaSuccess) return e;
        return cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
    };
    MPPB_CUDA(ctx, up(T.begin, d->path_begin, (K + 1) * sizeof(uint32_t)));
    MPPB_CUDA(ctx, up(T.x, d->x, n * sizeof(float)));
    MPPB_CUDA(ctx, up(T.y, d->y, n * sizeof(float)));
    MPPB_CUDA(ctx, up(T.phi, d->phi, n * sizeof(float)));
    MPPB_CUDA(ctx, up(T.dist, d->dist, n * sizeof(float)));
    MPPB_CUDA(ctx, up(T.v, d->v, n * sizeof(float)));
    MPPB_CUDA(ctx, up(T.w, d->w, n * sizeof(float)));
    MPPB_CUDA(ctx, up(T.cs, d->cos_phi, n * sizeof(double)));
    MPPB_CUDA(ctx, up(T.sn, d->sin_phi, n * sizeof(double)));
    MPPB_CUDA(ctx, up(T.initDist, init.data(), K * sizeof(double)));
    MPPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    T.set = true;
    return MPPB_OK;
}

int mppb_set_expand_params(mppb_ctx* ctx, const mppb_astar_params* a)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    if (!a || !a->ptg_sample_timestamps || a->num_timestamps == 0) return fail(ctx, MPPB_ERR_INVALID_ARG, "null parameters");
    expand_invalidate(ctx);
    ctx->expandParamsSet = false;
    if (a->num_timestamps > MPPB_MAX_TIMESTAMPS)
        return fail(ctx, MPPB_ERR_UNSUPPORTED, "too many ptg_sample_timestamps for the device expansion");
    ctx->expandParams = *a;
    for (uint32_t i = 0; i < a->num_timestamps; i++) ctx->expandTimestamps[i] = a->ptg_sample_timestamps[i];
    ctx->expandParams.ptg_sample_timestamps = ctx->expandTimestamps;
    ctx->expandParamsSet                    = true;
    return MPPB_OK;
}

uint32_t mppb_expand_max_neighbors(const mppb_ctx* ctx)
{
    if (!ctx) return 0;
    mppb_ctx* c = const_cast<mppb_ctx*>(ctx);
    const std::string keep = c->lastError;
    const int         rc   = expand_prepare(c);
    if (rc != MPPB_OK)
    {
        // "not available" is an answer here, not a failure of the context; the reason stays readable
        return 0;
    }
    c->lastError = keep;
    return static_cast<uint32_t>(ctx->expandHost.maxCand);
}

int mppb_expand_nodes(mppb_ctx* ctx, size_t n_nodes, const mppb_expand_node* nodes, int use_node_boxes, double clip_dist,
                      mppb_neighbor* out, uint32_t* out_counts)
{
    return expand_nodes_impl(ctx, n_nodes, nodes, use_node_boxes, clip_dist, out, out_counts, true);
}
int mppb_expand_nodes_begin(mppb_ctx* ctx, size_t n_nodes, const mppb_expand_node* nodes, int use_node_boxes,
                            double clip_dist, mppb_neighbor* out, uint32_t* out_counts)
{
    return expand_nodes_impl(ctx, n_nodes, nodes, use_node_boxes, clip_dist, out, out_counts, false);
}

int mppb_expand_last_rows(mppb_ctx* ctx, size_t n_nodes, float* out_rows)
{
    if (!ctx || !out_rows) return MPPB_ERR_INVALID_ARG;
    if (n_nodes > ctx->expandRowsNodes) return fail(ctx, MPPB_ERR_INVALID_ARG, "more rows asked for than the last expansion produced");
    MPPB_CUDA(ctx, cudaSetDevice(ctx->device));
    MPPB_CUDA(ctx, cudaMemcpyAsync(out_rows, ctx->dExpandRows.p, n_nodes * static_cast<size_t>(ctx->totalPaths) * sizeof(float),
                                   cudaMemcpyDeviceToHost, ctx->stream));
    MPPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MPPB_OK;
}

}  // extern "C"
