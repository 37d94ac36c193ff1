// Internal declarations of libmpp_b200 (context, device-side table layouts, launch helpers).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/mppb.h"

namespace mppb
{
// ---------------------------------------------------------------------------------------
// Device-side layouts
// ---------------------------------------------------------------------------------------

// Uniform-bin index over the obstacle cloud (global frame). Points are stored sorted by bin
// (row-major: y-row slowest) as interleaved float2, so the points of `ix0..ix1` of one bin
// row are ONE contiguous run: [binStart[iy*nbx+ix0], binStart[iy*nbx+ix1+1]).
struct CloudBins
{
    const float2*   pts      = nullptr;  // [n] sorted by bin
    const uint32_t* binStart = nullptr;  // [nbx*nby+1]
    float           minx = 0, miny = 0;  // bin grid origin
    float           invBin = 0;          // 1 / bin size
    int             nbx = 0, nby = 0;
    uint32_t        n = 0;
};

// The bin coordinate of a float coordinate: monotone non-decreasing in v, which is all the
// query needs for exactness (a point inside [lo,hi] has bin(lo) <= bin(v) <= bin(hi)).
__host__ __device__ inline int bin_coord(float v, float origin, float invBin, int nb)
{
    const float q = (v - origin) * invBin;
    int         i = q <= 0.f ? 0 : (q >= static_cast<float>(nb) ? nb - 1 : static_cast<int>(q));
    return i < nb ? i : nb - 1;
}

// Conservative float bounds of a double interval end for the bin-window look-up: at least one float ulp below / above
// v (the exact clip test is per point, in double). Cheaper than nextafterf on the device (~4 instructions against ~25
// with branches, four times per node or candidate); +-inf stay what they are.
#ifdef __CUDACC__
__device__ __forceinline__ float float_below(double v) { return __double2float_rd(v - 1e-7 * fabs(v) - 1e-30); }
__device__ __forceinline__ float float_above(double v) { return __double2float_ru(v + 1e-7 * fabs(v) + 1e-30); }
#endif

// A group of collision-grid PTGs (DiffDriveCollisionGridBased family) that share the grid
// geometry and the robot shape, as the kernel sees it: ONE merged CSR whose entries carry the
// output column (ptg offset + k) and the float bits of the TP-distance d. Points are mapped to
// a cell once per group instead of once per PTG.
// Tiled occupancy bitmap of a collision grid: word and bit of cell (cx, cy); tilesX = grid_tiles_x(nx).
__host__ __device__ inline int      grid_tiles_x(int nx) { return (nx + 7) >> 3; }
__host__ __device__ inline int      grid_tiles_y(int ny) { return (ny + 3) >> 2; }
__host__ __device__ inline uint32_t grid_tile_word(uint32_t cx, uint32_t cy, uint32_t tilesX)
{
    return (cy >> 2) * tilesX + (cx >> 3);
}
__host__ __device__ inline uint32_t grid_tile_bit(uint32_t cx, uint32_t cy) { return ((cy & 3u) << 3) | (cx & 7u); }

struct GridGroupDev
{
    double          xmin, ymin, res, invRes;  // CDynamicGrid extents/resolution of the collision grid
    int             nx, ny;
    const uint32_t* offsets;  // by-cell CSR [nx*ny+1]
    const uint2*    entries;  // by-cell entries (output column, float bits of d)
    // by-column view of the same table, for the column scan of tpobs_grid_kernel. The kernel's occupancy bitmap
    // is tiled: one 32-bit word covers 8 x 4 cells (grid_tile_word / grid_tile_bit below). Column j of the group
    // owns the CHUNKS [colStart[j], colStart[j+1]) and writes output column colOut[j]. A chunk is a run of the
    // column's cells in ascending-d order that touches at most 4 bitmap words: colWord[c] holds the 4 word
    // indices and colMask[c] the cells' bits in each (unused slots: mask 0), so one step of the scan tests the
    // whole run with 4 loads from shared memory. The d of the cell at (slot s, bit b) of chunk c is
    // colD[colDBase[c] + popc(mask[0..s-1]) + popc(mask[s] & ((1 << b) - 1))]; all d of chunk c are <= all d of
    // chunk c+1, so the first chunk with a hit holds the column's minimum.
    int             nCols;
    const uint32_t* colStart;
    const uint4*    colWord;
    const uint4*    colMask;
    const uint32_t* colDBase;
    const float*    colD;
    const int*      colOut;
    int             nverts;
    double          maxRadius;
    float           insideGuardR2;  // float upper bound of maxRadius^2 (quick reject of PNPOLY)
    double          shapeX[MPPB_MAX_SHAPE_VERTS];
    double          shapeY[MPPB_MAX_SHAPE_VERTS];
};

// HolonomicBlend PTG constants as the kernel sees them (indexed by the context's PTG index)
struct HoloPtgDev
{
    double robotRadius;  // CPTG_RobotShape_Circular radius
    double vMax;         // V_MAX (cruise-phase distance uses the untrimmed V_MAX, HolonomicBlend.cpp:835)
};

// One cost evaluator slot as the edge-cost kernel sees it
struct EvaluatorDev
{
    int           kind;  // 0 = empty, 1 = CostEvaluatorCostMap, 2 = CostEvaluatorPreferredWaypoint
    int           useAverage;
    // costmap (CDynamicGrid<double> geometry)
    double        xmin, ymin, res;
    int           nx, ny;
    const double* cells;
    // preferred waypoints (stored as float, like CSimplePointsMap)
    const float2* wps;
    int           nWps;
    double        radius, scale;
    float         radiusSqrF;
};

// Sampled paths of one collision-grid PTG as the expansion kernel sees them (expand.cu)
struct PathTableDev
{
    const uint32_t* begin;     // [K+1]
    const float *   x, *y, *phi, *dist, *v, *w;
    const double *  cs, *sn;   // cos / sin of (double)phi, host libm
    const double*   initDist;  // [K] initTPObstacleSingle: min(refDistance, dist of the last step)
    int             K, colOffset;
    double          dt;
};
// Everything of TPS_Astar_Parameters + PTG set that shapes the candidate set of an expansion
struct ExpandParamsDev
{
    double       resXY, resYaw, halfCell;
    int          nSeg, nPtgs, nT, maxCand;
    int          nK[MPPB_MAX_PTGS];       // size of the k subset per PTG
    int          candBase[MPPB_MAX_PTGS + 1];  // first candidate slot of each PTG: 1 + (nK + 1) * nT slots per PTG
    int          copies[MPPB_MAX_PTGS];   // identical per-speed copies each TPS point stands for
    uint16_t     kSubset[MPPB_MAX_PTGS][MPPB_MAX_EXPLORE_PATHS];  // ascending, no duplicates
    uint32_t     tsStep[MPPB_MAX_PTGS][MPPB_MAX_TIMESTAMPS];      // round(t / dt)
    PathTableDev paths[MPPB_MAX_PTGS];
};

// Fused merge of cloud-sharded rows (multigpu.cu): where tpobs_grid_kernel sends a finished row. Every rank owns
// `world` receive slots per epoch parity, slot r filled by rank r; flags[r] = last epoch rank r completed.
struct PeerRowsDev
{
    float*    rows[MPPB_MAX_RANKS];   // base of rank p's receive area: [2][world][maxNodes][cols]
    uint32_t* flags[MPPB_MAX_RANKS];  // rank p's flag array [world]
    int       world, rank;
    uint32_t  epoch;                  // this step (1, 2, ...); parity selects the half of the receive area
    size_t    slotStride;             // floats per slot = maxNodes * cols
    unsigned* done;                   // local counter of finished CTAs (reset by the last one)
};

// Near-first evaluation of large node batches (tpobs_grid_kernel, MODE 1 and 2). Every entry (k, d) of a collision-grid
// cell satisfies d >= dist(robot origin, cell) - slack, slack = max over the table of that difference (the robot's largest
// radius, measured on the tables when they are uploaded): an obstacle point farther than r from the node can only block
// a path at a distance >= r - cell diagonal - slack. A first pass therefore looks only at the cloud bins around the node
// (half-size r); if every path already has a minimum <= bound = r - cell diagonal - slack, no farther point can lower any
// of them and the row is final. Nodes that cannot be decided that way (free space ahead, or too few points nearby to
// be worth trying) are appended to a list and a second launch evaluates them on the whole clipping window.
struct NearPassDev
{
    double    r          = 0;        // MODE 1: half-size of the near window, global axes
    unsigned  boundBits  = 0;        // MODE 1: float bits of `bound`
    unsigned  minPoints  = 0;        // MODE 1: fewer points in the near bins -> straight to the list
    uint32_t* count      = nullptr;  // nodes in the list (MODE 1 appends, MODE 2 consumes)
    uint32_t* countOther = nullptr;  // the counter of this slot's next use; MODE 2 zeroes it (nobody reads it meanwhile)
    uint32_t* list       = nullptr;
    uint32_t* lastCount  = nullptr;  // host-visible copy of `count` of the last MODE 2 launch (adaptive switch)
    uint32_t  totalNodes = 0;        // rows the two launches produce together (peer epilogue)
};

// ---------------------------------------------------------------------------------------
// Host-side context
// ---------------------------------------------------------------------------------------
struct DevBuf
{
    void*  p     = nullptr;
    size_t bytes = 0;
    // grow-only device buffer
    cudaError_t reserve(size_t need)
    {
        if (need <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p             = nullptr;
        bytes         = 0;
        cudaError_t e = cudaMalloc(&p, need);
        if (e == cudaSuccess) bytes = need;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p     = nullptr;
        bytes = 0;
    }
    template <class T>
    T* as() const
    {
        return static_cast<T*>(p);
    }
};

struct PinnedBuf
{
    void*       p     = nullptr;
    size_t      bytes = 0;
    cudaError_t reserve(size_t need)
    {
        if (need <= bytes) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p             = nullptr;
        bytes         = 0;
        cudaError_t e = cudaHostAlloc(&p, need, cudaHostAllocDefault);
        if (e == cudaSuccess) bytes = need;
        return e;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p     = nullptr;
        bytes = 0;
    }
    template <class T>
    T* as() const
    {
        return static_cast<T*>(p);
    }
};

// One obstacle cloud on the device: raw SoA (concatenated sources) + its bin index
struct CloudStore
{
    DevBuf    rawX, rawY;
    size_t    rawN = 0;
    DevBuf    sortedPts, binStart, binCursor, scanScratch;
    CloudBins bins;
    double    binSize = 0;
    bool      hasBox  = false;
    float     box[4]  = {0, 0, 0, 0};  // minx, miny, maxx, maxy (inclusive, float compare)
    void      release()
    {
        rawX.release();
        rawY.release();
        sortedPts.release();
        binStart.release();
        binCursor.release();
        scanScratch.release();
        rawN = 0;
        bins = CloudBins();
    }
};

// scratch of mppb_build_collision_grid (colgrid_build.cu); the result's host arrays live here until the next build
struct ColgridBuildBufs
{
    DevBuf    dBegin, dX, dY, dDist, dCs, dDense, dCounts, dOffsets, dScan, dEntryK, dEntryD;
    PinnedBuf hCs, hOffsets, hEntryK, hEntryD;
    double    stats[4] = {0, 0, 0, 0};
    void      release()
    {
        for (DevBuf* b : {&dBegin, &dX, &dY, &dDist, &dCs, &dDense, &dCounts, &dOffsets, &dScan, &dEntryK, &dEntryD}) b->release();
        for (PinnedBuf* b : {&hCs, &hOffsets, &hEntryK, &hEntryD}) b->release();
    }
};

// host copy of one uploaded collision-grid PTG (kept to rebuild the merged groups)
struct PtgGridHost
{
    int                   K = 0, outOffset = 0;
    double                xmin = 0, ymin = 0, res = 0, maxRadius = 0;
    int                   nx = 0, ny = 0;
    std::vector<uint32_t> offsets;
    std::vector<uint16_t> entryK;
    std::vector<float>    entryD;
    std::vector<double>   shapeX, shapeY;
    bool sameGeometryAndShape(const PtgGridHost& o) const
    {
        return xmin == o.xmin && ymin == o.ymin && res == o.res && nx == o.nx && ny == o.ny &&
               maxRadius == o.maxRadius && shapeX == o.shapeX && shapeY == o.shapeY;
    }
};
// device copy of one PTG's sampled paths
struct PathTableHost
{
    bool   set = false;
    int    K   = 0;
    double dt = 0, refDistance = 0;
    DevBuf begin, x, y, phi, dist, v, w, cs, sn, initDist;
    std::vector<uint32_t> hostBegin;  // step counts are needed on the host to size the candidate set
    void   release()
    {
        for (DevBuf* b : {&begin, &x, &y, &phi, &dist, &v, &w, &cs, &sn, &initDist}) b->release();
        set = false;
    }
};
struct GridGroupHost
{
    GridGroupDev dev;
    DevBuf       offsets, entries, colStart, colWord, colMask, colDBase, colD, colOut;
    void         release()
    {
        offsets.release();
        entries.release();
        colStart.release();
        colWord.release();
        colMask.release();
        colDBase.release();
        colD.release();
        colOut.release();
    }
};

}  // namespace mppb

struct mppb_ctx
{
    int          device    = 0;
    cudaStream_t stream    = nullptr;
    bool         ownStream = false;
    int          numSMs    = 0;
    std::string  lastError;
    uint64_t     launches = 0;
    int          holoFilter = 1;  // tpobs_holo_kernel: 0 no pre-filter, 1 pre-filter + bin culling, 2 pre-filter only (mppb_set_holo_filter)
    bool         holoByNode = false; // host batches grouped by node run tpobs_holo_node_kernel (mppb_set_holo_by_node)
    cudaEvent_t  evBegin = nullptr, evEnd = nullptr;
    // second stream + fork/join events: the host entry points pipeline sub-batches over two streams
    cudaStream_t auxStream = nullptr;
    cudaEvent_t  evFork = nullptr, evJoin = nullptr;
    cudaStream_t auxMore[2] = {nullptr, nullptr};  // sub-batch streams 3 and 4 of mppb_eval_nodes
    cudaEvent_t  evJoinMore[2] = {nullptr, nullptr};

    // obstacle cloud: raw SoA (concatenated sources) + bin index
    mppb::CloudStore cloud;
    // second cloud used only by the costmap builder (it may differ from the planner's cloud)
    mppb::CloudStore auxCloud;
    mppb::DevBuf     bboxScratch;
    mppb::CloudBins& bins = cloud.bins;

    // PTGs in planner order: kind per index (0 = collision grid, 1 = HolonomicBlend)
    std::vector<int>                 ptgKinds;
    std::vector<mppb::PtgGridHost>   grids;   // one per collision-grid PTG, in PTG order
    std::vector<mppb::GridGroupHost> groups;  // merged tables, rebuilt when a PTG is added
    mppb::DevBuf                     groupDescs;  // GridGroupDev[numGroups] on the device
    int                              totalPaths = 0;
    // near-first evaluation (NearPassDev): table statistics from rebuild_groups, per-launch slots, adaptive switch
    double       nearSlackDiag = 0;     // max over groups of (slack + cell diagonal); 0: not applicable (e.g. a path without cells)
    int          gridNear      = 1;     // mppb_set_grid_near: 0 off, 1 adaptive (default), 2 always
    static constexpr int kNearSlots = 8;
    mppb::DevBuf nearList[kNearSlots];  // pending-node lists
    mppb::DevBuf nearCounters;          // [kNearSlots][2] count, arrived
    uint32_t*    nearLastCount = nullptr;   // pinned, [kNearSlots]
    uint32_t     nearNodesOfSlot[kNearSlots] = {};
    unsigned     nearSlot = 0, nearSkip = 0;
    uint8_t      nearParity[kNearSlots] = {};  // which of a slot's two counters its next use appends to
    uint64_t     nearLaunches = 0, nearPendingSeen = 0, nearNodesSeen = 0;  // statistics (mppb_grid_near_stats)
    std::vector<mppb::HoloPtgDev>    holos;      // indexed by PTG index (zero for other kinds)
    mppb::DevBuf                     holoDescs;  // HoloPtgDev[MPPB_MAX_PTGS]

    // cost evaluators
    mppb::EvaluatorDev evals[MPPB_MAX_EVALUATORS] = {};
    mppb::DevBuf       evalCells[MPPB_MAX_EVALUATORS], evalWps[MPPB_MAX_EVALUATORS];
    mppb::DevBuf       evalDescs;  // EvaluatorDev[MPPB_MAX_EVALUATORS]
    int                numEvals = 0;
    uint64_t           evalEpoch = 0;  // bumped by every change of the evaluator slots
    uint64_t           cloudEpoch = 0; // bumped by every change of the resident obstacle cloud (mppb_set_obstacles*)
    uint64_t           ptgEpoch   = 0; // bumped by mppb_clear_ptgs / mppb_set_ptg_*

    mppb::ColgridBuildBufs colgrid;

    // several GPUs (multigpu.cu): NCCL communicator (opened at run time) and the peer-memory receive slots
    void*        ncclComm  = nullptr;
    int          commWorld = 1, commRank = 0;
    mppb::DevBuf peerLocal;                 // this rank's receive area + flags (cudaMalloc, exported through IPC)
    size_t       peerMaxNodes = 0, peerCols = 0;
    void*        peerOpened[MPPB_MAX_RANKS] = {};  // cudaIpcOpenMemHandle mappings of the other ranks' areas
    mppb::PeerRowsDev peer{};               // device pointers as the kernel takes them (world == 0: not attached)
    uint32_t     peerEpoch = 0;
    mppb::DevBuf peerDone;

    // device-side expansion (expand.cu): sampled paths per PTG index, the planner parameters, staging
    mppb::PathTableHost  pathTables[MPPB_MAX_PTGS];
    bool                 expandParamsSet = false;
    mppb_astar_params    expandParams{};
    double               expandTimestamps[MPPB_MAX_TIMESTAMPS] = {};
    mppb::ExpandParamsDev expandHost{};   // what was last uploaded
    bool                 expandReady = false;  // expandDev matches PTGs + parameters
    mppb::DevBuf         expandDev;       // ExpandParamsDev on the device
    mppb::DevBuf         dExpandIn, dExpandRows, dExpandOut, dExpandCounts, dExpandNodes, dExpandBoxes;
    mppb::PinnedBuf      hExpandNodes, hExpandBoxes;
    size_t               expandRowsNodes = 0;  // rows held by dExpandRows (mppb_expand_last_rows)

    // staging for the host entry points
    void*           sincosHelper = nullptr;  // capi.cu: helper thread that takes part of a large batch's heading cos/sin
    unsigned        helperPause = 0, helperCalls = 0, helperLoses = 0;  // calls it sits out / large calls so far / probes it lost
    double          helperEmaWith = 0, helperEmaWithout = 0;            // running call times (us) with and without it
    int             helperAllowed = -1;  // -1: not decided yet; 0: the process has too few CPUs to spare one
    mppb::PinnedBuf hNodes, hOut, hStage, hNodesHolo;
    mppb::DevBuf    dNodes, dOut, dStage, dStage2, dStage3, dBoxes, dNodesHolo, dBoxesHolo, dCandBegin;
};

namespace mppb
{
int  fail(mppb_ctx* ctx, int code, const std::string& msg);
int  cuda_fail(mppb_ctx* ctx, cudaError_t e, const char* what);
void set_global_error(const std::string& msg);

#define MPPB_CUDA(ctx, call)                                           \
    do {                                                               \
        cudaError_t e__ = (call);                                      \
        if (e__ != cudaSuccess) return mppb::cuda_fail(ctx, e__, #call); \
    } while (0)

// kernels / launchers implemented in the .cu files
int build_cloud_bins(mppb_ctx* ctx, CloudStore& cs, double binSize);
bool near_window(const mppb_ctx* ctx, double clip, double& r, float& bound);  // near-first evaluation: window and bound
int launch_tpobs_grid(mppb_ctx* ctx, size_t nNodes, const double* dNodes, const float* dBoxes, double clip,
                      float* dOut, bool allowNear = true);
// same, each finished row also stored into slot `rank` of every attached peer (fused merge of cloud-sharded rows)
int launch_tpobs_grid_peers(mppb_ctx* ctx, size_t nNodes, const double* dNodes, double clip, const PeerRowsDev& peers);
void multigpu_release(mppb_ctx* ctx);
// nNodes = rows of dNodes if the caller knows it (1 lets the kernel fetch the row early), 0 otherwise
int launch_tpobs_holo(mppb_ctx* ctx, size_t nCands, const mppb_holo_cand* dCands, const double* dNodes,
                      const float* dBoxes, double clip, double* dOut, size_t nNodes = 0);
// one CTA per node: candidates grouped by node, dCandBegin[nNodes+1] offsets into dCands
int launch_tpobs_holo_by_node(mppb_ctx* ctx, size_t nNodes, const uint32_t* dCandBegin, size_t maxCandsPerNode,
                              const mppb_holo_cand* dCands, const double* dNodes, const float* dBoxes, double clip, double* dOut);
// nSamples = dOffsets[nEdges] if the caller knows it (selects the small-batch kernel), 0 otherwise
int launch_edge_costs(mppb_ctx* ctx, size_t nEdges, const double* dFrom, const uint32_t* dOffsets, const double* dRel,
                      double* dOut, size_t nSamples = 0);
int launch_costmap_d2(mppb_ctx* ctx, const CloudStore& cs, double xmin, double ymin, double res, int nx, int ny,
                      float clearance, float* dOut);
// expansion kernel over the collision rows dRows [n][totalPaths] (expand.cu)
int launch_expand(mppb_ctx* ctx, size_t nNodes, const mppb_expand_node* dIn, const double* dNodes4, const float* dRows,
                  mppb_neighbor* dOut, uint32_t* dCounts);
void expand_invalidate(mppb_ctx* ctx);  // PTGs or parameters changed: the device parameter block must be rebuilt

}  // namespace mppb
