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
// cell exactly as CDynamicGrid::x2idx does (double division + truncation; a reciprocal
// multiply decides all cases that are not within 1e-7 of a cell boundary, the true division
// decides the rest). cos/sin of the node heading come from the host C library. The per-path
// result is an order-independent min, reduced with atomicMin on the (non-negative) float bit
// pattern in shared memory.
//
// Filtered fast path: the chain above is only needed to pick the right collision-grid CELL. Each
// point is first transformed in FP32 (translation with a hi/lo split of the node position, so the
// error of the local coordinates is bounded by E = 3e-6 * max(clip,1) m); if the FP32 result is
// farther than that bound from every decision boundary (clip square, cell edges, grid border,
// robot bounding circle) the cell it yields is provably the reference's cell and the FP64 chain
// is skipped. The ~0.5 % of points near a boundary take the exact FP64 chain.
//
// Work decomposition: one CTA per node, two phases per PTG group.
//  Phase 1 (points -> occupancy bitmap): the node's clipping square selects a range of bin
//    rows of the cloud index; each row contributes one contiguous run of points. Warps take
//    rows round-robin, lanes take consecutive points (coalesced 8-byte loads), transform them
//    and set the bit of the collision-grid cell they fall in (shared-memory bitmap, one bit per
//    cell). The rare points inside the robot polygon do not set a bit; their cell's (k,d) list
//    is applied directly with the BACK_AWAY rule.
//  Phase 2 (bitmap -> per-path minimum): every output column (PTG, k) owns its cells in ascending-d
//    order, cut into chunks that touch at most 4 words of the bitmap (a word is an 8 x 4 tile of cells,
//    so the ~25 cells a path sweeps in a few steps share words); a thread scans two columns, one chunk
//    = 4 shared-memory word tests per step, and stops at the first chunk with an occupied cell, whose
//    smallest occupied d IS the column's minimum. Blocked paths exit after a chunk or two, a free path
//    costs ~35 steps, and no atomics or variable-length per-point lists remain on the common path.
//    Nodes with few points in their bin window scatter the by-cell lists of the occupied cells instead.
#include <algorithm>
#include <cfloat>
#include <cstdlib>
#include <cstring>

#include "mppb_internal.cuh"

namespace mppb
{
namespace
{
#ifndef MPPB_GRID_THREADS
#define MPPB_GRID_THREADS 128
#endif
constexpr int      kThreads = MPPB_GRID_THREADS;
constexpr int      kWarps   = kThreads / 32;
constexpr int      kRowsPerPass = 64;   // bin rows whose runs are flattened together
constexpr int      kMaxChunks   = 512;  // 32-point chunks indexed per pass
constexpr int      kMaxInside   = 128;  // queued points inside the robot shape per node and group
#ifndef MPPB_GRID_SCATTER_MAX
#define MPPB_GRID_SCATTER_MAX 256
#endif
constexpr int      kLongList    = 32;   // by-cell lists longer than this are folded by the whole block
constexpr int      kMaxLong     = 128;  // queue of such cells per node and group
constexpr int      kScatterMax  = MPPB_GRID_SCATTER_MAX;  // <= this many points in the bin window: sparse phase 2
constexpr int      kSplit         = 16;  // CTAs per node of a small batch
constexpr int      kSplitMaxNodes = 32;  // batches of at most this many nodes are split
const bool         g_noSplit      = std::getenv("MPPB_NO_SPLIT") != nullptr;  // A/B measurements only
constexpr unsigned kInfBits = 0x7f800000u;
constexpr unsigned kFull    = 0xffffffffu;

// reference predicate transform_pc_square_clipping.cpp:30-31 (true => point skipped)
__device__ __forceinline__ bool clipped_out(double g, double center, double D)
{
    return fabs(__dsub_rn(g, center)) > D;
}

// CDynamicGrid::x2idx: (int)((v - vmin) / res), exact. The reciprocal multiply is trusted
// unless it lands within 1e-7 of a cell boundary; then the true division decides.
__device__ __noinline__ int grid_index_exact(double a, double res) { return __double2int_rz(__ddiv_rn(a, res)); }
__device__ __forceinline__ int grid_index(double v, double vmin, double res, double invRes)
{
    const double a = __dsub_rn(v, vmin);
    const double q = __dmul_rn(a, invRes);
    if (__builtin_expect(fabs(q - rint(q)) < 1e-7, 0)) return grid_index_exact(a, res);
    return __double2int_rz(q);
}

// mrpt::math::TPolygon2D::contains (PNPOLY) in the reference's operation order.
__device__ __noinline__ bool polygon_contains(const GridGroupDev& g, double px, double py)
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

// An obstacle inside the robot shape (internal_TPObsDistancePostprocess, COLL_BEH_BACK_AWAY):
// for every (k,d) of its cell, d < maxRobotRadius -> ignored, else path k is blocked (distance 0).
__device__ __forceinline__ void apply_inside_cell(const GridGroupDev& G, uint32_t cell, unsigned* smin, int first,
                                                  int stride)
{
    const uint32_t eb = G.offsets[cell], ee = G.offsets[cell + 1];
    const double   R  = G.maxRadius;
    for (uint32_t e = eb + first; e < ee; e += stride)
    {
        const uint2 kd = G.entries[e];
        if (!(static_cast<double>(__uint_as_float(kd.y)) < R)) atomicMin(&smin[kd.x], 0u);
    }
}

// Minimum d (as float bits; d >= 0, so unsigned order is float order) over the occupied cells h0..h3 of a chunk of
// a by-column list (GridGroupDev): d values are stored by (slot, rank of the bit within the slot's mask).
__device__ __noinline__ unsigned chunk_min_d(const float* __restrict__ d, uint4 m, unsigned h0, unsigned h1,
                                             unsigned h2, unsigned h3)
{
    unsigned best = kInfBits;
    auto     fold = [&](unsigned h, unsigned mask, unsigned base) {
        while (h)
        {
            const unsigned b = __ffs(h) - 1;
            h &= h - 1;
            best = min(best, __float_as_uint(d[base + __popc(mask & ((1u << b) - 1u))]));
        }
    };
    const unsigned b1 = __popc(m.x), b2 = b1 + __popc(m.y), b3 = b2 + __popc(m.z);
    fold(h0, m.x, 0);
    fold(h1, m.y, b1);
    fold(h2, m.z, b2);
    fold(h3, m.w, b3);
    return best;
}

// Phase 2 for sparse clouds: scatter the (column, d) lists of the few occupied cells.
// A column scan walks a whole ascending-d list (hundreds of dependent steps) when nothing blocks the path; with few
// occupied cells it is far cheaper to visit those cells' by-cell lists instead: every thread walks its share of the
// bitmap words and folds the list of each set bit into smin (order-independent atomicMin). Kept out of line so that
// its registers do not weigh on the dense path of the kernel.
__device__ __noinline__ void sparse_phase2(const GridGroupDev& G, const unsigned* bitmap, int words, unsigned* smin,
                                           uint32_t* longCell, unsigned* nLongPtr)
{
    // a thread folds the short lists of its cells itself; long lists (cells near the origin carry one entry per
    // path) are queued and folded by the whole block afterwards
    auto foldCell = [&](uint32_t cell, uint32_t eb, uint32_t ee) {
        if (ee - eb > static_cast<uint32_t>(kLongList))
        {
            const unsigned slot = atomicAdd(nLongPtr, 1u);
            if (slot < static_cast<unsigned>(kMaxLong))
            {
                longCell[slot] = cell;
                return;
            }
        }
        // eight entries per round: the loads are issued together (one L2 latency per round, not per entry)
        for (uint32_t e = eb; e < ee; e += 8)
        {
            uint2 kd[8];
#pragma unroll
            for (int j = 0; j < 8; j++) kd[j] = G.entries[min(e + j, ee - 1)];
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (e + j < ee) atomicMin(&smin[kd[j].x], kd[j].y);
        }
    };
    const uint32_t gnx = static_cast<uint32_t>(G.nx), tilesX = static_cast<uint32_t>(grid_tiles_x(G.nx));
    for (int i = threadIdx.x; i < words; i += blockDim.x)
    {
        unsigned bits = bitmap[i];
        if (!bits) continue;
        // word i is the 8 x 4 tile (i % tilesX, i / tilesX); bit b of it is cell (b & 7, b >> 3) of the tile
        const uint32_t ty = static_cast<uint32_t>(i) / tilesX, tx = static_cast<uint32_t>(i) - ty * tilesX;
        const uint32_t cellBase = (ty << 2) * gnx + (tx << 3);
        while (bits)
        {
            // two cells at a time: their list bounds are independent loads
            uint32_t b = __ffs(bits) - 1;
            bits &= bits - 1;
            const uint32_t c0 = cellBase + (b >> 3) * gnx + (b & 7u);
            uint32_t       c1 = c0;
            if (bits)
            {
                b = __ffs(bits) - 1;
                bits &= bits - 1;
                c1 = cellBase + (b >> 3) * gnx + (b & 7u);
            }
            const uint32_t b0 = G.offsets[c0], e0 = G.offsets[c0 + 1];
            const uint32_t b1 = G.offsets[c1], e1 = G.offsets[c1 + 1];
            foldCell(c0, b0, e0);
            if (c1 != c0) foldCell(c1, b1, e1);
        }
    }
    __syncthreads();
    const unsigned nl = min(*nLongPtr, static_cast<unsigned>(kMaxLong));
    for (unsigned q = 0; q < nl; q++)
    {
        const uint32_t cell = longCell[q];
        const uint32_t eb = G.offsets[cell], ee = G.offsets[cell + 1];
        for (uint32_t e = eb + threadIdx.x; e < ee; e += blockDim.x)
        {
            const uint2 kd = G.entries[e];
            atomicMin(&smin[kd.x], kd.y);
        }
    }
}

// HAS_BOX: obstacle points outside the node's own world box (float compare, inclusive bounds:
// ObstacleSource.cpp:29-40) are ignored — the per-query clipping of independent planning queries
// that share one resident cloud.
#ifndef MPPB_GRID_MIN_BLOCKS
#define MPPB_GRID_MIN_BLOCKS 9  // 56 registers: 9 CTAs per SM (a 4096-node batch is then 3.07 waves of 148 x 9)
#endif
// SPLIT > 1 (small batches, e.g. the single node of one A* expansion): SPLIT CTAs per node. Each repeats phase 1 (a
// few microseconds for one node) and takes a contiguous 1/SPLIT of the output columns in phase 2, eight lanes per column looking
// at eight consecutive chunks at a time. A lone CTA is bound by the latency of its dependent loads (a free path is
// ~35 chunk loads in a row, a sparse window a chain of by-cell lists: 13-27 us measured on the reference's demo
// map); split, the whole scan is a handful of load round trips and there is no communication between the CTAs.
// STAGED: phase 1 reads the node's points from shared memory, where a producer warp puts them with bulk asynchronous
// copies (cp.async.bulk + mbarrier transaction counts, the 1-D form of TMA): the bin-row runs of the node's window are
// cut into 16-byte-aligned pieces and packed, piece after piece, into a ring of kStages buffers of kStagePts points; the
// other warps consume stage s while the copies of stages s+1.. are in flight. The point loop is then a dense loop over
// shared memory — no chunk table, no per-iteration look-up of the next run, no global-load latency inside it.
constexpr int kStages   = 3;
constexpr int kStagePts = 512;  // 4 KB per stage
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// PEERS: the cloud is sharded over several GPUs (multigpu.cu). Each finished row is stored not into `out` but into slot
// `rank` of EVERY rank's receive area — this GPU's own and, over NVLink, its peers' — while the other CTAs are still
// evaluating their nodes, so the exchange overlaps the computation node by node; the last CTA to finish raises this
// rank's flag on every peer. A short kernel on each rank then waits for all flags and takes the minimum over the slots.
// BITMAP = false: collision grids whose occupancy bitmap does not fit shared memory (beyond ~1260 x 1260 cells, e.g.
// a 40 m reference distance at 0.05 m cells). No bitmap and no column scan: every point that passes phase 1 folds the
// (column, d) list of its cell straight into the per-path minima (atomicMin), the plain form of
// updateTPObstacleSingle (DiffDriveCollisionGridBased.cpp:826-838). Slower per point, correct for any grid size.
// MODE 1 / MODE 2: near-first evaluation of a large batch in two launches (NearPassDev, mppb_internal.cuh). MODE 1 looks
// only at the cloud bins within near.r of the node; a node whose paths are all blocked within near.boundBits is final
// and its row is written, any other node is appended to near.list. MODE 2 is the plain kernel over the nodes of that
// list (CTAs beyond its length leave at once).
// THREADS: CTA size (128; 512 for the second launch of the near-first form, whose CTAs each walk a whole window).
template <bool HAS_BOX, int SPLIT, bool STAGED, bool PEERS, bool BITMAP, int MODE, int THREADS>
__device__ __forceinline__ void tpobs_grid_node(const CloudBins& bins, const GridGroupDev* __restrict__ groups, int numGroups,
                                                int totalPaths, int bitmapWords, const double* __restrict__ nodes,
                                                const float4* __restrict__ boxes, double clip, float* __restrict__ out,
                                                const PeerRowsDev& peers, const NearPassDev& near, const int node, const int part)
{
    static_assert(MODE != 1 || (SPLIT == 1 && !STAGED && BITMAP), "near pass: one CTA per node, bitmap form");
    static_assert(!STAGED || THREADS >= 64, "staged form: row tables are filled by the first 64 threads");
    constexpr int kThreads = THREADS, kWarps = THREADS / 32;  // (shadow the namespace-level defaults)
    extern __shared__ __align__(16) unsigned smem[];
    unsigned*           smin   = smem;               // [totalPaths] float bits
    unsigned*           bitmap = smem + totalPaths;  // [bitmapWords] one bit per grid cell, in 8 x 4 tiles per word
    // STAGED: the ring of point buffers follows (16-byte aligned), kStages x kStagePts float2
    float2* const stageBuf = reinterpret_cast<float2*>(smem + ((totalPaths + bitmapWords + 3) & ~3));
    __shared__ __align__(8) uint64_t fullBar[kStages], emptyBar[kStages];  // data landed / every warp done with the slot
    uint32_t stageCount = 0;  // stages consumed so far by this CTA (slot = count % kStages, parity = (count / kStages) & 1)
    if constexpr (STAGED)
    {
        if (threadIdx.x == 0)
        {
            for (int i = 0; i < kStages; i++)
            {
                mbar_init(&fullBar[i], 1);
                mbar_init(&emptyBar[i], kWarps);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    __shared__ uint32_t rowBeg[kRowsPerPass];
    __shared__ uint32_t rowPrefix[kRowsPerPass + 1];
    __shared__ uint32_t rowLen[kRowsPerPass];
    __shared__ uint32_t chunkStart[kMaxChunks];
    __shared__ uint8_t  chunkCnt[kMaxChunks];
    __shared__ uint32_t insideCell[kMaxInside];
    __shared__ unsigned nInside;
    __shared__ unsigned ptsTotal;  // points in the bin window: few => sparse form of phase 2
    __shared__ uint32_t longCell[kMaxLong];
    __shared__ unsigned nLong;

    const int    lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double x0 = nodes[4 * node + 0], y0 = nodes[4 * node + 1];
    const double c = nodes[4 * node + 2], s = nodes[4 * node + 3];

    float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
    if (HAS_BOX) box = boxes[node];

    for (int i = threadIdx.x; i < totalPaths; i += kThreads) smin[i] = kInfBits;

    // inverse pose (CPose2D unary minus): ix = -x*c - y*s ; iy = x*s - y*c
    const double ix = __dsub_rn(__dmul_rn(-x0, c), __dmul_rn(y0, s));
    const double iy = __dsub_rn(__dmul_rn(x0, s), __dmul_rn(y0, c));

    // FP32 fast path constants: hi/lo split of the node position, float cos/sin, error bound E
    const float x0h = __double2float_rn(x0), x0l = __double2float_rn(x0 - static_cast<double>(x0h));
    const float y0h = __double2float_rn(y0), y0l = __double2float_rn(y0 - static_cast<double>(y0h));
    const float cf = __double2float_rn(c), sf = __double2float_rn(s);
    const float clipF  = __double2float_rn(clip);
    const float errLoc = 3e-6f * fmaxf(clipF, 1.f);  // bound on |fast local coordinate - reference's|
    const float clipHi = clipF + errLoc + 1e-6f * clipF, clipLo = clipF - errLoc - 1e-6f * clipF;

    // conservative float window for the bin range (the exact test is per point, in double)
    int        bx0 = 0, bx1 = -1, by0 = 0, by1 = -1;
    const bool any = bins.n > 0 && isfinite(x0) && isfinite(y0) && clip >= 0;
    if (any)
    {
        // MODE 1: the bins that hold every point within near.r of the node (the clip test per point stays the reference's)
        const double win = MODE == 1 ? fmin(near.r, clip) : clip;
        const float  lox = float_below(x0 - win);
        const float  hix = float_above(x0 + win);
        const float  loy = float_below(y0 - win);
        const float  hiy = float_above(y0 + win);
        bx0             = bin_coord(lox, bins.minx, bins.invBin, bins.nbx);
        bx1             = bin_coord(hix, bins.minx, bins.invBin, bins.nbx);
        by0             = bin_coord(loy, bins.miny, bins.invBin, bins.nby);
        by1             = bin_coord(hiy, bins.miny, bins.invBin, bins.nby);
    }

    int  sparseWords   = 0;      // > 0: phase 2 of the last group is done in its sparse form after the loop
    bool nearUndecided = false;  // MODE 1: this node goes to the list (same value in all threads)
    for (int gi = 0; any && gi < numGroups; gi++)
    {
        const GridGroupDev& G = groups[gi];
        // group constants into registers (the descriptor lives in global memory)
        const double          gxmin = G.xmin, gymin = G.ymin, gres = G.res, ginv = G.invRes;
        const int             gnx = G.nx, gny = G.ny, nCols = G.nCols;
        const float           guardR2  = G.insideGuardR2;
        const float           gxminf = __double2float_rn(gxmin), gyminf = __double2float_rn(gymin);
        const float           ginvf  = __double2float_rn(ginv);
        // margin (in cells) inside which the FP32 cell index is not trusted
        const float mq = 1.25f * ((errLoc + fabsf(__double2float_rn(gxmin - static_cast<double>(gxminf))) +
                                   fabsf(__double2float_rn(gymin - static_cast<double>(gyminf)))) * ginvf +
                                  3e-7f * static_cast<float>(max(gnx, gny))) + 1e-5f;
        const bool  fastOk   = mq < 0.25f && isfinite(clipF) && clipLo > 0.f;
        const float guardHi  = guardR2 + 4.f * errLoc * (sqrtf(guardR2) + errLoc) + 1e-6f;
        // q = (c*dx + s*dy - gxmin) * inv  and  (c*dy - s*dx - gymin) * inv with pre-scaled constants
        // (their rounding adds < 2e-7 * q to the error, inside the 3e-7 * n term of mq)
        const float cq = cf * ginvf, sq = sf * ginvf, offx = -gxminf * ginvf, offy = -gyminf * ginvf;
        const float halfMinusMq = 0.5f - mq;
        const uint32_t        tilesX   = static_cast<uint32_t>(grid_tiles_x(gnx));
        const uint32_t* const colStart = G.colStart;
        const uint4* const    colWord  = G.colWord;
        const uint4* const    colMask  = G.colMask;
        const int* const      colOut   = G.colOut;

        __syncthreads();  // previous group's phase 2 has finished reading the bitmap
        if (threadIdx.x == 0)
        {
            nInside  = 0;
            ptsTotal = 0;
            nLong    = 0;
        }
        const int words = grid_tiles_x(gnx) * grid_tiles_y(gny);
        if constexpr (BITMAP)
            for (int i = threadIdx.x; i < words; i += kThreads) bitmap[i] = 0u;

        // an obstacle point in cell (cx, cy) outside the robot shape: its bit, or (no bitmap) its cell's list at once
        auto markCell = [&](unsigned cx, unsigned cy) {
            if constexpr (BITMAP)
                atomicOr(&bitmap[grid_tile_word(cx, cy, tilesX)], 1u << grid_tile_bit(cx, cy));
            else
            {
                const uint32_t cell = cx + cy * static_cast<unsigned>(gnx);
                const uint32_t eb = G.offsets[cell], ee = G.offsets[cell + 1];
                for (uint32_t e = eb; e < ee; e++)
                {
                    const uint2 kd = G.entries[e];
                    atomicMin(&smin[kd.x], kd.y);
                }
            }
        };
        // one point of the node's window: FP32 filtered path, exact FP64 chain for the undecided ones
        auto processPoint = [&](const float2 g) {
            if (HAS_BOX && (g.x < box.x || g.x > box.z || g.y < box.y || g.y > box.w)) return;
            if (fastOk)
            {
                // ---- FP32 filtered path ----
                const float dx = (g.x - x0h) - x0l, dy = (g.y - y0h) - y0l;
                const float am = fmaxf(fabsf(dx), fabsf(dy));
                if (am > clipHi) return;  // certainly outside the clipping square
                // cell coordinates q = (R^T d - gmin) / res, folded into two FMAs per axis
                const float qx = fmaf(dx, cq, fmaf(dy, sq, offx));
                const float qy = fmaf(dy, cq, fmaf(-dx, sq, offy));
                const float rx = qx - floorf(qx), ry = qy - floorf(qy);
                const float edge = fmaxf(fabsf(rx - 0.5f), fabsf(ry - 0.5f));  // > 0.5-mq: near a cell edge
                const bool  unsure = am >= clipLo || edge > halfMinusMq || fmaf(dx, dx, dy * dy) <= guardHi;
                if (!unsure)
                {
                    // x2idx truncates toward zero: q in (-1,0) still maps to cell 0
                    const unsigned cx = static_cast<unsigned>(__float2int_rz(qx));
                    const unsigned cy = static_cast<unsigned>(__float2int_rz(qy));
                    if (cx >= static_cast<unsigned>(gnx) || cy >= static_cast<unsigned>(gny)) return;
                    markCell(cx, cy);
                    return;
                }
            }
            // ---- exact FP64 chain (reference operation order, no FMA) ----
            const double gx = g.x, gy = g.y;
            if (clipped_out(gx, x0, clip) || clipped_out(gy, y0, clip)) return;
            // composePoint with the inverse pose
            const double ox = __dadd_rn(__dadd_rn(ix, __dmul_rn(gx, c)), __dmul_rn(gy, s));
            const double oy = __dadd_rn(__dsub_rn(iy, __dmul_rn(gx, s)), __dmul_rn(gy, c));
            const float  fx = __double2float_rn(ox), fy = __double2float_rn(oy);  // CPointsMap stores float
            const double lx = fx, ly = fy;
            const int    cx = grid_index(lx, gxmin, gres, ginv);
            const int    cy = grid_index(ly, gymin, gres, ginv);
            if (cx < 0 || cx >= gnx || cy < 0 || cy >= gny) return;
            const uint32_t cell = cx + cy * gnx;
            // isPointInsideRobotShape: only points within the bounding radius can be inside
            if (__builtin_expect(fx * fx + fy * fy <= guardR2, 0))
            {
                if (polygon_contains(G, lx, ly))
                {
                    // obstacle inside the robot shape: no bit; its cell is queued for phase 1b
                    const unsigned slot = atomicAdd(&nInside, 1u);
                    if (slot < kMaxInside)
                        insideCell[slot] = cell;
                    else
                        apply_inside_cell(G, cell, smin, 0, 1);  // queue overflow: do it right here
                    return;
                }
            }
            markCell(static_cast<unsigned>(cx), static_cast<unsigned>(cy));
        };

        // ---- phase 1: points -> occupancy bitmap ----
        if constexpr (STAGED)
        {
            for (int rowBase = by0; rowBase <= by1; rowBase += kRowsPerPass)
            {
                const int nRows = min(kRowsPerPass, by1 - rowBase + 1);
                __syncthreads();  // bitmap cleared / previous pass done with the row tables and the stage buffers
                if (threadIdx.x < nRows)
                {
                    const size_t   r = static_cast<size_t>(rowBase + threadIdx.x) * bins.nbx;
                    const uint32_t b = bins.binStart[r + bx0], e = bins.binStart[r + bx1 + 1];
                    // 16-byte-aligned piece around the run: at most one foreign point at either end, which the clip
                    // test rejects like any other point outside the window (past the cloud's end: a sentinel)
                    const uint32_t ba = b & ~1u, ea = (e + 1u) & ~1u;
                    rowBeg[threadIdx.x]        = ba;
                    rowPrefix[threadIdx.x + 1] = e != b ? ea - ba : 0u;  // points staged for this row
                    if (e != b) atomicAdd(&ptsTotal, e - b);
                }
                __syncthreads();
                if (warp == 0)
                {
                    // inclusive scan of up to 64 piece lengths, two per lane
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
                const uint32_t total   = rowPrefix[nRows];
                const uint32_t nStages = (total + kStagePts - 1) / kStagePts;
                // producer (warp 0): the pieces of stage st, one row per lane and round, into slot (stageCount+st) % kStages
                auto issue = [&](uint32_t st) {
                    const uint32_t slot = (stageCount + st) % kStages;
                    const uint32_t lo = st * kStagePts, hi = min(total, lo + kStagePts);
                    uint32_t       myBytes = 0;
                    for (int r = lane; r < nRows; r += 32)
                    {
                        const uint32_t plo = max(rowPrefix[r], lo), phi = min(rowPrefix[r + 1], hi);
                        if (plo < phi) myBytes += (phi - plo) * static_cast<uint32_t>(sizeof(float2));
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) myBytes += __shfl_xor_sync(kFull, myBytes, o);
                    if (lane == 0)
                    {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the slot was read through the generic proxy
                        mbar_arrive_expect_tx(&fullBar[slot], myBytes);
                    }
                    __syncwarp();
                    for (int r = lane; r < nRows; r += 32)
                    {
                        const uint32_t plo = max(rowPrefix[r], lo), phi = min(rowPrefix[r + 1], hi);
                        if (plo < phi)
                            bulk_copy_g2s(stageBuf + static_cast<size_t>(slot) * kStagePts + (plo - lo),
                                          bins.pts + rowBeg[r] + (plo - rowPrefix[r]),
                                          (phi - plo) * static_cast<uint32_t>(sizeof(float2)), &fullBar[slot]);
                    }
                };
                if (warp == 0)
                    for (uint32_t st = 0; st < min(nStages, static_cast<uint32_t>(kStages)); st++) issue(st);
                for (uint32_t st = 0; st < nStages; st++)
                {
                    const uint32_t seq = stageCount + st, slot = seq % kStages;
                    mbar_wait(&fullBar[slot], (seq / kStages) & 1u);
                    const uint32_t cnt = min(total - st * kStagePts, static_cast<uint32_t>(kStagePts));
                    const float2*  sp  = stageBuf + static_cast<size_t>(slot) * kStagePts;
                    for (uint32_t i = threadIdx.x; i < cnt; i += kThreads) processPoint(sp[i]);
                    // slot released warp by warp; only the producer warp waits for the others before it refills the
                    // slot, the consumers run on into the next stage (already in flight or landed)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&emptyBar[slot]);
                    if (warp == 0 && st + kStages < nStages)
                    {
                        mbar_wait(&emptyBar[slot], (seq / kStages) & 1u);
                        issue(st + kStages);
                    }
                }
                stageCount += nStages;
            }
        }
        else if constexpr (MODE == 1)
        {
            // near pass: a handful of short bin rows — no row tables, no chunk table, no barriers: warps take rows,
            // lanes stride over a row's run
            __syncthreads();  // bitmap cleared
            for (int row = by0 + warp; row <= by1; row += kWarps)
            {
                const size_t   r = static_cast<size_t>(row) * bins.nbx;
                const uint32_t b = bins.binStart[r + bx0], e = bins.binStart[r + bx1 + 1];
                if (lane == 0 && e != b) atomicAdd(&ptsTotal, e - b);
                for (uint32_t i = b + lane; i < e; i += 32) processPoint(bins.pts[i]);
            }
        }
        else
#ifdef MPPB_DBG_SKIP_PHASE1  // timing attribution experiments only
        for (int rowBase = by0; rowBase <= by1 && clip < 0; rowBase += kRowsPerPass)
#else
        for (int rowBase = by0; rowBase <= by1; rowBase += kRowsPerPass)
#endif
        {
            const int nRows = min(kRowsPerPass, by1 - rowBase + 1);
            __syncthreads();  // bitmap cleared / previous pass done with the row tables
            for (int t = threadIdx.x; t < nRows; t += kThreads)
            {
                const size_t   r = static_cast<size_t>(rowBase + t) * bins.nbx;
                const uint32_t b = bins.binStart[r + bx0], e = bins.binStart[r + bx1 + 1];
                rowBeg[t]        = b;
                rowLen[t]        = e - b;
                rowPrefix[t + 1] = (e - b + 31) >> 5;     // 32-point chunks of this row's run
                if (e != b) atomicAdd(&ptsTotal, e - b);  // points in the bin window: decides the form of phase 2
            }
            __syncthreads();
            if constexpr (MODE == 1)
            {
                // too few points nearby to decide all paths: not worth a near pass (block-uniform decision)
                if (gi == 0 && rowBase == by0 && nRows == by1 - by0 + 1 && ptsTotal < near.minPoints)
                {
                    nearUndecided = true;
                    break;
                }
            }
            if (warp == 0)
            {
                // inclusive scan of up to 64 chunk counts, two per lane
                static_assert(kRowsPerPass == 64, "scan below handles 64 rows");
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
            const uint32_t nChunks = rowPrefix[nRows];
            for (uint32_t cBase = 0; cBase < nChunks; cBase += kMaxChunks)
            {
                const uint32_t nc = min(static_cast<uint32_t>(kMaxChunks), nChunks - cBase);
                if (cBase) __syncthreads();
                // chunk table: chunks never straddle two rows, so a chunk is (first point, count <= 32)
                for (uint32_t ch = threadIdx.x; ch < nc; ch += kThreads)
                {
                    const uint32_t gch = cBase + ch;
                    int            lo = 0, hi = nRows;
                    while (hi - lo > 1)
                    {
                        const int mid = (lo + hi) >> 1;
                        if (rowPrefix[mid] <= gch)
                            lo = mid;
                        else
                            hi = mid;
                    }
                    const uint32_t j = (gch - rowPrefix[lo]) << 5;
                    chunkStart[ch]   = rowBeg[lo] + j;
                    chunkCnt[ch]     = static_cast<uint8_t>(min(32u, rowLen[lo] - j));
                }
                __syncthreads();
                // software pipeline: the point of the warp's next chunk is loaded before the current one is used
                float2 gNext     = make_float2(0.f, 0.f);
                bool   validNext = false;
                if (warp < nc && lane < chunkCnt[warp])
                {
                    gNext     = bins.pts[chunkStart[warp] + lane];
                    validNext = true;
                }
                for (uint32_t ch = warp; ch < nc; ch += kWarps)
                {
                    const float2 g     = gNext;
                    const bool   valid = validNext;
                    validNext          = false;
                    if (ch + kWarps < nc && lane < chunkCnt[ch + kWarps])
                    {
                        gNext     = bins.pts[chunkStart[ch + kWarps] + lane];
                        validNext = true;
                    }
                    if (!valid) continue;
                    processPoint(g);
                }
            }
        }
        if constexpr (MODE == 1)
        {
            // too few points nearby to decide all paths (block-uniform): the node goes to the list
            __syncthreads();
            if (gi == 0 && ptsTotal < near.minPoints) nearUndecided = true;
            if (nearUndecided) break;
        }
        if constexpr (!BITMAP)
        {
            // no bitmap, no column scan: the points' lists are folded already; only the queued inside-shape points remain
            __syncthreads();
            const unsigned nIn = min(nInside, static_cast<unsigned>(kMaxInside));
            for (unsigned q = 0; q < nIn; q++) apply_inside_cell(G, insideCell[q], smin, threadIdx.x, kThreads);
            continue;
        }
        if constexpr (SPLIT > 1)
        {
            __syncthreads();  // end of phase 1
            const unsigned nIn = min(nInside, static_cast<unsigned>(kMaxInside));
            for (unsigned q = 0; q < nIn; q++) apply_inside_cell(G, insideCell[q], smin, threadIdx.x, kThreads);
            if (nIn) __syncthreads();
            // ---- phase 2, column-split form: 8 lanes per column, 16 columns per pass ----
            constexpr int  kLanes = 8, kColsPerPass = kThreads / kLanes;
            const int      sub = threadIdx.x & (kLanes - 1), grp = threadIdx.x / kLanes;
            const unsigned grpShift = lane & ~(kLanes - 1);  // where the group's lanes sit in a warp ballot
            // this CTA's columns: a contiguous block (its part of the output row is then contiguous too)
            const int perPart = (nCols + SPLIT - 1) / SPLIT;
            const int colEnd  = min(nCols, (part + 1) * perPart);
            for (int c0 = part * perPart; c0 < colEnd; c0 += kColsPerPass)
            {
                const int col = c0 + grp;
                uint32_t  i = 0, end = 0;
                int       oc = 0;
                if (col < colEnd)
                {
                    i   = colStart[col];
                    end = colStart[col + 1];
                    oc  = colOut[col];
                    if (smin[oc] == 0u) end = i;  // already blocked at distance 0 by a point inside the robot
                }
                bool open = i < end;  // (uniform over the 8 lanes of the column)
                i += sub;
                while (__any_sync(kFull, open))
                {
                    unsigned h0 = 0, h1 = 0, h2 = 0, h3 = 0;
                    uint4    m  = make_uint4(0, 0, 0, 0);
                    if (open && i < end)
                    {
                        const uint4 w = colWord[i];
                        m             = colMask[i];
                        h0 = bitmap[w.x] & m.x, h1 = bitmap[w.y] & m.y, h2 = bitmap[w.z] & m.z, h3 = bitmap[w.w] & m.w;
                    }
                    const unsigned hits = (__ballot_sync(kFull, (h0 | h1 | h2 | h3) != 0u) >> grpShift) & 0xffu;
                    if (hits)
                    {
                        // chunks are in ascending-d order: the lowest lane with a hit holds the column's minimum
                        if (open && sub == __ffs(hits) - 1)
                        {
                            const unsigned v = chunk_min_d(G.colD + G.colDBase[i], m, h0, h1, h2, h3);
                            if (v < smin[oc]) smin[oc] = v;  // the only writer of this column
                        }
                        open = false;
                    }
                    else
                    {
                        i += kLanes;
                        if (i - sub >= end) open = false;
                    }
                }
            }
            continue;  // next group
        }
        // Phase 2 works on two columns per thread at a time (independent load chains). Their first chunks do
        // not depend on the bitmap: request them before the barrier that ends phase 1, so that the L2 latency
        // overlaps the barrier wait and phase 1b.
        struct ColScan
        {
            uint32_t i, end;  // current / one-past-last chunk of the column
            uint4    w, m;    // the current chunk: 4 bitmap words and the column's cells in each
            int      oc;      // output column, -1 = nothing to scan
        };
        auto openColumn = [&](int col, ColScan& cs) {
            cs.oc = -1;
            cs.i = cs.end = 0;
            cs.w = cs.m = make_uint4(0, 0, 0, 0);
            if (col >= nCols) return;
            cs.i   = colStart[col];
            cs.end = colStart[col + 1];
            if (cs.i >= cs.end) return;
            cs.oc = colOut[col];
            cs.w  = colWord[cs.i];
            cs.m  = colMask[cs.i];
        };
        ColScan A, B;
        openColumn(threadIdx.x, A);
        openColumn(threadIdx.x + kThreads, B);
        __syncthreads();

        // ---- phase 1b: points inside the robot polygon (COLL_BEH_BACK_AWAY), all threads per cell ----
        {
            const unsigned n = min(nInside, static_cast<unsigned>(kMaxInside));
            for (unsigned q = 0; q < n; q++) apply_inside_cell(G, insideCell[q], smin, threadIdx.x, kThreads);
            if (n) __syncthreads();  // the column scan below updates smin without atomics
        }

        // ---- phase 2, sparse clouds (few points in the bin window): scatter form ----
        // (only for the last group, after the loop: a call in here would keep the whole kernel state alive across it;
        // earlier groups of a multi-geometry PTG set take the column scan, which is correct for any density)
        // (not in MODE 1: its points lie around the robot, where the by-cell lists are at their longest)
        if (MODE != 1 && gi == numGroups - 1 && ptsTotal <= static_cast<unsigned>(kScatterMax))
        {
            sparseWords = words;
            break;
        }

        // ---- phase 2: per output column, the first chunk (ascending d) with an occupied cell ----
        // one step of a column scan: test the chunk in hand (4 words of the bitmap against the column's cells in
        // them); on a hit the minimum d over the occupied cells of the chunk is the column's result
        auto step = [&](ColScan& cs) -> bool {
            const unsigned h0 = bitmap[cs.w.x] & cs.m.x, h1 = bitmap[cs.w.y] & cs.m.y;
            const unsigned h2 = bitmap[cs.w.z] & cs.m.z, h3 = bitmap[cs.w.w] & cs.m.w;
            if (h0 | h1 | h2 | h3)
            {
                const unsigned v = chunk_min_d(G.colD + G.colDBase[cs.i], cs.m, h0, h1, h2, h3);
                if (v < smin[cs.oc]) smin[cs.oc] = v;  // this thread is the only writer of the column now
                return false;
            }
            if (++cs.i >= cs.end) return false;
            cs.w = colWord[cs.i];
            cs.m = colMask[cs.i];
            return true;
        };
#ifdef MPPB_DBG_SKIP_PHASE2  // timing attribution experiments only
        for (int base = 0; clip < 0;)
#else
        for (int base = 0;;)
#endif
        {
            // a column already blocked at distance 0 by a point inside the robot needs no scan
            bool actA = A.oc >= 0 && smin[A.oc] != 0u;
            bool actB = B.oc >= 0 && smin[B.oc] != 0u;
            while (actA || actB)
            {
                if (actA) actA = step(A);
                if (actB) actB = step(B);
            }
            base += 2 * kThreads;
            if (base >= nCols) break;
            openColumn(base + threadIdx.x, A);
            openColumn(base + threadIdx.x + kThreads, B);
        }
    }
    if constexpr (MODE != 1)
        if (sparseWords) sparse_phase2(groups[numGroups - 1], bitmap, sparseWords, smin, longCell, &nLong);
    __syncthreads();
    if constexpr (MODE == 1)
    {
        // final iff every path is blocked within the bound (a point outside the near window cannot lower such a minimum)
        bool open = nearUndecided;
        if (!open)
            for (int i = threadIdx.x; i < totalPaths; i += kThreads) open |= smin[i] > near.boundBits;
        if (__syncthreads_or(open))
        {
            if (threadIdx.x == 0) near.list[atomicAdd(near.count, 1u)] = static_cast<uint32_t>(node);
            return;
        }
    }
    // this CTA's columns of the row: all of it, or (SPLIT) its contiguous share of every group's columns
    auto forMyColumns = [&](auto&& f) {
        if constexpr (SPLIT > 1)
        {
            for (int gi = 0; gi < numGroups; gi++)
            {
                const GridGroupDev& G       = groups[gi];
                const int           perPart = (G.nCols + SPLIT - 1) / SPLIT;
                const int           colEnd  = min(G.nCols, (part + 1) * perPart);
                for (int col = part * perPart + static_cast<int>(threadIdx.x); col < colEnd; col += kThreads) f(G.colOut[col]);
            }
        }
        else
            for (int i = threadIdx.x; i < totalPaths; i += kThreads) f(i);
    };
    if constexpr (PEERS)
    {
        // slot `rank` of every rank's receive area (parity half of this epoch), coalesced stores; then, once every CTA
        // of this step has done so, the flag. A step evaluated near-first is two launches: a row the near pass decides
        // counts for all kSplit parts in which the second launch would have delivered it.
        const size_t slot = (static_cast<size_t>(peers.epoch & 1u) * peers.world + peers.rank) * peers.slotStride +
                            static_cast<size_t>(node) * totalPaths;
        for (int p = 0; p < peers.world; p++)
        {
            float* dst = peers.rows[p] + slot;
            forMyColumns([&](int i) { dst[i] = __uint_as_float(smin[i]); });
        }
        __syncthreads();  // the row stores of all threads are observed by thread 0 ...
        if (threadIdx.x == 0)
        {
            const unsigned weight   = MODE == 1 ? static_cast<unsigned>(kSplit) : 1u;
            const unsigned expected = MODE == 0 ? gridDim.x : near.totalNodes * static_cast<unsigned>(kSplit);
            __threadfence_system();  // ... whose fence orders them, system-wide, before its arrival (fence cumulativity)
            const unsigned arrived = atomicAdd(peers.done, weight);
            if (arrived + weight == expected)
            {
                *peers.done = 0;  // for the next step (stream order)
                __threadfence_system();
                for (int p = 0; p < peers.world; p++)
                    *reinterpret_cast<volatile uint32_t*>(peers.flags[p] + peers.rank) = peers.epoch;
            }
        }
        return;
    }
    float* o = out + static_cast<size_t>(node) * totalPaths;
    forMyColumns([&](int i) { o[i] = __uint_as_float(smin[i]); });
}

#ifndef MPPB_GRID_NEAR_MIN_BLOCKS
#define MPPB_GRID_NEAR_MIN_BLOCKS 20
#endif
#ifndef MPPB_GRID_NEAR128_MIN_BLOCKS
#define MPPB_GRID_NEAR128_MIN_BLOCKS 10  // the near pass keeps less state: 48 registers, 10 CTAs per SM (measured: 9 -> 10: -4 %, 12: spills)
#endif
template <bool HAS_BOX, int SPLIT, bool STAGED = false, bool PEERS = false, bool BITMAP = true, int MODE = 0,
          int THREADS = MPPB_GRID_THREADS>
__global__ void __launch_bounds__(THREADS, THREADS < 128 ? MPPB_GRID_NEAR_MIN_BLOCKS
                                           : THREADS > 128 ? 512 / THREADS
                                                           : (SPLIT > 1 ? 4 : (MODE == 1 ? MPPB_GRID_NEAR128_MIN_BLOCKS : MPPB_GRID_MIN_BLOCKS)))
    tpobs_grid_kernel(CloudBins bins, const GridGroupDev* __restrict__ groups, int numGroups, int totalPaths,
                      int bitmapWords, const double* __restrict__ nodes /* [n][4] x,y,cos,sin */,
                      const float4* __restrict__ boxes /* [n] minx,miny,maxx,maxy */, double clip,
                      float* __restrict__ out, const PeerRowsDev peers, const NearPassDev near)
{
    // A kernel launched behind this one with programmatic stream serialization (expand_kernel) may start its prologue
    // now; it waits for this grid's completion (and memory flush) before it reads the rows written below.
    cudaTriggerProgrammaticLaunchCompletion();
    if constexpr (MODE == 2)
    {
        // the nodes the near pass left in its list, SPLIT CTAs per node, this grid striding over them. Launched with
        // programmatic stream serialization: resident while the near pass drains, released when its list is complete.
        cudaGridDependencySynchronize();
        const uint32_t cnt = *reinterpret_cast<volatile uint32_t*>(near.count);
        if (blockIdx.x == 0 && threadIdx.x == 0)
        {
            *near.lastCount  = cnt;  // what the adaptive switch looks at (pinned host memory)
            *near.countOther = 0;    // the counter the NEXT near pass on this slot appends to
        }
        for (uint32_t item = blockIdx.x; item < cnt * static_cast<uint32_t>(SPLIT); item += gridDim.x)
        {
            tpobs_grid_node<HAS_BOX, SPLIT, STAGED, PEERS, BITMAP, MODE, THREADS>(
                bins, groups, numGroups, totalPaths, bitmapWords, nodes, boxes, clip, out, peers, near,
                static_cast<int>(near.list[item / SPLIT]), static_cast<int>(item % SPLIT));
            __syncthreads();  // the next node reuses the shared arrays
        }
    }
    else
        tpobs_grid_node<HAS_BOX, SPLIT, STAGED, PEERS, BITMAP, MODE, THREADS>(
            bins, groups, numGroups, totalPaths, bitmapWords, nodes, boxes, clip, out, peers, near,
            SPLIT > 1 ? blockIdx.x / SPLIT : blockIdx.x, SPLIT > 1 ? blockIdx.x % SPLIT : 0);
}
}  // namespace

// Near-first evaluation (NearPassDev): decides whether this launch takes the two-launch form and, if so, fills `np` and
// claims a slot of pending-list buffers. The adaptive switch (mode 1) looks at how many nodes of earlier batches needed
// the second launch — the count a finished second launch leaves in pinned memory — and keeps to the plain form for the
// next 256 calls whenever that was more than 2 % of the batch.
// half-size of the near window and the bound a launch with this clipping distance would use; false: not applicable
bool near_window(const mppb_ctx* ctx, double clip, double& r, float& bound)
{
    if (!(ctx->nearSlackDiag > 0) || ctx->bins.n == 0 || !(ctx->bins.invBin > 0)) return false;
    double maxRadius = 0;
    for (const auto& g : ctx->groups) maxRadius = std::max(maxRadius, g.dev.maxRadius);
    const double sd = ctx->nearSlackDiag, binSize = 1.0 / static_cast<double>(ctx->bins.invBin);
    static const double rFactor = std::getenv("MPPB_GRID_NEAR_R") ? std::atof(std::getenv("MPPB_GRID_NEAR_R")) : 4.0;  // experiments
    // 4 x (slack + cell diagonal): on a dense cloud (C3) 3 x leaves 11 of 4096 nodes to the second launch, whose
    // latency then shows; 4 x leaves none, 5 x makes the first launch slower
    r     = std::max({rFactor * sd, 2.0 * binSize, 2.0 * maxRadius + 0.1});
    bound = static_cast<float>(r - sd - 1e-3);  // (1 mm: float rounding of the local coordinates)
    return r <= 0.4 * clip && bound > 0.f;
}
constexpr size_t kNearMinNodes  = 256;
constexpr size_t kNearRestNodes = 16;   // listed nodes the second launch takes at a time (kSplit CTAs each)
static int near_pass_setup(mppb_ctx* ctx, size_t nNodes, double clip, bool eligible, NearPassDev& np, bool& use)
{
    use = false;
    static const bool envOff = std::getenv("MPPB_NO_GRID_NEAR") != nullptr;  // A/B measurements only
    double r     = 0;
    float  bound = 0;
    if (!eligible || envOff || ctx->gridNear == 0 || nNodes <= kNearMinNodes || !near_window(ctx, clip, r, bound)) return MPPB_OK;
    if (!ctx->nearLastCount)
    {
        MPPB_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void**>(&ctx->nearLastCount), mppb_ctx::kNearSlots * sizeof(uint32_t),
                                     cudaHostAllocMapped));
        std::memset(ctx->nearLastCount, 0, mppb_ctx::kNearSlots * sizeof(uint32_t));
        MPPB_CUDA(ctx, ctx->nearCounters.reserve(mppb_ctx::kNearSlots * 2 * sizeof(uint32_t)));
        MPPB_CUDA(ctx, cudaMemset(ctx->nearCounters.p, 0, mppb_ctx::kNearSlots * 2 * sizeof(uint32_t)));
    }
    // what earlier two-launch calls reported
    for (int i = 0; i < mppb_ctx::kNearSlots; i++)
    {
        const uint32_t nodes = ctx->nearNodesOfSlot[i], pend = ctx->nearLastCount[i];
        // a listed node costs the second launch kSplit CTAs that each walk the whole window: beyond a few per cent of
        // the batch the two launches together are slower than the plain form
        if (nodes && pend * 50ull > nodes && ctx->gridNear == 1) ctx->nearSkip = 256;
    }
    if (ctx->nearSkip > 0 && ctx->gridNear == 1)
    {
        ctx->nearSkip--;
        for (int i = 0; i < mppb_ctx::kNearSlots; i++)
            if (ctx->nearNodesOfSlot[i])
            {
                ctx->nearPendingSeen += ctx->nearLastCount[i];
                ctx->nearNodesSeen += ctx->nearNodesOfSlot[i];
                ctx->nearNodesOfSlot[i] = 0;
                ctx->nearLastCount[i]   = 0;
            }
        return MPPB_OK;
    }
    const unsigned slot = ctx->nearSlot++ % mppb_ctx::kNearSlots;
    if (ctx->nearNodesOfSlot[slot])
    {
        ctx->nearPendingSeen += ctx->nearLastCount[slot];
        ctx->nearNodesSeen += ctx->nearNodesOfSlot[slot];
    }
    MPPB_CUDA(ctx, ctx->nearList[slot].reserve(nNodes * sizeof(uint32_t)));
    ctx->nearLastCount[slot]   = 0;
    ctx->nearNodesOfSlot[slot] = static_cast<uint32_t>(nNodes);
    ctx->nearLaunches++;
    np.r          = r;
    np.boundBits  = 0;
    std::memcpy(&np.boundBits, &bound, sizeof(float));
    np.minPoints  = 48;
    const unsigned parity = ctx->nearParity[slot];
    ctx->nearParity[slot] ^= 1;
    np.count      = ctx->nearCounters.as<uint32_t>() + 2 * slot + parity;
    np.countOther = ctx->nearCounters.as<uint32_t>() + 2 * slot + (parity ^ 1u);
    np.list       = ctx->nearList[slot].as<uint32_t>();
    np.lastCount  = ctx->nearLastCount + slot;  // (mapped pinned memory: the same address on the device under UVA)
    np.totalNodes = static_cast<uint32_t>(nNodes);
    use           = true;
    return MPPB_OK;
}

int launch_tpobs_grid(mppb_ctx* ctx, size_t nNodes, const double* dNodes, const float* dBoxes, double clip,
                      float* dOut, bool allowNear)
{
    if (nNodes == 0) return MPPB_OK;
    if (ctx->grids.empty()) return fail(ctx, MPPB_ERR_STATE, "no PTG tables set (mppb_set_ptg_grid)");
    if (nNodes > 0x7fffffffull) return fail(ctx, MPPB_ERR_INVALID_ARG, "too many nodes in one batch");
    size_t words = 1;
    for (const auto& g : ctx->groups)
        words = std::max(words, static_cast<size_t>(grid_tiles_x(g.dev.nx)) * grid_tiles_y(g.dev.ny));
    // small batches: kSplit CTAs per node (see the kernel), else one
    const bool split = nNodes <= static_cast<size_t>(kSplitMaxNodes) && !g_noSplit;
    // the staged form (points through cp.async.bulk into a shared-memory ring) of the one-CTA-per-node kernel
    static const bool staged = std::getenv("MPPB_GRID_STAGED") != nullptr && std::atoi(std::getenv("MPPB_GRID_STAGED")) != 0;
    const bool        useStaged = staged && !split;
    size_t            smem = ((static_cast<size_t>(ctx->totalPaths) + words + 3) & ~static_cast<size_t>(3)) * sizeof(unsigned);
    if (useStaged) smem += static_cast<size_t>(kStages) * kStagePts * sizeof(float2);
    // a grid whose occupancy bitmap does not fit shared memory takes the form without one (see BITMAP)
    const bool noBitmap = smem > 200 * 1024;
    if (noBitmap) smem = ((static_cast<size_t>(ctx->totalPaths) + 3) & ~static_cast<size_t>(3)) * sizeof(unsigned);
    if (smem > 200 * 1024) return fail(ctx, MPPB_ERR_UNSUPPORTED, "too many PTG paths for the shared-memory minima");
    NearPassDev np;
    bool        useNear = false;
    {
        const int rcn = near_pass_setup(ctx, nNodes, clip, allowNear && !split && !useStaged && !noBitmap, np, useNear);
        if (rcn != MPPB_OK) return rcn;
    }
    auto       launch = [&](auto kernel, unsigned ctasPerNode, size_t gridNodes = 0, unsigned threads = kThreads, bool pdl = false) -> int {
        if (smem > 48 * 1024)
            MPPB_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim            = dim3(static_cast<unsigned>(gridNodes ? gridNodes : nNodes) * ctasPerNode);
        cfg.blockDim           = dim3(threads);
        cfg.dynamicSmemBytes   = smem;
        cfg.stream             = ctx->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs                                          = attr;
        cfg.numAttrs                                       = pdl ? 1 : 0;
        MPPB_CUDA(ctx, cudaLaunchKernelEx(&cfg, kernel, ctx->bins, static_cast<const GridGroupDev*>(ctx->groupDescs.as<GridGroupDev>()),
                                          static_cast<int>(ctx->groups.size()), ctx->totalPaths, static_cast<int>(words), dNodes,
                                          reinterpret_cast<const float4*>(dBoxes), clip, dOut, PeerRowsDev{}, static_cast<const NearPassDev&>(np)));
        ctx->launches++;
        return MPPB_OK;
    };
    int rc;
    if (useNear)
    {
        // rest pass: kSplit CTAs per listed node, a grid for kNearRestNodes of them at a time; 512 threads per CTA, since
        // every one of a node's CTAs walks the whole window (one listed node in a C3 batch: +16 us with 128 threads, +8 us
        // with 512)
        const size_t restNodes = std::min(nNodes, kNearRestNodes);
        rc = dBoxes ? launch(tpobs_grid_kernel<true, 1, false, false, true, 1>, 1) : launch(tpobs_grid_kernel<false, 1, false, false, true, 1>, 1);
        static const bool restPdl = std::getenv("MPPB_GRID_REST_NO_PDL") == nullptr;  // (A/B measurements)
        if (rc == MPPB_OK)
            rc = dBoxes ? launch(tpobs_grid_kernel<true, kSplit, false, false, true, 2, 512>, kSplit, restNodes, 512, restPdl)
                        : launch(tpobs_grid_kernel<false, kSplit, false, false, true, 2, 512>, kSplit, restNodes, 512, restPdl);
    }
    else if (noBitmap)
        rc = dBoxes ? launch(tpobs_grid_kernel<true, 1, false, false, false>, 1) : launch(tpobs_grid_kernel<false, 1, false, false, false>, 1);
    else if (split)
        rc = dBoxes ? launch(tpobs_grid_kernel<true, kSplit>, kSplit) : launch(tpobs_grid_kernel<false, kSplit>, kSplit);
    else if (useStaged)
        rc = dBoxes ? launch(tpobs_grid_kernel<true, 1, true>, 1) : launch(tpobs_grid_kernel<false, 1, true>, 1);
    else
        rc = dBoxes ? launch(tpobs_grid_kernel<true, 1>, 1) : launch(tpobs_grid_kernel<false, 1>, 1);
    if (rc != MPPB_OK) return rc;
    MPPB_CUDA(ctx, cudaGetLastError());
    return MPPB_OK;
}

// the one-CTA-per-node kernel with the peer-memory epilogue (see PEERS)
int launch_tpobs_grid_peers(mppb_ctx* ctx, size_t nNodes, const double* dNodes, double clip, const PeerRowsDev& peers)
{
    if (nNodes == 0) return MPPB_OK;
    if (ctx->grids.empty()) return fail(ctx, MPPB_ERR_STATE, "no PTG tables set (mppb_set_ptg_grid)");
    size_t words = 1;
    for (const auto& g : ctx->groups)
        words = std::max(words, static_cast<size_t>(grid_tiles_x(g.dev.nx)) * grid_tiles_y(g.dev.ny));
    const size_t smem = ((static_cast<size_t>(ctx->totalPaths) + words + 3) & ~static_cast<size_t>(3)) * sizeof(unsigned);
    if (smem > 200 * 1024)
        return fail(ctx, MPPB_ERR_UNSUPPORTED, "collision grid too large for the shared-memory occupancy bitmap");
    NearPassDev np;
    bool        useNear = false;
    {
        const int rcn = near_pass_setup(ctx, nNodes, clip, true, np, useNear);
        if (rcn != MPPB_OK) return rcn;
    }
    auto launch = [&](auto kernel, size_t grid, unsigned threads) -> int {
        if (smem > 48 * 1024)
            MPPB_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        kernel<<<static_cast<unsigned>(grid), threads, smem, ctx->stream>>>(
            ctx->bins, ctx->groupDescs.as<GridGroupDev>(), static_cast<int>(ctx->groups.size()), ctx->totalPaths,
            static_cast<int>(words), dNodes, nullptr, clip, nullptr, peers, np);
        ctx->launches++;
        return MPPB_OK;
    };
    int rc;
    if (useNear)
    {
        rc = launch(tpobs_grid_kernel<false, 1, false, true, true, 1>, nNodes, kThreads);
        if (rc == MPPB_OK)
            rc = launch(tpobs_grid_kernel<false, kSplit, false, true, true, 2, 512>, std::min(nNodes, kNearRestNodes) * kSplit, 512);
    }
    else
        rc = launch(tpobs_grid_kernel<false, 1, false, true>, nNodes, kThreads);
    if (rc != MPPB_OK) return rc;
    MPPB_CUDA(ctx, cudaGetLastError());
    return MPPB_OK;
}

}  // namespace mppb
