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

// A group of collision-grid PTGs (DiffDriveCollisionGridBased family) that share the grid
// geometry and the robot shape, as the kernel sees it: ONE merged CSR whose entries carry the
// output column (ptg offset + k) and the float bits of the TP-distance d. Points are mapped to
// a cell once per group instead of once per PTG.
struct GridGroupDev
{
    double          xmin, ymin, res, invRes;  // CDynamicGrid extents/resolution of the collision grid
    int             nx, ny;
    const uint32_t* offsets;  // by-cell CSR [nx*ny+1]
    const uint2*    entries;  // by-cell entries (output column, float bits of d)
    // by-column view of the same table: column j of the group owns [colStart[j], colStart[j+1])
    // of (colCell, colD), sorted by ascending d, and writes output column colOut[j].
    int             nCols;
    const uint32_t* colStart;
    const uint32_t* colCell;
    const float*    colD;
    const int*      colOut;
    int             nverts;
    double          maxRadius;
    float           insideGuardR2;  // float upper bound of maxRadius^2 (quick reject of PNPOLY)
    double          shapeX[MPPB_MAX_SHAPE_VERTS];
    double          shapeY[MPPB_MAX_SHAPE_VERTS];
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
struct GridGroupHost
{
    GridGroupDev dev;
    DevBuf       offsets, entries, colStart, colCell, colD, colOut;
    void         release()
    {
        offsets.release();
        entries.release();
        colStart.release();
        colCell.release();
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
    cudaEvent_t  evBegin = nullptr, evEnd = nullptr;
    // second stream + fork/join events: the host entry points pipeline sub-batches over two streams
    cudaStream_t auxStream = nullptr;
    cudaEvent_t  evFork = nullptr, evJoin = nullptr;

    // obstacle cloud: raw SoA (concatenated sources) + bin index
    mppb::DevBuf    rawX, rawY;
    size_t          rawN = 0;
    mppb::DevBuf    sortedPts, binStart, binCursor, bboxScratch;
    mppb::CloudBins bins;
    double          binSize = 0;

    // PTG tables
    std::vector<mppb::PtgGridHost>   grids;   // one per PTG, in PTG order
    std::vector<mppb::GridGroupHost> groups;  // merged tables, rebuilt when a PTG is added
    mppb::DevBuf                     groupDescs;  // GridGroupDev[numGroups] on the device
    int                            totalPaths = 0;

    // staging for the host entry points
    mppb::PinnedBuf hNodes, hOut;
    mppb::DevBuf    dNodes, dOut;
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
int build_cloud_bins(mppb_ctx* ctx, double binSize);
int launch_tpobs_grid(mppb_ctx* ctx, size_t nNodes, const double* dNodes, double clip, float* dOut);

}  // namespace mppb
