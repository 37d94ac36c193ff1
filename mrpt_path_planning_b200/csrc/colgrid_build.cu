// colgrid_sweep_kernel & co — construction of a collision-grid PTG's look-up table on the device
// (SURVEY.md §8f-2). Replaces the "else" branch of
//   DiffDriveCollisionGridBased::internal_initialize   (src/ptgs/DiffDriveCollisionGridBased.cpp:673-754)
// with its CCollisionGrid::updateCellInfo min-rule (:233-258): for every path k and every step n (all but the last),
// the robot polygon is placed at the path pose; every grid-cell CORNER (cell centre minus half a cell) inside the
// polygon marks that cell and its three lower-left neighbours with the step's TP-distance, keeping the minimum per
// (cell, k). The per-cell (k, d) lists come out in ascending k, which is the order the reference's k-outermost loop
// produces them in.
//
// Exactness: cos/sin of the (float) path headings are evaluated on the HOST with the C library — the same libm
// call the reference makes — and uploaded; everything on the device is IEEE double +,-,*,/ in the reference's
// operation order without FMA contraction (__dadd_rn/__dmul_rn/__ddiv_rn), so the table is bit-identical to the one
// the reference builds (tests/test_colgrid_build_gpu.py compares it with the checker's CSR entry by entry).
//
// Decomposition: one CTA per (path k, band of grid rows). The band's cells live in shared memory as float bits
// (200x200 cells = 160 KB: one band on B200's 227 KB); warps take the steps of the path round-robin, lanes take the
// cell corners of the step's bounding box, hits do shared-memory atomicMin (distances are >= 0, so the unsigned order
// of the bit patterns is the float order). The band is then written to a dense [K][cells] table; a second pass
// counts/compacts it per cell into the CSR (prefix sums by the three-launch scan of scan.cuh).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <thread>
#include <vector>

#include "mppb_internal.cuh"
#include "scan.cuh"

namespace mppb
{
namespace
{
constexpr int      kSweepThreads = 1024;
constexpr unsigned kInfBits      = 0x7f800000u;
constexpr size_t   kBandBytes    = 192 * 1024;  // shared-memory budget for one band of cells (+16 KB polygon scratch)

struct ColgridGeom
{
    double xmin, ymin, res, half;
    int    nx, ny;
    int    nverts;
    double sx[MPPB_MAX_SHAPE_VERTS], sy[MPPB_MAX_SHAPE_VERTS];
};

// CDynamicGrid::x2idx
__device__ __forceinline__ int x2idx(double v, double vmin, double res) { return __double2int_rz(__ddiv_rn(__dsub_rn(v, vmin), res)); }

__global__ void __launch_bounds__(kSweepThreads, 1)
    colgrid_sweep_kernel(ColgridGeom G, const uint32_t* __restrict__ pathBegin, const float* __restrict__ px,
                         const float* __restrict__ py, const double2* __restrict__ cs /* cos, sin per step */,
                         const float* __restrict__ pdist, int rowsPerBand, unsigned* __restrict__ dense /* [K][ny*nx] */)
{
    extern __shared__ unsigned cellMin[];  // [bandRows][nx]
    __shared__ double      svx[kSweepThreads / 32][MPPB_MAX_SHAPE_VERTS];  // the warp's transformed polygon
    __shared__ double      svy[kSweepThreads / 32][MPPB_MAX_SHAPE_VERTS];
    const int k     = blockIdx.x;
    const int row0  = blockIdx.y * rowsPerBand;  // the band owns cells of rows [row0, row1)
    const int row1  = min(row0 + rowsPerBand, G.ny);
    const int cells = (row1 - row0) * G.nx;
    for (int i = threadIdx.x; i < cells; i += kSweepThreads) cellMin[i] = kInfBits;
    __syncthreads();

    const uint32_t b = pathBegin[k], e = pathBegin[k + 1];
    const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = kSweepThreads / 32;
    const int      nv   = G.nverts;
    const double   shx = G.sx[lane], shy = G.sy[lane];  // lane m owns vertex m (nv <= 32)
    double* const  wvx = svx[warp];
    double* const  wvy = svy[warp];
    // all steps but the last one (":691  for (size_t n = 0; n < (nPoints - 1); n++)")
    for (uint32_t s = b + warp; s + 1 < e; s += nWarps)
    {
        const double  x = px[s], y = py[s];
        const double2 r = cs[s];
        // p.x + cos(p.phi) * vx - sin(p.phi) * vy ; p.y + sin(p.phi) * vx + cos(p.phi) * vy   (:705-710)
        const double tvx = __dsub_rn(__dadd_rn(x, __dmul_rn(r.x, shx)), __dmul_rn(r.y, shy));
        const double tvy = __dadd_rn(__dadd_rn(y, __dmul_rn(r.y, shx)), __dmul_rn(r.x, shy));
        double       bx0 = lane < nv ? tvx : DBL_MAX, bx1 = lane < nv ? tvx : -DBL_MAX;
        double       by0 = lane < nv ? tvy : DBL_MAX, by1 = lane < nv ? tvy : -DBL_MAX;
        __syncwarp();  // the previous step's readers are done with wvx/wvy
        if (lane < nv)
        {
            wvx[lane] = tvx;
            wvy[lane] = tvy;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1)
        {
            bx0 = fmin(bx0, __shfl_xor_sync(0xffffffffu, bx0, o));
            bx1 = fmax(bx1, __shfl_xor_sync(0xffffffffu, bx1, o));
            by0 = fmin(by0, __shfl_xor_sync(0xffffffffu, by0, o));
            by1 = fmax(by1, __shfl_xor_sync(0xffffffffu, by1, o));
        }
        __syncwarp();
        // range of cell corners that may lie inside this shape (:721-728); loops are "< max", as in the reference
        const int ix0 = max(0, x2idx(bx0, G.xmin, G.res) - 1), iy0 = max(0, x2idx(by0, G.ymin, G.res) - 1);
        const int ix1 = min(x2idx(bx1, G.xmin, G.res) + 1, G.nx - 1), iy1 = min(x2idx(by1, G.ymin, G.res) + 1, G.ny - 1);
        // corners of row iy touch cells of rows iy and iy-1: this band needs corners iy in [row0, row1]
        const int jy0 = max(iy0, row0), jy1 = min(iy1, row1 + 1);
        const int w = ix1 - ix0, h = jy1 - jy0;
        if (w <= 0 || h <= 0) continue;
        const unsigned dbits = __float_as_uint(pdist[s]);
        for (int t = lane; t < w * h; t += 32)
        {
            const int    ix = ix0 + t % w, iy = jy0 + t / w;
            const double cx = __dsub_rn(__dadd_rn(G.xmin, __dmul_rn(static_cast<double>(ix) + 0.5, G.res)), G.half);
            const double cy = __dsub_rn(__dadd_rn(G.ymin, __dmul_rn(static_cast<double>(iy) + 0.5, G.res)), G.half);
            // TPolygon2D::contains (PNPOLY): for (i = 0, j = n-1; i < n; j = i++), reference operation order
            bool   inside = false;
            double xj = wvx[nv - 1], yj = wvy[nv - 1];
            for (int i = 0; i < nv; i++)
            {
                const double xi = wvx[i], yi = wvy[i];
                if ((yi > cy) != (yj > cy))
                {
                    const double tt =
                        __dadd_rn(__ddiv_rn(__dmul_rn(__dsub_rn(xj, xi), __dsub_rn(cy, yi)), __dsub_rn(yj, yi)), xi);
                    if (cx < tt) inside = !inside;
                }
                xj = xi;
                yj = yi;
            }
            if (!inside) continue;
            // updateCellInfo(ix,iy), (ix-1,iy), (ix,iy-1), (ix-1,iy-1); out-of-grid cells are skipped (:237-238)
#pragma unroll
            for (int dy = 0; dy >= -1; dy--)
            {
                const int jy = iy + dy;
                if (jy < row0 || jy >= row1) continue;
#pragma unroll
                for (int dx = 0; dx >= -1; dx--)
                {
                    const int jx = ix + dx;
                    if (jx < 0) continue;
                    atomicMin(&cellMin[(jy - row0) * G.nx + jx], dbits);
                }
            }
        }
    }
    __syncthreads();
    unsigned* o = dense + static_cast<size_t>(k) * G.nx * G.ny + static_cast<size_t>(row0) * G.nx;
    for (int i = threadIdx.x; i < cells; i += kSweepThreads) o[i] = cellMin[i];
}

// per cell: number of paths that touch it
__global__ void colgrid_count_kernel(const unsigned* __restrict__ dense, int K, uint32_t nCells, uint32_t* __restrict__ counts)
{
    const uint32_t cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= nCells) return;
    uint32_t n = 0;
    for (int k = 0; k < K; k++) n += dense[static_cast<size_t>(k) * nCells + cell] != kInfBits;
    counts[cell] = n;
}

__global__ void colgrid_fill_kernel(const unsigned* __restrict__ dense, int K, uint32_t nCells,
                                    const uint32_t* __restrict__ offsets, uint16_t* __restrict__ entryK,
                                    float* __restrict__ entryD)
{
    const uint32_t cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= nCells) return;
    uint32_t pos = offsets[cell];
    for (int k = 0; k < K; k++)
    {
        const unsigned v = dense[static_cast<size_t>(k) * nCells + cell];
        if (v != kInfBits)
        {
            entryK[pos] = static_cast<uint16_t>(k);
            entryD[pos] = __uint_as_float(v);
            pos++;
        }
    }
}

int round_i(double v) { return static_cast<int>(std::lrint(v)); }  // mrpt::round
}  // namespace
}  // namespace mppb

using namespace mppb;

extern "C" int mppb_build_collision_grid(mppb_ctx* ctx, const mppb_colgrid_build_desc* d, mppb_colgrid* out)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    if (!d || !out) return fail(ctx, MPPB_ERR_INVALID_ARG, "null argument");
    if (d->num_paths <= 0 || d->num_paths > 65535 || !d->path_begin || !d->x || !d->y || !d->phi || !d->dist)
        return fail(ctx, MPPB_ERR_INVALID_ARG, "bad path tables");
    if (d->shape_nverts < 3 || d->shape_nverts > MPPB_MAX_SHAPE_VERTS || !d->shape_x || !d->shape_y)
        return fail(ctx, MPPB_ERR_INVALID_ARG, "robot shape needs 3..MPPB_MAX_SHAPE_VERTS vertices");
    if (!(d->ref_distance > 0) || !(d->resolution > 0)) return fail(ctx, MPPB_ERR_INVALID_ARG, "bad grid geometry");
    MPPB_CUDA(ctx, cudaSetDevice(ctx->device));

    // CDynamicGrid::setSize(-refDistance, refDistance, -refDistance, refDistance, resolution) (:657-658)
    ColgridGeom G;
    const double r  = d->resolution;
    G.xmin          = r * round_i(-d->ref_distance / r);
    G.ymin          = r * round_i(-d->ref_distance / r);
    const double xm = r * round_i(d->ref_distance / r), ym = r * round_i(d->ref_distance / r);
    G.res           = r;
    G.half          = r * 0.5;
    G.nx            = round_i((xm - G.xmin) / r);
    G.ny            = round_i((ym - G.ymin) / r);
    G.nverts        = d->shape_nverts;
    for (int i = 0; i < MPPB_MAX_SHAPE_VERTS; i++)
    {
        G.sx[i] = i < G.nverts ? d->shape_x[i] : 0.0;
        G.sy[i] = i < G.nverts ? d->shape_y[i] : 0.0;
    }
    if (G.nx <= 0 || G.ny <= 0) return fail(ctx, MPPB_ERR_INVALID_ARG, "empty collision grid");
    if (static_cast<size_t>(G.nx) * sizeof(unsigned) > kBandBytes)
        return fail(ctx, MPPB_ERR_UNSUPPORTED, "collision grid row too wide for a shared-memory band");
    const int      K      = d->num_paths;
    const uint32_t nSteps = d->path_begin[K];
    const size_t   nCells = static_cast<size_t>(G.nx) * G.ny;
    if (nCells * K > (size_t(1) << 31)) return fail(ctx, MPPB_ERR_UNSUPPORTED, "collision grid too large");

    // cos/sin of the float headings with the host C library (the reference's own libm call, :706-710)
    auto&      B  = ctx->colgrid;
    const auto t0 = std::chrono::steady_clock::now();
    MPPB_CUDA(ctx, B.hCs.reserve(static_cast<size_t>(nSteps) * sizeof(double2)));
    {
        double2*              hcs = B.hCs.as<double2>();
        const unsigned        nth = std::max(1u, std::min(std::thread::hardware_concurrency(), 16u));
        std::atomic<uint32_t> next{0};
        auto                  work = [&]() {
            for (;;)
            {
                const uint32_t lo = next.fetch_add(4096);
                if (lo >= nSteps) break;
                const uint32_t hi = std::min(nSteps, lo + 4096);
                for (uint32_t i = lo; i < hi; i++)
                {
                    const double ph = d->phi[i];
                    hcs[i].x        = std::cos(ph);
                    hcs[i].y        = std::sin(ph);
                }
            }
        };
        std::vector<std::thread> th;
        for (unsigned i = 1; i < nth; i++) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
    }
    const auto   t1 = std::chrono::steady_clock::now();
    cudaStream_t st = ctx->stream;
    MPPB_CUDA(ctx, B.dBegin.reserve((K + 1) * sizeof(uint32_t)));
    MPPB_CUDA(ctx, B.dX.reserve(nSteps * sizeof(float)));
    MPPB_CUDA(ctx, B.dY.reserve(nSteps * sizeof(float)));
    MPPB_CUDA(ctx, B.dDist.reserve(nSteps * sizeof(float)));
    MPPB_CUDA(ctx, B.dCs.reserve(static_cast<size_t>(nSteps) * sizeof(double2)));
    MPPB_CUDA(ctx, B.dDense.reserve(nCells * K * sizeof(unsigned)));
    MPPB_CUDA(ctx, B.dCounts.reserve(nCells * sizeof(uint32_t)));
    MPPB_CUDA(ctx, B.dOffsets.reserve((nCells + 1) * sizeof(uint32_t)));
    MPPB_CUDA(ctx, B.dScan.reserve((4096 + 1) * sizeof(uint32_t)));
    MPPB_CUDA(ctx, cudaMemcpyAsync(B.dBegin.p, d->path_begin, (K + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    MPPB_CUDA(ctx, cudaMemcpyAsync(B.dX.p, d->x, nSteps * sizeof(float), cudaMemcpyHostToDevice, st));
    MPPB_CUDA(ctx, cudaMemcpyAsync(B.dY.p, d->y, nSteps * sizeof(float), cudaMemcpyHostToDevice, st));
    MPPB_CUDA(ctx, cudaMemcpyAsync(B.dDist.p, d->dist, nSteps * sizeof(float), cudaMemcpyHostToDevice, st));
    MPPB_CUDA(ctx, cudaMemcpyAsync(B.dCs.p, B.hCs.p, static_cast<size_t>(nSteps) * sizeof(double2), cudaMemcpyHostToDevice, st));

    MPPB_CUDA(ctx, cudaEventRecord(ctx->evBegin, st));
    const int    rowsPerBand = static_cast<int>(std::min<size_t>(G.ny, kBandBytes / (G.nx * sizeof(unsigned))));
    const int    nBands      = (G.ny + rowsPerBand - 1) / rowsPerBand;
    const size_t smem        = static_cast<size_t>(rowsPerBand) * G.nx * sizeof(unsigned);
    MPPB_CUDA(ctx, cudaFuncSetAttribute(colgrid_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    colgrid_sweep_kernel<<<dim3(K, nBands), kSweepThreads, smem, st>>>(
        G, B.dBegin.as<uint32_t>(), B.dX.as<float>(), B.dY.as<float>(), B.dCs.as<double2>(), B.dDist.as<float>(),
        rowsPerBand, B.dDense.as<unsigned>());
    const unsigned cblocks = static_cast<unsigned>((nCells + 255) / 256);
    colgrid_count_kernel<<<cblocks, 256, 0, st>>>(B.dDense.as<unsigned>(), K, static_cast<uint32_t>(nCells), B.dCounts.as<uint32_t>());
    if (device_exclusive_scan(B.dCounts.as<uint32_t>(), static_cast<uint32_t>(nCells), B.dOffsets.as<uint32_t>(),
                              B.dScan.as<uint32_t>(), nullptr, st) < 0)
        return fail(ctx, MPPB_ERR_UNSUPPORTED, "collision grid too large");
    ctx->launches += 5;
    MPPB_CUDA(ctx, cudaGetLastError());
    MPPB_CUDA(ctx, cudaEventRecord(ctx->evEnd, st));
    MPPB_CUDA(ctx, B.hOffsets.reserve((nCells + 1) * sizeof(uint32_t)));
    MPPB_CUDA(ctx, cudaMemcpyAsync(B.hOffsets.p, B.dOffsets.p, (nCells + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    MPPB_CUDA(ctx, cudaStreamSynchronize(st));
    const uint32_t nEntries = B.hOffsets.as<uint32_t>()[nCells];
    MPPB_CUDA(ctx, B.dEntryK.reserve(std::max<size_t>(1, nEntries) * sizeof(uint16_t)));
    MPPB_CUDA(ctx, B.dEntryD.reserve(std::max<size_t>(1, nEntries) * sizeof(float)));
    MPPB_CUDA(ctx, B.hEntryK.reserve(std::max<size_t>(1, nEntries) * sizeof(uint16_t)));
    MPPB_CUDA(ctx, B.hEntryD.reserve(std::max<size_t>(1, nEntries) * sizeof(float)));
    colgrid_fill_kernel<<<cblocks, 256, 0, st>>>(B.dDense.as<unsigned>(), K, static_cast<uint32_t>(nCells), B.dOffsets.as<uint32_t>(),
                                                  B.dEntryK.as<uint16_t>(), B.dEntryD.as<float>());
    ctx->launches++;
    MPPB_CUDA(ctx, cudaGetLastError());
    if (nEntries)
    {
        MPPB_CUDA(ctx, cudaMemcpyAsync(B.hEntryK.p, B.dEntryK.p, nEntries * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
        MPPB_CUDA(ctx, cudaMemcpyAsync(B.hEntryD.p, B.dEntryD.p, nEntries * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    MPPB_CUDA(ctx, cudaStreamSynchronize(st));
    float sweepMs = 0.f;
    cudaEventElapsedTime(&sweepMs, ctx->evBegin, ctx->evEnd);
    B.stats[0] = std::chrono::duration<double>(t1 - t0).count();                                // host cos/sin
    B.stats[1] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();  // uploads, kernels, result
    B.stats[2] = 1e-3 * sweepMs;                                                                // sweep+count+scan kernels
    B.stats[3] = static_cast<double>(nSteps);
    out->x_min        = G.xmin;
    out->y_min        = G.ymin;
    out->res          = G.res;
    out->nx           = G.nx;
    out->ny           = G.ny;
    out->n_entries    = nEntries;
    out->cell_offsets = B.hOffsets.as<uint32_t>();
    out->entry_k      = B.hEntryK.as<uint16_t>();
    out->entry_d      = B.hEntryD.as<float>();
    return MPPB_OK;
}

extern "C" int mppb_colgrid_build_stats(const mppb_ctx* ctx, double* out4)
{
    if (!ctx || !out4) return MPPB_ERR_INVALID_ARG;
    for (int i = 0; i < 4; i++) out4[i] = ctx->colgrid.stats[i];
    return MPPB_OK;
}
