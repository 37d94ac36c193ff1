// C ABI of libmpp_b200 (see include/mppb.h). Host-side plumbing only: context, uploads,
// staging and launches. There is deliberately no CPU fallback anywhere in this file.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <sched.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "mppb_internal.cuh"

namespace mppb
{
static std::mutex  g_errMtx;
static std::string g_globalError;

void set_global_error(const std::string& msg)
{
    std::lock_guard<std::mutex> lk(g_errMtx);
    g_globalError = msg;
}
int fail(mppb_ctx* ctx, int code, const std::string& msg)
{
    if (ctx)
        ctx->lastError = msg;
    else
        set_global_error(msg);
    return code;
}
int cuda_fail(mppb_ctx* ctx, cudaError_t e, const char* what)
{
    return fail(ctx, MPPB_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what);
}
}  // namespace mppb

using namespace mppb;

// Device-visible alias of a PINNED host buffer (mppb_host_alloc / cudaHostRegister): under UVA a kernel can read and
// write it directly over PCIe. Small latency-bound batches (one A* expansion) use this instead of staging copies:
// it removes two to four cudaMemcpyAsync calls and their DMA set-up per batch. Returns nullptr for pageable memory.
static const bool g_noZeroCopy = std::getenv("MPPB_NO_ZEROCOPY") != nullptr;  // A/B measurements only
template <class T>
static T* mapped_alias(const T* hostPtr)
{
    if (g_noZeroCopy || !hostPtr) return nullptr;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, hostPtr) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
        return static_cast<T*>(attr.devicePointer);
    cudaGetLastError();  // pageable memory may be reported as an error by older drivers: not ours
    return nullptr;
}
// inputs of at most this many bytes are read in place by the kernels; larger ones are staged into HBM by DMA
static constexpr size_t kZeroCopyInputBytes = 256 * 1024;

extern "C" {

const char* mppb_version(void) { return "mpp-b200 0.1 (sm_100a)"; }

const char* mppb_last_global_error(void)
{
    std::lock_guard<std::mutex> lk(g_errMtx);
    static thread_local std::string copy;
    copy = g_globalError;
    return copy.c_str();
}

int mppb_ctx_create(int device, void* cuda_stream, mppb_ctx** out_ctx)
{
    if (!out_ctx) return fail(nullptr, MPPB_ERR_INVALID_ARG, "out_ctx is null");
    *out_ctx  = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, MPPB_ERR_CUDA,
                    std::string("no usable CUDA device (libmpp_b200 has no CPU fallback): ") +
                        (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    if (device < 0 || device >= count) return fail(nullptr, MPPB_ERR_INVALID_ARG, "device ordinal out of range");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    auto* ctx   = new mppb_ctx();
    ctx->device = device;
    cudaDeviceGetAttribute(&ctx->numSMs, cudaDevAttrMultiProcessorCount, device);
    if (cuda_stream)
    {
        ctx->stream    = static_cast<cudaStream_t>(cuda_stream);
        ctx->ownStream = false;
    }
    else
    {
        e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess)
        {
            delete ctx;
            return cuda_fail(nullptr, e, "cudaStreamCreateWithFlags");
        }
        ctx->ownStream = true;
    }
    cudaEventCreate(&ctx->evBegin);
    cudaEventCreate(&ctx->evEnd);
    *out_ctx = ctx;
    return MPPB_OK;
}

static void sincos_helper_destroy(mppb_ctx* ctx);  // (helper thread of the host-buffer entry point, below)
void mppb_ctx_destroy(mppb_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->cloud.release();
    ctx->auxCloud.release();
    ctx->bboxScratch.release();
    for (auto& g : ctx->groups) g.release();
    ctx->groupDescs.release();
    sincos_helper_destroy(ctx);
    for (auto& b : ctx->nearList) b.release();
    ctx->nearCounters.release();
    if (ctx->nearLastCount) cudaFreeHost(ctx->nearLastCount);
    ctx->holoDescs.release();
    for (int i = 0; i < MPPB_MAX_EVALUATORS; i++)
    {
        ctx->evalCells[i].release();
        ctx->evalWps[i].release();
    }
    ctx->evalDescs.release();
    ctx->colgrid.release();
    multigpu_release(ctx);
    for (auto& t : ctx->pathTables) t.release();
    ctx->expandDev.release();
    for (mppb::DevBuf* b : {&ctx->dExpandIn, &ctx->dExpandRows, &ctx->dExpandOut, &ctx->dExpandCounts, &ctx->dExpandNodes, &ctx->dExpandBoxes})
        b->release();
    ctx->hExpandNodes.release();
    ctx->hExpandBoxes.release();
    ctx->hNodes.release();
    ctx->hOut.release();
    ctx->hStage.release();
    ctx->hNodesHolo.release();
    ctx->dNodesHolo.release();
    ctx->dBoxesHolo.release();
    ctx->dCandBegin.release();
    ctx->dNodes.release();
    ctx->dOut.release();
    ctx->dStage.release();
    ctx->dStage2.release();
    ctx->dStage3.release();
    ctx->dBoxes.release();
    if (ctx->auxStream) cudaStreamDestroy(ctx->auxStream);
    if (ctx->evFork) cudaEventDestroy(ctx->evFork);
    if (ctx->evJoin) cudaEventDestroy(ctx->evJoin);
    for (int i = 0; i < 2; i++)
    {
        if (ctx->auxMore[i]) cudaStreamDestroy(ctx->auxMore[i]);
        if (ctx->evJoinMore[i]) cudaEventDestroy(ctx->evJoinMore[i]);
    }
    if (ctx->evBegin) cudaEventDestroy(ctx->evBegin);
    if (ctx->evEnd) cudaEventDestroy(ctx->evEnd);
    if (ctx->ownStream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* mppb_last_error(const mppb_ctx* ctx) { return ctx ? ctx->lastError.c_str() : mppb_last_global_error(); }

int mppb_sync(mppb_ctx* ctx)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    MPPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->peer.world > 1)
    {
        // fused multi-GPU merge: a peer that never raised its flag ends the wait kernel with an error word, not a hang
        const size_t areaFloats = 2 * static_cast<size_t>(ctx->peer.world) * ctx->peerMaxNodes * ctx->peerCols;
        int          status     = 0;
        MPPB_CUDA(ctx, cudaMemcpy(&status, reinterpret_cast<const uint32_t*>(ctx->peerLocal.as<float>() + areaFloats) + MPPB_MAX_RANKS,
                                  sizeof(int), cudaMemcpyDeviceToHost));
        if (status != 0) return fail(ctx, MPPB_ERR_STATE, "multi-GPU merge timed out waiting for a peer's rows");
    }
    return MPPB_OK;
}
int mppb_set_holo_filter(mppb_ctx* ctx, int on)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    ctx->holoFilter = on == 0 ? 0 : (on == 2 ? 2 : 1);
    return MPPB_OK;
}
// the adaptive switch of the near-first evaluation starts afresh (new cloud, new tables, new mode): what earlier calls
// reported goes into the statistics
static void near_reset_adaptive(mppb_ctx* ctx)
{
    ctx->nearSkip = 0;
    for (int i = 0; i < mppb_ctx::kNearSlots; i++)
        if (ctx->nearNodesOfSlot[i])
        {
            ctx->nearPendingSeen += ctx->nearLastCount[i];
            ctx->nearNodesSeen += ctx->nearNodesOfSlot[i];
            ctx->nearNodesOfSlot[i] = 0;
            ctx->nearLastCount[i]   = 0;
        }
}
int mppb_set_grid_near(mppb_ctx* ctx, int mode)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    ctx->gridNear = mode == 0 ? 0 : (mode == 2 ? 2 : 1);
    near_reset_adaptive(ctx);
    return MPPB_OK;
}
int mppb_grid_near_stats(const mppb_ctx* ctx, uint64_t* out3)
{
    if (!ctx || !out3) return MPPB_ERR_INVALID_ARG;
    // the counts of launches still in flight are not in yet; the caller synchronises first if it wants them exact
    uint64_t pend = ctx->nearPendingSeen, nodes = ctx->nearNodesSeen;
    if (ctx->nearLastCount)
        for (int i = 0; i < mppb_ctx::kNearSlots; i++)
            if (ctx->nearNodesOfSlot[i])
            {
                pend += ctx->nearLastCount[i];
                nodes += ctx->nearNodesOfSlot[i];
            }
    out3[0] = ctx->nearLaunches;
    out3[1] = nodes;
    out3[2] = pend;
    return MPPB_OK;
}
int mppb_grid_near_window(const mppb_ctx* ctx, double clip_dist, double* half_size, double* bound)
{
    if (!ctx || !half_size || !bound) return MPPB_ERR_INVALID_ARG;
    double r = 0;
    float  b = 0;
    const bool ok = near_window(ctx, clip_dist, r, b);
    *half_size    = ok ? r : 0.0;
    *bound        = ok ? static_cast<double>(b) : 0.0;
    return MPPB_OK;
}
int mppb_set_holo_by_node(mppb_ctx* ctx, int on)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    ctx->holoByNode = on != 0;
    return MPPB_OK;
}
void*    mppb_stream(mppb_ctx* ctx) { return ctx ? static_cast<void*>(ctx->stream) : nullptr; }
int      mppb_ctx_device(const mppb_ctx* ctx) { return ctx ? ctx->device : -1; }
uint64_t mppb_launch_count(const mppb_ctx* ctx) { return ctx ? ctx->launches : 0; }

int mppb_host_alloc(size_t bytes, void** out_ptr)
{
    if (!out_ptr) return MPPB_ERR_INVALID_ARG;
    cudaError_t e = cudaHostAlloc(out_ptr, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaHostAlloc");
    return MPPB_OK;
}
int mppb_host_free(void* ptr)
{
    if (!ptr) return MPPB_OK;
    cudaError_t e = cudaFreeHost(ptr);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaFreeHost");
    return MPPB_OK;
}

int mppb_timer_begin(mppb_ctx* ctx)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    MPPB_CUDA(ctx, cudaEventRecord(ctx->evBegin, ctx->stream));
    return MPPB_OK;
}
int mppb_timer_end(mppb_ctx* ctx, float* out_ms)
{
    if (!ctx || !out_ms) return MPPB_ERR_INVALID_ARG;
    MPPB_CUDA(ctx, cudaEventRecord(ctx->evEnd, ctx->stream));
    MPPB_CUDA(ctx, cudaEventSynchronize(ctx->evEnd));
    MPPB_CUDA(ctx, cudaEventElapsedTime(out_ms, ctx->evBegin, ctx->evEnd));
    return MPPB_OK;
}

// ---- obstacle cloud -------------------------------------------------------------------
static int set_cloud(mppb_ctx* ctx, CloudStore& cs, const float* xs, const float* ys, size_t n, double bin_size,
                     int append, const float* box, cudaMemcpyKind kind)
{
    if (n > 0 && (!xs || !ys)) return fail(ctx, MPPB_ERR_INVALID_ARG, "null cloud pointers");
    MPPB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (&cs == &ctx->cloud)
    {
        ctx->cloudEpoch++;  // whatever happens below, the resident cloud is no longer the old one
        near_reset_adaptive(ctx);
    }
    const size_t keep  = append ? cs.rawN : 0;
    const size_t total = keep + n;
    if (!append)
    {
        cs.hasBox = box != nullptr;
        if (box) std::memcpy(cs.box, box, 4 * sizeof(float));
    }
    else if (box)
    {
        if (!cs.hasBox || std::memcmp(cs.box, box, 4 * sizeof(float)) != 0)
            return fail(ctx, MPPB_ERR_INVALID_ARG, "appended obstacle sources must use the same clipping box");
    }
    if (total * sizeof(float) > cs.rawX.bytes)
    {
        // grow, preserving the already-uploaded sources
        DevBuf nx, ny;
        MPPB_CUDA(ctx, nx.reserve(std::max<size_t>(total, 1) * sizeof(float)));
        MPPB_CUDA(ctx, ny.reserve(std::max<size_t>(total, 1) * sizeof(float)));
        if (keep)
        {
            MPPB_CUDA(ctx, cudaMemcpyAsync(nx.p, cs.rawX.p, keep * sizeof(float), cudaMemcpyDeviceToDevice,
                                           ctx->stream));
            MPPB_CUDA(ctx, cudaMemcpyAsync(ny.p, cs.rawY.p, keep * sizeof(float), cudaMemcpyDeviceToDevice,
                                           ctx->stream));
            MPPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
        cs.rawX.release();
        cs.rawY.release();
        cs.rawX = nx;
        cs.rawY = ny;
    }
    if (n)
    {
        MPPB_CUDA(ctx, cudaMemcpyAsync(cs.rawX.as<float>() + keep, xs, n * sizeof(float), kind, ctx->stream));
        MPPB_CUDA(ctx, cudaMemcpyAsync(cs.rawY.as<float>() + keep, ys, n * sizeof(float), kind, ctx->stream));
    }
    cs.rawN = total;
    return build_cloud_bins(ctx, cs, bin_size > 0 ? bin_size : (cs.binSize > 0 && append ? cs.binSize : 0.5));
}

int mppb_set_obstacles(mppb_ctx* ctx, const float* xs, const float* ys, size_t n, double bin_size, int append)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    return set_cloud(ctx, ctx->cloud, xs, ys, n, bin_size, append, nullptr, cudaMemcpyHostToDevice);
}
int mppb_set_obstacles_device(mppb_ctx* ctx, const float* d_xs, const float* d_ys, size_t n, double bin_size,
                              int append)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    return set_cloud(ctx, ctx->cloud, d_xs, d_ys, n, bin_size, append, nullptr, cudaMemcpyDeviceToDevice);
}
int mppb_set_obstacles_clipped(mppb_ctx* ctx, const float* xs, const float* ys, size_t n, double bin_size,
                               int append, const float* box)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    return set_cloud(ctx, ctx->cloud, xs, ys, n, bin_size, append, box, cudaMemcpyHostToDevice);
}
size_t   mppb_num_obstacles(const mppb_ctx* ctx) { return ctx ? ctx->bins.n : 0; }
uint64_t mppb_obstacles_epoch(const mppb_ctx* ctx) { return ctx ? ctx->cloudEpoch : 0; }
uint64_t mppb_ptgs_epoch(const mppb_ctx* ctx) { return ctx ? ctx->ptgEpoch : 0; }

// ---- PTG tables -------------------------------------------------------------------------
static void release_groups(mppb_ctx* ctx)
{
    for (auto& g : ctx->groups) g.release();
    ctx->groups.clear();
}

int mppb_clear_ptgs(mppb_ctx* ctx)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    MPPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->ptgEpoch++;
    expand_invalidate(ctx);
    for (auto& t : ctx->pathTables) t.release();
    release_groups(ctx);
    ctx->grids.clear();
    ctx->holos.clear();
    ctx->ptgKinds.clear();
    ctx->totalPaths = 0;
    return MPPB_OK;
}

// Merge PTGs with identical grid geometry and robot shape into one CSR per group and upload.
static int rebuild_groups(mppb_ctx* ctx)
{
    MPPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    release_groups(ctx);
    std::vector<std::vector<size_t>> members;
    for (size_t i = 0; i < ctx->grids.size(); i++)
    {
        bool placed = false;
        for (auto& m : members)
            if (ctx->grids[m[0]].sameGeometryAndShape(ctx->grids[i]))
            {
                m.push_back(i);
                placed = true;
                break;
            }
        if (!placed) members.push_back({i});
    }
    std::vector<GridGroupDev> descs;
    near_reset_adaptive(ctx);
    ctx->nearSlackDiag = 0;
    bool nearApplies   = !members.empty();
    for (const auto& m : members)
    {
        const PtgGridHost& g0     = ctx->grids[m[0]];
        const size_t       ncells = static_cast<size_t>(g0.nx) * g0.ny;
        std::vector<uint32_t> off(ncells + 1, 0);
        for (size_t c = 0; c < ncells; c++)
        {
            uint32_t n = 0;
            for (size_t i : m) n += ctx->grids[i].offsets[c + 1] - ctx->grids[i].offsets[c];
            off[c + 1] = off[c] + n;
        }
        std::vector<uint2> ent(off[ncells]);
        for (size_t c = 0; c < ncells; c++)
        {
            uint32_t pos = off[c];
            for (size_t i : m)
            {
                const PtgGridHost& g = ctx->grids[i];
                for (uint32_t e = g.offsets[c]; e < g.offsets[c + 1]; e++)
                {
                    uint32_t bits;
                    std::memcpy(&bits, &g.entryD[e], 4);
                    ent[pos++] = make_uint2(static_cast<uint32_t>(g.outOffset) + g.entryK[e], bits);
                }
            }
        }
        // near-first evaluation (NearPassDev): slack = max over the entries of dist(origin, cell) - d, measured here
        {
            double slack = 0;
            for (size_t c = 0; c < ncells; c++)
            {
                if (off[c + 1] == off[c]) continue;
                const double xa = g0.xmin + static_cast<double>(c % g0.nx) * g0.res, ya = g0.ymin + static_cast<double>(c / g0.nx) * g0.res;
                const double nx = std::min(std::max(0.0, xa), xa + g0.res), ny = std::min(std::max(0.0, ya), ya + g0.res);
                const double dc = std::hypot(nx, ny);
                for (size_t i : m)
                {
                    const PtgGridHost& g = ctx->grids[i];
                    for (uint32_t e = g.offsets[c]; e < g.offsets[c + 1]; e++)
                        slack = std::max(slack, dc - static_cast<double>(g.entryD[e]));
                }
            }
            ctx->nearSlackDiag = std::max(ctx->nearSlackDiag, slack + g0.res * 1.4142136);
        }
        GridGroupHost G;
        MPPB_CUDA(ctx, G.offsets.reserve((ncells + 1) * sizeof(uint32_t)));
        MPPB_CUDA(ctx, G.entries.reserve(std::max<size_t>(ent.size(), 1) * sizeof(uint2)));
        MPPB_CUDA(ctx, cudaMemcpyAsync(G.offsets.p, off.data(), (ncells + 1) * sizeof(uint32_t),
                                       cudaMemcpyHostToDevice, ctx->stream));
        if (!ent.empty())
            MPPB_CUDA(ctx, cudaMemcpyAsync(G.entries.p, ent.data(), ent.size() * sizeof(uint2),
                                           cudaMemcpyHostToDevice, ctx->stream));
        // by-column view (GridGroupDev): per output column the cells in ascending-d order, cut greedily into
        // chunks that touch at most 4 words of the kernel's tiled occupancy bitmap
        std::vector<uint32_t> colStart(1, 0), colDBase;
        std::vector<uint4>    colWord, colMask;
        std::vector<float>    colD;
        std::vector<int>      colOut;
        const uint32_t        tilesX = static_cast<uint32_t>(grid_tiles_x(g0.nx));
        for (size_t i : m)
        {
            const PtgGridHost&                            g = ctx->grids[i];
            std::vector<std::vector<std::pair<float, uint32_t>>> perK(g.K);
            for (size_t c = 0; c < ncells; c++)
                for (uint32_t e = g.offsets[c]; e < g.offsets[c + 1]; e++)
                    perK[g.entryK[e]].emplace_back(g.entryD[e], static_cast<uint32_t>(c));
            for (int k = 0; k < g.K; k++)
            {
                if (perK[k].empty()) nearApplies = false;  // a path without cells can never be decided by near points
                std::sort(perK[k].begin(), perK[k].end());
                uint32_t w[4] = {0, 0, 0, 0}, mk[4] = {0, 0, 0, 0};
                float    dOf[4][32];
                int      used = 0, cells = 0, chunkOfColumn = 0;
                // The first chunk of a column is kept short (16 cells): on a dense cloud the scan ends in it, and the
                // kernel looks up the d of every occupied cell of the chunk that ends the scan (2^23-point cloud:
                // 0.427 -> 0.400 ms per 4096 nodes; C3 unchanged; caps of 4/8/16 or 8/16 cost C3 1.5-3 %).
                auto cellCap = [&]() { return chunkOfColumn == 0 ? 16 : 1 << 30; };
                auto flush   = [&]() {
                    if (!used) return;
                    chunkOfColumn++;
                    cells = 0;
                    colWord.push_back(make_uint4(w[0], w[1], w[2], w[3]));
                    colMask.push_back(make_uint4(mk[0], mk[1], mk[2], mk[3]));
                    colDBase.push_back(static_cast<uint32_t>(colD.size()));
                    for (int sl = 0; sl < used; sl++)
                        for (int bit = 0; bit < 32; bit++)
                            if (mk[sl] >> bit & 1u) colD.push_back(dOf[sl][bit]);
                    used = 0;
                    for (int sl = 0; sl < 4; sl++) w[sl] = mk[sl] = 0;
                };
                for (const auto& dc : perK[k])
                {
                    const uint32_t cx = dc.second % static_cast<uint32_t>(g0.nx), cy = dc.second / static_cast<uint32_t>(g0.nx);
                    const uint32_t word = grid_tile_word(cx, cy, tilesX), bit = grid_tile_bit(cx, cy);
                    int            sl = 0;
                    while (sl < used && w[sl] != word) sl++;
                    if (cells == cellCap() || (sl == used && used == 4))
                    {
                        flush();
                        sl = 0;
                    }
                    if (sl == used)
                    {
                        w[sl] = word;
                        used  = sl + 1;
                    }
                    mk[sl] |= 1u << bit;
                    dOf[sl][bit] = dc.first;
                    cells++;
                }
                flush();
                colStart.push_back(static_cast<uint32_t>(colWord.size()));
                colOut.push_back(g.outOffset + k);
            }
        }
        MPPB_CUDA(ctx, G.colStart.reserve(colStart.size() * sizeof(uint32_t)));
        MPPB_CUDA(ctx, G.colWord.reserve(std::max<size_t>(colWord.size(), 1) * sizeof(uint4)));
        MPPB_CUDA(ctx, G.colMask.reserve(std::max<size_t>(colMask.size(), 1) * sizeof(uint4)));
        MPPB_CUDA(ctx, G.colDBase.reserve(std::max<size_t>(colDBase.size(), 1) * sizeof(uint32_t)));
        MPPB_CUDA(ctx, G.colD.reserve(std::max<size_t>(colD.size(), 1) * sizeof(float)));
        MPPB_CUDA(ctx, G.colOut.reserve(colOut.size() * sizeof(int)));
        MPPB_CUDA(ctx, cudaMemcpyAsync(G.colStart.p, colStart.data(), colStart.size() * sizeof(uint32_t),
                                       cudaMemcpyHostToDevice, ctx->stream));
        if (!colWord.empty())
        {
            MPPB_CUDA(ctx, cudaMemcpyAsync(G.colWord.p, colWord.data(), colWord.size() * sizeof(uint4),
                                           cudaMemcpyHostToDevice, ctx->stream));
            MPPB_CUDA(ctx, cudaMemcpyAsync(G.colMask.p, colMask.data(), colMask.size() * sizeof(uint4),
                                           cudaMemcpyHostToDevice, ctx->stream));
            MPPB_CUDA(ctx, cudaMemcpyAsync(G.colDBase.p, colDBase.data(), colDBase.size() * sizeof(uint32_t),
                                           cudaMemcpyHostToDevice, ctx->stream));
            MPPB_CUDA(ctx, cudaMemcpyAsync(G.colD.p, colD.data(), colD.size() * sizeof(float),
                                           cudaMemcpyHostToDevice, ctx->stream));
        }
        MPPB_CUDA(ctx, cudaMemcpyAsync(G.colOut.p, colOut.data(), colOut.size() * sizeof(int),
                                       cudaMemcpyHostToDevice, ctx->stream));
        MPPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        GridGroupDev& v = G.dev;
        std::memset(&v, 0, sizeof(v));
        v.nCols    = static_cast<int>(colOut.size());
        v.colStart = G.colStart.as<uint32_t>();
        v.colWord  = G.colWord.as<uint4>();
        v.colMask  = G.colMask.as<uint4>();
        v.colDBase = G.colDBase.as<uint32_t>();
        v.colD     = G.colD.as<float>();
        v.colOut   = G.colOut.as<int>();
        v.xmin      = g0.xmin;
        v.ymin      = g0.ymin;
        v.res       = g0.res;
        v.invRes    = 1.0 / g0.res;
        v.nx        = g0.nx;
        v.ny        = g0.ny;
        v.offsets   = G.offsets.as<uint32_t>();
        v.entries   = G.entries.as<uint2>();
        v.nverts    = static_cast<int>(g0.shapeX.size());
        v.maxRadius = g0.maxRadius;
        // float upper bound of R^2 with a relative guard band far above any rounding of fx*fx+fy*fy
        v.insideGuardR2 = static_cast<float>(g0.maxRadius * g0.maxRadius * 1.0001 + 1e-12);
        for (int i = 0; i < v.nverts; i++)
        {
            v.shapeX[i] = g0.shapeX[i];
            v.shapeY[i] = g0.shapeY[i];
        }
        ctx->groups.push_back(G);
        descs.push_back(v);
    }
    if (!nearApplies) ctx->nearSlackDiag = 0;
    MPPB_CUDA(ctx, ctx->groupDescs.reserve(std::max<size_t>(descs.size(), 1) * sizeof(GridGroupDev)));
    MPPB_CUDA(ctx, cudaMemcpyAsync(ctx->groupDescs.p, descs.data(), descs.size() * sizeof(GridGroupDev),
                                   cudaMemcpyHostToDevice, ctx->stream));
    MPPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MPPB_OK;
}

int mppb_set_ptg_grid(mppb_ctx* ctx, int ptg_index, const mppb_ptg_grid_desc* d)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    if (!d) return fail(ctx, MPPB_ERR_INVALID_ARG, "null PTG descriptor");
    if (ptg_index != static_cast<int>(ctx->ptgKinds.size()))
        return fail(ctx, MPPB_ERR_INVALID_ARG, "ptg_index must be the next free index");
    if (ptg_index >= MPPB_MAX_PTGS) return fail(ctx, MPPB_ERR_UNSUPPORTED, "too many PTGs");
    if (d->num_paths <= 0 || d->num_paths > 65535) return fail(ctx, MPPB_ERR_INVALID_ARG, "bad num_paths");
    if (d->grid_nx <= 0 || d->grid_ny <= 0 || !(d->grid_res > 0) || !d->cell_offsets)
        return fail(ctx, MPPB_ERR_INVALID_ARG, "bad collision grid geometry");
    if (d->shape_nverts < 3 || d->shape_nverts > MPPB_MAX_SHAPE_VERTS || !d->shape_x || !d->shape_y)
        return fail(ctx, MPPB_ERR_INVALID_ARG, "robot shape must have 3..32 vertices");
    const size_t   ncells = static_cast<size_t>(d->grid_nx) * d->grid_ny;
    const uint32_t nent   = d->cell_offsets[ncells];
    if (nent && (!d->entry_k || !d->entry_d)) return fail(ctx, MPPB_ERR_INVALID_ARG, "null entry arrays");
    for (size_t c = 0; c < ncells; c++)
        if (d->cell_offsets[c] > d->cell_offsets[c + 1]) return fail(ctx, MPPB_ERR_INVALID_ARG, "offsets not sorted");
    for (uint32_t i = 0; i < nent; i++)
    {
        if (d->entry_k[i] >= d->num_paths) return fail(ctx, MPPB_ERR_INVALID_ARG, "entry_k out of range");
        if (!(d->entry_d[i] >= 0.f)) return fail(ctx, MPPB_ERR_INVALID_ARG, "entry_d must be >= 0");
    }
    MPPB_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->ptgEpoch++;
    expand_invalidate(ctx);
    PtgGridHost g;
    g.K         = d->num_paths;
    g.outOffset = ctx->totalPaths;
    g.xmin      = d->grid_x_min;
    g.ymin      = d->grid_y_min;
    g.res       = d->grid_res;
    g.nx        = d->grid_nx;
    g.ny        = d->grid_ny;
    g.maxRadius = d->max_radius;
    g.offsets.assign(d->cell_offsets, d->cell_offsets + ncells + 1);
    g.entryK.assign(d->entry_k, d->entry_k + nent);
    g.entryD.assign(d->entry_d, d->entry_d + nent);
    g.shapeX.assign(d->shape_x, d->shape_x + d->shape_nverts);
    g.shapeY.assign(d->shape_y, d->shape_y + d->shape_nverts);
    ctx->grids.push_back(std::move(g));
    ctx->ptgKinds.push_back(0);
    ctx->holos.push_back(HoloPtgDev{0, 0});
    ctx->totalPaths += d->num_paths;
    return rebuild_groups(ctx);
}
int mppb_num_ptgs(const mppb_ctx* ctx) { return ctx ? static_cast<int>(ctx->ptgKinds.size()) : 0; }
int mppb_total_paths(const mppb_ctx* ctx) { return ctx ? ctx->totalPaths : 0; }
int mppb_ptg_column_offset(const mppb_ctx* ctx, int ptg_index)
{
    if (!ctx || ptg_index < 0 || ptg_index >= static_cast<int>(ctx->ptgKinds.size())) return -1;
    if (ctx->ptgKinds[ptg_index] != 0) return -1;
    int g = 0;
    for (int i = 0; i < ptg_index; i++) g += ctx->ptgKinds[i] == 0;
    return ctx->grids[g].outOffset;
}

int mppb_set_ptg_holo(mppb_ctx* ctx, int ptg_index, const mppb_ptg_holo_desc* d)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    if (!d) return fail(ctx, MPPB_ERR_INVALID_ARG, "null PTG descriptor");
    if (ptg_index != static_cast<int>(ctx->ptgKinds.size()))
        return fail(ctx, MPPB_ERR_INVALID_ARG, "ptg_index must be the next free index");
    if (ptg_index >= MPPB_MAX_PTGS) return fail(ctx, MPPB_ERR_UNSUPPORTED, "too many PTGs");
    if (!(d->robot_radius > 0) || !(d->v_max > 0) || d->num_paths <= 0)
        return fail(ctx, MPPB_ERR_INVALID_ARG, "bad HolonomicBlend constants");
    MPPB_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->ptgEpoch++;
    expand_invalidate(ctx);
    ctx->ptgKinds.push_back(1);
    ctx->holos.push_back(HoloPtgDev{d->robot_radius, d->v_max});
    HoloPtgDev table[MPPB_MAX_PTGS] = {};
    for (size_t i = 0; i < ctx->holos.size(); i++) table[i] = ctx->holos[i];
    MPPB_CUDA(ctx, ctx->holoDescs.reserve(sizeof(table)));
    MPPB_CUDA(ctx, cudaMemcpyAsync(ctx->holoDescs.p, table, sizeof(table), cudaMemcpyHostToDevice, ctx->stream));
    MPPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MPPB_OK;
}

// ---- evaluation ---------------------------------------------------------------------------
void mppb_nodes_xycs_from_poses(size_t n, const double* poses, double* out);
namespace
{
// The host-buffer entry point needs cos/sin of every node heading from the host C library (they must be the
// reference's values, bit for bit): 15 ns per node on one thread, 62 us for 4096 nodes — as long as the GPU takes for
// the whole batch. A helper thread takes the batch from its second sub-batch on while the calling thread does the
// first and issues the copies and launches. The helper spins for 2 ms after a request (a sleeping thread takes longer
// to wake than the work it would save) and sleeps on a condition variable after that.
struct SincosHelper
{
    std::thread             th;
    std::mutex              m;
    std::condition_variable cv;
    std::atomic<uint64_t>   ticket{0};
    std::atomic<size_t>     progress{0};  // nodes [begin, progress) of the current request are done
    std::atomic<bool>       stop{false};
    const double*           poses = nullptr;
    double*                 out   = nullptr;
    size_t                  begin = 0, end = 0;

    void run()
    {
        uint64_t seen = 0;
        for (;;)
        {
            const auto t0 = std::chrono::steady_clock::now();
            unsigned   spins = 0;
            while (ticket.load(std::memory_order_acquire) == seen && !stop.load(std::memory_order_relaxed))
            {
                if ((++spins & 1023u) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(2))
                {
                    std::unique_lock<std::mutex> lk(m);
                    cv.wait(lk, [&] { return ticket.load(std::memory_order_acquire) != seen || stop.load(); });
                    break;
                }
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
            }
            if (stop.load()) return;
            seen = ticket.load(std::memory_order_acquire);
            for (size_t i = begin; i < end; i += 128)
            {
                const size_t e = std::min(end, i + 128);
                mppb_nodes_xycs_from_poses(e - i, poses + 3 * i, out + 4 * i);
                progress.store(e, std::memory_order_release);
            }
        }
    }
    void post(const double* p, double* o, size_t b, size_t e)
    {
        poses = p;
        out   = o;
        begin = b;
        end   = e;
        progress.store(b, std::memory_order_relaxed);
        ticket.fetch_add(1, std::memory_order_release);
        {
            std::lock_guard<std::mutex> lk(m);
        }
        cv.notify_one();
    }
    void wait_for(size_t upTo) const
    {
        while (progress.load(std::memory_order_acquire) < upTo)
        {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
    }
    ~SincosHelper()
    {
        stop.store(true);
        {
            std::lock_guard<std::mutex> lk(m);
        }
        cv.notify_one();
        if (th.joinable()) th.join();
    }
};
SincosHelper* sincos_helper(mppb_ctx* ctx)
{
    if (!ctx->sincosHelper)
    {
        auto* h = new SincosHelper();
        h->th   = std::thread([h] { h->run(); });
        ctx->sincosHelper = h;
    }
    return static_cast<SincosHelper*>(ctx->sincosHelper);
}
}  // namespace
int mppb_selftest_sincos_helper(uint64_t seed, size_t n, int rounds)
{
    if (n < 8) return -1;
    SincosHelper h;
    h.th = std::thread([&h] { h.run(); });
    std::vector<double> poses(3 * n), got(4 * n), want(4 * n);
    uint64_t            x = seed * 6364136223846793005ull + 1442695040888963407ull;
    auto                rnd = [&]() {
        x ^= x << 13;
        x ^= x >> 7;
        x ^= x << 17;
        return static_cast<double>(x >> 11) / 9007199254740992.0;
    };
    int bad = 0;
    for (int r = 0; r < rounds; r++)
    {
        for (size_t i = 0; i < n; i++)
        {
            poses[3 * i + 0] = 200.0 * rnd();
            poses[3 * i + 1] = 200.0 * rnd();
            poses[3 * i + 2] = (rnd() - 0.5) * (r % 3 == 2 ? 1e6 : 6.283185307179586);
        }
        std::fill(got.begin(), got.end(), -7.0);
        const size_t split = 1 + static_cast<size_t>(rnd() * static_cast<double>(n - 2));
        h.post(poses.data(), got.data(), split, n);
        mppb_nodes_xycs_from_poses(split, poses.data(), got.data());
        // (the entry point asks for sub-batch after sub-batch)
        for (size_t upTo = split; upTo < n; upTo = std::min(n, upTo + n / 4 + 1)) h.wait_for(upTo);
        h.wait_for(n);
        mppb_nodes_xycs_from_poses(n, poses.data(), want.data());
        for (size_t i = 0; i < n; i++) bad += std::memcmp(&got[4 * i], &want[4 * i], 4 * sizeof(double)) != 0;
        if (r % 4 == 1) std::this_thread::sleep_for(std::chrono::microseconds(static_cast<int>(5000 * rnd())));
    }
    return bad;
}
static void sincos_helper_destroy(mppb_ctx* ctx)
{
    delete static_cast<SincosHelper*>(ctx->sincosHelper);
    ctx->sincosHelper = nullptr;
}
void mppb_nodes_xycs_from_poses(size_t n, const double* poses, double* out)
{
    for (size_t i = 0; i < n; i++)
    {
        out[4 * i + 0] = poses[3 * i + 0];
        out[4 * i + 1] = poses[3 * i + 1];
        out[4 * i + 2] = std::cos(poses[3 * i + 2]);
        out[4 * i + 3] = std::sin(poses[3 * i + 2]);
    }
}

int mppb_eval_nodes_device(mppb_ctx* ctx, size_t n_nodes, const double* d_nodes_xycs, double clip_dist, float* d_out)
{
    return mppb_eval_nodes_boxed_device(ctx, n_nodes, d_nodes_xycs, nullptr, clip_dist, d_out);
}
int mppb_eval_nodes_boxed_device(mppb_ctx* ctx, size_t n_nodes, const double* d_nodes_xycs, const float* d_node_boxes,
                                 double clip_dist, float* d_out)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    if (n_nodes && (!d_nodes_xycs || !d_out)) return fail(ctx, MPPB_ERR_INVALID_ARG, "null device pointers");
    return launch_tpobs_grid(ctx, n_nodes, d_nodes_xycs, d_node_boxes, clip_dist, d_out);
}

int mppb_eval_nodes(mppb_ctx* ctx, size_t n_nodes, const double* poses_xyphi, double clip_dist, float* out)
{
    return mppb_eval_nodes_boxed(ctx, n_nodes, poses_xyphi, nullptr, clip_dist, out);
}

static int eval_nodes_boxed_impl(mppb_ctx* ctx, size_t n_nodes, const double* poses_xyphi, const float* node_boxes,
                                 double clip_dist, float* out, bool wait);
int mppb_eval_nodes_boxed(mppb_ctx* ctx, size_t n_nodes, const double* poses_xyphi, const float* node_boxes,
                          double clip_dist, float* out)
{
    return eval_nodes_boxed_impl(ctx, n_nodes, poses_xyphi, node_boxes, clip_dist, out, true);
}
int mppb_eval_nodes_boxed_begin(mppb_ctx* ctx, size_t n_nodes, const double* poses_xyphi, const float* node_boxes,
                                double clip_dist, float* out)
{
    return eval_nodes_boxed_impl(ctx, n_nodes, poses_xyphi, node_boxes, clip_dist, out, false);
}

static int eval_nodes_boxed_impl(mppb_ctx* ctx, size_t n_nodes, const double* poses_xyphi, const float* node_boxes,
                                 double clip_dist, float* out, bool wait)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    if (n_nodes == 0) return MPPB_OK;
    if (!poses_xyphi || !out) return fail(ctx, MPPB_ERR_INVALID_ARG, "null host pointers");
    if (ctx->grids.empty()) return fail(ctx, MPPB_ERR_STATE, "no PTG tables set (mppb_set_ptg_grid)");
    MPPB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t cols     = static_cast<size_t>(ctx->totalPaths);
    const size_t inBytes  = n_nodes * 4 * sizeof(double);
    const size_t outBytes = n_nodes * cols * sizeof(float);
    MPPB_CUDA(ctx, ctx->hNodes.reserve(inBytes));
    MPPB_CUDA(ctx, ctx->dNodes.reserve(inBytes));
    const bool   nodesInPlace = inBytes <= kZeroCopyInputBytes && n_nodes <= 1024 && mapped_alias(ctx->hNodes.as<double>());
    const float* dBoxes       = nullptr;
    if (node_boxes)
    {
        if (nodesInPlace) dBoxes = mapped_alias(node_boxes);
        if (!dBoxes)
        {
            MPPB_CUDA(ctx, ctx->dBoxes.reserve(n_nodes * 4 * sizeof(float)));
            MPPB_CUDA(ctx, cudaMemcpyAsync(ctx->dBoxes.p, node_boxes, n_nodes * 4 * sizeof(float), cudaMemcpyHostToDevice,
                                           ctx->stream));
            dBoxes = ctx->dBoxes.as<float>();
        }
    }
    if (!ctx->auxStream)
    {
        MPPB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->auxStream, cudaStreamNonBlocking));
        MPPB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->evFork, cudaEventDisableTiming));
        MPPB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->evJoin, cudaEventDisableTiming));
        for (int i = 0; i < 2; i++)
        {
            MPPB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->auxMore[i], cudaStreamNonBlocking));
            MPPB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->evJoinMore[i], cudaEventDisableTiming));
        }
    }
    // Software pipeline over sub-batches on several streams: while the GPU evaluates sub-batch j
    // the host prepares cos/sin of j+1 and the copy engine drains the result of j-1.
    // D2H goes straight into the caller's buffer: a true async DMA if it is pinned
    // (mppb_host_alloc), a staged copy inside the driver otherwise.
    // If the caller's buffer is pinned (mppb_host_alloc / cudaHostRegister) it is device-accessible under UVA: the
    // kernel then stores its rows straight into it over PCIe (posted writes, 968 contiguous bytes per node), which
    // overlaps the result transfer with the computation CTA by CTA and removes the D2H copies altogether.
    static const bool noHelper  = std::getenv("MPPB_E2E_NO_HELPER") != nullptr;   // A/B measurements
    float* const outDev = mapped_alias(out);
    if (!outDev) MPPB_CUDA(ctx, ctx->dOut.reserve(outBytes));
    // <= 5 sub-batches of >= 820 nodes (API calls cost ~5 us each), round-robin over 4 streams: a sub-batch on a
    // stream of its own can start while the tails of the two before it are still draining (measured: 2 streams
    // 0.175 ms, 4 streams 0.165-0.170 ms per 4096-node call)
    constexpr int    kSubStreams = 4;
    if (ctx->helperAllowed < 0)
    {
        // a spinning helper needs a core of its own: with fewer than 8 CPUs in the process's affinity mask (e.g. one of
        // eight ranks on a 32-CPU box) it would take it from this thread or from a neighbour's (measured: slower)
        cpu_set_t set;
        CPU_ZERO(&set);
        ctx->helperAllowed = (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) >= 8) ? 1 : 0;
    }
    // Whether the helper pays depends on where the scheduler puts it (next to this thread on one core it costs more than
    // it saves): the 4th, the 7th and every 16th large call run without it, and when the calls with it are not at least 3 % faster in two
    // such comparisons in a row it sits out the next 1000 calls.
    const bool helperOk   = wait && !noHelper && ctx->helperAllowed == 1 && n_nodes >= 2048 && (ctx->helperPause == 0 || --ctx->helperPause == 0);
    const bool helperIsNew = ctx->sincosHelper == nullptr;  // (its first call starts the thread: not a measurement)
    const bool probePlain  = helperOk && (++ctx->helperCalls % 16u == 0 || ctx->helperCalls == 4 || ctx->helperCalls == 7);
    const bool wantHelper = helperOk && !probePlain;
    const auto tCall0     = std::chrono::steady_clock::now();
    // (5 sub-batches without the helper thread, 3 with it: measured 157-165 us and 141-144 us per 4096-node call)
    static const int envSubs     = std::getenv("MPPB_E2E_SUBS") ? std::max(1, std::atoi(std::getenv("MPPB_E2E_SUBS"))) : 0;
    const int        kSubBatches = envSubs ? envSubs : (wantHelper ? 3 : 5);
    const size_t  sub = n_nodes <= 1024 ? n_nodes : std::max<size_t>(820, (n_nodes + kSubBatches - 1) / kSubBatches);
    cudaStream_t  streams[kSubStreams] = {ctx->stream, ctx->auxStream, ctx->auxMore[0], ctx->auxMore[1]};
    const bool    single   = n_nodes <= sub;  // one sub-batch: no second stream, no fork/join events
    const int     nStreams = std::min<int>(kSubStreams, static_cast<int>((n_nodes + sub - 1) / sub));
    if (!single)
    {
        MPPB_CUDA(ctx, cudaEventRecord(ctx->evFork, ctx->stream));
        for (int i = 1; i < nStreams; i++) MPPB_CUDA(ctx, cudaStreamWaitEvent(streams[i], ctx->evFork, 0));
    }
    cudaStream_t saved = ctx->stream;
    int          rc    = MPPB_OK;
    size_t       j     = 0;
    // cos/sin of the headings: this thread takes the first sub-batch, the helper thread everything behind it
    SincosHelper* helper    = (wantHelper && !single) ? sincos_helper(ctx) : nullptr;
    if (helper) helper->post(poses_xyphi, ctx->hNodes.as<double>(), sub, n_nodes);
    for (size_t lo = 0; lo < n_nodes && rc == MPPB_OK; lo += sub, j++)
    {
        const size_t  n  = std::min(sub, n_nodes - lo);
        cudaStream_t  st = streams[j % nStreams];
        double*       hN = ctx->hNodes.as<double>() + 4 * lo;
        double*       dN = ctx->dNodes.as<double>() + 4 * lo;
        float*        dO = outDev ? outDev + lo * cols : ctx->dOut.as<float>() + lo * cols;
        if (helper && lo > 0)
            helper->wait_for(lo + n);
        else
            mppb_nodes_xycs_from_poses(n, poses_xyphi + 3 * lo, hN);
        cudaError_t e = cudaSuccess;
        if (nodesInPlace)
            dN = mapped_alias(hN);  // one small batch: the kernel reads the 32 B per node from the pinned staging
        else
            e = cudaMemcpyAsync(dN, hN, n * 4 * sizeof(double), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess)
        {
            rc = cuda_fail(ctx, e, "H2D node batch");
            break;
        }
        ctx->stream = st;  // launch on the sub-batch's stream
        // (plain form: this path is bound by the host's cos/sin and the rows' way over PCIe, measured; the second
        // launch of the near-first form would only add API calls)
        rc          = launch_tpobs_grid(ctx, n, dN, dBoxes ? dBoxes + 4 * lo : nullptr, clip_dist, dO, /*allowNear=*/false);
        ctx->stream = saved;
        if (rc != MPPB_OK) break;
        if (outDev) continue;  // rows already land in the caller's pinned buffer
        e = cudaMemcpyAsync(out + lo * cols, dO, n * cols * sizeof(float), cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) rc = cuda_fail(ctx, e, "D2H result");
    }
    if (helper) helper->wait_for(n_nodes);  // (also on the error paths: it reads the caller's poses until it is through)
    if (!single)
    {
        cudaEvent_t evs[3] = {ctx->evJoin, ctx->evJoinMore[0], ctx->evJoinMore[1]};
        for (int i = 1; i < nStreams; i++)
        {
            cudaEventRecord(evs[i - 1], streams[i]);
            cudaStreamWaitEvent(ctx->stream, evs[i - 1], 0);
        }
    }
    if (wait) MPPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // _begin: the caller's mppb_sync() completes the call
    if (helperOk && rc == MPPB_OK && !(helperIsNew && !probePlain))
    {
        const double us  = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - tCall0).count();
        double&      ema = probePlain ? ctx->helperEmaWithout : ctx->helperEmaWith;
        ema              = ema > 0 ? 0.75 * ema + 0.25 * us : us;
        if (probePlain && ctx->helperEmaWith > 0)
        {
            ctx->helperLoses = ctx->helperEmaWith > 0.97 * ctx->helperEmaWithout ? ctx->helperLoses + 1 : 0;
            if (ctx->helperLoses >= 2)
            {
                ctx->helperPause = 1000;
                ctx->helperLoses = 0;
                ctx->helperEmaWith = ctx->helperEmaWithout = 0;
            }
        }
    }
    return rc;
}


// ---- HolonomicBlend candidates --------------------------------------------------------------
int mppb_eval_holo_candidates_device(mppb_ctx* ctx, size_t n_cands, const mppb_holo_cand* d_cands,
                                     const double* d_nodes_xycs, double clip_dist, double* d_out)
{
    return mppb_eval_holo_candidates_boxed_device(ctx, n_cands, d_cands, d_nodes_xycs, nullptr, clip_dist, d_out);
}
int mppb_eval_holo_candidates_boxed_device(mppb_ctx* ctx, size_t n_cands, const mppb_holo_cand* d_cands,
                                           const double* d_nodes_xycs, const float* d_node_boxes, double clip_dist,
                                           double* d_out)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    if (n_cands && (!d_cands || !d_nodes_xycs || !d_out)) return fail(ctx, MPPB_ERR_INVALID_ARG, "null device pointers");
    return launch_tpobs_holo(ctx, n_cands, d_cands, d_nodes_xycs, d_node_boxes, clip_dist, d_out);
}

int mppb_eval_holo_candidates(mppb_ctx* ctx, size_t n_nodes, const double* poses_xyphi, size_t n_cands,
                              const mppb_holo_cand* cands, double clip_dist, double* out)
{
    return mppb_eval_holo_candidates_boxed(ctx, n_nodes, poses_xyphi, nullptr, n_cands, cands, clip_dist, out);
}

static int eval_holo_boxed_impl(mppb_ctx* ctx, size_t n_nodes, const double* poses_xyphi, const float* node_boxes,
                                size_t n_cands, const mppb_holo_cand* cands, double clip_dist, double* out, bool wait);
int mppb_eval_holo_candidates_boxed(mppb_ctx* ctx, size_t n_nodes, const double* poses_xyphi, const float* node_boxes,
                                    size_t n_cands, const mppb_holo_cand* cands, double clip_dist, double* out)
{
    return eval_holo_boxed_impl(ctx, n_nodes, poses_xyphi, node_boxes, n_cands, cands, clip_dist, out, true);
}
int mppb_eval_holo_candidates_boxed_begin(mppb_ctx* ctx, size_t n_nodes, const double* poses_xyphi,
                                          const float* node_boxes, size_t n_cands, const mppb_holo_cand* cands,
                                          double clip_dist, double* out)
{
    return eval_holo_boxed_impl(ctx, n_nodes, poses_xyphi, node_boxes, n_cands, cands, clip_dist, out, false);
}

// (node rows and boxes are staged in buffers of their own, so that a collision-grid call and a HolonomicBlend call
// of the same expansion can be in flight together)
static int eval_holo_boxed_impl(mppb_ctx* ctx, size_t n_nodes, const double* poses_xyphi, const float* node_boxes,
                                size_t n_cands, const mppb_holo_cand* cands, double clip_dist, double* out, bool wait)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    if (n_cands == 0) return MPPB_OK;
    if (!poses_xyphi || !cands || !out || n_nodes == 0) return fail(ctx, MPPB_ERR_INVALID_ARG, "null host pointers");
    for (size_t i = 0; i < n_cands; i++)
    {
        const mppb_holo_cand& c = cands[i];
        if (c.node < 0 || static_cast<size_t>(c.node) >= n_nodes)
            return fail(ctx, MPPB_ERR_INVALID_ARG, "candidate node index out of range");
        if (c.ptg_index < 0 || c.ptg_index >= static_cast<int>(ctx->ptgKinds.size()) || ctx->ptgKinds[c.ptg_index] != 1)
            return fail(ctx, MPPB_ERR_INVALID_ARG, "candidate ptg_index is not a HolonomicBlend PTG");
        if (!(c.T_ramp > 0)) return fail(ctx, MPPB_ERR_INVALID_ARG, "candidate T_ramp must be > 0");
    }
    MPPB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t nodeBytes = n_nodes * 4 * sizeof(double), candBytes = n_cands * sizeof(mppb_holo_cand);
    const size_t outBytes  = n_cands * sizeof(double);
    MPPB_CUDA(ctx, ctx->hNodesHolo.reserve(nodeBytes));
    mppb_nodes_xycs_from_poses(n_nodes, poses_xyphi, ctx->hNodesHolo.as<double>());
    // small batch with pinned caller buffers: the kernel works on them in place (no staging copies)
    const bool            small  = nodeBytes + candBytes <= kZeroCopyInputBytes;
    const double*         dNodes = small ? mapped_alias(ctx->hNodesHolo.as<double>()) : nullptr;
    const mppb_holo_cand* dCands = small ? mapped_alias(cands) : nullptr;
    const float*          dBoxes = (small && node_boxes) ? mapped_alias(node_boxes) : nullptr;
    double*               dOut   = mapped_alias(out);
    if (!dNodes)
    {
        MPPB_CUDA(ctx, ctx->dNodesHolo.reserve(nodeBytes));
        MPPB_CUDA(ctx, cudaMemcpyAsync(ctx->dNodesHolo.p, ctx->hNodesHolo.p, nodeBytes, cudaMemcpyHostToDevice, ctx->stream));
        dNodes = ctx->dNodesHolo.as<double>();
    }
    if (!dCands)
    {
        MPPB_CUDA(ctx, ctx->dStage.reserve(candBytes));
        MPPB_CUDA(ctx, cudaMemcpyAsync(ctx->dStage.p, cands, candBytes, cudaMemcpyHostToDevice, ctx->stream));
        dCands = ctx->dStage.as<mppb_holo_cand>();
    }
    if (node_boxes && !dBoxes)
    {
        MPPB_CUDA(ctx, ctx->dBoxesHolo.reserve(n_nodes * 4 * sizeof(float)));
        MPPB_CUDA(ctx, cudaMemcpyAsync(ctx->dBoxesHolo.p, node_boxes, n_nodes * 4 * sizeof(float), cudaMemcpyHostToDevice,
                                       ctx->stream));
        dBoxes = ctx->dBoxesHolo.as<float>();
    }
    const bool outInPlace = dOut != nullptr;
    if (!outInPlace)
    {
        MPPB_CUDA(ctx, ctx->dStage2.reserve(outBytes));
        dOut = ctx->dStage2.as<double>();
    }
    // Diagnostic alternative (mppb_set_holo_by_node; measured no faster, see tpobs_holo.cu): candidates grouped by node
    // -> one CTA per node, which walks the node's window once for all its candidates.
    bool grouped = ctx->holoByNode && n_nodes >= 2;
    for (size_t i = 1; grouped && i < n_cands; i++) grouped = cands[i].node >= cands[i - 1].node;
    int rc;
    if (grouped)
    {
        std::vector<uint32_t> begin(n_nodes + 1, 0);
        for (size_t i = 0; i < n_cands; i++) begin[static_cast<size_t>(cands[i].node) + 1]++;
        uint32_t maxPerNode = 0;
        for (size_t b = 0; b < n_nodes; b++)
        {
            maxPerNode = std::max(maxPerNode, begin[b + 1]);
            begin[b + 1] += begin[b];
        }
        MPPB_CUDA(ctx, ctx->dCandBegin.reserve((n_nodes + 1) * sizeof(uint32_t)));
        MPPB_CUDA(ctx, cudaMemcpyAsync(ctx->dCandBegin.p, begin.data(), (n_nodes + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice,
                                       ctx->stream));  // (pageable source: staged by the driver before the call returns)
        rc = launch_tpobs_holo_by_node(ctx, n_nodes, ctx->dCandBegin.as<uint32_t>(), maxPerNode, dCands, dNodes, dBoxes, clip_dist, dOut);
    }
    else
        rc = launch_tpobs_holo(ctx, n_cands, dCands, dNodes, dBoxes, clip_dist, dOut, n_nodes);
    if (rc != MPPB_OK) return rc;
    if (!outInPlace) MPPB_CUDA(ctx, cudaMemcpyAsync(out, dOut, outBytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (wait) MPPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // _begin: the caller's mppb_sync() completes the call
    return MPPB_OK;
}

// ---- cost evaluators ----------------------------------------------------------------------------
static int upload_evaluators(mppb_ctx* ctx)
{
    int n = 0;
    for (int i = 0; i < MPPB_MAX_EVALUATORS; i++)
        if (ctx->evals[i].kind != 0) n = i + 1;
    for (int i = 0; i < n; i++)
        if (ctx->evals[i].kind == 0) return fail(ctx, MPPB_ERR_STATE, "cost evaluator slots must be filled in order");
    ctx->numEvals = n;
    MPPB_CUDA(ctx, ctx->evalDescs.reserve(sizeof(ctx->evals)));
    MPPB_CUDA(ctx, cudaMemcpyAsync(ctx->evalDescs.p, ctx->evals, sizeof(ctx->evals), cudaMemcpyHostToDevice,
                                   ctx->stream));
    MPPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MPPB_OK;
}

int mppb_clear_evaluators(mppb_ctx* ctx)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    MPPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < MPPB_MAX_EVALUATORS; i++) ctx->evals[i] = EvaluatorDev{};
    ctx->numEvals = 0;
    ctx->evalEpoch++;
    return MPPB_OK;
}
int      mppb_num_evaluators(const mppb_ctx* ctx) { return ctx ? ctx->numEvals : 0; }
uint64_t mppb_evaluators_epoch(const mppb_ctx* ctx) { return ctx ? ctx->evalEpoch : 0; }

int mppb_set_costmap(mppb_ctx* ctx, int slot, const mppb_costmap_desc* d)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    if (slot < 0 || slot >= MPPB_MAX_EVALUATORS) return fail(ctx, MPPB_ERR_INVALID_ARG, "evaluator slot out of range");
    if (!d || !d->cells || d->nx <= 0 || d->ny <= 0 || !(d->res > 0))
        return fail(ctx, MPPB_ERR_INVALID_ARG, "bad costmap descriptor");
    MPPB_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->evalEpoch++;
    MPPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const size_t bytes = static_cast<size_t>(d->nx) * d->ny * sizeof(double);
    MPPB_CUDA(ctx, ctx->evalCells[slot].reserve(bytes));
    MPPB_CUDA(ctx, cudaMemcpyAsync(ctx->evalCells[slot].p, d->cells, bytes, cudaMemcpyHostToDevice, ctx->stream));
    EvaluatorDev E{};
    E.kind       = 1;
    E.useAverage = d->use_average != 0;
    E.xmin       = d->x_min;
    E.ymin       = d->y_min;
    E.res        = d->res;
    E.nx         = d->nx;
    E.ny         = d->ny;
    E.cells      = ctx->evalCells[slot].as<double>();
    ctx->evals[slot] = E;
    return upload_evaluators(ctx);
}

int mppb_set_waypoints(mppb_ctx* ctx, int slot, const double* xs, const double* ys, size_t n, double influence_radius,
                       double cost_scale, int use_average)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    if (slot < 0 || slot >= MPPB_MAX_EVALUATORS) return fail(ctx, MPPB_ERR_INVALID_ARG, "evaluator slot out of range");
    if (n && (!xs || !ys)) return fail(ctx, MPPB_ERR_INVALID_ARG, "null waypoint arrays");
    if (!(influence_radius > 0)) return fail(ctx, MPPB_ERR_INVALID_ARG, "influence radius must be > 0");
    ctx->evalEpoch++;
    if (n > 0x7fffffffull) return fail(ctx, MPPB_ERR_INVALID_ARG, "too many waypoints");
    MPPB_CUDA(ctx, cudaSetDevice(ctx->device));
    MPPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<float2> w(n);
    for (size_t i = 0; i < n; i++) w[i] = make_float2(static_cast<float>(xs[i]), static_cast<float>(ys[i]));
    MPPB_CUDA(ctx, ctx->evalWps[slot].reserve(std::max<size_t>(n, 1) * sizeof(float2)));
    if (n)
        MPPB_CUDA(ctx, cudaMemcpyAsync(ctx->evalWps[slot].p, w.data(), n * sizeof(float2), cudaMemcpyHostToDevice,
                                       ctx->stream));
    MPPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    EvaluatorDev E{};
    E.kind       = 2;
    E.useAverage = use_average != 0;
    E.wps        = ctx->evalWps[slot].as<float2>();
    E.nWps       = static_cast<int>(n);
    E.radius     = influence_radius;
    E.scale      = cost_scale;
    E.radiusSqrF = static_cast<float>(influence_radius * influence_radius);
    ctx->evals[slot] = E;
    return upload_evaluators(ctx);
}

int mppb_eval_edge_costs_device(mppb_ctx* ctx, size_t n_edges, const double* d_from_xycs,
                                const uint32_t* d_sample_offsets, const double* d_rel_xy, double* d_out)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    if (n_edges && ctx->numEvals && (!d_from_xycs || !d_sample_offsets || !d_rel_xy || !d_out))
        return fail(ctx, MPPB_ERR_INVALID_ARG, "null device pointers");
    return launch_edge_costs(ctx, n_edges, d_from_xycs, d_sample_offsets, d_rel_xy, d_out);
}

int mppb_eval_edge_costs(mppb_ctx* ctx, size_t n_edges, const double* from_xyphi, const uint32_t* sample_offsets,
                         const double* rel_xy, double* out)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    if (n_edges == 0 || ctx->numEvals == 0) return MPPB_OK;
    if (!from_xyphi || !sample_offsets || !rel_xy || !out) return fail(ctx, MPPB_ERR_INVALID_ARG, "null host pointers");
    for (size_t e = 0; e < n_edges; e++)
        if (sample_offsets[e] > sample_offsets[e + 1]) return fail(ctx, MPPB_ERR_INVALID_ARG, "sample_offsets not sorted");
    MPPB_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t m         = sample_offsets[n_edges];
    const size_t fromBytes = n_edges * 4 * sizeof(double), offBytes = (n_edges + 1) * sizeof(uint32_t);
    const size_t relBytes  = std::max<size_t>(m, 1) * 2 * sizeof(double);
    const size_t outBytes  = n_edges * ctx->numEvals * sizeof(double);
    MPPB_CUDA(ctx, ctx->hNodes.reserve(fromBytes));
    mppb_nodes_xycs_from_poses(n_edges, from_xyphi, ctx->hNodes.as<double>());
    // small batch with pinned caller buffers: the kernel works on them in place (no staging copies)
    const bool      small = fromBytes + offBytes + relBytes <= kZeroCopyInputBytes;
    const double*   dFrom = small ? mapped_alias(ctx->hNodes.as<double>()) : nullptr;
    const uint32_t* dOff  = small ? mapped_alias(sample_offsets) : nullptr;
    const double*   dRel  = small ? mapped_alias(rel_xy) : nullptr;
    double*         dOut  = mapped_alias(out);
    if (!dFrom)
    {
        MPPB_CUDA(ctx, ctx->dNodes.reserve(fromBytes));
        MPPB_CUDA(ctx, cudaMemcpyAsync(ctx->dNodes.p, ctx->hNodes.p, fromBytes, cudaMemcpyHostToDevice, ctx->stream));
        dFrom = ctx->dNodes.as<double>();
    }
    if (!dOff)
    {
        MPPB_CUDA(ctx, ctx->dStage.reserve(offBytes));
        MPPB_CUDA(ctx, cudaMemcpyAsync(ctx->dStage.p, sample_offsets, offBytes, cudaMemcpyHostToDevice, ctx->stream));
        dOff = ctx->dStage.as<uint32_t>();
    }
    if (!dRel)
    {
        MPPB_CUDA(ctx, ctx->dStage2.reserve(relBytes));
        if (m) MPPB_CUDA(ctx, cudaMemcpyAsync(ctx->dStage2.p, rel_xy, m * 2 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        dRel = ctx->dStage2.as<double>();
    }
    const bool outInPlace = dOut != nullptr;
    if (!outInPlace)
    {
        MPPB_CUDA(ctx, ctx->dStage3.reserve(outBytes));
        dOut = ctx->dStage3.as<double>();
    }
    const int rc = launch_edge_costs(ctx, n_edges, dFrom, dOff, dRel, dOut, m);
    if (rc != MPPB_OK) return rc;
    if (!outInPlace) MPPB_CUDA(ctx, cudaMemcpyAsync(out, dOut, outBytes, cudaMemcpyDeviceToHost, ctx->stream));
    MPPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MPPB_OK;
}

int mppb_costmap_nearest_d2(mppb_ctx* ctx, const float* xs, const float* ys, size_t n, double x_min, double y_min,
                            double res, int nx, int ny, float clearance, float* out_d2)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    if (nx <= 0 || ny <= 0 || !(res > 0) || !(clearance > 0) || !out_d2)
        return fail(ctx, MPPB_ERR_INVALID_ARG, "bad costmap geometry");
    // bins of about the clearance distance keep the per-cell search window at a few bins
    const double bin = std::max(0.05, static_cast<double>(clearance));
    int          rc  = set_cloud(ctx, ctx->auxCloud, xs, ys, n, bin, 0, nullptr, cudaMemcpyHostToDevice);
    if (rc != MPPB_OK) return rc;
    const size_t bytes = static_cast<size_t>(nx) * ny * sizeof(float);
    MPPB_CUDA(ctx, ctx->dStage3.reserve(bytes));
    rc = launch_costmap_d2(ctx, ctx->auxCloud, x_min, y_min, res, nx, ny, clearance, ctx->dStage3.as<float>());
    if (rc != MPPB_OK) return rc;
    MPPB_CUDA(ctx, cudaMemcpyAsync(out_d2, ctx->dStage3.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MPPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MPPB_OK;
}

}  // extern "C"
