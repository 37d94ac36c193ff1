// Several GPUs: cloud sharding and the min-merge of partial TP-obstacle rows (include/mppb.h, "several GPUs";
// SURVEY.md 8e, config C5). The reference is single-threaded and has no counterpart; what makes the split exact is
// that tp_obstacles_single_path (src/algos/tp_obstacles_single_path.cpp:24-34) folds an order-independent minimum
// over the obstacle points, so the element-wise minimum of the rows computed against disjoint shards of the cloud is
// the row computed against the whole cloud, bit for bit.
//
// Two merges:
//  * NCCL: ncclAllReduce(ncclMin) on the context's stream. libnccl.so.2 is opened with dlopen the first time it is
//    needed (the copy already in the process, e.g. torch's, or the system's), so the library keeps no link-time
//    dependency on it.
//  * fused: tpobs_grid_kernel<PEERS> stores each finished row into slot `rank` of every rank's receive area (peer
//    memory mapped through CUDA IPC, NVLink underneath) as soon as the row is complete, i.e. while the rest of the
//    batch is still being evaluated; the last CTA raises the rank's flag on every peer; merge_slots_kernel on each
//    rank waits for all flags of the epoch and writes min over the slots into the caller's rows. No atomics, no
//    initialisation of the receive area (every slot is overwritten in full each step); two parity halves make a slot
//    safe to overwrite: a rank starts step s+1 only after it has seen every flag of step s, i.e. after every rank has
//    finished WRITING step s, and nobody reads half (s+1)&1 = (s-1)&1 any more, because step s-1 was merged before
//    step s was launched (stream order).
#include <dlfcn.h>

#include <algorithm>
#include <cstring>

#include "mppb_internal.cuh"

namespace mppb
{
namespace
{
// ---- NCCL through dlopen --------------------------------------------------------------------------------------------
struct NcclUniqueId
{
    char internal[MPPB_COMM_ID_BYTES];
};
static_assert(MPPB_COMM_ID_BYTES == 128, "NCCL_UNIQUE_ID_BYTES");
constexpr int kNcclFloat32 = 7, kNcclMin = 3;  // ncclDataType_t / ncclRedOp_t values of nccl.h (stable across 2.x)
struct NcclApi
{
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId*)                                                           = nullptr;
    int (*CommInitRank)(void**, int, NcclUniqueId, int)                                         = nullptr;
    int (*CommDestroy)(void*)                                                                   = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t)                 = nullptr;
    const char* (*GetErrorString)(int)                                                          = nullptr;
    std::string error;
};
NcclApi& nccl()
{
    static NcclApi api = [] {
        NcclApi a;
        for (const char* name : {"libnccl.so.2", "libnccl.so"})
        {
            a.lib = dlopen(name, RTLD_NOW | RTLD_NOLOAD);  // a copy the process already holds (torch's) first
            if (!a.lib) a.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (a.lib) break;
        }
        if (!a.lib)
        {
            const char* why = dlerror();  // (cleared by the call: read once)
            a.error         = std::string("libnccl.so.2 not found: ") + (why ? why : "");
            return a;
        }
        a.GetUniqueId    = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(a.lib, "ncclGetUniqueId"));
        a.CommInitRank   = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(a.lib, "ncclCommInitRank"));
        a.CommDestroy    = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(a.lib, "ncclCommDestroy"));
        a.AllReduce      = reinterpret_cast<decltype(a.AllReduce)>(dlsym(a.lib, "ncclAllReduce"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(a.lib, "ncclGetErrorString"));
        if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce) a.error = "libnccl.so.2 lacks the expected symbols";
        return a;
    }();
    return api;
}
int nccl_fail(mppb_ctx* ctx, int rc, const char* what)
{
    NcclApi& a = nccl();
    return fail(ctx, MPPB_ERR_CUDA, std::string("NCCL error in ") + what + ": " + (a.GetErrorString ? a.GetErrorString(rc) : "?"));
}

// ---- fused merge: wait for every rank's flag of this epoch, then min over the slots ---------------------------------
__global__ void __launch_bounds__(256)
    merge_slots_kernel(const float* __restrict__ area, const volatile uint32_t* __restrict__ flags, int world, uint32_t epoch,
                       size_t slotStride, size_t count, float* __restrict__ out, int* __restrict__ status)
{
    __shared__ int ok;
    if (threadIdx.x == 0)
    {
        // every block waits by itself (no grid-wide dependency); bounded, so that a lost peer ends as an error code
        // and not as a hung GPU
        ok = 1;
        const long long t0 = clock64();
        for (int r = 0; r < world; r++)
            while (static_cast<int32_t>(flags[r] - epoch) < 0)
                if (clock64() - t0 > 20000000000ll)  // ~10 s
                {
                    ok = 0;
                    break;
                }
        __threadfence_system();
    }
    __syncthreads();
    if (!ok)
    {
        if (threadIdx.x == 0) atomicExch(status, 1);
        return;
    }
    const float* half = area + static_cast<size_t>(epoch & 1u) * world * slotStride;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < count; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    {
        // (values are 0, a table float or +inf: fminf is the reference's min; no NaNs occur)
        float v = __ldcv(half + i);
        for (int r = 1; r < world; r++) v = fminf(v, __ldcv(half + static_cast<size_t>(r) * slotStride + i));
        out[i] = v;
    }
}
}  // namespace

void multigpu_release(mppb_ctx* ctx)
{
    for (int p = 0; p < MPPB_MAX_RANKS; p++)
        if (ctx->peerOpened[p])
        {
            cudaIpcCloseMemHandle(ctx->peerOpened[p]);
            ctx->peerOpened[p] = nullptr;
        }
    ctx->peerLocal.release();
    ctx->peerDone.release();
    ctx->peer = PeerRowsDev{};
    if (ctx->ncclComm && nccl().CommDestroy) nccl().CommDestroy(ctx->ncclComm);
    ctx->ncclComm = nullptr;
}
}  // namespace mppb

using namespace mppb;

extern "C" {

int mppb_comm_unique_id(uint8_t* out_id)
{
    if (!out_id) return MPPB_ERR_INVALID_ARG;
    NcclApi& a = nccl();
    if (!a.error.empty()) return fail(nullptr, MPPB_ERR_UNSUPPORTED, a.error);
    NcclUniqueId id;
    const int    rc = a.GetUniqueId(&id);
    if (rc != 0) return nccl_fail(nullptr, rc, "ncclGetUniqueId");
    std::memcpy(out_id, id.internal, MPPB_COMM_ID_BYTES);
    return MPPB_OK;
}

int mppb_comm_init(mppb_ctx* ctx, int world, int rank, const uint8_t* id)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    if (!id || world < 1 || world > MPPB_MAX_RANKS || rank < 0 || rank >= world) return fail(ctx, MPPB_ERR_INVALID_ARG, "bad communicator arguments");
    NcclApi& a = nccl();
    if (!a.error.empty()) return fail(ctx, MPPB_ERR_UNSUPPORTED, a.error);
    MPPB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->ncclComm) a.CommDestroy(ctx->ncclComm);
    ctx->ncclComm = nullptr;
    NcclUniqueId uid;
    std::memcpy(uid.internal, id, MPPB_COMM_ID_BYTES);
    const int rc = a.CommInitRank(&ctx->ncclComm, world, uid, rank);
    if (rc != 0) return nccl_fail(ctx, rc, "ncclCommInitRank");
    ctx->commWorld = world;
    ctx->commRank  = rank;
    return MPPB_OK;
}
int mppb_comm_world(const mppb_ctx* ctx) { return ctx ? ctx->commWorld : 1; }
int mppb_comm_rank(const mppb_ctx* ctx) { return ctx ? ctx->commRank : 0; }

int mppb_set_obstacles_shard(mppb_ctx* ctx, const float* xs, const float* ys, size_t n, double bin_size, int shard, int n_shards)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    if (n_shards < 1 || shard < 0 || shard >= n_shards) return fail(ctx, MPPB_ERR_INVALID_ARG, "bad shard index");
    if (n && (!xs || !ys)) return fail(ctx, MPPB_ERR_INVALID_ARG, "null cloud pointers");
    std::vector<float> sx, sy;
    sx.reserve(n / n_shards + 1);
    sy.reserve(n / n_shards + 1);
    for (size_t i = static_cast<size_t>(shard); i < n; i += static_cast<size_t>(n_shards))
    {
        sx.push_back(xs[i]);
        sy.push_back(ys[i]);
    }
    return mppb_set_obstacles(ctx, sx.data(), sy.data(), sx.size(), bin_size, 0);
}

int mppb_allreduce_min_device(mppb_ctx* ctx, float* d_rows, size_t count)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    if (count == 0) return MPPB_OK;
    if (!d_rows) return fail(ctx, MPPB_ERR_INVALID_ARG, "null device pointer");
    if (!ctx->ncclComm) return ctx->commWorld == 1 ? MPPB_OK : fail(ctx, MPPB_ERR_STATE, "no communicator (mppb_comm_init)");
    MPPB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int rc = nccl().AllReduce(d_rows, d_rows, count, kNcclFloat32, kNcclMin, ctx->ncclComm, ctx->stream);
    if (rc != 0) return nccl_fail(ctx, rc, "ncclAllReduce");
    return MPPB_OK;
}

// receive area of a rank: [2 parity halves][world slots][maxNodes][cols] floats, then flags [MPPB_MAX_RANKS] + status
static size_t peer_area_floats(size_t world, size_t maxNodes, size_t cols) { return 2 * world * maxNodes * cols; }

int mppb_peer_rows_create(mppb_ctx* ctx, size_t max_nodes, uint8_t* out_handle)
{
    if (!ctx || !out_handle) return MPPB_ERR_INVALID_ARG;
    if (max_nodes == 0 || ctx->totalPaths == 0) return fail(ctx, MPPB_ERR_STATE, "set the PTG tables first and give a batch size");
    if (ctx->commWorld < 2) return fail(ctx, MPPB_ERR_STATE, "no communicator of more than one rank (mppb_comm_init)");
    MPPB_CUDA(ctx, cudaSetDevice(ctx->device));
    MPPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const size_t cols  = static_cast<size_t>(ctx->totalPaths);
    const size_t bytes = peer_area_floats(ctx->commWorld, max_nodes, cols) * sizeof(float) + (MPPB_MAX_RANKS + 4) * sizeof(uint32_t);
    ctx->peerLocal.release();  // (a fresh cudaMalloc: IPC handles refer to whole allocations)
    MPPB_CUDA(ctx, ctx->peerLocal.reserve(bytes));
    MPPB_CUDA(ctx, cudaMemsetAsync(ctx->peerLocal.p, 0, bytes, ctx->stream));
    MPPB_CUDA(ctx, ctx->peerDone.reserve(sizeof(unsigned)));
    MPPB_CUDA(ctx, cudaMemsetAsync(ctx->peerDone.p, 0, sizeof(unsigned), ctx->stream));
    MPPB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->peerMaxNodes = max_nodes;
    ctx->peerCols     = cols;
    ctx->peerEpoch    = 0;
    ctx->peer         = PeerRowsDev{};
    cudaIpcMemHandle_t h;
    static_assert(sizeof(cudaIpcMemHandle_t) == MPPB_IPC_HANDLE_BYTES, "CUDA IPC handle size");
    MPPB_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->peerLocal.p));
    std::memcpy(out_handle, &h, sizeof(h));
    return MPPB_OK;
}

int mppb_peer_rows_attach(mppb_ctx* ctx, int world, int rank, const uint8_t* handles)
{
    if (!ctx || !handles) return MPPB_ERR_INVALID_ARG;
    if (world != ctx->commWorld || rank != ctx->commRank) return fail(ctx, MPPB_ERR_INVALID_ARG, "world/rank differ from the communicator's");
    if (!ctx->peerLocal.p) return fail(ctx, MPPB_ERR_STATE, "mppb_peer_rows_create first");
    MPPB_CUDA(ctx, cudaSetDevice(ctx->device));
    PeerRowsDev P{};
    P.world      = world;
    P.rank       = rank;
    P.slotStride = ctx->peerMaxNodes * ctx->peerCols;
    P.done       = ctx->peerDone.as<unsigned>();
    const size_t areaFloats = peer_area_floats(world, ctx->peerMaxNodes, ctx->peerCols);
    for (int p = 0; p < world; p++)
    {
        void* base = nullptr;
        if (p == rank)
            base = ctx->peerLocal.p;
        else
        {
            cudaIpcMemHandle_t h;
            std::memcpy(&h, handles + static_cast<size_t>(p) * MPPB_IPC_HANDLE_BYTES, sizeof(h));
            if (ctx->peerOpened[p]) cudaIpcCloseMemHandle(ctx->peerOpened[p]);
            ctx->peerOpened[p] = nullptr;
            MPPB_CUDA(ctx, cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
            ctx->peerOpened[p] = base;
        }
        P.rows[p]  = static_cast<float*>(base);
        P.flags[p] = reinterpret_cast<uint32_t*>(static_cast<float*>(base) + areaFloats);
    }
    ctx->peer = P;
    return MPPB_OK;
}

int mppb_sharded_mode(const mppb_ctx* ctx)
{
    if (!ctx) return 0;
    if (ctx->peer.world > 1) return 2;
    return ctx->ncclComm && ctx->commWorld > 1 ? 1 : 0;
}

int mppb_eval_nodes_sharded_device(mppb_ctx* ctx, size_t n_nodes, const double* d_nodes_xycs, double clip_dist, float* d_out)
{
    if (!ctx) return MPPB_ERR_INVALID_ARG;
    if (n_nodes == 0) return MPPB_OK;
    if (!d_nodes_xycs || !d_out) return fail(ctx, MPPB_ERR_INVALID_ARG, "null device pointers");
    const int mode = mppb_sharded_mode(ctx);
    if (mode != 2)
    {
        const int rc = launch_tpobs_grid(ctx, n_nodes, d_nodes_xycs, nullptr, clip_dist, d_out);
        if (rc != MPPB_OK || mode == 0) return rc;
        return mppb_allreduce_min_device(ctx, d_out, n_nodes * static_cast<size_t>(ctx->totalPaths));
    }
    if (n_nodes > ctx->peerMaxNodes) return fail(ctx, MPPB_ERR_INVALID_ARG, "batch larger than the receive slots (mppb_peer_rows_create)");
    if (static_cast<size_t>(ctx->totalPaths) != ctx->peerCols) return fail(ctx, MPPB_ERR_STATE, "PTG tables changed since mppb_peer_rows_create");
    MPPB_CUDA(ctx, cudaSetDevice(ctx->device));
    PeerRowsDev P = ctx->peer;
    P.epoch       = ++ctx->peerEpoch;
    int rc        = launch_tpobs_grid_peers(ctx, n_nodes, d_nodes_xycs, clip_dist, P);
    if (rc != MPPB_OK) return rc;
    const size_t count      = n_nodes * ctx->peerCols;
    const size_t areaFloats = peer_area_floats(P.world, ctx->peerMaxNodes, ctx->peerCols);
    uint32_t*    flags      = reinterpret_cast<uint32_t*>(ctx->peerLocal.as<float>() + areaFloats);
    int*         status     = reinterpret_cast<int*>(flags + MPPB_MAX_RANKS);
    const unsigned blocks   = static_cast<unsigned>(std::min<size_t>((count + 255) / 256, static_cast<size_t>(ctx->numSMs) * 4));
    merge_slots_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->peerLocal.as<float>(), flags, P.world, P.epoch, P.slotStride, count, d_out, status);
    ctx->launches++;
    MPPB_CUDA(ctx, cudaGetLastError());
    return MPPB_OK;
}

}  // extern "C"
