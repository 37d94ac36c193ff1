// TEST INFRASTRUCTURE — not part of the product, never shipped, never loaded by the product.
//
// A stand-in for the DEVICE entry points of include/mppb.h (the ones libmpp_b200's host code reaches through the PLT)
// that answers them on the CPU with the oracle (oracle/_build/liboracle.so). Pre-loaded (LD_PRELOAD) into a test
// process it lets the product's HOST logic — mpp::TPS_Astar's phases A, C, D, F, its open set, lattice, motion tree,
// the split collision calls, the deferred tree insertions — run without a GPU, so that `-m "not gpu"` can compare the
// product planner's motion tree with the oracle's. Nothing here is a CPU fallback of the product: libmpp_b200.so itself
// still fails loudly without CUDA (mppb_ctx_create), and the GPU parity of the kernels is the `-m gpu` tests' job.
//
// Supported: collision-grid PTGs (answered by orc_eval_nodes on oracle PTG objects the harness registers), cost
// evaluators (orc_eval_edge_costs on oracle evaluators the harness registers per slot), one obstacle cloud with the
// ObstacleSource.cpp:29-40 clipping box, and HolonomicBlend candidates: the harness registers a table
// (vxf, vyf, vf) -> (k, trimmed speed) made with the product's host PTG, the candidate record is matched against it bit
// for bit, and an oracle HolonomicBlend object of the stand-in's own (not the oracle planner's) is put into the
// candidate's initial velocity and trimmed speed and folded over the node's local obstacles.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <set>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/mppb.h"

extern "C" {
// oracle/oracle_capi.cpp
double orc_eval_nodes(void** ptgs, int nptg, const float* xs, const float* ys, size_t n, const double* poses_xyphi,
                      size_t B, double clip, double* out, int mode, int nthreads);
int    orc_eval_edge_costs(void* evaluator, int kind, size_t nEdges, const double* from_xyphi,
                           const uint32_t* sample_offsets, const double* rel_xy, double* out_cost);
long   orc_local_obstacles(const float* xs, const float* ys, size_t n, const double* pose_xyphi, double clip, float* out_x,
                           float* out_y);
int    orc_ptg_update_dyn_state(void* h, const double* dyn7, int force);
void   orc_ptg_set_trim_speed(void* h, double s);
double orc_ptg_update_tpobs_single(void* h, double ox, double oy, int k, double tp_obstacle_k);
}

struct mppb_ctx
{
    std::string        lastError;
    std::vector<int>   numPaths;  // per uploaded collision-grid PTG
    std::vector<float> xs, ys;    // the (clipped) cloud
    bool               haveBox = false;
    float              box[4]  = {0, 0, 0, 0};
    int                holoPtgs = 0;
    int                numEvals = 0;
    uint64_t           evalEpoch = 0, cloudEpoch = 0, ptgEpoch = 0;
    uint64_t           calls[4]  = {0, 0, 0, 0};  // eval_nodes, eval_nodes_begin, eval_edge_costs, sync
    uint64_t           expandCalls = 0, expandNodes = 0;
    // device-side expansion: the product's sampled paths (copied at mppb_set_ptg_paths) and the planner parameters
    struct Paths
    {
        bool                  set = false;
        int                   K   = 0;
        double                dt = 0, refDistance = 0;
        std::vector<uint32_t> begin;
        std::vector<float>    x, y, phi, dist, v, w;
        std::vector<double>   cs, sn;
    } paths[MPPB_MAX_PTGS];
    bool                paramsSet = false;
    double              resXY = 0, resYaw = 0;
    unsigned            M = 0, S = 0, nSeg = 0;
    std::vector<double> timestamps;
    std::vector<float>  lastRows;
};

namespace
{
std::vector<void*> g_oraclePtgs;
void*              g_oracleEval[MPPB_MAX_EVALUATORS]     = {};
int                g_oracleEvalKind[MPPB_MAX_EVALUATORS] = {};
struct HoloKey
{
    double vxf, vyf, vf, speed;
    int    k;
};
std::vector<HoloKey> g_holoTable;
void*                g_oracleHolo = nullptr;  // the stand-in's own oracle HolonomicBlend
std::string        g_globalError;
mppb_ctx*          g_lastCtx = nullptr;

int fail(mppb_ctx* ctx, const char* msg)
{
    if (ctx) ctx->lastError = msg;
    g_globalError = msg;
    return MPPB_ERR_STATE;
}
int eval_nodes(mppb_ctx* ctx, size_t n, const double* poses, const float* boxes, double clip, float* out)
{
    if (g_oraclePtgs.size() + ctx->holoPtgs != ctx->numPaths.size())
        return fail(ctx, "mppb_sim: register the oracle PTGs first");
    size_t cols = 0;
    for (int k : ctx->numPaths) cols += static_cast<size_t>(k);
    std::vector<double> tmp(n * cols);
    const int           nptg = static_cast<int>(g_oraclePtgs.size());
    if (!boxes)
    {
        if (orc_eval_nodes(g_oraclePtgs.data(), nptg, ctx->xs.data(), ctx->ys.data(), ctx->xs.size(), poses, n, clip,
                           tmp.data(), 1, 1) < 0)
            return fail(ctx, "mppb_sim: orc_eval_nodes failed");
    }
    else
    {
        // per-node world box (ObstacleSource.cpp:29-40 applied per query): the node sees the points inside its box
        std::vector<float> bx, by;
        for (size_t i = 0; i < n; i++)
        {
            const float* b = boxes + 4 * i;
            bx.clear();
            by.clear();
            for (size_t p = 0; p < ctx->xs.size(); p++)
            {
                const float x = ctx->xs[p], y = ctx->ys[p];
                if (x < b[0] || x > b[2] || y < b[1] || y > b[3]) continue;
                bx.push_back(x);
                by.push_back(y);
            }
            if (orc_eval_nodes(g_oraclePtgs.data(), nptg, bx.data(), by.data(), bx.size(), poses + 3 * i, 1, clip,
                               tmp.data() + i * cols, 1, 1) < 0)
                return fail(ctx, "mppb_sim: orc_eval_nodes failed");
        }
    }
    // the oracle returns min(init_k, c): 0, a table float or the reference distance — all exact in float; the
    // planner applies min(init_k, .) again, which changes nothing
    for (size_t i = 0; i < tmp.size(); i++) out[i] = static_cast<float>(tmp[i]);
    return MPPB_OK;
}
}  // namespace

extern "C" {

// ---- harness controls ----
void mppbsim_set_holo(void* oracle_holo_ptg, size_t n, const double* vxf, const double* vyf, const double* vf,
                      const int* k, const double* speed)
{
    g_oracleHolo = oracle_holo_ptg;
    g_holoTable.clear();
    for (size_t i = 0; i < n; i++) g_holoTable.push_back(HoloKey{vxf[i], vyf[i], vf[i], speed[i], k[i]});
}
void mppbsim_set_oracle_ptgs(void** handles, int n) { g_oraclePtgs.assign(handles, handles + n); }
void mppbsim_set_oracle_evaluator(int slot, void* handle, int kind)
{
    g_oracleEval[slot]     = handle;
    g_oracleEvalKind[slot] = kind;
}
void mppbsim_call_counts(uint64_t* out4)
{
    for (int i = 0; i < 4; i++) out4[i] = g_lastCtx ? g_lastCtx->calls[i] : 0;
}

// ---- device entry points of include/mppb.h ----
const char* mppb_last_global_error(void) { return g_globalError.c_str(); }
int         mppb_ctx_create(int, void*, mppb_ctx** out)
{
    *out      = new mppb_ctx();
    g_lastCtx = *out;
    return MPPB_OK;
}
void mppb_ctx_destroy(mppb_ctx* ctx)
{
    if (g_lastCtx == ctx) g_lastCtx = nullptr;
    delete ctx;
}
const char* mppb_last_error(const mppb_ctx* ctx) { return ctx ? ctx->lastError.c_str() : ""; }
int         mppb_sync(mppb_ctx* ctx)
{
    ctx->calls[3]++;
    return MPPB_OK;
}
int mppb_host_alloc(size_t bytes, void** out)
{
    return posix_memalign(out, 256, bytes ? bytes : 256) == 0 ? MPPB_OK : MPPB_ERR_STATE;
}
int mppb_host_free(void* p)
{
    free(p);
    return MPPB_OK;
}

int mppb_set_obstacles_clipped(mppb_ctx* ctx, const float* xs, const float* ys, size_t n, double, int append,
                               const float* box)
{
    ctx->cloudEpoch++;
    if (!append)
    {
        ctx->xs.clear();
        ctx->ys.clear();
        ctx->haveBox = box != nullptr;
        if (box) std::memcpy(ctx->box, box, sizeof(ctx->box));
    }
    for (size_t i = 0; i < n; i++)
    {
        const float x = xs[i], y = ys[i];
        if (!std::isfinite(x) || !std::isfinite(y)) continue;
        if (ctx->haveBox && (x < ctx->box[0] || x > ctx->box[2] || y < ctx->box[1] || y > ctx->box[3])) continue;
        ctx->xs.push_back(x);
        ctx->ys.push_back(y);
    }
    return MPPB_OK;
}
int mppb_set_obstacles(mppb_ctx* ctx, const float* xs, const float* ys, size_t n, double bin, int append)
{
    return mppb_set_obstacles_clipped(ctx, xs, ys, n, bin, append, nullptr);
}

uint64_t mppb_obstacles_epoch(const mppb_ctx* ctx) { return ctx->cloudEpoch; }
uint64_t mppb_ptgs_epoch(const mppb_ctx* ctx) { return ctx->ptgEpoch; }
int      mppb_ctx_device(const mppb_ctx*) { return 0; }

int mppb_clear_ptgs(mppb_ctx* ctx)
{
    ctx->ptgEpoch++;
    ctx->holoPtgs = 0;
    ctx->numPaths.clear();
    return MPPB_OK;
}
int mppb_set_ptg_grid(mppb_ctx* ctx, int ptg_index, const mppb_ptg_grid_desc* desc)
{
    if (ptg_index != static_cast<int>(ctx->numPaths.size())) return fail(ctx, "mppb_sim: PTGs must be set in order");
    ctx->ptgEpoch++;
    ctx->numPaths.push_back(desc->num_paths);
    return MPPB_OK;
}
int mppb_set_ptg_holo(mppb_ctx* ctx, int ptg_index, const mppb_ptg_holo_desc*)
{
    if (ptg_index != static_cast<int>(ctx->numPaths.size())) return fail(ctx, "mppb_sim: PTGs must be set in order");
    ctx->ptgEpoch++;
    ctx->numPaths.push_back(0);  // no columns in the collision-grid rows
    ctx->holoPtgs++;
    return MPPB_OK;
}
int mppb_num_ptgs(const mppb_ctx* ctx) { return static_cast<int>(ctx->numPaths.size()); }
int mppb_total_paths(const mppb_ctx* ctx)
{
    int s = 0;
    for (int k : ctx->numPaths) s += k;
    return s;
}
int mppb_ptg_column_offset(const mppb_ctx* ctx, int ptg_index)
{
    int s = 0;
    for (int i = 0; i < ptg_index && i < static_cast<int>(ctx->numPaths.size()); i++) s += ctx->numPaths[i];
    return s;
}

int mppb_eval_nodes_boxed(mppb_ctx* ctx, size_t n, const double* poses, const float* boxes, double clip, float* out)
{
    ctx->calls[0]++;
    return eval_nodes(ctx, n, poses, boxes, clip, out);
}
int mppb_eval_nodes_boxed_begin(mppb_ctx* ctx, size_t n, const double* poses, const float* boxes, double clip,
                                float* out)
{
    ctx->calls[1]++;
    return eval_nodes(ctx, n, poses, boxes, clip, out);  // complete at once; mppb_sync is then a no-op
}
static int eval_holo(mppb_ctx* ctx, size_t n_nodes, const double* poses, const float* boxes, size_t n_cands,
                     const mppb_holo_cand* cands, double clip, double* out)
{
    if (!g_oracleHolo) return fail(ctx, "mppb_sim: register the oracle HolonomicBlend and its table first");
    std::vector<float> lx(ctx->xs.size() + 1), ly(ctx->xs.size() + 1), bx, by;
    long               nLocal = 0;
    int                localOf = -1;
    for (size_t c = 0; c < n_cands; c++)
    {
        const mppb_holo_cand& cd = cands[c];
        if (cd.node < 0 || static_cast<size_t>(cd.node) >= n_nodes) return fail(ctx, "mppb_sim: bad candidate node");
        const HoloKey* key = nullptr;
        for (const HoloKey& h : g_holoTable)
            if (h.vxf == cd.vxf && h.vyf == cd.vyf && h.vf == cd.vf) key = &h;
        if (!key) return fail(ctx, "mppb_sim: candidate parameters are not in the registered (k, speed) table");
        if (cd.node != localOf)  // candidates come grouped by node
        {
            const float *px = ctx->xs.data(), *py = ctx->ys.data();
            size_t       np = ctx->xs.size();
            if (boxes)
            {
                const float* b = boxes + 4 * cd.node;
                bx.clear();
                by.clear();
                for (size_t p = 0; p < np; p++)
                    if (!(px[p] < b[0] || px[p] > b[2] || py[p] < b[1] || py[p] > b[3]))
                    {
                        bx.push_back(px[p]);
                        by.push_back(py[p]);
                    }
                px = bx.data(), py = by.data(), np = bx.size();
            }
            nLocal = orc_local_obstacles(px, py, np, poses + 3 * cd.node, clip, lx.data(), ly.data());
            if (nLocal < 0) return fail(ctx, "mppb_sim: orc_local_obstacles failed");
            localOf = cd.node;
        }
        // the oracle object computes its parameters from (k, trimmed speed, initial velocity): put it into the
        // candidate's (the |V|, |omega| formulas of the tested configurations read nothing else)
        const double dyn[7] = {cd.vxi, cd.vyi, 0.0, 1.0, 0.0, 0.0, 0.0};  // (some target: it is not read)
        orc_ptg_set_trim_speed(g_oracleHolo, key->speed);
        if (orc_ptg_update_dyn_state(g_oracleHolo, dyn, 1) != 0) return fail(ctx, "mppb_sim: oracle dyn state failed");
        orc_ptg_set_trim_speed(g_oracleHolo, key->speed);
        double raw = std::numeric_limits<double>::infinity();
        for (long i = 0; i < nLocal; i++) raw = orc_ptg_update_tpobs_single(g_oracleHolo, lx[i], ly[i], key->k, raw);
        out[c] = raw;
    }
    return MPPB_OK;
}
int mppb_eval_holo_candidates_boxed(mppb_ctx* ctx, size_t n_nodes, const double* poses, const float* boxes, size_t n_cands,
                                    const mppb_holo_cand* cands, double clip, double* out)
{
    ctx->calls[0]++;
    return eval_holo(ctx, n_nodes, poses, boxes, n_cands, cands, clip, out);
}
int mppb_eval_holo_candidates_boxed_begin(mppb_ctx* ctx, size_t n_nodes, const double* poses, const float* boxes,
                                          size_t n_cands, const mppb_holo_cand* cands, double clip, double* out)
{
    ctx->calls[1]++;
    return eval_holo(ctx, n_nodes, poses, boxes, n_cands, cands, clip, out);
}

int mppb_clear_evaluators(mppb_ctx* ctx)
{
    ctx->numEvals = 0;
    ctx->evalEpoch++;
    return MPPB_OK;
}
int mppb_set_costmap(mppb_ctx* ctx, int slot, const mppb_costmap_desc*)
{
    ctx->numEvals = std::max(ctx->numEvals, slot + 1);
    ctx->evalEpoch++;
    return MPPB_OK;
}
int mppb_set_waypoints(mppb_ctx* ctx, int slot, const double*, const double*, size_t, double, double, int)
{
    ctx->numEvals = std::max(ctx->numEvals, slot + 1);
    ctx->evalEpoch++;
    return MPPB_OK;
}
int      mppb_num_evaluators(const mppb_ctx* ctx) { return ctx->numEvals; }
uint64_t mppb_evaluators_epoch(const mppb_ctx* ctx) { return ctx->evalEpoch; }
int      mppb_eval_edge_costs(mppb_ctx* ctx, size_t n_edges, const double* from, const uint32_t* off, const double* rel,
                              double* out)
{
    ctx->calls[2]++;
    const int           E = ctx->numEvals;
    std::vector<double> col(n_edges);
    for (int s = 0; s < E; s++)
    {
        if (!g_oracleEval[s]) return fail(ctx, "mppb_sim: register an oracle evaluator for every slot");
        if (orc_eval_edge_costs(g_oracleEval[s], g_oracleEvalKind[s], n_edges, from, off, rel, col.data()) != 0)
            return fail(ctx, "mppb_sim: orc_eval_edge_costs failed");
        for (size_t e = 0; e < n_edges; e++) out[e * E + s] = col[e];
    }
    return MPPB_OK;
}
// Only used by the product's host costmap construction, whose cells this stand-in never reads (costs come from the
// registered oracle evaluators): a plain nearest search, exact below the clearance like the kernel's.
int mppb_costmap_nearest_d2(mppb_ctx*, const float* xs, const float* ys, size_t n, double x_min, double y_min, double res,
                            int nx, int ny, float, float* out)
{
    for (int cy = 0; cy < ny; cy++)
        for (int cx = 0; cx < nx; cx++)
        {
            const float x = static_cast<float>(x_min + (cx + 0.5) * res), y = static_cast<float>(y_min + (cy + 0.5) * res);
            float       best = std::numeric_limits<float>::infinity();
            for (size_t i = 0; i < n; i++)
            {
                const float dx = x - xs[i], dy = y - ys[i], d2 = dx * dx + dy * dy;
                best = d2 < best ? d2 : best;
            }
            out[cx + static_cast<size_t>(cy) * nx] = best;
        }
    return MPPB_OK;
}

// ---- device-side expansion (include/mppb.h, "one A* expansion per node"), answered on the CPU in the REFERENCE's own
// sequential form: std::set of trajectory indices, TPS points in a vector, the goal-cell set, a map of best paths with
// the order of first insertion, std::map of time-keyed samples (TPS_Astar.cpp:538-826, :269-345,
// edge_interpolated_path.cpp:51-72). Collision rows and evaluator values come from the oracle.
int mppb_set_ptg_paths(mppb_ctx* ctx, int ptg_index, const mppb_ptg_paths_desc* d)
{
    if (ptg_index < 0 || ptg_index >= static_cast<int>(ctx->numPaths.size())) return fail(ctx, "mppb_sim: bad ptg_index");
    mppb_ctx::Paths& P = ctx->paths[ptg_index];
    const size_t     K = static_cast<size_t>(d->num_paths), n = d->path_begin[K];
    P.K                = d->num_paths;
    P.dt               = d->step_duration;
    P.refDistance      = d->ref_distance;
    P.begin.assign(d->path_begin, d->path_begin + K + 1);
    P.x.assign(d->x, d->x + n);
    P.y.assign(d->y, d->y + n);
    P.phi.assign(d->phi, d->phi + n);
    P.dist.assign(d->dist, d->dist + n);
    P.v.assign(d->v, d->v + n);
    P.w.assign(d->w, d->w + n);
    P.cs.assign(d->cos_phi, d->cos_phi + n);
    P.sn.assign(d->sin_phi, d->sin_phi + n);
    P.set = true;
    return MPPB_OK;
}
int mppb_set_expand_params(mppb_ctx* ctx, const mppb_astar_params* a)
{
    ctx->resXY  = a->grid_resolution_xy;
    ctx->resYaw = a->grid_resolution_yaw;
    ctx->M      = a->max_ptg_trajectories_to_explore;
    ctx->S      = a->max_ptg_speeds_to_explore;
    ctx->nSeg   = a->path_interpolated_segments;
    ctx->timestamps.assign(a->ptg_sample_timestamps, a->ptg_sample_timestamps + a->num_timestamps);
    ctx->paramsSet = true;
    return MPPB_OK;
}
static bool sim_expand_available(const mppb_ctx* ctx)
{
    if (!ctx->paramsSet || ctx->numPaths.empty() || ctx->holoPtgs > 0) return false;
    if (std::getenv("MPPBSIM_NO_EXPAND")) return false;
    for (size_t p = 0; p < ctx->numPaths.size(); p++)
        if (!ctx->paths[p].set) return false;
    return true;
}
uint32_t mppb_expand_max_neighbors(const mppb_ctx* ctx)
{
    if (!sim_expand_available(ctx)) return 0;
    uint32_t n = 0;
    for (size_t p = 0; p < ctx->numPaths.size(); p++) n += 1 + (ctx->M + 1) * static_cast<uint32_t>(ctx->timestamps.size());
    return n;
}

namespace
{
double sim_wrap_2pi(double a)
{
    const bool neg = a < 0;
    a              = std::fmod(a, 2.0 * M_PI);
    return neg ? a + 2.0 * M_PI : a;
}
double sim_wrap_pi(double a) { return sim_wrap_2pi(a + M_PI) - M_PI; }
using Cell = std::tuple<int32_t, int32_t, int32_t>;
}  // namespace

int mppb_expand_nodes(mppb_ctx* ctx, size_t n_nodes, const mppb_expand_node* nodes, int use_node_boxes, double clip,
                      mppb_neighbor* out, uint32_t* out_counts)
{
    if (!sim_expand_available(ctx)) return fail(ctx, "mppb_sim: expansion not available");
    ctx->expandCalls++;
    ctx->expandNodes += n_nodes;
    const uint32_t stride = mppb_expand_max_neighbors(ctx);
    size_t         cols   = 0;
    for (int k : ctx->numPaths) cols += static_cast<size_t>(k);
    ctx->lastRows.assign(n_nodes * cols, 0.f);
    const int nP = static_cast<int>(ctx->numPaths.size());
    for (size_t i = 0; i < n_nodes; i++)
    {
        const mppb_expand_node& N = nodes[i];
        float* row = ctx->lastRows.data() + i * cols;
        float  box[4] = {static_cast<float>(N.bbox_min[0]), static_cast<float>(N.bbox_min[1]),
                         static_cast<float>(N.bbox_max[0]), static_cast<float>(N.bbox_max[1])};
        if (eval_nodes(ctx, 1, N.pose, use_node_boxes ? box : nullptr, clip, row) != MPPB_OK) return MPPB_ERR_STATE;
        const double fromC = std::cos(N.pose[2]), fromS = std::sin(N.pose[2]);
        const double halfCell = ctx->resXY * 0.5;
        struct Best
        {
            int      ptg = -1, k = -1;
            uint32_t step = 0;
            double   dist = std::numeric_limits<double>::max();
            double   rel[3] = {0, 0, 0};
        };
        std::map<Cell, Best> bestPaths;
        std::vector<Cell>    order;  // cells in the order of their first insertion into bestPaths
        uint32_t             tps = 0;
        size_t               colBase = 0;
        for (int p = 0; p < nP; p++)
        {
            const mppb_ctx::Paths& T = ctx->paths[p];
            std::set<int>          trajIdxs;
            for (size_t m = 0; m < ctx->M; m++)
            {
                const size_t q = m * static_cast<size_t>(T.K - 1) / (ctx->M - 1);
                trajIdxs.insert(static_cast<int>(std::lrint(static_cast<double>(q))));
            }
            unsigned nSpeeds = 0;
            for (double sp = 1.0 / ctx->S; sp < 1.001; sp += 1.0 / ctx->S) nSpeeds++;
            struct Tps
            {
                int      k;
                uint32_t step;
                bool     target;
            };
            std::vector<Tps> pts;
            if (N.goal_k[p] >= 0)
            {
                if (N.goal_step[p] != 0xffffffffu) pts.push_back({N.goal_k[p], N.goal_step[p], true});
                trajIdxs.insert(N.goal_k[p]);
            }
            // (the per-speed copies of a DiffDrive_C TPS point are identical: one is evaluated, nSpeeds are counted)
            for (int k : trajIdxs)
                for (double t : ctx->timestamps)
                {
                    const uint32_t step = static_cast<uint32_t>(static_cast<int>(std::lrint(t / T.dt)));
                    if (step >= T.begin[k + 1] - T.begin[k]) continue;
                    pts.push_back({k, step, false});
                }
            std::set<Cell> goalNodeCoords;
            for (const Tps& tp : pts)
            {
                if (tp.step == 0) continue;
                const size_t idx = static_cast<size_t>(T.begin[tp.k]) + tp.step;
                const double relTrgDist = T.dist[idx];
                const double rx = T.x[idx], ry = T.y[idx], rphi = T.phi[idx];
                const double ax = N.pose[0] + rx * fromC - ry * fromS, ay = N.pose[1] + rx * fromS + ry * fromC;
                const double aphi = sim_wrap_pi(N.pose[2] + rphi);
                if (ax < N.bbox_min[0] || ay < N.bbox_min[1] || aphi < N.bbox_min[2]) continue;
                if (ax > N.bbox_max[0] - halfCell || ay > N.bbox_max[1] - halfCell || aphi > N.bbox_max[2]) continue;
                tps += nSpeeds;
                const double init = std::min(T.refDistance, static_cast<double>(T.dist[T.begin[tp.k + 1] - 1]));
                const double freeDistance = std::min(init, static_cast<double>(row[colBase + tp.k]));
                if (relTrgDist >= freeDistance) continue;
                // TPS_Astar.h:224-230
                const float fx = static_cast<float>(ax), fy = static_cast<float>(ay), fyaw = static_cast<float>(aphi);
                const float pi = static_cast<float>(M_PI), twoPi = static_cast<float>(2.0 * M_PI);
                float       a   = fyaw + pi;
                const bool  neg = a < 0;
                a               = std::fmod(a, twoPi);
                if (neg) a += twoPi;
                const Cell nc(static_cast<int32_t>(fx / ctx->resXY), static_cast<int32_t>(fy / ctx->resXY),
                              static_cast<int32_t>((a - pi) / ctx->resYaw));
                if (tp.target)
                    goalNodeCoords.insert(nc);
                else if (goalNodeCoords.count(nc))
                    continue;
                auto it = bestPaths.find(nc);
                if (it == bestPaths.end())
                {
                    it = bestPaths.emplace(nc, Best()).first;
                    order.push_back(nc);
                }
                if (relTrgDist < it->second.dist)
                {
                    Best& b = it->second;
                    b.dist = relTrgDist, b.ptg = p, b.k = tp.k, b.step = tp.step;
                    b.rel[0] = rx, b.rel[1] = ry, b.rel[2] = rphi;
                }
            }
            colBase += static_cast<size_t>(T.K);
        }
        if (order.size() > stride) return fail(ctx, "mppb_sim: more neighbours than the row holds");
        mppb_neighbor* o = out + i * stride;
        for (size_t j = 0; j < order.size(); j++)
        {
            const Best&            b = bestPaths.at(order[j]);
            const mppb_ctx::Paths& T = ctx->paths[b.ptg];
            const size_t           base = T.begin[b.k], idx = base + b.step;
            mppb_neighbor          r{};
            r.cell[0] = std::get<0>(order[j]), r.cell[1] = std::get<1>(order[j]), r.cell[2] = std::get<2>(order[j]);
            r.ptg_index = static_cast<uint8_t>(b.ptg);
            r.k         = static_cast<uint16_t>(b.k);
            r.step      = b.step;
            r.ptg_dist  = static_cast<float>(b.dist);
            r.pose[0]   = N.pose[0] + b.rel[0] * fromC - b.rel[1] * fromS;
            r.pose[1]   = N.pose[1] + b.rel[0] * fromS + b.rel[1] * fromC;
            r.pose[2]   = sim_wrap_pi(N.pose[2] + b.rel[2]);
            const double v = T.v[idx], w = T.w[idx];
            const double lx = v * T.cs[idx] - 0.0 * T.sn[idx], ly = v * T.sn[idx] + 0.0 * T.cs[idx];
            r.vel[0] = lx * fromC - ly * fromS;
            r.vel[1] = lx * fromS + ly * fromC;
            r.vel[2] = w;
            r.in_bbox = !(r.pose[0] < N.bbox_min[0] || r.pose[1] < N.bbox_min[1] || r.pose[2] < N.bbox_min[2] ||
                          r.pose[0] > N.bbox_max[0] || r.pose[1] > N.bbox_max[1] || r.pose[2] > N.bbox_max[2]);
            // edge_interpolated_path: a std::map keyed by time, later assignments overwrite
            std::map<double, std::pair<double, double>> ip;
            ip[0 * T.dt] = {0.0, 0.0};
            for (size_t s2 = 0; s2 < ctx->nSeg; s2++)
            {
                const size_t iStep = ((s2 + 1) * static_cast<size_t>(b.step)) / (ctx->nSeg + 2);
                ip[iStep * T.dt]   = {T.x[base + iStep], T.y[base + iStep]};
            }
            ip[static_cast<size_t>(b.step) * T.dt] = {b.rel[0], b.rel[1]};
            std::vector<double> rel;
            for (const auto& kv : ip)
            {
                rel.push_back(kv.second.first);
                rel.push_back(kv.second.second);
            }
            const uint32_t off[2] = {0, static_cast<uint32_t>(ip.size())};
            double         cost   = static_cast<size_t>(b.step) * T.dt;
            for (int e = 0; e < ctx->numEvals; e++)
            {
                if (!g_oracleEval[e]) return fail(ctx, "mppb_sim: register an oracle evaluator for every slot");
                double c1 = 0;
                if (orc_eval_edge_costs(g_oracleEval[e], g_oracleEvalKind[e], 1, N.pose, off, rel.data(), &c1) != 0)
                    return fail(ctx, "mppb_sim: orc_eval_edge_costs failed");
                cost += c1;
            }
            r.cost = cost;
            o[j]   = r;
        }
        out_counts[2 * i + 0] = static_cast<uint32_t>(order.size());
        out_counts[2 * i + 1] = tps;
    }
    return MPPB_OK;
}
int mppb_expand_nodes_begin(mppb_ctx* ctx, size_t n_nodes, const mppb_expand_node* nodes, int use_node_boxes, double clip,
                            mppb_neighbor* out, uint32_t* out_counts)
{
    return mppb_expand_nodes(ctx, n_nodes, nodes, use_node_boxes, clip, out, out_counts);
}
int mppb_expand_last_rows(mppb_ctx* ctx, size_t n_nodes, float* out_rows)
{
    size_t cols = 0;
    for (int k : ctx->numPaths) cols += static_cast<size_t>(k);
    std::memcpy(out_rows, ctx->lastRows.data(), n_nodes * cols * sizeof(float));
    return MPPB_OK;
}
void mppbsim_expand_counts(mppb_ctx* ctx, uint64_t* out2)
{
    out2[0] = ctx->expandCalls;
    out2[1] = ctx->expandNodes;
}

}  // extern "C"
