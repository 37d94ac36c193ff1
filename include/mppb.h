/*
 * mppb.h — C ABI of libmpp_b200.so: the B200-native replacement of the hot path of
 * MRPT/mrpt_path_planning (batched TP-space collision + cost evaluation of candidate
 * motion-primitive edges). Plain C: pointers and sizes only, no C++/torch types.
 *
 * Reference interfaces replaced (paths relative to mrpt_path_planning/ in the reference):
 *   mppb_set_obstacles*      <- the obstacle clouds handed to
 *                               TPS_Astar::find_feasible_paths_to_neighbors
 *                               (src/algos/TPS_Astar.cpp:124-137, 542) and consumed by
 *                               transform_pc_square_clipping (include/mpp/algos/
 *                               transform_pc_square_clipping.h:16-19)
 *   mppb_set_ptg_grid        <- DiffDriveCollisionGridBased collision grid + robot shape
 *                               (include/mpp/ptgs/DiffDriveCollisionGridBased.h:162-215)
 *   mppb_build_collision_grid <- DiffDriveCollisionGridBased::internal_initialize, the table
 *                               construction (src/ptgs/DiffDriveCollisionGridBased.cpp:673-754)
 *   mppb_eval_nodes*         <- cached_local_obstacles + tp_obstacles_single_path for every
 *                               (node, PTG, k) (src/algos/TPS_Astar.cpp:561-562,744-745;
 *                               include/mpp/algos/tp_obstacles_single_path.h:17-19)
 *   mppb_set_ptg_holo,
 *   mppb_eval_holo_candidates <- HolonomicBlend::updateTPObstacleSingle folded over the local
 *                               obstacles (src/ptgs/HolonomicBlend.cpp:719-840) for explicit
 *                               (node, k, speed) candidates
 *   mppb_set_costmap / mppb_set_waypoints / mppb_eval_edge_costs
 *                            <- CostEvaluatorCostMap::operator() (src/algos/
 *                               CostEvaluatorCostMap.cpp:122-162) and
 *                               CostEvaluatorPreferredWaypoint::operator() (src/algos/
 *                               CostEvaluatorPreferredWaypoint.cpp:59-120) over the interpolated
 *                               samples of an edge (src/algos/edge_interpolated_path.cpp:51-72)
 *   mppb_costmap_nearest_d2  <- the nearest-obstacle search of CostEvaluatorCostMap::
 *                               FromStaticPointObstacles (CostEvaluatorCostMap.cpp:91-97)
 *   mppb_set_obstacles_clipped <- ObstacleSource::apply_clipping_box (src/interfaces/
 *                               ObstacleSource.cpp:19-42)
 *   mppb_comm_* / mppb_set_obstacles_shard / mppb_eval_nodes_sharded_device
 *                            <- (no reference counterpart: the reference is single-threaded) cloud sharding over
 *                               several GPUs with a min-merge of the partial TP-obstacle rows
 *   mppb_set_ptg_paths / mppb_expand_nodes
 *                            <- the candidate set, end poses, blocked test, per-cell selection and edge costs of
 *                               TPS_Astar::find_feasible_paths_to_neighbors (src/algos/TPS_Astar.cpp:538-826) and
 *                               of the per-neighbour block of plan() (:269-345), over the sampled PTG paths
 *                               (src/ptgs/DiffDriveCollisionGridBased.cpp:766-811)
 *   mppb_planner_*           <- mpp::TPS_Astar::plan (src/algos/TPS_Astar.cpp:86-475) behind the
 *                               mpp::Planner interface (include/mpp/algos/Planner.h:23-62)
 *
 * Conventions
 *   - every function returns 0 (MPPB_OK) or a negative error code; mppb_last_error(ctx)
 *     returns the message of the last failure on that context.
 *   - a context is bound to ONE CUDA device and ONE stream and is not thread-safe; use one
 *     context per host thread (TPS_Astar::plan is single-threaded per planner object).
 *   - "host" entry points take host pointers, copy in/out and return after the result is
 *     in the caller's buffer; "_device" entry points take device pointers, only enqueue
 *     work on the context's stream and return immediately (call mppb_sync()).
 *   - there is NO CPU fallback: if no CUDA device is usable, mppb_ctx_create fails.
 */
#ifndef MPPB_H_
#define MPPB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPPB_OK 0
#define MPPB_ERR_INVALID_ARG (-1)
#define MPPB_ERR_CUDA (-2)
#define MPPB_ERR_STATE (-3)
#define MPPB_ERR_UNSUPPORTED (-4)

#define MPPB_MAX_PTGS 8
#define MPPB_MAX_SHAPE_VERTS 32
#define MPPB_MAX_EVALUATORS 4

typedef struct mppb_ctx mppb_ctx;

/* ---- library / context --------------------------------------------------------------- */
const char* mppb_version(void);
/* message of the last failure that had no context (e.g. mppb_ctx_create) */
const char* mppb_last_global_error(void);

/* device: CUDA ordinal. cuda_stream: a cudaStream_t to run on (e.g. the caller's
 * framework stream), or NULL to let the context create its own non-blocking stream. */
int         mppb_ctx_create(int device, void* cuda_stream, mppb_ctx** out_ctx);
void        mppb_ctx_destroy(mppb_ctx* ctx);
const char* mppb_last_error(const mppb_ctx* ctx);
int         mppb_sync(mppb_ctx* ctx);
/* the cudaStream_t the context enqueues on */
void*       mppb_stream(mppb_ctx* ctx);
/* the CUDA ordinal the context is bound to (-1 for a null context) */
int         mppb_ctx_device(const mppb_ctx* ctx);
/* number of kernels this context has launched so far (for launch accounting) */
uint64_t    mppb_launch_count(const mppb_ctx* ctx);

/* pinned host memory for zero-staging transfers (cudaHostAlloc / cudaFreeHost) */
int mppb_host_alloc(size_t bytes, void** out_ptr);
int mppb_host_free(void* ptr);

/* CUDA-event stopwatch on the context's stream: begin records, end records + waits and
 * returns the elapsed device time in milliseconds. */
int mppb_timer_begin(mppb_ctx* ctx);
int mppb_timer_end(mppb_ctx* ctx, float* out_ms);

/* Roofline denominators not in MEASURED_PEAKS.json: CUDA-core FMA throughput (2 FLOP per
 * FMA) measured on this device, best of 4 runs after one warm-up. */
int mppb_measure_fp32_tflops(mppb_ctx* ctx, double* out_tflops);
int mppb_measure_fp64_tflops(mppb_ctx* ctx, double* out_tflops);

/* ---- obstacle cloud -------------------------------------------------------------------
 * Float SoA cloud in the global frame, exactly the layout of mrpt::maps::CPointsMap
 * buffers read at transform_pc_square_clipping.cpp:14-16. The call copies the points to
 * HBM and builds a uniform-bin index over them (bin_size metres; <=0 selects the default).
 * Several clouds (obstacle sources) are concatenated by calling with append != 0.
 * Non-finite points are dropped (they can never produce a TP-obstacle). */
int mppb_set_obstacles(mppb_ctx* ctx, const float* xs, const float* ys, size_t n, double bin_size, int append);
int mppb_set_obstacles_device(mppb_ctx* ctx, const float* d_xs, const float* d_ys, size_t n, double bin_size,
                              int append);
/* Same, keeping only the points inside box = {minx, miny, maxx, maxy} (float comparisons,
 * inclusive bounds: ObstacleSource.cpp:29-40). The box applies to this and to appended sources
 * until the cloud is replaced. box == NULL means no clipping. */
int mppb_set_obstacles_clipped(mppb_ctx* ctx, const float* xs, const float* ys, size_t n, double bin_size,
                               int append, const float* box);
/* number of finite (and in-box) points currently indexed */
size_t mppb_num_obstacles(const mppb_ctx* ctx);
/* counter bumped by every mppb_set_obstacles* call on this context: lets a caller that keeps the cloud resident
 * across plans (the reference re-reads ObstacleSource::obstacles() on every plan(), TPS_Astar.cpp:125-137) notice
 * that somebody else replaced it */
uint64_t mppb_obstacles_epoch(const mppb_ctx* ctx);

/* ---- PTG tables: collision-grid family ------------------------------------------------
 * One descriptor per PTG of the DiffDriveCollisionGridBased family. All arrays are host
 * pointers and are copied. cell (cx,cy) owns entries [cell_offsets[cx+cy*nx],
 * cell_offsets[cx+cy*nx+1]) of (entry_k, entry_d): the (k, min TP-distance) list of
 * DiffDriveCollisionGridBased::TCollisionCell (.h:162). */
typedef struct mppb_ptg_grid_desc
{
    int32_t         num_paths;    /* K */
    double          ref_distance; /* PTG refDistance */
    double          grid_x_min, grid_y_min, grid_res;
    int32_t         grid_nx, grid_ny;
    const uint32_t* cell_offsets; /* [nx*ny+1] */
    const uint16_t* entry_k;      /* [cell_offsets[nx*ny]] */
    const float*    entry_d;      /* [cell_offsets[nx*ny]] */
    int32_t         shape_nverts; /* polygonal robot shape (<= MPPB_MAX_SHAPE_VERTS) */
    const double*   shape_x;
    const double*   shape_y;
    double          max_radius;   /* getMaxRobotRadius() */
} mppb_ptg_grid_desc;

/* Construction of that table on the device: the recompute branch of
 * DiffDriveCollisionGridBased::internal_initialize (src/ptgs/DiffDriveCollisionGridBased.cpp:673-754)
 * with CCollisionGrid::updateCellInfo's min rule (:233-258). Inputs are the simulated paths
 * (simulateTrajectories, :103-176): path k owns steps [path_begin[k], path_begin[k+1]) of the float
 * arrays x, y, phi, dist (TCPoint::x,y,phi,dist). The grid is CDynamicGrid::setSize(-ref_distance,
 * ref_distance, -ref_distance, ref_distance, resolution). cos/sin of the headings are taken with
 * the host C library inside this call; the device does the polygon-in-cell tests in IEEE double in
 * the reference's operation order, so the result is bit-identical to the reference's table, with
 * each cell's entries in ascending k (the order the reference produces). The arrays of `out` are
 * owned by the context and stay valid until the next call or mppb_ctx_destroy. */
typedef struct mppb_colgrid_build_desc
{
    int32_t         num_paths;
    const uint32_t* path_begin; /* [num_paths+1] */
    const float*    x;
    const float*    y;
    const float*    phi;
    const float*    dist;
    double          ref_distance, resolution;
    int32_t         shape_nverts;
    const double*   shape_x;
    const double*   shape_y;
} mppb_colgrid_build_desc;
typedef struct mppb_colgrid
{
    double          x_min, y_min, res;
    int32_t         nx, ny;
    uint64_t        n_entries;
    const uint32_t* cell_offsets; /* [nx*ny+1] */
    const uint16_t* entry_k;      /* [n_entries] */
    const float*    entry_d;      /* [n_entries] */
} mppb_colgrid;
int mppb_build_collision_grid(mppb_ctx* ctx, const mppb_colgrid_build_desc* desc, mppb_colgrid* out);
/* timings of the last build: out4 = host cos/sin s, uploads+kernels+result s, kernels s (device clock), steps */
int mppb_colgrid_build_stats(const mppb_ctx* ctx, double* out4);

int mppb_clear_ptgs(mppb_ctx* ctx);
/* ptg_index must be the next free index (0,1,2,...) over PTGs of all kinds */
int mppb_set_ptg_grid(mppb_ctx* ctx, int ptg_index, const mppb_ptg_grid_desc* desc);
int mppb_num_ptgs(const mppb_ctx* ctx);
/* counter bumped by mppb_clear_ptgs and every mppb_set_ptg_* call (same purpose as mppb_obstacles_epoch) */
uint64_t mppb_ptgs_epoch(const mppb_ctx* ctx);
/* sum over the collision-grid PTGs of num_paths = row length of the mppb_eval_nodes output */
int mppb_total_paths(const mppb_ctx* ctx);
/* first output column of collision-grid PTG ptg_index, or -1 if it is of another kind */
int mppb_ptg_column_offset(const mppb_ctx* ctx, int ptg_index);

/* ---- PTG constants: HolonomicBlend ------------------------------------------------------
 * Closed-form PTG: nothing is tabulated. The device needs the circular robot radius and V_MAX;
 * everything that depends on (k, speed, dynamic state) travels with each candidate. */
typedef struct mppb_ptg_holo_desc
{
    int32_t num_paths;
    double  ref_distance;
    double  robot_radius;
    double  v_max;
} mppb_ptg_holo_desc;
int mppb_set_ptg_holo(mppb_ctx* ctx, int ptg_index, const mppb_ptg_holo_desc* desc);

/* One (node, trajectory, speed) candidate of a HolonomicBlend PTG: InternalParams of
 * HolonomicBlend::internal_params_from_dir_and_dynstate (HolonomicBlend.cpp:942-973), computed by
 * the caller (mppb_ptg_holo_params does it for a host PTG object). */
typedef struct mppb_holo_cand
{
    int32_t node;      /* index into the node batch */
    int32_t ptg_index; /* a PTG set with mppb_set_ptg_holo */
    double  T_ramp, vxi, vyi, vxf, vyf, vf;
} mppb_holo_cand;

/* ---- stage (a)+(b): free TP-distance per (node, PTG, k) --------------------------------
 * For every node pose and every path k of every PTG: the minimum, over all obstacle
 * points inside the node's +-clip_dist square (global axes), of the TP-space collision
 * distance, as tp_obstacles_single_path() computes it BEFORE the initial value is
 * applied:   out[node][ptg_offset + k] = min_p c(p,k)   or +INF if no point constrains k.
 * The reference value is  freeDistance = min(init_k, (double)out)  with
 * init_k = min(refDistance, pathDist(k, last)) [initTPObstacleSingle]; the caller holds
 * init_k. out is row-major [n_nodes][mppb_total_paths]. */
int mppb_eval_nodes(mppb_ctx* ctx, size_t n_nodes, const double* poses_xyphi /* host [n][3] */,
                    double clip_dist, float* out /* host [n][total_paths] */);
/* Device-resident variant: d_nodes is [n][4] doubles (x, y, cos(phi), sin(phi)) with the
 * cos/sin evaluated by the caller's libm (glibc for bit parity with the reference). */
int mppb_eval_nodes_device(mppb_ctx* ctx, size_t n_nodes, const double* d_nodes_xycs, double clip_dist,
                           float* d_out);
/* Same with a world box per node: node_boxes is NULL or [n][4] floats {minx, miny, maxx, maxy};
 * obstacle points outside node i's box are ignored for node i (float comparisons, inclusive
 * bounds). This is ObstacleSource::apply_clipping_box (ObstacleSource.cpp:29-40) applied per
 * query, so that independent queries with different world boxes share one resident cloud. */
int mppb_eval_nodes_boxed(mppb_ctx* ctx, size_t n_nodes, const double* poses_xyphi, const float* node_boxes,
                          double clip_dist, float* out);
int mppb_eval_nodes_boxed_device(mppb_ctx* ctx, size_t n_nodes, const double* d_nodes_xycs, const float* d_node_boxes,
                                 double clip_dist, float* d_out);
/* Split form of the host call, for callers that have host work of their own to do while the GPU
 * evaluates (TPS_Astar reconstructs the end poses of its candidate TPS points, TPS_Astar.cpp:725-736,
 * which need the node pose only): _begin enqueues the same work as mppb_eval_nodes_boxed and returns;
 * mppb_sync(ctx) completes it. Until then poses_xyphi, node_boxes and out must stay alive and
 * untouched, and no second _begin of the same entry point may be issued on the context. */
int mppb_eval_nodes_boxed_begin(mppb_ctx* ctx, size_t n_nodes, const double* poses_xyphi, const float* node_boxes,
                                double clip_dist, float* out);
/* Fill [n][4] (x, y, cos, sin) from [n][3] (x, y, phi) on the host with the C library. */
void mppb_nodes_xycs_from_poses(size_t n_nodes, const double* poses_xyphi, double* out_xycs);

/* Diagnostic switch: on == 0 runs the HolonomicBlend kernel without its geometric pre-filter (every in-window point
 * goes through the exact FP64 chain); on == 2 keeps the per-point pre-filter but walks the whole clipping window
 * instead of only the cloud bins the candidate's path can reach. Results are identical in all three modes; the tests
 * compare them. Default: 1 (pre-filter and bin culling). */
int mppb_set_holo_filter(mppb_ctx* ctx, int on);
/* Near-first evaluation of node batches against collision-grid PTGs (more than 256 nodes per call). Every entry (k, d)
 * of a collision-grid cell has d >= distance(robot origin, cell) - (largest robot radius), which is measured on the
 * tables when they are set; so once every path of a node is blocked within `bound` by obstacle points of the cloud
 * bins around the node, no farther point can change the row, and the rest of the clipping window is not looked at.
 * Nodes that are not decided that way are evaluated on the whole window by a second launch behind the first. Results
 * are identical in all modes (tested). mode 0: off; 1 (default): on, and switched off for the next 256 calls whenever
 * more than 2 % of a batch needed the second launch (maps with free space around the nodes); 2: always on.
 * mppb_grid_near_stats: out3 = {calls that took the two-launch form, their nodes, nodes that needed the second launch}. */
int mppb_set_grid_near(mppb_ctx* ctx, int mode);
int mppb_grid_near_stats(const mppb_ctx* ctx, uint64_t* out3);
/* The near window's half-size and the bound (both in metres) a batch with this clipping distance would use with the
 * tables and cloud now resident; both 0 when the near-first form does not apply. */
int mppb_grid_near_window(const mppb_ctx* ctx, double clip_dist, double* half_size, double* bound);
/* Diagnostic switch: on != 0 runs host batches of mppb_eval_holo_candidates* whose candidates are grouped by node on
 * the one-CTA-per-node form of the kernel (the node's window is walked once for all its candidates). Results are
 * identical either way; the tests compare the two. Measured no faster on B200, so the default is off. */
int mppb_set_holo_by_node(mppb_ctx* ctx, int on);

/* HolonomicBlend: out[c] = min over the in-window points of node cands[c].node of the
 * TP-space collision distance of candidate c (+INF if unconstrained), in double.
 * freeDistance = min(init, out[c]) with init from mppb_ptg_init_dist (the caller holds it). */
int mppb_eval_holo_candidates(mppb_ctx* ctx, size_t n_nodes, const double* poses_xyphi /* host [n][3] */,
                              size_t n_cands, const mppb_holo_cand* cands /* host */, double clip_dist,
                              double* out /* host [n_cands] */);
int mppb_eval_holo_candidates_device(mppb_ctx* ctx, size_t n_cands, const mppb_holo_cand* d_cands,
                                     const double* d_nodes_xycs, double clip_dist, double* d_out);
/* per-node world boxes as in mppb_eval_nodes_boxed */
int mppb_eval_holo_candidates_boxed(mppb_ctx* ctx, size_t n_nodes, const double* poses_xyphi, const float* node_boxes,
                                    size_t n_cands, const mppb_holo_cand* cands, double clip_dist, double* out);
int mppb_eval_holo_candidates_boxed_device(mppb_ctx* ctx, size_t n_cands, const mppb_holo_cand* d_cands,
                                           const double* d_nodes_xycs, const float* d_node_boxes, double clip_dist,
                                           double* d_out);
/* split form (see mppb_eval_nodes_boxed_begin); may be in flight together with one mppb_eval_nodes_boxed_begin */
int mppb_eval_holo_candidates_boxed_begin(mppb_ctx* ctx, size_t n_nodes, const double* poses_xyphi,
                                          const float* node_boxes, size_t n_cands, const mppb_holo_cand* cands,
                                          double clip_dist, double* out);

/* ---- free-function seams -------------------------------------------------------------------------------------
 * mpp::transform_pc_square_clipping (include/mpp/algos/transform_pc_square_clipping.h:16-19; src/algos/
 * transform_pc_square_clipping.cpp:9-43) for ONE pose, as a single call: the points of (xs, ys) inside the +-max_dist
 * square around the pose (global axes, FP64 compare), composed with the inverse pose in FP64 and stored as float, in
 * input order. Host pointers; out_x / out_y hold up to n points; *out_n = points written. (The batched kernels fuse
 * this step and never materialise the local cloud; this entry point exists for callers of the reference's function.)
 * The single-call form of mpp::tp_obstacles_single_path (.h:17-19) is mpp::tp_obstacles_single_path() of the C++
 * mirror (csrc/host/mpp.h), built on mppb_set_obstacles + mppb_eval_nodes / mppb_eval_holo_candidates. */
int mppb_transform_pc_square_clipping(mppb_ctx* ctx, const float* xs, const float* ys, size_t n, const double* pose_xyphi,
                                      double max_dist, float* out_x, float* out_y, size_t* out_n);

/* ---- several GPUs: cloud sharding and the min-merge of partial rows -------------------
 * A cloud too large for one GPU (or simply large) is split by point index over the ranks; every rank evaluates ALL
 * nodes against its shard, and because a path's free distance is an order-independent minimum over points, the
 * element-wise minimum of the partial rows is the single-GPU result bit for bit. One process per GPU; the caller
 * carries a few bytes between the ranks once (MPI, torch.distributed, files: its choice), after that everything
 * runs on the contexts' streams.
 *   merge by NCCL:  mppb_comm_unique_id (rank 0) -> mppb_comm_init (every rank) -> mppb_eval_nodes_sharded_device
 *   fused merge:    additionally mppb_peer_rows_create (every rank) -> mppb_peer_rows_attach (every rank, with all
 *                   ranks' handles): tpobs_grid_kernel then stores each finished row straight into every peer's
 *                   memory over NVLink while the other nodes are still being evaluated (a slot per source rank, no
 *                   atomics), a flag per rank closes the step, and a short kernel takes the minimum over the slots.
 * libnccl.so.2 is opened at run time (dlopen); nothing here is needed for single-GPU use. */
#define MPPB_COMM_ID_BYTES 128
#define MPPB_IPC_HANDLE_BYTES 64
#define MPPB_MAX_RANKS 16
/* ncclGetUniqueId: called on ONE rank; the bytes go to every rank's mppb_comm_init */
int mppb_comm_unique_id(uint8_t* out_id /* [MPPB_COMM_ID_BYTES] */);
/* ncclCommInitRank on the context's device (collective: every rank must call it) */
int mppb_comm_init(mppb_ctx* ctx, int world, int rank, const uint8_t* id /* [MPPB_COMM_ID_BYTES] */);
int mppb_comm_world(const mppb_ctx* ctx); /* 1 without a communicator */
int mppb_comm_rank(const mppb_ctx* ctx);
/* this rank's shard of the cloud: the points i with i % n_shards == shard (strided, so that every shard covers the
 * whole map and the ranks stay balanced wherever the nodes are); xs, ys: the WHOLE cloud, host pointers */
int mppb_set_obstacles_shard(mppb_ctx* ctx, const float* xs, const float* ys, size_t n, double bin_size, int shard,
                             int n_shards);
/* in-place element-wise minimum over the ranks of a device array of floats (ncclAllReduce, ncclMin) on the context's
 * stream; asynchronous */
int mppb_allreduce_min_device(mppb_ctx* ctx, float* d_rows, size_t count);
/* fused merge, set-up: allocates this rank's receive slots for batches of up to max_nodes nodes and returns the IPC
 * handle the other ranks open; then every rank attaches to all of them (handles in rank order, its own included) */
int mppb_peer_rows_create(mppb_ctx* ctx, size_t max_nodes, uint8_t* out_handle /* [MPPB_IPC_HANDLE_BYTES] */);
int mppb_peer_rows_attach(mppb_ctx* ctx, int world, int rank, const uint8_t* handles /* [world][MPPB_IPC_HANDLE_BYTES] */);
/* mppb_eval_nodes_device against this rank's shard + the merge: d_out [n_nodes][total_paths] holds the merged rows
 * on every rank when the stream reaches the end of the call. Uses the fused path if peers are attached, NCCL if only
 * a communicator is, and is a plain mppb_eval_nodes_device without either. */
int mppb_eval_nodes_sharded_device(mppb_ctx* ctx, size_t n_nodes, const double* d_nodes_xycs, double clip_dist, float* d_out);
/* 0 = plain, 1 = NCCL merge, 2 = fused peer-memory merge: what mppb_eval_nodes_sharded_device will do */
int mppb_sharded_mode(const mppb_ctx* ctx);

/* ---- one A* expansion per node on the device (collision-grid PTGs) ------------------------
 * Everything TPS_Astar::find_feasible_paths_to_neighbors (src/algos/TPS_Astar.cpp:538-826) and the per-neighbour
 * block of TPS_Astar::plan (:269-345) compute for an expanded node that does not depend on the search state:
 *   - the candidate TPS points in the reference's order (:614-697): per PTG the direct-to-goal point first, then
 *     for k in (k-subset + goal k, ascending) / for t in ptg_sample_timestamps, with trjStep = round(t / dt);
 *   - end pose, world-box test and lattice cell of each (:725-736; getPathPose/getPathDist are look-ups into the
 *     sampled paths, DiffDriveCollisionGridBased.cpp:766-811);
 *   - blocked iff relTrgDist >= freeDistance against the node's collision row (:749), computed by the same call;
 *   - goal-cell rule (:765-778) and the best candidate per reached lattice cell, strict < (:783-796);
 *   - per reached cell: neighbour pose, twist (getPathTwist rotated into the global frame, :290-297), the
 *     neighbour's own world-box test (:303-308), the interpolated samples (edge_interpolated_path.cpp:51-72) and
 *     cost_path_segment = estimatedExecTime + evaluators in slot order (Planner.cpp:15-29).
 * What stays with the caller is what depends on the search: the insertion into its per-expansion hash map (whose
 * iteration order assigns node ids), the visited test, relaxation, open set and tree. Results are returned per node
 * in the order in which the reference's map would see each cell for the first time.
 * Requires: every PTG of the context is of the collision-grid kind and has its sampled paths set. */
typedef struct mppb_ptg_paths_desc
{
    int32_t         num_paths;
    const uint32_t* path_begin;   /* [num_paths+1]: path k owns steps [path_begin[k], path_begin[k+1]) */
    const float*    x;            /* TCPoint::x, y, phi, dist, v, w per step (DiffDriveCollisionGridBased.h:18-28) */
    const float*    y;
    const float*    phi;
    const float*    dist;
    const float*    v;
    const float*    w;
    const double*   cos_phi;      /* cos / sin of (double)phi per step, taken with the host C library */
    const double*   sin_phi;
    double          step_duration; /* getPathStepDuration() */
    double          ref_distance;
} mppb_ptg_paths_desc;
/* sampled paths of PTG ptg_index, which must already be set as a collision-grid PTG (mppb_set_ptg_grid) */
int mppb_set_ptg_paths(mppb_ctx* ctx, int ptg_index, const mppb_ptg_paths_desc* desc);

#define MPPB_MAX_TIMESTAMPS 8
#define MPPB_MAX_EXPLORE_PATHS 64
struct mppb_astar_params;
/* planner parameters the expansion depends on: grid_resolution_xy/yaw, max_ptg_trajectories_to_explore,
 * ptg_sample_timestamps, path_interpolated_segments (the other fields are ignored) */
int mppb_set_expand_params(mppb_ctx* ctx, const struct mppb_astar_params* params);
/* upper bound of neighbours per node with the current PTGs and parameters = row length of mppb_expand_nodes' output;
 * 0 if the expansion is not available (PTG of another kind, paths or parameters missing, limits exceeded) */
uint32_t mppb_expand_max_neighbors(const mppb_ctx* ctx);

typedef struct mppb_expand_node
{
    double   pose[3];                    /* x y phi of the node */
    double   bbox_min[3], bbox_max[3];   /* world box of the node's query: x y phi */
    int32_t  goal_k[MPPB_MAX_PTGS];      /* per PTG: path straight to the goal if inverseMap_WS2TP is exact, else -1 */
    uint32_t goal_step[MPPB_MAX_PTGS];   /* its step if getPathStepForDist found one, else 0xffffffff */
    double   cos_phi, sin_phi;           /* cos / sin of pose[2] from the caller's C library, or both 0: the call takes them */
} mppb_expand_node;

typedef struct mppb_neighbor
{
    int32_t  cell[3];    /* lattice cell x, y, yaw (TPS_Astar.h:224-230) */
    uint8_t  ptg_index;
    uint8_t  in_bbox;    /* the neighbour pose passes the test of TPS_Astar.cpp:303-308 */
    uint16_t k;          /* chosen path */
    uint32_t step;       /* relTrgStep */
    float    ptg_dist;   /* relTrgDist: a float of the path table */
    double   pose[3];    /* x_i.pose */
    double   vel[3];     /* x_i.vel (global frame) */
    double   cost;       /* cost_path_segment of the edge */
} mppb_neighbor;

/* out: [n_nodes][mppb_expand_max_neighbors()] records, of which the first out_counts[2*i] are valid for node i;
 * out_counts[2*i+1] = TPS points evaluated for node i (with the reference's count of per-speed copies).
 * use_node_boxes != 0: obstacle points outside a node's world box are ignored for it, as in mppb_eval_nodes_boxed
 * (box = bbox_min/max narrowed to float). Host pointers; pinned buffers are used in place. */
int mppb_expand_nodes(mppb_ctx* ctx, size_t n_nodes, const mppb_expand_node* nodes, int use_node_boxes, double clip_dist,
                      mppb_neighbor* out, uint32_t* out_counts);
/* split form: enqueues and returns; mppb_sync() completes it (buffers must stay alive and untouched until then) */
int mppb_expand_nodes_begin(mppb_ctx* ctx, size_t n_nodes, const mppb_expand_node* nodes, int use_node_boxes,
                            double clip_dist, mppb_neighbor* out, uint32_t* out_counts);
/* the collision rows [n_nodes][mppb_total_paths] of the last mppb_expand_nodes call (device -> host copy; for tests) */
int mppb_expand_last_rows(mppb_ctx* ctx, size_t n_nodes, float* out_rows);

/* ---- stage (c): cost evaluators ---------------------------------------------------------
 * Evaluator slots are evaluated in slot order, like Planner::costEvaluators_ (Planner.cpp:22-26). */
typedef struct mppb_costmap_desc
{
    double        x_min, y_min, res; /* CDynamicGrid<double> geometry */
    int32_t       nx, ny;
    const double* cells;       /* host, row-major [ny][nx] */
    int32_t       use_average; /* CostEvaluatorCostMap::Parameters::useAverageOfPath */
} mppb_costmap_desc;
int mppb_clear_evaluators(mppb_ctx* ctx);
int mppb_set_costmap(mppb_ctx* ctx, int slot, const mppb_costmap_desc* desc);
/* waypoints are stored as float (CSimplePointsMap::insertPoint) */
int mppb_set_waypoints(mppb_ctx* ctx, int slot, const double* xs, const double* ys, size_t n, double influence_radius,
                       double cost_scale, int use_average);
int mppb_num_evaluators(const mppb_ctx* ctx);
/* counter bumped by every mppb_clear_evaluators / mppb_set_costmap / mppb_set_waypoints on this context: lets a caller
 * that keeps evaluators resident across plans notice that somebody else changed the slots */
uint64_t mppb_evaluators_epoch(const mppb_ctx* ctx);
/* Edge e starts at from_xyphi[e] and has the interpolated samples
 * [sample_offsets[e], sample_offsets[e+1]) of rel_xy (poses relative to the start, in time
 * order: MoveEdgeSE2_TPS::interpolatedPath). out[e][slot] = evaluator(edge), row length
 * mppb_num_evaluators(). The edge cost of Planner::cost_path_segment is
 * estimatedExecTime + out[e][0] + out[e][1] + ... */
int mppb_eval_edge_costs(mppb_ctx* ctx, size_t n_edges, const double* from_xyphi /* host [n][3] */,
                         const uint32_t* sample_offsets /* host [n+1] */, const double* rel_xy /* host [m][2] */,
                         double* out /* host [n][evaluators] */);
int mppb_eval_edge_costs_device(mppb_ctx* ctx, size_t n_edges, const double* d_from_xycs /* [n][4] */,
                                const uint32_t* d_sample_offsets, const double* d_rel_xy, double* d_out);

/* Nearest-obstacle search of the costmap builder: for every cell centre of the grid
 * (x = (float)(x_min + (cx+0.5) res), likewise y) the minimum float squared distance to the
 * cloud (xs, ys), exact whenever sqrt(d2) < clearance and +INF or any value >= clearance^2
 * otherwise (such cells have zero cost). out_d2 is host, row-major [ny][nx]. */
int mppb_costmap_nearest_d2(mppb_ctx* ctx, const float* xs, const float* ys, size_t n, double x_min, double y_min,
                            double res, int nx, int ny, float clearance, float* out_d2);


/* ---- host-side PTG objects ---------------------------------------------------------------
 * Mirror of the PTG "operator" the planner calls (mrpt::nav::CParameterizedTrajectoryGenerator
 * virtuals at src/algos/TPS_Astar.cpp:120,279-292,578-795). A PTG object owns the host
 * tables (trajectories, collision grid) and can upload itself into a context. */
typedef struct mppb_ptg mppb_ptg;

typedef struct mppb_diffdrive_c_params
{
    int32_t       num_paths;
    double        ref_distance;
    double        resolution;             /* collision grid cell size [m] */
    double        v_max_mps;
    double        w_max_rps;              /* rad/s (ini w_max_dps * pi / 180) */
    double        K;                      /* +1 forward, -1 backward */
    double        turning_radius_ref;     /* default 0.10 */
    int32_t       shape_nverts;
    const double* shape_x;
    const double* shape_y;
} mppb_diffdrive_c_params;

/* mpp::ptg::DiffDrive_C (src/ptgs/DiffDrive_C.cpp): simulates the trajectories and builds
 * the collision grid (DiffDriveCollisionGridBased.cpp:615-764). */
int  mppb_ptg_create_diffdrive_c(const mppb_diffdrive_c_params* params, mppb_ptg** out_ptg);
/* Same PTG with the collision grid built on the GPU of `ctx` (mppb_build_collision_grid) instead of
 * the multi-threaded host builder; identical tables. */
int  mppb_ptg_create_diffdrive_c_on(mppb_ctx* ctx, const mppb_diffdrive_c_params* params, mppb_ptg** out_ptg);
void mppb_ptg_destroy(mppb_ptg* ptg);
/* message of the last failure on a PTG object or PTG constructor */
const char* mppb_ptg_last_error(void);

int    mppb_ptg_num_paths(const mppb_ptg* ptg);
double mppb_ptg_ref_distance(const mppb_ptg* ptg);
double mppb_ptg_step_duration(const mppb_ptg* ptg);
double mppb_ptg_max_radius(const mppb_ptg* ptg);
/* getPathStepCount; <0 on error */
int64_t mppb_ptg_step_count(const mppb_ptg* ptg, int k);
/* getPathPose + getPathDist: out4 = x, y, phi, dist */
int mppb_ptg_path_pose(const mppb_ptg* ptg, int k, uint32_t step, double* out4);
/* getPathStepForDist: returns 1 found / 0 not found / <0 error */
int mppb_ptg_step_for_dist(const mppb_ptg* ptg, int k, double dist, uint32_t* out_step);
/* inverseMap_WS2TP: returns 1 exact / 0 not / <0 error; out_d is normalised by refDistance */
int mppb_ptg_inverse_map(const mppb_ptg* ptg, double x, double y, double tol, int* out_k, double* out_d);
/* initTPObstacleSingle value for path k */
double mppb_ptg_init_dist(const mppb_ptg* ptg, int k);
/* collision-grid tables (pointers stay valid while the PTG object lives); error if the PTG
 * is not of the collision-grid family */
int mppb_ptg_grid_export(const mppb_ptg* ptg, mppb_ptg_grid_desc* out_desc);
/* sampled paths of a collision-grid PTG object (pointers stay valid while the PTG object lives); the descriptor type
 * is declared with mppb_set_ptg_paths */
int mppb_ptg_paths_export(const mppb_ptg* ptg, mppb_ptg_paths_desc* out_desc);
/* upload this PTG's tables (collision grid and sampled paths) as the next PTG of the context */
int mppb_ctx_add_ptg(mppb_ctx* ctx, const mppb_ptg* ptg);

typedef struct mppb_holo_params
{
    int32_t     num_paths;
    double      ref_distance;
    double      T_ramp_max;
    double      v_max_mps;
    double      w_max_rps;
    double      robot_radius;
    const char* expr_V; /* NULL or "" = "V_MAX" */
    const char* expr_W; /* NULL or "" = "W_MAX" */
} mppb_holo_params;
/* mpp::ptg::HolonomicBlend (src/ptgs/HolonomicBlend.cpp) */
int mppb_ptg_create_holo(const mppb_holo_params* params, mppb_ptg** out_ptg);
/* updateNavDynamicState: dyn = vxLocal, vyLocal, omega, relTargetX, relTargetY, relTargetPhi, targetRelSpeed */
int mppb_ptg_update_dyn_state(mppb_ptg* ptg, const double* dyn7, int force);
/* SpeedTrimmablePTG::trimmableSpeed_ */
int mppb_ptg_set_trim_speed(mppb_ptg* ptg, double speed);
/* getPathTwist: out3 = vx, vy, omega */
int mppb_ptg_path_twist(const mppb_ptg* ptg, int k, uint32_t step, double* out3);
/* HolonomicBlend: candidate parameters of path k at the current dynamic state and trimmed speed */
int mppb_ptg_holo_params(const mppb_ptg* ptg, int k, mppb_holo_cand* out_cand);

/* ---- planner ----------------------------------------------------------------------------------
 * mpp::TPS_Astar behind a C surface (reference: mpp::Planner::plan, include/mpp/algos/Planner.h:36;
 * TPS_Astar_Parameters, include/mpp/algos/TPS_Astar.h:24-57; the set-up sequence of
 * apps/path-planner-cli/path-planner-cli.cpp:200-341). The A* search runs on the host; every
 * expanded node's collision rows and every tentative edge's evaluator costs are computed on the
 * GPU of `ctx`. */
typedef struct mppb_planner mppb_planner;

typedef struct mppb_astar_params
{
    double        grid_resolution_xy;
    double        grid_resolution_yaw; /* radians */
    double        SE2_metricAngleWeight;
    double        heuristic_heading_weight;
    uint32_t      max_ptg_trajectories_to_explore;
    uint32_t      max_ptg_speeds_to_explore;
    uint32_t      path_interpolated_segments;
    double        maximum_computation_time; /* seconds; <= 0: unlimited */
    uint64_t      max_expansions;           /* extension for deterministic runs; 0: unlimited */
    const double* ptg_sample_timestamps;
    uint32_t      num_timestamps;
} mppb_astar_params;

typedef struct mppb_plan_query
{
    double  start[6]; /* x y phi vx vy omega (global frame) */
    double  goal[6];  /* x y phi vx vy omega; phi ignored if goal_is_point */
    int32_t goal_is_point;
    double  bbox_min[3], bbox_max[3]; /* world box: x y phi */
} mppb_plan_query;

typedef struct mppb_plan_result
{
    int32_t  success;
    double   path_cost;
    uint64_t tree_nodes, tree_parents, nodes_expanded, tps_points, edges_costed;
    int64_t  best_node, goal_node;
    double   computation_time, best_cost_to_goal;
    uint32_t path_len; /* nodes on the path root -> best node */
} mppb_plan_result;

int         mppb_planner_create(mppb_ctx* ctx, mppb_planner** out_planner);
void        mppb_planner_destroy(mppb_planner* p);
const char* mppb_planner_last_error(const mppb_planner* p);
int         mppb_planner_set_params(mppb_planner* p, const mppb_astar_params* params);
/* the PTG object is borrowed and must outlive the planner; PTGs are used in the order added */
int mppb_planner_add_ptg(mppb_planner* p, mppb_ptg* ptg);
/* one static obstacle source (ObstacleSource::FromStaticPointcloud); the points are copied */
int mppb_planner_add_obstacles(mppb_planner* p, const float* xs, const float* ys, size_t n);
/* new points for obstacle source #source, written into the source's existing cloud object (the next plan sees
 * them, as the reference re-reads ObstacleSource::obstacles() on every plan(), TPS_Astar.cpp:125-137) */
int mppb_planner_update_obstacles(mppb_planner* p, size_t source, const float* xs, const float* ys, size_t n);
/* CostEvaluatorCostMap::FromStaticPointObstacles + push_back into costEvaluators_ */
int mppb_planner_add_costmap(mppb_planner* p, const float* xs, const float* ys, size_t n, double resolution,
                             double preferred_clearance, double max_cost, int use_average, double max_radius_from_robot,
                             const double* robot_pose_xyphi /* NULL if max_radius_from_robot <= 0 */);
/* geometry (x_min, y_min, res, nx, ny) and cells [ny][nx] of costmap evaluator #evaluator; either may be NULL */
int mppb_planner_costmap_cells(mppb_planner* p, int evaluator, double* info5, double* cells);
/* CostEvaluatorPreferredWaypoint + setPreferredWaypoints + push_back into costEvaluators_ */
int mppb_planner_add_waypoints(mppb_planner* p, const double* xs, const double* ys, size_t n, double influence_radius,
                               double cost_scale, int use_average);
/* n_queries == 1: TPS_Astar::plan. n_queries > 1: independent queries over the same obstacles,
 * advanced in lock-step with one GPU batch per A* iteration (n_threads host threads, 0 = all);
 * each result equals what plan() returns for that query alone. */
int mppb_planner_plan(mppb_planner* p, size_t n_queries, const mppb_plan_query* queries, int n_threads,
                      mppb_plan_result* results);
/* Independent queries over several GPUs of ONE process: planners[i] was created on a context of its own (normally one
 * per device) and set up like the others (same PTGs, obstacles, evaluators, parameters). Query q is planned by planner
 * q % n_planners as its local query q / n_planners; the planners run at once, each on a thread of its own with
 * n_threads_per_planner host threads (0 = all). results[q] is what mppb_planner_plan returns for that query; paths and
 * trees are read from the owning planner with the local query index. */
int mppb_planner_plan_multi(mppb_planner** planners, int n_planners, size_t n_queries, const mppb_plan_query* queries,
                            int n_threads_per_planner, mppb_plan_result* results);
/* refine_trajectory on the backtracked path of a query (src/algos/refine_trajectory.cpp:14-89) */
int mppb_planner_refine(mppb_planner* p, size_t query);
/* nodes [path_len][8] = id x y phi vx vy omega cost; edges [path_len-1][8] = ptgIndex k step ptgDist cost
 * estimatedExecTime trimSpeed nInterpolated; either may be NULL */
int mppb_planner_path(mppb_planner* p, size_t query, double* nodes, double* edges);
/* whole motion tree, rows [id parent(-1 for the root) x y phi cost edgeCost]; returns the node count */
int64_t mppb_planner_tree(mppb_planner* p, size_t query, double* out, int64_t cap);
/* plan_to_trajectory + save_to_txt text (src/algos/trajectories.cpp:15-106); returns the text length */
int64_t mppb_planner_trajectory_txt(mppb_planner* p, size_t query, double sample_period, char* buf, int64_t cap);
/* Work of the last plan, out14 = collision batches, cost batches, nodes, holo candidates, edges, seconds in GPU
 * calls, then wall seconds per phase: set-up, F relax + pop + A enumerate, B collision (GPU; for a single query it
 * includes the second half of A, which the host runs while the kernels do), C+D select/build, E assemble,
 * E cost (GPU), teardown; [13] = HolonomicBlend free/blocked decisions taken within 1e-9 relative of a tie (the guard
 * on the device's acos/cos in the quartic's trigonometric branch; expected 0) */
int mppb_planner_gpu_stats(const mppb_planner* p, double* out14);
/* Host self-test (no GPU) of the worker pool behind plan_batch: `rounds` times every index of [0,count) must run
 * exactly once on n_threads workers, and an exception thrown by a work item must reach the caller. Returns the number
 * of indices that did not (0 = correct), negative on a lost exception. */
int mppb_selftest_worker_pool(int n_threads, size_t count, int rounds);
/* Host self-test (no GPU) of the planner's open set, a binary heap ordered by (key, insertion number) that stands in
 * for the reference's std::multimap<cost, Node*> (TPS_Astar.cpp: inserted into and popped at begin() only): `ops`
 * random insertions with frequent equal keys, interleaved with pops. Returns the number of pops that differ from a
 * std::multimap driven the same way (0 = same order). */
int mppb_selftest_open_set(uint64_t seed, size_t ops, int distinct_keys);
/* Host self-test (no GPU) of the planner's SE(2) lattice index, an open-addressing table over arena nodes that
 * stands in for the reference's std::unordered_map<NodeCoords, Node> (TPS_Astar.h:222; only ever indexed): `ops`
 * look-ups of random coordinates in [-range/2, range/2) through several table growths; every coordinate must keep
 * returning the node it got the first time, distinct coordinates distinct nodes. Returns the number of violations. */
int mppb_selftest_lattice_index(uint64_t seed, size_t ops, int range);

/* Host self-test (no GPU) of the planner's replay of the ORDER in which the reference's std::unordered_map of reached
 * cells (TPS_Astar.cpp:565-566, 806-814) is iterated — the order that assigns node ids: `trials` random sets of up to
 * max_keys distinct lattice cells are fed to the replay and to a real std::unordered_map with the reference's hash.
 * Returns the number of sets iterated in a different order (0 = identical). */
int mppb_selftest_cell_order(uint64_t seed, size_t trials, int max_keys);

/* Host self-test (no GPU) of the helper thread behind mppb_eval_nodes' heading cos/sin (large host batches): `rounds`
 * batches of n random poses are split between the calling thread and the helper as the entry point splits them, with
 * pauses of 0 to 5 ms between them (spinning and sleeping helper); every (x, y, cos, sin) row must equal the
 * single-threaded mppb_nodes_xycs_from_poses bit for bit. Returns the number of rows that differ (0 = correct). */
int mppb_selftest_sincos_helper(uint64_t seed, size_t n, int rounds);

#ifdef __cplusplus
}
#endif
#endif /* MPPB_H_ */
