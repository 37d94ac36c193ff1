// mpp:: planner API mirror (product code, host side).
//
// Source-level mirror of the part of the reference's public API that surrounds the hot path, so
// that the GPU path is a drop-in. Same names, argument meaning and error
// behaviour (std::exception on bad input); MRPT types are replaced by the small value types of
// mpp_math.h. Reference headers mirrored (paths under mrpt_path_planning/include/mpp/):
//   data/SE2_KinState.h:32-87, data/MoveEdgeSE2_TPS.h:20-68, data/MotionPrimitivesTree.h:72-300,
//   data/PlannerInput.h:16-29, data/PlannerOutput.h:19-54, data/TrajectoriesAndRobotShape.h:20-43,
//   data/ProgressCallbackData.h:19-34, data/trajectory_t.h, interfaces/ObstacleSource.h:19-112,
//   algos/CostEvaluator.h:22-43, algos/CostEvaluatorCostMap.h, algos/CostEvaluatorPreferredWaypoint.h,
//   algos/Planner.h:23-62, algos/TPS_Astar.h:24-310, algos/edge_interpolated_path.h:14-18,
//   algos/refine_trajectory.h, algos/trajectories.h.
// What differs by design: TPS_Astar evaluates collisions and evaluator costs through libmpp_b200
// (mppb_* C ABI) in batches, and offers plan_batch() for many independent queries in lock-step.
#pragma once
#include <algorithm>
#include <functional>
#include <iterator>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <optional>
#include <string>
#include <vector>

#include "ptg.h"

namespace mpp
{
using TNodeID            = uint64_t;
using cost_t             = double;
using distance_t         = double;
using duration_seconds_t = double;
using normalized_speed_t = double;
using trajectory_index_t = int;
using ptg_step_t         = uint32_t;

// ---- configuration files (config.cpp) ---------------------------------------------------------
/** INI reader with `@define NAME value` / `${NAME}` substitution and `//`, `#`, `;` comments
 *  ([MRPT] CConfigFile as share/ptgs_*.ini uses it). */
class ConfigFile
{
   public:
    explicit ConfigFile(const std::string& fileName);
    static ConfigFile FromText(const std::string& text);
    bool              has(const std::string& section, const std::string& key) const;
    std::string       read_string(const std::string& s, const std::string& k, const std::string& def, bool must = false) const;
    double            read_double(const std::string& s, const std::string& k, double def, bool must = false) const;
    int               read_int(const std::string& s, const std::string& k, int def, bool must = false) const;
    std::vector<double> read_vector(const std::string& s, const std::string& k) const;

   private:
    ConfigFile() = default;
    void parse(const std::string& text);
    std::map<std::string, std::map<std::string, std::string>> data_;  // lower-case section/key
};

/** YAML subset: a top-level map of scalars and flow sequences, or a top-level block sequence of
 *  such maps (enough for the files in share/). */
class Yaml
{
   public:
    static Yaml FromFile(const std::string& fileName);
    static Yaml FromText(const std::string& text);
    bool        isMap() const { return !isSeq_; }
    bool        has(const std::string& key) const { return map_.count(key) != 0; }
    double      getDouble(const std::string& key) const;
    bool        getBool(const std::string& key) const;
    std::vector<double> getDoubleVector(const std::string& key) const;
    const std::vector<Yaml>& sequence() const { return seq_; }
    template <class T>
    T getOrDefault(const std::string& key, const T& def) const
    {
        return has(key) ? static_cast<T>(getDouble(key)) : def;
    }

   private:
    bool                               isSeq_ = false;
    std::map<std::string, std::string> map_;
    std::vector<Yaml>                  seq_;
};

// ---- data model --------------------------------------------------------------------------------
/** SE(2) pose or R(2) point goal (PoseOrPoint variant of the reference) */
struct PoseOrPoint
{
    PoseOrPoint() = default;
    PoseOrPoint(const TPoint2D& p) : kind_(2), pose_(p.x, p.y, 0) {}
    PoseOrPoint(const TPose2D& p) : kind_(1), pose_(p) {}
    /** "[x y]" -> point, "[x y phi_degrees]" -> pose */
    static PoseOrPoint FromString(const std::string& s);
    bool               isEmpty() const { return kind_ == 0; }
    bool               isPose() const { return kind_ == 1; }
    bool               isPoint() const { return kind_ == 2; }
    const TPose2D&     pose() const;
    TPoint2D           point() const;
    std::string        asString() const;

   private:
    int     kind_ = 0;
    TPose2D pose_;
};

struct SE2_KinState
{
    TPose2D     pose{0, 0, 0};
    TTwist2D    vel{0, 0, 0};
    std::string asString() const;
};
struct SE2orR2_KinState
{
    PoseOrPoint  state;
    TTwist2D     vel{0, 0, 0};
    SE2_KinState asSE2KinState() const;
    std::string  asString() const;
};

/** Time-keyed pose samples of an edge: a sorted flat map with inline storage and the subset of the
 *  std::map interface the reference uses on MoveEdgeSE2_TPS::interpolatedPath (ordered iteration over
 *  (time, pose) pairs, operator[] insert-or-assign, size/empty/clear, find). An edge holds at most
 *  nSeg + 2 samples (9 after refine_trajectory), so no heap allocation is needed: motion trees with
 *  millions of edges are built without touching the allocator for their samples. */
class TimedPoseMap
{
   public:
    using key_type       = duration_seconds_t;
    using mapped_type    = TPose2D;
    using value_type     = std::pair<duration_seconds_t, TPose2D>;
    using iterator       = value_type*;
    using const_iterator = const value_type*;
    static constexpr size_t kInline = 10;

    TimedPoseMap() = default;
    TimedPoseMap(const TimedPoseMap& o) { *this = o; }
    TimedPoseMap& operator=(const TimedPoseMap& o)
    {
        if (this == &o) return *this;
        n_ = o.n_;
        if (o.heap_.empty())
        {
            heap_.clear();
            std::copy(o.inl_, o.inl_ + n_, inl_);
        }
        else
            heap_ = o.heap_;
        return *this;
    }
    iterator       begin() { return data(); }
    iterator       end() { return data() + n_; }
    const_iterator begin() const { return data(); }
    const_iterator end() const { return data() + n_; }
    std::reverse_iterator<const_iterator> rbegin() const { return std::reverse_iterator<const_iterator>(end()); }
    std::reverse_iterator<const_iterator> rend() const { return std::reverse_iterator<const_iterator>(begin()); }
    size_t size() const { return n_; }
    bool   empty() const { return n_ == 0; }
    void   clear()
    {
        n_ = 0;
        heap_.clear();
    }
    const_iterator find(key_type k) const
    {
        const_iterator it = std::lower_bound(begin(), end(), k, [](const value_type& v, key_type x) { return v.first < x; });
        return (it != end() && it->first == k) ? it : end();
    }
    size_t count(key_type k) const { return find(k) != end() ? 1 : 0; }
    const mapped_type& at(key_type k) const
    {
        const_iterator it = find(k);
        if (it == end()) throw std::out_of_range("TimedPoseMap::at");
        return it->second;
    }
    mapped_type& operator[](key_type k)
    {
        value_type* d  = data();
        size_t      lo = std::lower_bound(d, d + n_, k, [](const value_type& v, key_type x) { return v.first < x; }) - d;
        if (lo < n_ && d[lo].first == k) return d[lo].second;
        if (heap_.empty() && n_ == kInline) heap_.assign(inl_, inl_ + n_);  // rare: spill to the heap
        if (!heap_.empty())
        {
            heap_.insert(heap_.begin() + lo, value_type(k, TPose2D()));
            n_++;
            return heap_[lo].second;
        }
        for (size_t i = n_; i > lo; i--) inl_[i] = inl_[i - 1];
        inl_[lo] = value_type(k, TPose2D());
        n_++;
        return inl_[lo].second;
    }

   private:
    value_type*       data() { return heap_.empty() ? inl_ : heap_.data(); }
    const value_type* data() const { return heap_.empty() ? inl_ : heap_.data(); }
    value_type              inl_[kInline];
    std::vector<value_type> heap_;
    size_t                  n_ = 0;
};

struct MoveEdgeSE2_TPS
{
    TNodeID            parentId = static_cast<TNodeID>(-1);
    SE2_KinState       stateFrom, stateTo;
    double             cost         = std::numeric_limits<double>::max();
    int8_t             ptgIndex     = -1;
    int16_t            ptgPathIndex = -1;
    distance_t         ptgDist      = std::numeric_limits<distance_t>::max();
    normalized_speed_t ptgTrimmableSpeed = 1.0;
    TPose2D            ptgFinalRelativeGoal;
    normalized_speed_t ptgFinalGoalRelSpeed = .0;
    duration_seconds_t estimatedExecTime    = .0;
    /** std::map<duration_seconds_t, TPose2D> in the reference; same ordered-map semantics, flat storage */
    TimedPoseMap interpolatedPath;
    /** extension: the PTG step the edge ends at (relTrgStep of the accepted neighbour) */
    ptg_step_t ptgStep = 0;

    TNavDynamicState getPTGDynState() const;
    std::string      asString() const;
};

class MotionPrimitivesTreeSE2
{
   public:
    struct node_t : public SE2_KinState
    {
        TNodeID                nodeID_ = static_cast<TNodeID>(-1);
        std::optional<TNodeID> parentID_;
        cost_t                 cost_ = 0;
    };
    struct TEdgeInfo
    {
        TNodeID         id;
        bool            reverse = false;
        MoveEdgeSE2_TPS data;
    };
    using TListEdges      = std::list<TEdgeInfo>;
    using node_map_t      = std::map<TNodeID, node_t>;
    using path_t          = std::list<node_t>;
    using edge_sequence_t = std::list<MoveEdgeSE2_TPS*>;
    using edge_t          = MoveEdgeSE2_TPS;

    TNodeID                       root = 0;
    std::map<TNodeID, TListEdges> edges_to_children;

    void insert_root_node(TNodeID id, const SE2_KinState& data);
    void insert_node_and_edge(TNodeID parentId, TNodeID childId, const SE2_KinState& childData, const MoveEdgeSE2_TPS& e);
    void rewire_node_parent(TNodeID nodeId, const MoveEdgeSE2_TPS& newEdgeFromParent);
    /** extension: insert_node_and_edge for a caller that adds several children of ONE parent in a row (an A* expansion):
     *  the parent's edge list and node are looked up once by the caller (`edges_to_children[parentId]`,
     *  `nodes().at(parentId)`; both stay valid across insertions), and a child id above all present ids goes in with an
     *  end hint. Same resulting tree as insert_node_and_edge. */
    void insert_child_of(TListEdges& parentEdges, const node_t& parentNode, TNodeID parentId, TNodeID childId,
                         const SE2_KinState& childData, const MoveEdgeSE2_TPS& e);
    /** same, taking the edge over (its interpolated path is not copied) */
    void insert_child_of(TListEdges& parentEdges, const node_t& parentNode, TNodeID parentId, TNodeID childId,
                         const SE2_KinState& childData, MoveEdgeSE2_TPS&& e);
    const MoveEdgeSE2_TPS& edge_to_parent(TNodeID nodeId) const;
    const node_map_t&      nodes() const { return nodes_; }
    SE2_KinState&          node_state(TNodeID id) { return nodes_.at(id); }
    std::tuple<path_t, edge_sequence_t> backtrack_path(TNodeID target) const;

   private:
    node_map_t nodes_;
};

/** float SoA obstacle cloud (the CPointsMap buffers the hot path reads) */
struct PointCloud
{
    std::vector<float> xs, ys;
    size_t             size() const { return xs.size(); }
    bool               empty() const { return xs.empty(); }
    void               insertPoint(float x, float y)
    {
        xs.push_back(x);
        ys.push_back(y);
    }
    /** text file with one "x y" pair per line ([MRPT] load2D_from_text_file) */
    bool load2D_from_text_file(const std::string& fileName);
    void boundingBox(float& minx, float& miny, float& maxx, float& maxy) const;
    using Ptr = std::shared_ptr<PointCloud>;
};

class ObstacleSource
{
   public:
    using Ptr                 = std::shared_ptr<ObstacleSource>;
    virtual ~ObstacleSource() = default;
    static Ptr FromStaticPointcloud(const PointCloud::Ptr& pc);
    /** the (unclipped) cloud; clipping to the world box is applied on the GPU at upload time */
    virtual PointCloud::Ptr obstacles() = 0;
    virtual bool            dynamic() const { return false; }
};

using robot_radius_t = double;
struct TrajectoriesAndRobotShape
{
    std::vector<std::shared_ptr<ptg_t>> ptgs;
    std::vector<double>                 shape_xs, shape_ys;  // polygonal shape (may be empty)
    std::optional<robot_radius_t>       shape_radius;        // circular shape
    /** buildCtx (extension): collision grids are built on that context's GPU (mppb_build_collision_grid) instead of
     *  the host builder; the tables are the same. */
    void initFromConfigFile(const ConfigFile& cfg, const std::string& section = "SelfDriving", mppb_ctx* buildCtx = nullptr);
    /** extension: adopt already constructed PTG objects */
    void initFromPtgs(std::vector<std::shared_ptr<ptg_t>> list)
    {
        ptgs         = std::move(list);
        initialized_ = !ptgs.empty();
    }
    bool initialized() const { return initialized_; }
    void clear() { *this = TrajectoriesAndRobotShape(); }

   private:
    bool initialized_ = false;
};

struct PlannerInput
{
    SE2_KinState                     stateStart;
    SE2orR2_KinState                 stateGoal;
    TPose2D                          worldBboxMin, worldBboxMax;
    std::vector<ObstacleSource::Ptr> obstacles;
    TrajectoriesAndRobotShape        ptgs;
};

struct PlannerOutput
{
    PlannerInput            originalInput;
    bool                    success         = false;
    double                  computationTime = 0;
    double                  pathCost        = std::numeric_limits<double>::max();
    std::optional<TNodeID>  goalNodeId, bestNodeId;
    cost_t                  bestNodeIdCostToGoal = std::numeric_limits<cost_t>::max();
    MotionPrimitivesTreeSE2 motionTree;
    /** extension: work counters of this plan */
    size_t nodesExpanded = 0, tpsPointsEvaluated = 0, edgesCosted = 0;
    /** extension: size of the motion tree and nodes on the path root -> best node (valid whether or not motionTree has
     *  been built yet) */
    size_t treeNodes = 0, treeParents = 0, pathLength = 0;
    /** extension: TPS_Astar::plan_batch(..., materializeTrees = false) leaves motionTree empty — a search of thousands
     *  of queries holds tens of millions of tree nodes that most callers never look at — and puts a builder here;
     *  materializeTree() runs it once. plan() and plan_batch() by default return with the tree built. */
    std::function<void(PlannerOutput&)> treeBuilder;
    void materializeTree()
    {
        if (!treeBuilder) return;
        auto f      = std::move(treeBuilder);
        treeBuilder = nullptr;
        f(*this);
    }
};

// ---- cost evaluators -----------------------------------------------------------------------------
/** Plugin interface (CostEvaluator.h:22-43). Evaluators of the two known types are evaluated on
 *  the GPU in batches by the planner; any other subclass is called on the host per edge. */
class CostEvaluator
{
   public:
    using Ptr                                                       = std::shared_ptr<CostEvaluator>;
    virtual ~CostEvaluator()                                        = default;
    virtual double operator()(const MoveEdgeSE2_TPS& edge) const    = 0;
};

namespace detail
{
/** The private one-edge GPU context behind operator() of the two built-in evaluators. A planner keeps ITS evaluator
 *  slots resident in its own context; a direct operator() call (a user calling the plugin, Planner::cost_path_segment,
 *  or the planner's host-plugin path for evaluators beyond MPPB_MAX_EVALUATORS) must not disturb them, and may come
 *  from any thread: it runs on a context of its own, created on first use on the same device, behind a mutex. */
struct EvaluatorGpu
{
    std::mutex m;
    mppb_ctx*  ctx      = nullptr;
    bool       uploaded = false;
    double     key[6]   = {0, 0, 0, 0, 0, 0};  // the public parameters the resident copy was made from
    ~EvaluatorGpu()
    {
        if (ctx) mppb_ctx_destroy(ctx);
    }
};
}  // namespace detail

class CostEvaluatorCostMap : public CostEvaluator
{
   public:
    using Ptr = std::shared_ptr<CostEvaluatorCostMap>;
    struct Parameters
    {
        double resolution = 0.05, preferredClearanceDistance = 0.4, maxCost = 2.0, maxRadiusFromRobot = .0;
        bool   useAverageOfPath = false;
        static Parameters FromYAML(const Yaml& c);
    };
    /** nearest-obstacle search on the GPU (ctx), pow() on the host (libm, bit parity with the reference) */
    static Ptr FromStaticPointObstacles(mppb_ctx* ctx, const PointCloud& obsPts, const Parameters& p,
                                        const std::optional<TPose2D>& curRobotPose = std::nullopt);
    double     operator()(const MoveEdgeSE2_TPS& edge) const override;  // one-edge GPU batch

    Parameters          params_;
    double              x_min = 0, y_min = 0, x_max = 0, y_max = 0;
    int                 size_x = 0, size_y = 0;
    std::vector<double> cells;  // row-major [size_y][size_x]
    mppb_ctx*           ctx_ = nullptr;  // the context the table was built on (tells operator() which device to use)

   private:
    std::shared_ptr<detail::EvaluatorGpu> gpu_ = std::make_shared<detail::EvaluatorGpu>();
};

class CostEvaluatorPreferredWaypoint : public CostEvaluator
{
   public:
    using Ptr = std::shared_ptr<CostEvaluatorPreferredWaypoint>;
    struct Parameters
    {
        double waypointInfluenceRadius = 1.5, costScale = 10.0;
        bool   useAverageOfPath = true;
        static Parameters FromYAML(const Yaml& c);
    };
    explicit CostEvaluatorPreferredWaypoint(mppb_ctx* ctx) : ctx_(ctx) {}
    void   setPreferredWaypoints(const std::vector<TPoint2D>& pts);
    double operator()(const MoveEdgeSE2_TPS& edge) const override;  // one-edge GPU batch

    Parameters            params_;
    std::vector<TPoint2D> waypoints_;
    mppb_ctx*             ctx_ = nullptr;  // tells operator() which device to use

   private:
    std::shared_ptr<detail::EvaluatorGpu> gpu_ = std::make_shared<detail::EvaluatorGpu>();
};

/** reads the `target: [x, y]` entries of a waypoint-sequence YAML (src/data/Waypoints.cpp:85-105) */
std::vector<TPoint2D> waypoint_targets_from_yaml(const Yaml& y);

// ---- planner ---------------------------------------------------------------------------------------
struct ProgressCallbackData
{
    cost_t                                   bestCostFromStart = std::numeric_limits<cost_t>::max();
    cost_t                                   bestCostToGoal    = std::numeric_limits<cost_t>::max();
    std::optional<TNodeID>                   bestFinalNode;
    MotionPrimitivesTreeSE2::edge_sequence_t bestPath;
    const MotionPrimitivesTreeSE2*           tree              = nullptr;
    const PlannerInput*                      originalPlanInput = nullptr;
    const std::vector<CostEvaluator::Ptr>*   costEvaluators    = nullptr;
};
using planner_progress_callback_t = std::function<void(const ProgressCallbackData&)>;

/** Named timing sections ([MRPT] mrpt::system::CTimeLogger as Planner::profiler_() of the reference exposes it,
 *  Planner.h:44-46): number of calls and accumulated seconds per section name. */
class TimeLogger
{
   public:
    struct Section
    {
        size_t calls   = 0;
        double seconds = 0;
    };
    void   clear() { sections_.clear(); }
    void   registerUserMeasure(const std::string& name, double seconds, size_t calls = 1)
    {
        Section& s = sections_[name];
        s.calls += calls;
        s.seconds += seconds;
    }
    const std::map<std::string, Section>& sections() const { return sections_; }
    double      getMeanTime(const std::string& name) const
    {
        const auto it = sections_.find(name);
        return it == sections_.end() || !it->second.calls ? 0.0 : it->second.seconds / static_cast<double>(it->second.calls);
    }
    std::string getStatsAsText() const;

   private:
    std::map<std::string, Section> sections_;
};

class Planner
{
   public:
    using Ptr          = std::shared_ptr<Planner>;
    virtual ~Planner() = default;
    virtual PlannerOutput plan(const PlannerInput& in) = 0;
    virtual void          params_from_yaml(const Yaml& c) = 0;
    virtual std::string   params_as_yaml_text() const     = 0;

    std::vector<CostEvaluator::Ptr> costEvaluators_;
    planner_progress_callback_t     progressCallback_;
    double                          progressCallbackCallPeriod_ = 0.1;

    /** estimatedExecTime + sum of the evaluators (Planner.cpp:15-29); known evaluator types run on the GPU */
    cost_t cost_path_segment(const MoveEdgeSE2_TPS& edge) const;

    /** timing sections of the last plan(s) under the reference's section names (TPS_Astar.cpp:89,195,548,574,605,
     *  701,740,803,833): see TPS_Astar::plan */
    TimeLogger& profiler_() { return profiler__; }

   protected:
    TimeLogger profiler__;
};

struct TPS_Astar_Parameters
{
    double   grid_resolution_xy              = 0.20;
    double   grid_resolution_yaw             = 5.0 * M_PI / 180.0;
    double   SE2_metricAngleWeight           = 1.0;
    double   heuristic_heading_weight        = 0.1;
    uint32_t max_ptg_trajectories_to_explore = 20;
    std::vector<duration_seconds_t> ptg_sample_timestamps = {1.0, 3.0, 5.0};
    uint32_t max_ptg_speeds_to_explore = 3;
    size_t   pathInterpolatedSegments  = 5;
    duration_seconds_t maximumComputationTime = std::numeric_limits<duration_seconds_t>::max();
    /** extension (deterministic benchmarks): stop after this many node expansions */
    size_t maxExpansions = std::numeric_limits<size_t>::max();

    void                        load_from_yaml(const Yaml& c);
    static TPS_Astar_Parameters FromYAML(const Yaml& c);
    std::string                 as_yaml_text() const;
};

using astar_heuristic_t = std::function<cost_t(const SE2_KinState&, const SE2orR2_KinState&)>;

/** A* over an SE(2) lattice with PTG motion primitives (TPS_Astar.h). The search logic stays on
 *  the host, restructured in phases so that every expanded node's collision rows and every
 *  candidate edge's evaluator costs are computed by the GPU library in one batch per phase. */
class TPS_Astar : public Planner
{
   public:
    /** ctx == nullptr: a context on CUDA device 0 is created on first use (throws without a GPU) */
    explicit TPS_Astar(mppb_ctx* ctx = nullptr);
    ~TPS_Astar() override;

    TPS_Astar_Parameters params_;

    PlannerOutput plan(const PlannerInput& in) override;
    /** Many independent queries over the SAME obstacle sources, advanced in lock-step: one GPU
     *  batch per A* iteration covers the current node of every unfinished query. Results are
     *  identical to calling plan() on each input alone. nThreads host threads (0 = all). */
    std::vector<PlannerOutput> plan_batch(const std::vector<PlannerInput>& ins, int nThreads = 0, bool materializeTrees = true);

    void        params_from_yaml(const Yaml& c) override { params_.load_from_yaml(c); }
    std::string params_as_yaml_text() const override { return params_.as_yaml_text(); }

    cost_t default_heuristic(const SE2_KinState& from, const SE2orR2_KinState& goal) const;
    cost_t default_heuristic_SE2(const SE2_KinState& from, const TPose2D& goal) const;
    cost_t default_heuristic_R2(const SE2_KinState& from, const TPoint2D& goal) const;
    astar_heuristic_t heuristic;

    mppb_ctx* context();
    /** device work of the last plan()/plan_batch(): GPU batches issued and seconds spent in them */
    struct GpuStats
    {
        size_t collisionBatches = 0, costBatches = 0, nodesEvaluated = 0, holoCandidates = 0, edgesCosted = 0;
        /** HolonomicBlend TPS points whose free/blocked decision (relTrgDist >= freeDistance, TPS_Astar.cpp:749) was
         *  taken with the two distances within 1e-9 relative of each other — the only place where the device's
         *  acos/cos (CUDA libm, not glibc) could tip a decision. Expected: 0. */
        size_t holoNearTies = 0;
        double gpuSeconds = 0;
        /** wall seconds per phase: [0] set-up and uploads, [1] F relax + pop + A enumerate (fused per query), [2] B collision (GPU),
         *  [3] C+D select/build, [4] E assemble, [5] E cost (GPU), [6] result hand-over and teardown */
        double phaseSeconds[7] = {0, 0, 0, 0, 0, 0, 0};
    } lastGpuStats;

   private:
    struct Impl;
    std::unique_ptr<Impl> impl_;
};

// ---- class registry and plug-ins ---------------------------------------------------------------------
/** What MRPT's RTTI class factory does for the reference: classes register under their names when a library is
 *  loaded (src/registerClasses.cpp:16-37), `--planner <name>` instantiates by name (path-planner-cli.cpp:256-265),
 *  `--plugins a.so,b.so` loads further libraries whose initialisers register more classes (:452-461), and the
 *  `PTGn_Type` keys of a PTG configuration name PTG classes (TrajectoriesAndRobotShape.cpp:58-64). */
class ClassRegistry
{
   public:
    using PlannerFactory = std::function<std::shared_ptr<Planner>(mppb_ctx*)>;
    /** builds PTG #index of section `section`; shape: the robot shape read from the same section; buildCtx: see
     *  TrajectoriesAndRobotShape::initFromConfigFile */
    using PtgFactory = std::function<std::shared_ptr<ptg_t>(const ConfigFile& cfg, const std::string& section, const std::string& keyPrefix,
                                                            const TrajectoriesAndRobotShape& shape, mppb_ctx* buildCtx)>;
    static ClassRegistry& instance();
    void                  registerPlanner(const std::string& name, PlannerFactory f);
    void                  registerPtg(const std::string& name, PtgFactory f);
    /** nullptr if no class of that name is registered */
    std::shared_ptr<Planner> createPlanner(const std::string& name, mppb_ctx* ctx) const;
    std::shared_ptr<ptg_t>   createPtg(const std::string& name, const ConfigFile& cfg, const std::string& section,
                                       const std::string& keyPrefix, const TrajectoriesAndRobotShape& shape, mppb_ctx* buildCtx) const;
    std::vector<std::string> plannerNames() const;
    std::vector<std::string> ptgNames() const;

   private:
    ClassRegistry();
    std::map<std::string, PlannerFactory> planners_;
    std::map<std::string, PtgFactory>     ptgs_;
};
/** Loads the shared objects of a comma-separated list (RTLD_NOW | RTLD_GLOBAL); their static initialisers call
 *  ClassRegistry::instance().register...(). Returns false and fills `error` if one cannot be loaded. */
bool loadPluginModules(const std::string& commaSeparatedList, std::string* error = nullptr);
/** in a plug-in: MPP_REGISTER_PLANNER("my::Planner", MyPlanner)  (MyPlanner constructible from mppb_ctx*) */
#define MPP_REGISTER_PLANNER(NAME, CLASS)                                                                              \
    namespace                                                                                                          \
    {                                                                                                                  \
    const bool mpp_registered_##CLASS = (::mpp::ClassRegistry::instance().registerPlanner(                            \
                                             NAME, [](mppb_ctx* c) { return std::static_pointer_cast<::mpp::Planner>(std::make_shared<CLASS>(c)); }), \
                                         true);                                                                        \
    }

// ---- free functions ----------------------------------------------------------------------------------
void edge_interpolated_path(MoveEdgeSE2_TPS& edge, const TrajectoriesAndRobotShape& trs,
                            const std::optional<TPose2D>& reconstrRelPoseOpt = std::nullopt,
                            const std::optional<size_t>& ptg_stepOpt = std::nullopt,
                            const std::optional<size_t>& numSegments = std::nullopt);

/** mpp::transform_pc_square_clipping (transform_pc_square_clipping.h:16-19) on the GPU of `ctx`: the points of
 *  inMap inside the +-maxDistXY square around the pose, in its local frame, appended to (or replacing) outMap. */
void transform_pc_square_clipping(mppb_ctx* ctx, const PointCloud& inMap, const TPose2D& asSeenFrom, double maxDistXY, PointCloud& outMap,
                                  bool appendToOutMap = true);
/** mpp::tp_obstacles_single_path (tp_obstacles_single_path.h:17-19) on the GPU of `ctx`: free TP-distance of path k
 *  against obstacle points given in the robot frame, initTPObstacleSingle's value included. One call uploads the
 *  points and, if they are not the context's current PTG set, the PTG's tables: this is the reference's single-call
 *  interface; a planner should use the batched entry points. */
distance_t tp_obstacles_single_path(mppb_ctx* ctx, trajectory_index_t tp_space_k_direction, const PointCloud& localObstacles, const ptg_t& ptg);

void refine_trajectory(const MotionPrimitivesTreeSE2::path_t& inPath, MotionPrimitivesTreeSE2::edge_sequence_t& edgesToRefine,
                       const TrajectoriesAndRobotShape& ptgInfo, const double ptg_tolerance_dist = 0.10);

struct trajectory_state_t
{
    SE2_KinState state;
    size_t       ptgIndex     = 0;
    int          ptgPathIndex = 0;
    uint32_t     ptgStep      = 0;
};
using trajectory_t = std::map<duration_seconds_t, trajectory_state_t>;
trajectory_t plan_to_trajectory(const MotionPrimitivesTreeSE2::edge_sequence_t& planEdges,
                                const TrajectoriesAndRobotShape& ptgInfo, const duration_seconds_t samplePeriod = 50e-3);
std::string  trajectory_as_text(const trajectory_t& traj);
bool         save_to_txt(const trajectory_t& traj, const std::string& fileName);

bool within_bbox(const TPose2D& p, const TPose2D& max, const TPose2D& min);
bool within_bbox(const TPoint2D& p, const TPose2D& max, const TPose2D& min);

}  // namespace mpp
