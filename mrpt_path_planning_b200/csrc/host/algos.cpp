// Data-model methods and the host-side free functions of the mpp:: mirror (product code).
// Reference files followed: src/data/SE2_KinState.cpp, src/data/MoveEdgeSE2_TPS.cpp,
// include/mpp/data/MotionPrimitivesTree.h, src/data/TrajectoriesAndRobotShape.cpp,
// src/algos/edge_interpolated_path.cpp, src/algos/refine_trajectory.cpp, src/algos/trajectories.cpp.
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>

#include "mpp.h"

namespace mpp
{
namespace
{
std::string fmt(const char* f, double a, double b, double c)
{
    char buf[160];
    snprintf(buf, sizeof(buf), f, a, b, c);
    return buf;
}
std::string pose_str(const TPose2D& p) { return fmt("(%.4f,%.4f,%.2fdeg)", p.x, p.y, p.phi * 180.0 / M_PI); }
std::string twist_str(const TTwist2D& t) { return fmt("(%.4f,%.4f,%.2f deg/s)", t.vx, t.vy, t.omega * 180.0 / M_PI); }
}  // namespace

// ---- PoseOrPoint / kinematic states -------------------------------------------------------------
PoseOrPoint PoseOrPoint::FromString(const std::string& s)
{
    std::string t = s;
    for (auto& c : t)
        if (c == '[' || c == ']' || c == ',' || c == ';') c = ' ';
    std::istringstream  ss(t);
    std::vector<double> v;
    double              x;
    while (ss >> x) v.push_back(x);
    std::string rest;
    if (!ss.eof() && (ss.clear(), ss >> rest, !rest.empty()))
        throw std::runtime_error("Malformed expression in FromString, s=\"" + s + "\"");
    if (v.size() == 2) return PoseOrPoint(TPoint2D{v[0], v[1]});
    if (v.size() == 3) return PoseOrPoint(TPose2D(v[0], v[1], v[2] * M_PI / 180.0));
    throw std::runtime_error("Wrong size of vector in FromString (expected 1x2 or 1x3)");
}
const TPose2D& PoseOrPoint::pose() const
{
    if (!isPose()) throw std::runtime_error("PoseOrPoint: not a pose");
    return pose_;
}
TPoint2D PoseOrPoint::point() const
{
    if (!isPoint()) throw std::runtime_error("PoseOrPoint: not a point");
    return {pose_.x, pose_.y};
}
std::string PoseOrPoint::asString() const
{
    if (isPoint()) return fmt("(%.4f,%.4f)%.0s", pose_.x, pose_.y, 0);
    if (isPose()) return pose_str(pose_);
    return "(empty)";
}
std::string  SE2_KinState::asString() const { return "p=" + pose_str(pose) + " v=" + twist_str(vel); }
std::string  SE2orR2_KinState::asString() const { return "p=" + state.asString() + " v=" + twist_str(vel); }
SE2_KinState SE2orR2_KinState::asSE2KinState() const
{
    SE2_KinState s;
    s.vel = vel;
    if (state.isPoint())
        s.pose = TPose2D(state.point().x, state.point().y, .0);
    else if (state.isPose())
        s.pose = state.pose();
    else
        throw std::runtime_error("Called with undefined state.");
    return s;
}

TNavDynamicState MoveEdgeSE2_TPS::getPTGDynState() const
{
    TNavDynamicState d;
    d.curVelLocal = stateFrom.vel;
    d.curVelLocal.rotate(-stateFrom.pose.phi);  // global to local velocity
    d.relTarget      = ptgFinalRelativeGoal;
    d.targetRelSpeed = ptgFinalGoalRelSpeed;
    return d;
}
std::string MoveEdgeSE2_TPS::asString() const
{
    std::stringstream ss;
    ss << "-\n  parentId: " << parentId << "\n  from: " << stateFrom.asString() << "\n  to: " << stateTo.asString()
       << "\n  {cost: " << cost << ", ptgIndex: " << static_cast<int>(ptgIndex) << ", path: " << ptgPathIndex
       << ", ptgDist: " << ptgDist << ", trimSpeed: " << ptgTrimmableSpeed
       << ", finalRelGoal: " << pose_str(ptgFinalRelativeGoal) << ", finalGoalSpeed: " << ptgFinalGoalRelSpeed
       << ", estimatedExecTime: " << estimatedExecTime << "}\n"
       << "  interpolatedPathSize: " << interpolatedPath.size() << "\n";
    return ss.str();
}

// ---- motion tree ----------------------------------------------------------------------------------
void MotionPrimitivesTreeSE2::insert_root_node(TNodeID id, const SE2_KinState& data)
{
    if (!nodes_.empty()) throw std::runtime_error("insert_root_node() called on a non-empty tree");
    node_t n;
    static_cast<SE2_KinState&>(n) = data;
    n.nodeID_                     = id;
    n.cost_                       = 0;
    nodes_[id]                    = n;
}
void MotionPrimitivesTreeSE2::insert_node_and_edge(TNodeID parentId, TNodeID childId, const SE2_KinState& childData,
                                                   const MoveEdgeSE2_TPS& e)
{
    edges_to_children[parentId].push_back(TEdgeInfo{childId, false, e});
    node_t n;
    static_cast<SE2_KinState&>(n) = childData;
    n.nodeID_                     = childId;
    n.parentID_                   = parentId;
    n.cost_                       = nodes_.at(parentId).cost_ + e.cost;
    nodes_[childId]               = n;
}
void MotionPrimitivesTreeSE2::insert_child_of(TListEdges& parentEdges, const node_t& parentNode, TNodeID parentId, TNodeID childId,
                                              const SE2_KinState& childData, const MoveEdgeSE2_TPS& e)
{
    insert_child_of(parentEdges, parentNode, parentId, childId, childData, MoveEdgeSE2_TPS(e));
}
void MotionPrimitivesTreeSE2::insert_child_of(TListEdges& parentEdges, const node_t& parentNode, TNodeID parentId, TNodeID childId,
                                              const SE2_KinState& childData, MoveEdgeSE2_TPS&& e)
{
    parentEdges.emplace_back();
    TEdgeInfo& ei = parentEdges.back();
    ei.id         = childId;
    ei.reverse    = false;
    const cost_t edgeCost = e.cost;
    ei.data               = std::move(e);
    node_t n;
    static_cast<SE2_KinState&>(n) = childData;
    n.nodeID_                     = childId;
    n.parentID_                   = parentId;
    n.cost_                       = parentNode.cost_ + edgeCost;
    if (nodes_.empty() || childId > nodes_.rbegin()->first)
        nodes_.emplace_hint(nodes_.end(), childId, n);
    else
        nodes_[childId] = n;
}
void MotionPrimitivesTreeSE2::rewire_node_parent(TNodeID nodeId, const MoveEdgeSE2_TPS& newEdgeFromParent)
{
    node_t& node     = nodes_.at(nodeId);
    auto&   oldEdges = edges_to_children[*node.parentID_];
    auto    it = std::find_if(oldEdges.begin(), oldEdges.end(), [&](const TEdgeInfo& e) { return e.id == nodeId; });
    if (it == oldEdges.end())
        throw std::runtime_error("[rewire_node_parent] Error: Could not find edge from former parent #" +
                                 std::to_string(*node.parentID_) + " -> node #" + std::to_string(nodeId));
    oldEdges.erase(it);
    const TNodeID parentId = newEdgeFromParent.parentId;
    edges_to_children[parentId].push_back(TEdgeInfo{nodeId, false, newEdgeFromParent});
    const cost_t newCost = nodes_.at(parentId).cost_ + newEdgeFromParent.cost;
    MPP_ASSERT(newCost <= node.cost_);
    node.parentID_ = parentId;
    node.cost_     = newCost;
}
const MoveEdgeSE2_TPS& MotionPrimitivesTreeSE2::edge_to_parent(TNodeID nodeId) const
{
    const node_t& node = nodes_.at(nodeId);
    for (const auto& e : edges_to_children.at(*node.parentID_))
        if (e.id == nodeId) return e.data;
    throw std::runtime_error("Could not find edge to parent for node #" + std::to_string(nodeId));
}
std::tuple<MotionPrimitivesTreeSE2::path_t, MotionPrimitivesTreeSE2::edge_sequence_t>
    MotionPrimitivesTreeSE2::backtrack_path(TNodeID target) const
{
    path_t          outPath;
    edge_sequence_t edgeList;
    auto            it = nodes_.find(target);
    if (it == nodes_.end()) throw std::runtime_error("backtrackPath: target_node not found in tree!");
    for (const node_t* node = &it->second;;)
    {
        outPath.push_front(*node);
        if (!node->parentID_.has_value()) break;
        edgeList.push_front(const_cast<MoveEdgeSE2_TPS*>(&edge_to_parent(node->nodeID_)));
        auto up = nodes_.find(*node->parentID_);
        if (up == nodes_.end()) throw std::runtime_error("backtrackPath: Node ID not found during tree traversal!");
        node = &up->second;
    }
    return {outPath, edgeList};
}

// ---- obstacle clouds ----------------------------------------------------------------------------------
bool PointCloud::load2D_from_text_file(const std::string& fileName)
{
    std::ifstream f(fileName);
    if (!f.is_open()) return false;
    xs.clear();
    ys.clear();
    std::string line;
    while (std::getline(f, line))
    {
        for (auto& c : line)
            if (c == ',' || c == ';') c = ' ';
        std::istringstream ls(line);
        float              x, y;
        if (ls >> x >> y) insertPoint(x, y);
    }
    return true;
}
void PointCloud::boundingBox(float& minx, float& miny, float& maxx, float& maxy) const
{
    if (xs.empty())
    {
        minx = miny = maxx = maxy = 0;
        return;
    }
    minx = maxx = xs[0];
    miny = maxy = ys[0];
    for (size_t i = 1; i < xs.size(); i++)
    {
        minx = std::min(minx, xs[i]);
        maxx = std::max(maxx, xs[i]);
        miny = std::min(miny, ys[i]);
        maxy = std::max(maxy, ys[i]);
    }
}
namespace
{
class StaticCloudSource final : public ObstacleSource
{
   public:
    explicit StaticCloudSource(PointCloud::Ptr pc) : pc_(std::move(pc)) {}
    PointCloud::Ptr obstacles() override { return pc_; }

   private:
    PointCloud::Ptr pc_;
};
}  // namespace
ObstacleSource::Ptr ObstacleSource::FromStaticPointcloud(const PointCloud::Ptr& pc)
{
    return std::make_shared<StaticCloudSource>(pc);
}

// ---- PTG set from an INI file ------------------------------------------------------------------------
void TrajectoriesAndRobotShape::initFromConfigFile(const ConfigFile& c, const std::string& s, mppb_ctx* buildCtx)
{
    const int count = c.read_int(s, "PTG_COUNT", 0, true);
    // the reference reads the polygon as float (TrajectoriesAndRobotShape.cpp:25-27)
    shape_xs = c.read_vector(s, "RobotModel_shape2D_xs");
    shape_ys = c.read_vector(s, "RobotModel_shape2D_ys");
    for (auto* v : {&shape_xs, &shape_ys})
        for (double& x : *v) x = static_cast<double>(static_cast<float>(x));
    if (shape_xs.size() != shape_ys.size())
        throw std::runtime_error(
            "Config parameters `RobotModel_shape2D_xs` and `RobotModel_shape2D_ys` must have the same length!");
    if (const double r = c.read_double(s, "RobotModel_circular_shape_radius", -1.0); r > 0) shape_radius = r;
    ptgs.clear();
    for (int n = 0; n < count; n++)
    {
        const std::string pre  = "PTG" + std::to_string(n) + "_";
        const std::string type = c.read_string(s, pre + "Type", "", true);
        auto              ptg  = ClassRegistry::instance().createPtg(type, c, s, pre, *this, buildCtx);
        if (!ptg) throw std::runtime_error("Unknown PTG class name: " + type);
        ptgs.push_back(std::move(ptg));
    }
    initialized_ = true;
}

// ---- edge_interpolated_path ------------------------------------------------------------------------------
void edge_interpolated_path(MoveEdgeSE2_TPS& edge, const TrajectoriesAndRobotShape& trs,
                            const std::optional<TPose2D>& reconstrRelPoseOpt, const std::optional<size_t>& ptg_stepOpt,
                            const std::optional<size_t>& numSegments)
{
    size_t nSeg = 0;
    if (numSegments.has_value())
        nSeg = *numSegments;
    else
    {
        MPP_ASSERT(edge.interpolatedPath.size() > 1);
        nSeg = edge.interpolatedPath.size();
    }
    if (!reconstrRelPoseOpt.has_value() || !ptg_stepOpt.has_value()) throw std::runtime_error("To do");
    const auto&              ptg      = trs.ptgs.at(edge.ptgIndex);
    const duration_seconds_t dt       = ptg->getPathStepDuration();
    const size_t             ptg_step = *ptg_stepOpt;
    auto&                    ip       = edge.interpolatedPath;
    ip.clear();
    ip[0 * dt] = TPose2D(0, 0, 0);
    // equal keys collapse: later samples overwrite earlier ones (std::map semantics of the reference)
    for (size_t i = 0; i < nSeg; i++)
    {
        const size_t iStep = ((i + 1) * ptg_step) / (nSeg + 2);
        ip[iStep * dt]     = ptg->getPathPose(static_cast<uint16_t>(edge.ptgPathIndex), static_cast<uint32_t>(iStep));
    }
    ip[ptg_step * dt]      = *reconstrRelPoseOpt;
    edge.estimatedExecTime = ptg_step * dt;
}

// ---- single-call seams of the reference's free functions ---------------------------------------------------------
void transform_pc_square_clipping(mppb_ctx* ctx, const PointCloud& inMap, const TPose2D& asSeenFrom, double maxDistXY, PointCloud& outMap,
                                  bool appendToOutMap)
{
    if (!ctx) throw std::runtime_error("transform_pc_square_clipping needs a libmpp_b200 context (no CPU fallback)");
    if (!appendToOutMap)
    {
        outMap.xs.clear();
        outMap.ys.clear();
    }
    const size_t       n = inMap.size();
    std::vector<float> ox(n), oy(n);
    const double       pose[3] = {asSeenFrom.x, asSeenFrom.y, asSeenFrom.phi};
    size_t             m       = 0;
    if (mppb_transform_pc_square_clipping(ctx, inMap.xs.data(), inMap.ys.data(), n, pose, maxDistXY, ox.data(), oy.data(), &m) != MPPB_OK)
        throw std::runtime_error(std::string("libmpp_b200: ") + mppb_last_error(ctx));
    outMap.xs.insert(outMap.xs.end(), ox.begin(), ox.begin() + static_cast<std::ptrdiff_t>(m));
    outMap.ys.insert(outMap.ys.end(), oy.begin(), oy.begin() + static_cast<std::ptrdiff_t>(m));
}

distance_t tp_obstacles_single_path(mppb_ctx* ctx, trajectory_index_t k, const PointCloud& localObstacles, const ptg_t& ptg)
{
    if (!ctx) throw std::runtime_error("tp_obstacles_single_path needs a libmpp_b200 context (no CPU fallback)");
    auto ck = [ctx](int rc) {
        if (rc != MPPB_OK) throw std::runtime_error(std::string("libmpp_b200: ") + mppb_last_error(ctx));
    };
    MPP_ASSERT(k >= 0 && k < static_cast<int>(ptg.getPathCount()));
    // the context's PTG set becomes this PTG alone (its tables travel once per call: single-call interface)
    ck(mppb_clear_ptgs(ctx));
    static const float none = 0;
    const size_t       n    = localObstacles.size();
    ck(mppb_set_obstacles(ctx, n ? localObstacles.xs.data() : &none, n ? localObstacles.ys.data() : &none, n, 0.0, 0));
    // the points are local already: the node is the origin, and the clipping square is wide open
    const double     origin[3] = {0, 0, 0};
    const double     wide      = 1e30;
    const distance_t init      = ptg.initTPObstacleSingle(static_cast<uint16_t>(k));
    if (auto* dd = dynamic_cast<const DiffDrive_C*>(&ptg))
    {
        const mppb_ptg_grid_desc d = dd->gridDesc();
        ck(mppb_set_ptg_grid(ctx, 0, &d));
        std::vector<float> row(static_cast<size_t>(d.num_paths));
        ck(mppb_eval_nodes(ctx, 1, origin, wide, row.data()));
        return std::min(init, static_cast<double>(row[static_cast<size_t>(k)]));
    }
    if (auto* hb = dynamic_cast<const HolonomicBlend*>(&ptg))
    {
        const mppb_ptg_holo_desc d = hb->holoDesc();
        ck(mppb_set_ptg_holo(ctx, 0, &d));
        mppb_holo_cand c;
        hb->holoParams(static_cast<uint16_t>(k), c);
        c.node      = 0;
        c.ptg_index = 0;
        double out  = 0;
        ck(mppb_eval_holo_candidates(ctx, 1, origin, 1, &c, wide, &out));
        return std::min(init, out);
    }
    throw std::runtime_error("tp_obstacles_single_path: PTG of a kind the GPU library does not know");
}

// ---- refine_trajectory --------------------------------------------------------------------------------------
void refine_trajectory(const MotionPrimitivesTreeSE2::path_t& inPath, MotionPrimitivesTreeSE2::edge_sequence_t& edgesToRefine,
                       const TrajectoriesAndRobotShape& ptgInfo, const double ptg_tolerance_dist)
{
    const size_t nEdges = edgesToRefine.size();
    MPP_ASSERT(inPath.size() == nEdges + 1);
    auto itPath = inPath.begin();
    auto itEdge = edgesToRefine.begin();
    for (size_t i = 0; i < nEdges; i++, ++itPath, ++itEdge)
    {
        MoveEdgeSE2_TPS& edge = **itEdge;
        const auto&      ptg  = ptgInfo.ptgs.at(edge.ptgIndex);
        ptg->updateNavDynamicState(edge.getPTGDynState());
        if (ptg->isSpeedTrimmable()) ptg->trimmableSpeed_ = edge.ptgTrimmableSpeed;
        auto itNext = itPath;
        ++itNext;
        const TPose2D deltaNodes = itNext->pose - itPath->pose;
        if (deltaNodes.x == 0 && deltaNodes.y == 0) continue;
        int    newK        = -1;
        double newNormDist = 0;
        if (!ptg->inverseMap_WS2TP(deltaNodes.x, deltaNodes.y, newK, newNormDist, ptg_tolerance_dist))
        {
            std::cerr << "[refine_trajectory] Warning: Could not refine this path segment:\n"
                      << " - deltaNodes: " << pose_str(deltaNodes) << "\n - edge: " << edge.asString() << std::endl;
            continue;
        }
        const distance_t newDist    = newNormDist * ptg->getRefDistance();
        uint32_t         newPtgStep = 0;
        ptg->getPathStepForDist(static_cast<uint16_t>(newK), newDist, newPtgStep);
        edge.ptgPathIndex = static_cast<int16_t>(newK);
        edge.ptgDist      = newDist;
        edge.ptgStep      = newPtgStep;
        edge_interpolated_path(edge, ptgInfo, deltaNodes, newPtgStep);
    }
}

// ---- plan_to_trajectory / save_to_txt -----------------------------------------------------------------------
trajectory_t plan_to_trajectory(const MotionPrimitivesTreeSE2::edge_sequence_t& planEdges,
                                const TrajectoriesAndRobotShape& ptgInfo, const duration_seconds_t samplePeriod)
{
    MPP_ASSERT(ptgInfo.initialized());
    MPP_ASSERT(samplePeriod > 0.);
    trajectory_t       out;
    duration_seconds_t edgeStartTime = .0;
    TPose2D            endOfLastEdge(0, 0, 0);
    for (const MoveEdgeSE2_TPS* edge : planEdges)
    {
        const auto&  ptg    = ptgInfo.ptgs.at(edge->ptgIndex);
        const double ptg_dt = ptg->getPathStepDuration();
        MPP_ASSERT(ptg_dt > 0.);
        ptg->updateNavDynamicState(edge->getPTGDynState());
        if (ptg->isSpeedTrimmable()) ptg->trimmableSpeed_ = edge->ptgTrimmableSpeed;
        const uint16_t k         = static_cast<uint16_t>(edge->ptgPathIndex);
        uint32_t       finalStep = 0;
        const bool     ok        = ptg->getPathStepForDist(k, edge->ptgDist, finalStep);
        MPP_ASSERT(ok);
        const uint32_t stepIncr = std::max<uint32_t>(1, static_cast<uint32_t>(round_i(samplePeriod / ptg_dt)));
        for (uint32_t step = 0;; step += stepIncr)
        {
            if (step > finalStep) step = finalStep;
            trajectory_state_t& ts = out[edgeStartTime + step * ptg_dt];
            ts.state.pose          = endOfLastEdge + ptg->getPathPose(k, step);
            ts.state.vel           = ptg->getPathTwist(k, step);
            ts.ptgIndex            = edge->ptgIndex;
            ts.ptgPathIndex        = edge->ptgPathIndex;
            ts.ptgStep             = step;
            if (step == finalStep) break;
        }
        edgeStartTime += finalStep * ptg_dt;
        endOfLastEdge = edge->stateTo.pose - planEdges.front()->stateFrom.pose;
    }
    return out;
}

std::string trajectory_as_text(const trajectory_t& traj)
{
    std::string s;
    char        buf[512];
    snprintf(buf, sizeof(buf), "%% %15s  %15s %15s %15s  %15s %15s %15s %15s %15s %15s\n", "Time [s]", "x_global [m]",
             "y_global [m]", "phi [rad]", "vx_local [m]", "vy_local[m]", "omega [rad/s]", "PTG_index", "PTG_traj_index",
             "PTG_step");
    s += buf;
    for (const auto& kv : traj)
    {
        const trajectory_state_t& ts = kv.second;
        snprintf(buf, sizeof(buf), "%15.03f %15.03f %15.03f %15.03f  %15.03f %15.03f %15.03f   %15u %15u %15u\n", kv.first,
                 ts.state.pose.x, ts.state.pose.y, ts.state.pose.phi, ts.state.vel.vx, ts.state.vel.vy, ts.state.vel.omega,
                 static_cast<unsigned int>(ts.ptgIndex), static_cast<unsigned int>(ts.ptgPathIndex),
                 static_cast<unsigned int>(ts.ptgStep));
        s += buf;
    }
    return s;
}
bool save_to_txt(const trajectory_t& traj, const std::string& fileName)
{
    std::ofstream f(fileName);
    if (!f.is_open()) return false;
    f << trajectory_as_text(traj);
    return true;
}

// ---- class registry ---------------------------------------------------------------------------------------------
ClassRegistry& ClassRegistry::instance()
{
    static ClassRegistry r;
    return r;
}
// the classes of this library (src/registerClasses.cpp:16-37 of the reference)
ClassRegistry::ClassRegistry()
{
    const PlannerFactory astar = [](mppb_ctx* ctx) { return std::static_pointer_cast<Planner>(std::make_shared<TPS_Astar>(ctx)); };
    planners_["mpp::TPS_Astar"]      = astar;
    planners_["mpp::TPS_Astar_B200"] = astar;
    // CPTG_DiffDrive_C names MRPT's upstream class; it is served by the in-repo fork's semantics
    const PtgFactory diffDrive = [](const ConfigFile& c, const std::string& s, const std::string& pre, const TrajectoriesAndRobotShape& shape,
                                    mppb_ctx* buildCtx) -> std::shared_ptr<ptg_t> {
        DiffDriveCParams p;
        p.num_paths   = static_cast<unsigned>(c.read_int(s, pre + "num_paths", 0, true));
        p.refDistance = c.read_double(s, pre + "refDistance", 0, true);
        p.resolution  = c.read_double(s, pre + "resolution", 0.05, true);
        p.v_max_mps   = c.read_double(s, pre + "v_max_mps", 0, true);
        p.w_max_rps   = c.read_double(s, pre + "w_max_dps", 0, true) * M_PI / 180.0;
        p.K           = c.read_double(s, pre + "K", 1.0, true);
        p.turningRadiusReference = c.read_double(s, pre + "turningRadiusReference", 0.10);
        if (shape.shape_xs.size() < 3) throw std::runtime_error("PTG DiffDrive_C needs a polygonal robot shape");
        p.shape_x = shape.shape_xs;
        p.shape_y = shape.shape_ys;
        return std::make_shared<DiffDrive_C>(p, buildCtx);
    };
    for (const char* n : {"CPTG_DiffDrive_C", "mpp::ptg::DiffDrive_C", "mpp::DiffDrive_C"}) ptgs_[n] = diffDrive;
    const PtgFactory holo = [](const ConfigFile& c, const std::string& s, const std::string& pre, const TrajectoriesAndRobotShape& shape,
                               mppb_ctx*) -> std::shared_ptr<ptg_t> {
        HolonomicBlendParams p;
        p.num_paths   = static_cast<unsigned>(c.read_int(s, pre + "num_paths", 0, true));
        p.refDistance = c.read_double(s, pre + "refDistance", 0, true);
        p.T_ramp_max  = c.read_double(s, pre + "T_ramp_max", 0, true);
        p.v_max_mps   = c.read_double(s, pre + "v_max_mps", 0, true);
        p.w_max_rps   = c.read_double(s, pre + "w_max_dps", 0, true) * M_PI / 180.0;
        p.expr_V      = c.read_string(s, pre + "expr_V", "V_MAX");
        p.expr_W      = c.read_string(s, pre + "expr_W", "W_MAX");
        if (!shape.shape_radius) throw std::runtime_error("PTG HolonomicBlend needs RobotModel_circular_shape_radius");
        p.robot_radius = *shape.shape_radius;
        return std::make_shared<HolonomicBlend>(p);
    };
    for (const char* n : {"mpp::ptg::HolonomicBlend", "HolonomicBlend"}) ptgs_[n] = holo;
}
void ClassRegistry::registerPlanner(const std::string& name, PlannerFactory f) { planners_[name] = std::move(f); }
void ClassRegistry::registerPtg(const std::string& name, PtgFactory f) { ptgs_[name] = std::move(f); }
std::shared_ptr<Planner> ClassRegistry::createPlanner(const std::string& name, mppb_ctx* ctx) const
{
    const auto it = planners_.find(name);
    return it == planners_.end() ? nullptr : it->second(ctx);
}
std::shared_ptr<ptg_t> ClassRegistry::createPtg(const std::string& name, const ConfigFile& cfg, const std::string& section,
                                                const std::string& keyPrefix, const TrajectoriesAndRobotShape& shape,
                                                mppb_ctx* buildCtx) const
{
    const auto it = ptgs_.find(name);
    return it == ptgs_.end() ? nullptr : it->second(cfg, section, keyPrefix, shape, buildCtx);
}
std::vector<std::string> ClassRegistry::plannerNames() const
{
    std::vector<std::string> v;
    for (const auto& kv : planners_) v.push_back(kv.first);
    return v;
}
std::vector<std::string> ClassRegistry::ptgNames() const
{
    std::vector<std::string> v;
    for (const auto& kv : ptgs_) v.push_back(kv.first);
    return v;
}
bool loadPluginModules(const std::string& list, std::string* error)
{
    std::stringstream ss(list);
    std::string       lib;
    while (std::getline(ss, lib, ','))
    {
        while (!lib.empty() && (lib.front() == ' ' || lib.front() == '\t')) lib.erase(lib.begin());
        while (!lib.empty() && (lib.back() == ' ' || lib.back() == '\t')) lib.pop_back();
        if (lib.empty()) continue;
        if (!dlopen(lib.c_str(), RTLD_NOW | RTLD_GLOBAL))
        {
            const char* why = dlerror();  // (cleared by the call: read once)
            if (error) *error = std::string("cannot load plug-in module '") + lib + "': " + (why ? why : "?");
            return false;
        }
    }
    return true;
}

std::string TimeLogger::getStatsAsText() const
{
    std::string out;
    char        buf[256];
    snprintf(buf, sizeof(buf), "%-45s %12s %14s %14s\n", "section", "calls", "total [s]", "mean [s]");
    out += buf;
    for (const auto& kv : sections_)
    {
        snprintf(buf, sizeof(buf), "%-45s %12zu %14.6f %14.3e\n", kv.first.c_str(), kv.second.calls, kv.second.seconds,
                 kv.second.calls ? kv.second.seconds / static_cast<double>(kv.second.calls) : 0.0);
        out += buf;
    }
    return out;
}

bool within_bbox(const TPose2D& p, const TPose2D& max, const TPose2D& min)
{
    return p.x < max.x && p.y < max.y && p.phi < max.phi + 1e-6 && p.x > min.x && p.y > min.y && p.phi > min.phi - 1e-6;
}
bool within_bbox(const TPoint2D& p, const TPose2D& max, const TPose2D& min)
{
    return p.x < max.x && p.y < max.y && p.x > min.x && p.y > min.y;
}

}  // namespace mpp
