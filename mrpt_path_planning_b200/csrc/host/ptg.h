// Host-side PTG objects of the mpp:: mirror (product code).
//
// Mirrors the part of mrpt::nav::CParameterizedTrajectoryGenerator that the planner calls
// (reference call sites: src/algos/TPS_Astar.cpp:120,279-292,578-795) and the in-repo PTG
// families (src/ptgs/DiffDrive_C.cpp, DiffDriveCollisionGridBased.cpp, HolonomicBlend.cpp).
// The per-point collision arithmetic (updateTPObstacleSingle) is NOT implemented on the
// host: it lives in the CUDA kernels, which consume the tables these classes build.
#pragma once
#include <array>
#include <atomic>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "../../../include/mppb.h"
#include "expr.h"
#include "mpp_math.h"

namespace mpp
{
struct TNavDynamicState
{
    TTwist2D curVelLocal;
    TPose2D  relTarget{20.0, 0, 0};
    double   targetRelSpeed = 0;
    bool     operator==(const TNavDynamicState& o) const
    {
        return curVelLocal == o.curVelLocal && relTarget == o.relTarget && targetRelSpeed == o.targetRelSpeed;
    }
};

enum class PtgKind
{
    CollisionGrid,
    HolonomicBlend
};

/** Abstract PTG (the "ptg_t" of the reference). */
class ptg_t
{
   public:
    virtual ~ptg_t() = default;
    virtual PtgKind kind() const = 0;

    uint16_t getPathCount() const { return static_cast<uint16_t>(numPaths_); }
    double   getRefDistance() const { return refDistance_; }
    bool     isInitialized() const { return initialized_; }

    double index2alpha(unsigned k) const { return M_PI * (-1.0 + 2.0 * (k + 0.5) / numPaths_); }
    int    alpha2index(double alpha) const
    {
        int k = round_i(0.5 * (numPaths_ * (1.0 + alpha / M_PI) - 1.0));
        if (k < 0) k = 0;
        if (k >= static_cast<int>(numPaths_)) k = static_cast<int>(numPaths_) - 1;
        return k;
    }

    virtual size_t   getPathStepCount(uint16_t k) const                               = 0;
    virtual TPose2D  getPathPose(uint16_t k, uint32_t step) const                     = 0;
    virtual double   getPathDist(uint16_t k, uint32_t step) const                     = 0;
    virtual bool     getPathStepForDist(uint16_t k, double dist, uint32_t& step) const = 0;
    virtual double   getPathStepDuration() const                                      = 0;
    virtual TTwist2D getPathTwist(uint16_t k, uint32_t step) const;
    virtual bool     inverseMap_WS2TP(double x, double y, int& k, double& d, double tol = 0.10) const = 0;
    virtual bool     supportSpeedAtTarget() const { return false; }
    virtual bool     isSpeedTrimmable() const { return false; }
    /** false for a SpeedTrimmablePTG whose paths ignore trimmableSpeed_ (DiffDrive_C: its steering function does not
     *  read it, DiffDrive_C.cpp:70-80): the planner's per-speed TPS points are then exact copies of each other. */
    virtual bool     pathsDependOnTrimmableSpeed() const { return isSpeedTrimmable(); }
    virtual bool     isPointInsideRobotShape(double x, double y) const = 0;
    virtual double   getMaxRobotRadius() const                         = 0;

    /** value tp_obstacles_single_path() starts from: min(refDistance, dist of the last step) */
    double initTPObstacleSingle(uint16_t k) const
    {
        return std::min(refDistance_, getPathDist(k, static_cast<uint32_t>(getPathStepCount(k) - 1)));
    }

    void                    updateNavDynamicState(const TNavDynamicState& ds, bool force = false);
    const TNavDynamicState& getCurrentNavDynamicState() const { return dynState_; }

    /** speed modulation in (0,1] of SpeedTrimmablePTG (only meaningful if isSpeedTrimmable()) */
    double trimmableSpeed_ = 1.0;

    /** A PTG object with its own mutable state (dynamic state, trimmed speed, caches) that shares
     *  the immutable tables of this one. Independent planning queries each use their own clone
     *  so that they can run interleaved / concurrently; `this` must outlive the clone. */
    virtual std::unique_ptr<ptg_t> cloneForQuery() const = 0;

    /** Process-wide unique id of this object's tables: every constructed PTG object draws a fresh one, so a new
     *  object that happens to live at the address of a destroyed one is never mistaken for it by a planner that
     *  keeps tables resident on the GPU. */
    uint64_t revision() const { return revision_; }

   protected:
    virtual void     onNewNavDynamicState() {}
    unsigned         numPaths_    = 0;
    double           refDistance_ = 0;
    bool             initialized_ = false;
    TNavDynamicState dynState_;
    int              target_k_ = -1;

   private:
    static uint64_t newRevision()
    {
        static std::atomic<uint64_t> next{1};
        return next.fetch_add(1);
    }
    uint64_t revision_ = newRevision();
};

/** Parameters of mpp::ptg::DiffDrive_C (ini keys of share/ptgs_ackermann_vehicle.ini). */
struct DiffDriveCParams
{
    unsigned            num_paths              = 121;
    double              refDistance            = 5.0;
    double              resolution             = 0.05;
    double              v_max_mps              = 1.0;
    double              w_max_rps              = 1.0;
    double              K                      = 1.0;
    double              turningRadiusReference = 0.10;
    std::vector<double> shape_x, shape_y;
};

/** Constant-curvature "C" PTG with a precomputed collision look-up grid.
 *  Tables are laid out flat (SoA per path, CSR per grid cell) so they upload as-is. */
class DiffDrive_C : public ptg_t
{
   public:
    /** buildCtx == nullptr: collision grid from the multi-threaded host builder; otherwise it is built on that
     *  context's GPU (mppb_build_collision_grid). Both give the same table. */
    explicit DiffDrive_C(const DiffDriveCParams& p, mppb_ctx* buildCtx = nullptr);
    PtgKind kind() const override { return PtgKind::CollisionGrid; }

    size_t   getPathStepCount(uint16_t k) const override { return pathBegin_.at(k + 1) - pathBegin_.at(k); }
    TPose2D  getPathPose(uint16_t k, uint32_t step) const override;
    double   getPathDist(uint16_t k, uint32_t step) const override;
    bool     getPathStepForDist(uint16_t k, double dist, uint32_t& step) const override;
    double   getPathStepDuration() const override { return stepDuration_; }
    TTwist2D getPathTwist(uint16_t k, uint32_t step) const override;
    bool     inverseMap_WS2TP(double x, double y, int& k, double& d, double tol = 0.10) const override;
    bool     isSpeedTrimmable() const override { return true; }
    bool     pathsDependOnTrimmableSpeed() const override { return false; }
    bool     isPointInsideRobotShape(double x, double y) const override;
    double   getMaxRobotRadius() const override { return maxRadius_; }

    /** descriptor of the collision grid for mppb_set_ptg_grid (pointers owned by this object) */
    mppb_ptg_grid_desc gridDesc() const;
    /** descriptor of the sampled trajectories for mppb_set_ptg_paths (pointers owned by this object) */
    mppb_ptg_paths_desc pathsDesc() const;
    std::unique_ptr<ptg_t> cloneForQuery() const override;

    const DiffDriveCParams& params() const { return p_; }

   private:
    void simulate();
    void buildCollisionGrid();
    void buildCollisionGridOn(mppb_ctx* ctx);
    size_t at(uint16_t k, uint32_t step) const;

    DiffDriveCParams p_;
    double           maxRadius_    = 0;
    double           stepDuration_ = 0;
    // trajectories: flat SoA, path k owns [pathBegin_[k], pathBegin_[k+1])
    std::vector<size_t> pathBegin_;
    std::vector<float>  tx_, ty_, tphi_, tdist_, tv_, tw_;
    std::vector<uint32_t> pathBegin32_;
    // cos/sin of every step's heading, (double)phi through the C library: getPathTwist() rotates the step's (v, 0, w) by
    // it, and the GPU expansion kernel needs the same values (device cos/sin differ from glibc's in the last bits)
    std::vector<double> tcos_, tsin_;
    // collision grid (CDynamicGrid geometry) in CSR form, entries sorted by k inside a cell
    double                gxmin_ = 0, gymin_ = 0, gres_ = 0;
    int                   gnx_ = 0, gny_ = 0;
    std::vector<uint32_t> cellOffsets_;
    std::vector<uint16_t> entryK_;
    std::vector<float>    entryD_;
};

/** Parameters of mpp::ptg::HolonomicBlend (ini keys of share/ptgs_holonomic_robot.ini). */
struct HolonomicBlendParams
{
    unsigned    num_paths    = 191;
    double      refDistance  = 5.0;
    double      T_ramp_max   = 1.0;
    double      v_max_mps    = 1.0;
    double      w_max_rps    = 1.0;
    double      robot_radius = 0.15;
    std::string expr_V = "V_MAX", expr_W = "W_MAX";
};

/** Closed-form holonomic PTG with a velocity ramp (src/ptgs/HolonomicBlend.cpp). The host side
 *  provides the path geometry the planner asks for; the per-obstacle collision arithmetic
 *  (updateTPObstacleSingle) runs on the GPU from the per-candidate parameters of holoParams(). */
class HolonomicBlend : public ptg_t
{
   public:
    explicit HolonomicBlend(const HolonomicBlendParams& p);
    HolonomicBlend(const HolonomicBlend&)            = delete;
    HolonomicBlend& operator=(const HolonomicBlend&) = delete;

    PtgKind kind() const override { return PtgKind::HolonomicBlend; }

    static constexpr double PATH_TIME_STEP = 10e-3;
    static constexpr double eps            = 1e-4;

    size_t   getPathStepCount(uint16_t k) const override;
    TPose2D  getPathPose(uint16_t k, uint32_t step) const override;
    double   getPathDist(uint16_t k, uint32_t step) const override;
    bool     getPathStepForDist(uint16_t k, double dist, uint32_t& step) const override;
    double   getPathStepDuration() const override { return PATH_TIME_STEP; }
    bool     inverseMap_WS2TP(double x, double y, int& k, double& d, double tol = 0.10) const override;
    bool     supportSpeedAtTarget() const override { return true; }
    bool     isSpeedTrimmable() const override { return true; }
    bool     isPointInsideRobotShape(double x, double y) const override { return ::hypot(x, y) < p_.robot_radius; }
    double   getMaxRobotRadius() const override { return p_.robot_radius; }
    std::unique_ptr<ptg_t> cloneForQuery() const override { return std::make_unique<HolonomicBlend>(p_); }

    /** (vxi, vyi, vxf, vyf, vf, T_ramp) of path k at the current dynamic state and trimmed speed */
    void holoParams(uint16_t k, mppb_holo_cand& out) const;
    mppb_ptg_holo_desc holoDesc() const
    {
        return {static_cast<int32_t>(numPaths_), refDistance_, p_.robot_radius, p_.v_max_mps};
    }
    const HolonomicBlendParams& params() const { return p_; }

   protected:
    void onNewNavDynamicState() override;

   private:
    struct Internal
    {
        double vxi, vyi, vxf, vyf, vf, wf, T_ramp;
    };
    Internal      paramsForDir(double dir) const;
    double        pathDistAt(uint32_t step, double T_ramp, double vxf, double vyf) const;
    static double rampDistance(double k2, double k4, double vxi, double vyi, double t);
    static double rampIntegral(double T, double a, double b, double c);

    HolonomicBlendParams     p_;
    RuntimeExpr              exprV_, exprW_;
    mutable double           exprDir_    = 0;
    double                   targetDir_  = 0;
    double                   targetDist_ = 0;
    mutable std::vector<int> stepCountCache_;
    // paramsForDir() costs two user expressions plus cos/sin of the direction, and the planner asks for the same
    // (k, speed) several times per expansion (candidate parameters, then distance, pose and initial TP-obstacle value
    // of every sample timestamp) and for the same few dozen (k, speed) in every expansion. Its results are kept in a
    // direct-mapped table keyed by (dir, trimmableSpeed_); a slot that another key took is simply recomputed. The
    // initial velocity is not part of an entry (it is the current state's). If the expressions read the dynamic
    // state, every entry is dropped (generation counter) when the state changes; the shipped configurations
    // (functions of dir and the trimmed speed) keep theirs for the lifetime of the object.
    struct Memo
    {
        double   dir = 0, speed = 0;
        Internal q{};
        uint32_t gen = 0;
    };
    mutable std::array<Memo, 512> memo_{};
    mutable uint32_t              memoGen_ = 1;
    bool                          exprsReadDynState_ = true;
};

}  // namespace mpp
