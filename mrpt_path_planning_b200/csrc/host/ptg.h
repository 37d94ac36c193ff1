// Host-side PTG objects of the mpp:: mirror (product code).
//
// Mirrors the part of mrpt::nav::CParameterizedTrajectoryGenerator that the planner calls
// (reference call sites: src/algos/TPS_Astar.cpp:120,279-292,578-795) and the in-repo PTG
// families (src/ptgs/DiffDrive_C.cpp, DiffDriveCollisionGridBased.cpp, HolonomicBlend.cpp).
// The per-point collision arithmetic (updateTPObstacleSingle) is NOT implemented on the
// host: it lives in the CUDA kernels, which consume the tables these classes build.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "../../../include/mppb.h"
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

   protected:
    virtual void     onNewNavDynamicState() {}
    unsigned         numPaths_    = 0;
    double           refDistance_ = 0;
    bool             initialized_ = false;
    TNavDynamicState dynState_;
    int              target_k_ = -1;
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
    explicit DiffDrive_C(const DiffDriveCParams& p);
    PtgKind kind() const override { return PtgKind::CollisionGrid; }

    size_t   getPathStepCount(uint16_t k) const override { return pathBegin_.at(k + 1) - pathBegin_.at(k); }
    TPose2D  getPathPose(uint16_t k, uint32_t step) const override;
    double   getPathDist(uint16_t k, uint32_t step) const override;
    bool     getPathStepForDist(uint16_t k, double dist, uint32_t& step) const override;
    double   getPathStepDuration() const override { return stepDuration_; }
    TTwist2D getPathTwist(uint16_t k, uint32_t step) const override;
    bool     inverseMap_WS2TP(double x, double y, int& k, double& d, double tol = 0.10) const override;
    bool     isSpeedTrimmable() const override { return true; }
    bool     isPointInsideRobotShape(double x, double y) const override;
    double   getMaxRobotRadius() const override { return maxRadius_; }

    /** descriptor of the collision grid for mppb_set_ptg_grid (pointers owned by this object) */
    mppb_ptg_grid_desc gridDesc() const;

    const DiffDriveCParams& params() const { return p_; }

   private:
    void simulate();
    void buildCollisionGrid();
    size_t at(uint16_t k, uint32_t step) const;

    DiffDriveCParams p_;
    double           maxRadius_    = 0;
    double           stepDuration_ = 0;
    // trajectories: flat SoA, path k owns [pathBegin_[k], pathBegin_[k+1])
    std::vector<size_t> pathBegin_;
    std::vector<float>  tx_, ty_, tphi_, tdist_, tv_, tw_;
    // collision grid (CDynamicGrid geometry) in CSR form, entries sorted by k inside a cell
    double                gxmin_ = 0, gymin_ = 0, gres_ = 0;
    int                   gnx_ = 0, gny_ = 0;
    std::vector<uint32_t> cellOffsets_;
    std::vector<uint16_t> entryK_;
    std::vector<float>    entryD_;
};

}  // namespace mpp
