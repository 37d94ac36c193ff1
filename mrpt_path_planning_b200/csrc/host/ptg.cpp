// Host-side PTG table builders (product code). See ptg.h.
//
// DiffDrive_C: the trajectory tables and the collision look-up grid are what the reference
// builds in DiffDriveCollisionGridBased::simulateTrajectories (.cpp:103-176) and
// internal_initialize (.cpp:657-754); here they are produced directly in the flat layouts
// the GPU consumes (SoA paths, CSR grid). The float/double mix of the reference arithmetic
// is kept operation by operation, because a one-ulp change moves table entries.
#include "ptg.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <limits>
#include <thread>

namespace mpp
{
// ---------------------------------------------------------------------------------------
// ptg_t
// ---------------------------------------------------------------------------------------
TTwist2D ptg_t::getPathTwist(uint16_t k, uint32_t step) const
{
    if (step == 0) return dynState_.curVelLocal;
    const TPose2D a = getPathPose(k, step - 1), b = getPathPose(k, step);
    const double  dt = getPathStepDuration();
    return {(b.x - a.x) / dt, (b.y - a.y) / dt, (b.phi - a.phi) / dt};
}

void ptg_t::updateNavDynamicState(const TNavDynamicState& ds, bool force)
{
    if (!force && ds == dynState_) return;
    dynState_ = ds;
    target_k_ = -1;
    onNewNavDynamicState();
    if (supportSpeedAtTarget())
    {
        int    k = 0;
        double d = 0;
        if (inverseMap_WS2TP(ds.relTarget.x, ds.relTarget.y, k, d, 1.0) && d > 0.01 && d < 0.99)
        {
            target_k_ = k;
            onNewNavDynamicState();
        }
    }
}

// ---------------------------------------------------------------------------------------
// DiffDrive_C
// ---------------------------------------------------------------------------------------
namespace
{
// crossing-number point-in-polygon (TPolygon2D::contains)
bool pnpoly(const double* vx, const double* vy, size_t n, double px, double py)
{
    bool in = false;
    for (size_t i = 0, j = n - 1; i < n; j = i++)
        if (((vy[i] > py) != (vy[j] > py)) && (px < (vx[j] - vx[i]) * (py - vy[i]) / (vy[j] - vy[i]) + vx[i]))
            in = !in;
    return in;
}

// CDynamicGrid::setSize extents snapping
struct GridGeom
{
    double xmin, ymin, res;
    int    nx, ny;
    GridGeom(double x0, double x1, double y0, double y1, double r)
    {
        xmin             = r * round_i(x0 / r);
        ymin             = r * round_i(y0 / r);
        const double xmx = r * round_i(x1 / r), ymx = r * round_i(y1 / r);
        res = r;
        nx  = round_i((xmx - xmin) / r);
        ny  = round_i((ymx - ymin) / r);
    }
    int    x2idx(double x) const { return static_cast<int>((x - xmin) / res); }
    int    y2idx(double y) const { return static_cast<int>((y - ymin) / res); }
    double idx2x(int i) const { return xmin + (i + 0.5) * res; }
    double idx2y(int i) const { return ymin + (i + 0.5) * res; }
};
}  // namespace

DiffDrive_C::DiffDrive_C(const DiffDriveCParams& p, mppb_ctx* buildCtx) : p_(p)
{
    MPP_ASSERT(p_.shape_x.size() == p_.shape_y.size() && p_.shape_x.size() >= 3);
    MPP_ASSERT(p_.shape_x.size() <= MPPB_MAX_SHAPE_VERTS);
    MPP_ASSERT(p_.refDistance > 0 && p_.v_max_mps > 0 && p_.w_max_rps > 0 && p_.resolution > 0);
    MPP_ASSERT(p_.num_paths > 0 && p_.num_paths < 65536);
    numPaths_    = p_.num_paths;
    refDistance_ = p_.refDistance;
    for (size_t i = 0; i < p_.shape_x.size(); i++)
        maxRadius_ = std::max(maxRadius_, std::sqrt(p_.shape_x[i] * p_.shape_x[i] + p_.shape_y[i] * p_.shape_y[i]));
    simulate();
    if (buildCtx)
        buildCollisionGridOn(buildCtx);
    else
        buildCollisionGrid();
    initialized_ = true;
}

// Collision grid on the GPU: the paths go up, the CSR comes back (csrc/colgrid_build.cu).
void DiffDrive_C::buildCollisionGridOn(mppb_ctx* ctx)
{
    std::vector<uint32_t> begin(pathBegin_.size());
    for (size_t i = 0; i < pathBegin_.size(); i++) begin[i] = static_cast<uint32_t>(pathBegin_[i]);
    mppb_colgrid_build_desc d{};
    d.num_paths    = static_cast<int32_t>(numPaths_);
    d.path_begin   = begin.data();
    d.x            = tx_.data();
    d.y            = ty_.data();
    d.phi          = tphi_.data();
    d.dist         = tdist_.data();
    d.ref_distance = refDistance_;
    d.resolution   = p_.resolution;
    d.shape_nverts = static_cast<int32_t>(p_.shape_x.size());
    d.shape_x      = p_.shape_x.data();
    d.shape_y      = p_.shape_y.data();
    mppb_colgrid g{};
    if (mppb_build_collision_grid(ctx, &d, &g) != MPPB_OK)
        throw std::runtime_error(std::string("mppb_build_collision_grid: ") + mppb_last_error(ctx));
    gxmin_ = g.x_min;
    gymin_ = g.y_min;
    gres_  = g.res;
    gnx_   = g.nx;
    gny_   = g.ny;
    cellOffsets_.assign(g.cell_offsets, g.cell_offsets + static_cast<size_t>(g.nx) * g.ny + 1);
    entryK_.assign(g.entry_k, g.entry_k + g.n_entries);
    entryD_.assign(g.entry_d, g.entry_d + g.n_entries);
}

// Euler integration with float state, dt = 1e-3 s, until dist >= refDistance, t >= 100 s or
// |turned| >= 1.95*pi. cos()/sin() of the float heading are the C double functions (what
// unqualified cos(float) resolves to under <cmath>), so each increment is formed in double
// and rounded once on the store.
void DiffDrive_C::simulate()
{
    const float max_time = 100.0f, max_dist = static_cast<float>(refDistance_), dt = 1.0e-3f;
    stepDuration_ = dt;
    const int sgn = sign_of(p_.K);
    // paths are independent: each one is integrated into its own buffer by a pool of host threads,
    // then the buffers are concatenated in k order
    struct Path
    {
        std::vector<float> x, y, phi, dist, v, w;
    };
    std::vector<Path>     paths(numPaths_);
    std::atomic<unsigned> next{0};
    auto                  worker = [&]() {
        for (;;)
        {
            const unsigned k = next.fetch_add(1);
            if (k >= numPaths_) break;
            Path&       P     = paths[k];
            const float alpha = static_cast<float>(index2alpha(k));
            const float v     = static_cast<float>(p_.v_max_mps * sgn);
            const float w     = static_cast<float>((alpha / M_PI) * p_.w_max_rps * sgn);
            float       t = 0, dist = 0, turned = 0, x = 0, y = 0, phi = 0;
            auto        push = [&](float pv, float pw) {
                P.x.push_back(x);
                P.y.push_back(y);
                P.phi.push_back(phi);
                P.dist.push_back(dist);
                P.v.push_back(pv);
                P.w.push_back(pw);
            };
            push(0.f, 0.f);
            bool moved = false;
            while (t < max_time && dist < max_dist && std::fabs(static_cast<double>(turned)) < 1.95 * M_PI)
            {
                x += std::cos(static_cast<double>(phi)) * v * dt;
                y += std::sin(static_cast<double>(phi)) * v * dt;
                phi += w * dt;
                turned += w * dt;
                const double wr   = w * p_.turningRadiusReference;
                const float  vtps = static_cast<float>(std::sqrt((v * v) + wr * wr));
                dist += vtps * dt;
                t += dt;
                P.v.back() = v;  // the command that leaves the previous record
                P.w.back() = w;
                push(v, w);
                moved = true;
            }
            // the reference appends the final state once more
            P.v.back() = moved ? v : 0.f;
            P.w.back() = moved ? w : 0.f;
            push(moved ? v : 0.f, moved ? w : 0.f);
        }
    };
    unsigned nth = std::max(1u, std::min(std::thread::hardware_concurrency(), 16u));
    nth          = std::min(nth, numPaths_);
    std::vector<std::thread> th;
    for (unsigned i = 1; i < nth; i++) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();

    pathBegin_.assign(1, 0);
    size_t total = 0;
    for (const auto& P : paths) total += P.x.size();
    for (auto* v : {&tx_, &ty_, &tphi_, &tdist_, &tv_, &tw_}) v->reserve(total);
    for (const auto& P : paths)
    {
        tx_.insert(tx_.end(), P.x.begin(), P.x.end());
        ty_.insert(ty_.end(), P.y.begin(), P.y.end());
        tphi_.insert(tphi_.end(), P.phi.begin(), P.phi.end());
        tdist_.insert(tdist_.end(), P.dist.begin(), P.dist.end());
        tv_.insert(tv_.end(), P.v.begin(), P.v.end());
        tw_.insert(tw_.end(), P.w.begin(), P.w.end());
        pathBegin_.push_back(tx_.size());
    }
    pathBegin32_.assign(pathBegin_.begin(), pathBegin_.end());
    // heading cos/sin per step (TTwist2D::rotate(phi) of getPathTwist takes them of the float heading widened to double)
    tcos_.resize(total);
    tsin_.resize(total);
    std::atomic<size_t> nextBlock{0};
    auto                trig = [&]() {
        for (;;)
        {
            const size_t b = nextBlock.fetch_add(1) * 8192;
            if (b >= total) break;
            for (size_t i = b; i < std::min(total, b + 8192); i++)
            {
                tcos_[i] = std::cos(static_cast<double>(tphi_[i]));
                tsin_[i] = std::sin(static_cast<double>(tphi_[i]));
            }
        }
    };
    th.clear();
    for (unsigned i = 1; i < nth; i++) th.emplace_back(trig);
    trig();
    for (auto& t : th) t.join();
}

size_t DiffDrive_C::at(uint16_t k, uint32_t step) const
{
    MPP_ASSERT(k < numPaths_);
    const size_t i = pathBegin_[k] + step;
    MPP_ASSERT(i < pathBegin_[k + 1]);
    return i;
}
TPose2D DiffDrive_C::getPathPose(uint16_t k, uint32_t step) const
{
    const size_t i = at(k, step);
    return {tx_[i], ty_[i], tphi_[i]};
}
double DiffDrive_C::getPathDist(uint16_t k, uint32_t step) const { return tdist_[at(k, step)]; }
bool   DiffDrive_C::getPathStepForDist(uint16_t k, double dist, uint32_t& step) const
{
    MPP_ASSERT(k < numPaths_);
    const size_t b = pathBegin_[k], n = pathBegin_[k + 1] - b;
    // first i with dist(i+1) >= dist (DiffDriveCollisionGridBased.cpp:800-807): the accumulated TP-distance never
    // decreases along a path, so the reference's linear scan is a lower bound search
    const float* first = tdist_.data() + b + 1;
    const float* last  = tdist_.data() + b + n;
    const float* it    = std::lower_bound(first, last, dist, [](float v, double d) { return !(static_cast<double>(v) >= d); });
    if (it != last)
    {
        step = static_cast<uint32_t>(it - first);
        return true;
    }
    step = static_cast<uint32_t>(n - 1);
    return false;
}
TTwist2D DiffDrive_C::getPathTwist(uint16_t k, uint32_t step) const
{
    const size_t i = at(k, step);
    TTwist2D     tw(tv_[i], 0, tw_[i]);
    tw.rotateCS(tcos_[i], tsin_[i]);  // rotate(phi) with the tabulated cos/sin of the same argument
    return tw;
}
bool DiffDrive_C::isPointInsideRobotShape(double x, double y) const
{
    return pnpoly(p_.shape_x.data(), p_.shape_y.data(), p_.shape_x.size(), x, y);
}

// Closed-form inverse map of the C PTG (arc through the origin and (x,y)).
bool DiffDrive_C::inverseMap_WS2TP(double x, double y, int& k_out, double& d_out, double) const
{
    bool exact = true;
    if (y != 0)
    {
        double       R    = (x * x + y * y) / (2 * y);
        const double Rmin = std::abs(p_.v_max_mps / p_.w_max_rps);
        const double sx   = p_.K > 0 ? x : -x;
        double       th   = y > 0 ? std::atan2(sx, std::fabs(R) - y) : std::atan2(sx, y + std::fabs(R));
        th                = wrap_to_2pi(th);
        d_out             = static_cast<float>(th * (std::fabs(R) + p_.turningRadiusReference));
        if (std::abs(R) < Rmin)
        {
            exact = false;
            R     = Rmin * sign_of(R);
        }
        const double a = M_PI * p_.v_max_mps / (p_.w_max_rps * R);
        k_out          = alpha2index(static_cast<float>(a));
    }
    else if (sign_of(x) == sign_of(p_.K))
    {
        k_out = alpha2index(0);
        d_out = x;
    }
    else
    {
        k_out = static_cast<int>(numPaths_) - 1;
        d_out = 1e+3;
        exact = false;
    }
    d_out = d_out / refDistance_;
    MPP_ASSERT(k_out >= 0 && k_out < static_cast<int>(numPaths_));
    return exact;
}

// Collision grid. For each path, the robot polygon is swept over the steps; a grid-cell
// corner inside the polygon marks that cell and its three lower-left neighbours with the
// step's TP-distance, keeping the minimum per (cell, path). Built path-parallel: each
// worker rasterises whole paths into a private per-cell min array and emits that path's
// (cell, d) list; a counting sort by cell then yields the CSR with k ascending in a cell.
void DiffDrive_C::buildCollisionGrid()
{
    const GridGeom G(-refDistance_, refDistance_, -refDistance_, refDistance_, p_.resolution);
    gxmin_ = G.xmin;
    gymin_ = G.ymin;
    gres_  = G.res;
    gnx_   = G.nx;
    gny_   = G.ny;
    MPP_ASSERT(gnx_ > 0 && gny_ > 0);
    const size_t ncells = static_cast<size_t>(gnx_) * gny_;
    const size_t nv     = p_.shape_x.size();
    const double half   = G.res * 0.5;
    const float  INF    = std::numeric_limits<float>::infinity();

    struct PathHits
    {
        std::vector<uint32_t> cell;
        std::vector<float>    d;
    };
    std::vector<PathHits> hits(numPaths_);

    std::atomic<unsigned> next{0};
    auto                  worker = [&]() {
        std::vector<float>    cellMin(ncells, INF);
        std::vector<uint32_t> touched;
        std::vector<double>   vx(nv), vy(nv);
        for (;;)
        {
            const unsigned k = next.fetch_add(1);
            if (k >= numPaths_) break;
            touched.clear();
            const size_t b = pathBegin_[k], n = pathBegin_[k + 1] - b;
            for (size_t s = 0; s + 1 < n; s++)
            {
                const double px = tx_[b + s], py = ty_[b + s], ph = tphi_[b + s];
                const double c = std::cos(ph), sn = std::sin(ph);
                double       x0 = std::numeric_limits<double>::max(), y0 = x0, x1 = -x0, y1 = -x0;
                for (size_t m = 0; m < nv; m++)
                {
                    vx[m] = px + c * p_.shape_x[m] - sn * p_.shape_y[m];
                    vy[m] = py + sn * p_.shape_x[m] + c * p_.shape_y[m];
                    x0    = std::min(x0, vx[m]);
                    x1    = std::max(x1, vx[m]);
                    y0    = std::min(y0, vy[m]);
                    y1    = std::max(y1, vy[m]);
                }
                const int ix0 = std::max(0, G.x2idx(x0) - 1), iy0 = std::max(0, G.y2idx(y0) - 1);
                const int ix1 = std::min(G.x2idx(x1) + 1, gnx_ - 1), iy1 = std::min(G.y2idx(y1) + 1, gny_ - 1);
                const float d = tdist_[b + s];
                for (int ix = ix0; ix < ix1; ix++)
                {
                    const double cx = G.idx2x(ix) - half;
                    for (int iy = iy0; iy < iy1; iy++)
                    {
                        const double cy = G.idx2y(iy) - half;
                        if (!pnpoly(vx.data(), vy.data(), nv, cx, cy)) continue;
                        for (int dx = -1; dx <= 0; dx++)
                            for (int dy = -1; dy <= 0; dy++)
                            {
                                const int jx = ix + dx, jy = iy + dy;
                                if (jx < 0 || jy < 0) continue;  // (upper bounds hold: ix < nx-1, iy < ny-1)
                                float& m = cellMin[jx + static_cast<size_t>(jy) * gnx_];
                                if (m == INF) touched.push_back(jx + jy * gnx_);
                                if (d < m) m = d;
                            }
                    }
                }
            }
            std::sort(touched.begin(), touched.end());
            PathHits& h = hits[k];
            h.cell      = touched;
            h.d.resize(touched.size());
            for (size_t i = 0; i < touched.size(); i++)
            {
                h.d[i]              = cellMin[touched[i]];
                cellMin[touched[i]] = INF;
            }
        }
    };
    unsigned nth = std::max(1u, std::min(std::thread::hardware_concurrency(), 16u));
    nth          = std::min(nth, numPaths_);
    std::vector<std::thread> th;
    for (unsigned i = 1; i < nth; i++) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();

    // counting sort of (cell, k, d) by cell; paths are visited in ascending k
    cellOffsets_.assign(ncells + 1, 0);
    for (const auto& h : hits)
        for (uint32_t c : h.cell) cellOffsets_[c + 1]++;
    for (size_t c = 0; c < ncells; c++) cellOffsets_[c + 1] += cellOffsets_[c];
    entryK_.resize(cellOffsets_[ncells]);
    entryD_.resize(cellOffsets_[ncells]);
    std::vector<uint32_t> cur(cellOffsets_.begin(), cellOffsets_.end() - 1);
    for (unsigned k = 0; k < numPaths_; k++)
        for (size_t i = 0; i < hits[k].cell.size(); i++)
        {
            const uint32_t pos = cur[hits[k].cell[i]]++;
            entryK_[pos]       = static_cast<uint16_t>(k);
            entryD_[pos]       = hits[k].d[i];
        }
}

namespace
{
// A DiffDrive_C with its own mutable state (dynamic state, trimmed speed) over the tables of a
// shared, immutable one. None of the table look-ups depends on that state.
class DiffDriveView final : public ptg_t
{
   public:
    explicit DiffDriveView(const DiffDrive_C* base) : base_(base)
    {
        numPaths_    = base->getPathCount();
        refDistance_ = base->getRefDistance();
        initialized_ = true;
    }
    PtgKind  kind() const override { return PtgKind::CollisionGrid; }
    size_t   getPathStepCount(uint16_t k) const override { return base_->getPathStepCount(k); }
    TPose2D  getPathPose(uint16_t k, uint32_t step) const override { return base_->getPathPose(k, step); }
    double   getPathDist(uint16_t k, uint32_t step) const override { return base_->getPathDist(k, step); }
    bool     getPathStepForDist(uint16_t k, double d, uint32_t& s) const override
    {
        return base_->getPathStepForDist(k, d, s);
    }
    double   getPathStepDuration() const override { return base_->getPathStepDuration(); }
    TTwist2D getPathTwist(uint16_t k, uint32_t step) const override { return base_->getPathTwist(k, step); }
    bool     inverseMap_WS2TP(double x, double y, int& k, double& d, double tol) const override
    {
        return base_->inverseMap_WS2TP(x, y, k, d, tol);
    }
    bool   isSpeedTrimmable() const override { return true; }
    bool   pathsDependOnTrimmableSpeed() const override { return false; }
    bool   isPointInsideRobotShape(double x, double y) const override { return base_->isPointInsideRobotShape(x, y); }
    double getMaxRobotRadius() const override { return base_->getMaxRobotRadius(); }
    std::unique_ptr<ptg_t> cloneForQuery() const override { return std::make_unique<DiffDriveView>(base_); }

   private:
    const DiffDrive_C* base_;
};
}  // namespace

std::unique_ptr<ptg_t> DiffDrive_C::cloneForQuery() const { return std::make_unique<DiffDriveView>(this); }

mppb_ptg_paths_desc DiffDrive_C::pathsDesc() const
{
    mppb_ptg_paths_desc d;
    std::memset(&d, 0, sizeof(d));
    d.num_paths     = static_cast<int32_t>(numPaths_);
    d.path_begin    = pathBegin32_.data();
    d.x             = tx_.data();
    d.y             = ty_.data();
    d.phi           = tphi_.data();
    d.dist          = tdist_.data();
    d.v             = tv_.data();
    d.w             = tw_.data();
    d.cos_phi       = tcos_.data();
    d.sin_phi       = tsin_.data();
    d.step_duration = stepDuration_;
    d.ref_distance  = refDistance_;
    return d;
}

mppb_ptg_grid_desc DiffDrive_C::gridDesc() const
{
    mppb_ptg_grid_desc d;
    std::memset(&d, 0, sizeof(d));
    d.num_paths    = static_cast<int32_t>(numPaths_);
    d.ref_distance = refDistance_;
    d.grid_x_min   = gxmin_;
    d.grid_y_min   = gymin_;
    d.grid_res     = gres_;
    d.grid_nx      = gnx_;
    d.grid_ny      = gny_;
    d.cell_offsets = cellOffsets_.data();
    d.entry_k      = entryK_.data();
    d.entry_d      = entryD_.data();
    d.shape_nverts = static_cast<int32_t>(p_.shape_x.size());
    d.shape_x      = p_.shape_x.data();
    d.shape_y      = p_.shape_y.data();
    d.max_radius   = maxRadius_;
    return d;
}

}  // namespace mpp
