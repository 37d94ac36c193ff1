// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product.
//
// CPU restatement of the MRPT math/container semantics that the hot path of
// mrpt_path_planning touches. MRPT itself is NOT vendored in
// /root/reference and is not installed here, so everything in this file is the
// *assumed* third-party behaviour ("parity unpinned": no golden vectors exist for it,
// see DESIGN.md §Oracle). Each item names the reference call site that relies on it.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may use anything under oracle/.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc
{
#define ORC_ASSERT(cond)                                                                   \
    do {                                                                                   \
        if (!(cond))                                                                       \
            throw std::runtime_error(                                                      \
                std::string("oracle assert failed: " #cond " at ") + __FILE__ + ":" +      \
                std::to_string(__LINE__));                                                 \
    } while (0)

// mrpt::round(): lrint-based (round-half-even in the default FP mode).
// Call sites: TPS_Astar.cpp:621,685; HolonomicBlend.cpp:712; CDynamicGrid::setSize.
inline int mrpt_round(double v) { return static_cast<int>(std::lrint(v)); }

inline double square(double x) { return x * x; }

inline int sign_nz(double x) { return x < 0 ? -1 : 1; }  // mrpt::sign
inline int sign_with_zero(double x) { return x == 0 ? 0 : (x < 0 ? -1 : 1); }  // mrpt::signWithZero

// mrpt::math::wrapTo2Pi / wrapToPi. wrapToPi result lies in [-pi, pi).
inline double wrapTo2Pi(double a)
{
    const bool was_neg = a < 0;
    a                  = std::fmod(a, 2.0 * M_PI);
    if (was_neg) a += 2.0 * M_PI;
    return a;
}
inline double wrapToPi(double a) { return wrapTo2Pi(a + M_PI) - M_PI; }

// mrpt::math::angDistance (TPS_Astar.cpp:491,508)
inline double angDistance(double from, double to)
{
    from     = wrapToPi(from);
    to       = wrapToPi(to);
    double d = to - from;
    if (d > M_PI)
        d -= 2.0 * M_PI;
    else if (d < -M_PI)
        d += 2.0 * M_PI;
    return d;
}

struct TPoint2D
{
    double x = 0, y = 0;
    double norm() const { return std::sqrt(x * x + y * y); }
};

// mrpt::math::TTwist2D (vx, vy, omega) with rotate(): TPS_Astar.cpp:297,589
struct TTwist2D
{
    double vx = 0, vy = 0, omega = 0;
    void   rotate(double ang)
    {
        const double nvx = vx * std::cos(ang) - vy * std::sin(ang);
        const double nvy = vx * std::sin(ang) + vy * std::cos(ang);
        vx               = nvx;
        vy               = nvy;
    }
    bool operator==(const TTwist2D& o) const { return vx == o.vx && vy == o.vy && omega == o.omega; }
};

// mrpt::math::TPose2D with SE(2) composition "+" and inverse composition "-".
// Call sites: TPS_Astar.cpp:290,329,556,727; CostEvaluatorCostMap.cpp:150.
struct TPose2D
{
    double x = 0, y = 0, phi = 0;
    TPose2D() = default;
    TPose2D(double x_, double y_, double phi_) : x(x_), y(y_), phi(phi_) {}

    TPose2D operator+(const TPose2D& b) const
    {
        const double c = std::cos(phi), s = std::sin(phi);
        return TPose2D(x + b.x * c - b.y * s, y + b.x * s + b.y * c, wrapToPi(phi + b.phi));
    }
    TPose2D operator-(const TPose2D& b) const
    {
        const double c = std::cos(b.phi), s = std::sin(b.phi);
        return TPose2D(
            (x - b.x) * c + (y - b.y) * s, -(x - b.x) * s + (y - b.y) * c, wrapToPi(phi - b.phi));
    }
    double norm() const { return std::sqrt(x * x + y * y); }
    bool   operator==(const TPose2D& o) const { return x == o.x && y == o.y && phi == o.phi; }
    bool   operator!=(const TPose2D& o) const { return !(*this == o); }
};

// mrpt::poses::CPose2D pieces used by transform_pc_square_clipping.cpp:22,39:
// unary minus (inverse) and composePoint, with cached cos/sin of the heading.
struct CPose2D
{
    double x = 0, y = 0, phi = 0, c = 1, s = 0;
    CPose2D() = default;
    explicit CPose2D(const TPose2D& p) : x(p.x), y(p.y), phi(p.phi), c(std::cos(p.phi)), s(std::sin(p.phi)) {}
    CPose2D inverse() const
    {
        CPose2D r;
        r.x   = -x * c - y * s;
        r.y   = x * s - y * c;
        r.phi = -phi;
        r.c   = std::cos(r.phi);
        r.s   = std::sin(r.phi);
        return r;
    }
    void composePoint(double lx, double ly, double& gx, double& gy) const
    {
        gx = x + lx * c - ly * s;
        gy = y + lx * s + ly * c;
    }
};

// mrpt::containers::CDynamicGrid<T> (subset). Call sites:
// DiffDriveCollisionGridBased.cpp:182-198,225,237,657-736; CostEvaluatorCostMap.cpp:84-106,160.
template <class T>
struct DynGrid
{
    double         x_min = 0, x_max = 0, y_min = 0, y_max = 0, res = 1;
    int            size_x = 0, size_y = 0;
    std::vector<T> cells;

    void setSize(double xmin, double xmax, double ymin, double ymax, double resolution, const T& fill = T())
    {
        x_min  = resolution * mrpt_round(xmin / resolution);
        y_min  = resolution * mrpt_round(ymin / resolution);
        x_max  = resolution * mrpt_round(xmax / resolution);
        y_max  = resolution * mrpt_round(ymax / resolution);
        res    = resolution;
        size_x = mrpt_round((x_max - x_min) / res);
        size_y = mrpt_round((y_max - y_min) / res);
        cells.assign(static_cast<size_t>(size_x) * size_y, fill);
    }
    int    x2idx(double x) const { return static_cast<int>((x - x_min) / res); }
    int    y2idx(double y) const { return static_cast<int>((y - y_min) / res); }
    double idx2x(int cx) const { return x_min + (cx + 0.5) * res; }
    double idx2y(int cy) const { return y_min + (cy + 0.5) * res; }
    T*     cellByIndex(int cx, int cy)
    {
        if (cx < 0 || cx >= size_x || cy < 0 || cy >= size_y) return nullptr;
        return &cells[cx + static_cast<size_t>(cy) * size_x];
    }
    const T* cellByIndex(int cx, int cy) const { return const_cast<DynGrid*>(this)->cellByIndex(cx, cy); }
    T*       cellByPos(double x, double y) { return cellByIndex(x2idx(x), y2idx(y)); }
    const T* cellByPos(double x, double y) const { return cellByIndex(x2idx(x), y2idx(y)); }
};

// mrpt::math::TPolygon2D::contains — crossing-number (PNPOLY) test.
// Call sites: DiffDriveCollisionGridBased.cpp:738 and, through
// CPTG_RobotShape_Polygonal::isPointInsideRobotShape, the TP-obstacle post-process.
struct Polygon2D
{
    std::vector<double> vx, vy;
    size_t              size() const { return vx.size(); }
    bool                contains(double px, double py) const
    {
        const size_t n = vx.size();
        bool         c = false;
        for (size_t i = 0, j = n - 1; i < n; j = i++)
        {
            if (((vy[i] > py) != (vy[j] > py)) &&
                (px < (vx[j] - vx[i]) * (py - vy[i]) / (vy[j] - vy[i]) + vx[i]))
                c = !c;
        }
        return c;
    }
};

// Float SoA point cloud (mrpt::maps::CSimplePointsMap subset).
struct PointCloud
{
    std::vector<float> xs, ys;
    size_t             size() const { return xs.size(); }
    void               clear()
    {
        xs.clear();
        ys.clear();
    }
    void insertPointFast(float x, float y)
    {
        xs.push_back(x);
        ys.push_back(y);
    }
    // nanoflann L2_Simple in float: sum over dims of (q - p)^2, accumulated in float.
    // CostEvaluatorCostMap.cpp:97 (closest), CostEvaluatorPreferredWaypoint.cpp:106 (radius).
    float closestPointSqrDist(float qx, float qy) const
    {
        float best = std::numeric_limits<float>::max();
        for (size_t i = 0; i < xs.size(); i++)
        {
            const float dx = qx - xs[i], dy = qy - ys[i];
            const float d  = dx * dx + dy * dy;
            if (d < best) best = d;
        }
        return best;
    }
};

}  // namespace orc
