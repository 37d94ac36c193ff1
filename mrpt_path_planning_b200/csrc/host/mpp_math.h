// Host-side value types of the mpp:: mirror (product code). Minimal stand-ins for the
// MRPT types that cross the planner API: SE(2) poses, twists, angle
// wrapping, integer rounding. Semantics follow MRPT so that host decisions (lattice cell,
// candidate step, heuristic) are reproducible.
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>

namespace mpp
{
#define MPP_ASSERT(cond)                                                                        \
    do {                                                                                        \
        if (!(cond)) throw std::runtime_error(std::string("Assert failed: " #cond " (") + __FILE__ + ":" + \
                                              std::to_string(__LINE__) + ")");                  \
    } while (0)

/** mrpt::round: nearest integer, ties to even (lrint in the default rounding mode). */
inline int round_i(double v) { return static_cast<int>(std::lrint(v)); }
inline int sign_of(double v) { return v < 0 ? -1 : 1; }
inline int sign_or_zero(double v) { return v == 0 ? 0 : (v < 0 ? -1 : 1); }

inline double wrap_to_2pi(double a)
{
    const bool neg = a < 0;
    // fmod is exact and returns its first argument unchanged when that is smaller in magnitude than the second: for
    // |a| < 2 pi — every angle the planner meets — the library call (~50-100 cycles, three per heuristic evaluation)
    // can be skipped without changing a bit of the result
    if (!(std::fabs(a) < 2.0 * M_PI)) a = std::fmod(a, 2.0 * M_PI);
    return neg ? a + 2.0 * M_PI : a;
}
/** result in [-pi, pi) */
inline double wrap_to_pi(double a) { return wrap_to_2pi(a + M_PI) - M_PI; }
inline double ang_distance(double from, double to)
{
    double d = wrap_to_pi(to) - wrap_to_pi(from);
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

struct TTwist2D
{
    double vx = 0, vy = 0, omega = 0;
    TTwist2D() = default;
    TTwist2D(double vx_, double vy_, double w_) : vx(vx_), vy(vy_), omega(w_) {}
    /** rotate the linear part by `ang` radians */
    void rotate(double ang) { rotateCS(std::cos(ang), std::sin(ang)); }
    /** rotate() with the angle's cosine and sine taken by the caller (the same angle for many twists) */
    void rotateCS(double c, double s)
    {
        const double nx = vx * c - vy * s, ny = vx * s + vy * c;
        vx = nx;
        vy = ny;
    }
    bool operator==(const TTwist2D& o) const { return vx == o.vx && vy == o.vy && omega == o.omega; }
};

struct TPose2D
{
    double x = 0, y = 0, phi = 0;
    TPose2D() = default;
    TPose2D(double x_, double y_, double phi_) : x(x_), y(y_), phi(phi_) {}
    /** pose composition this (+) b */
    TPose2D operator+(const TPose2D& b) const { return composeCS(std::cos(phi), std::sin(phi), b); }
    /** operator+ with cos(phi), sin(phi) of this pose taken by the caller (one pose composed with many) */
    TPose2D composeCS(double c, double s, const TPose2D& b) const
    {
        return {x + b.x * c - b.y * s, y + b.x * s + b.y * c, wrap_to_pi(phi + b.phi)};
    }
    /** inverse composition this (-) b: this pose as seen from b */
    TPose2D operator-(const TPose2D& b) const
    {
        const double c = std::cos(b.phi), s = std::sin(b.phi);
        const double dx = x - b.x, dy = y - b.y;
        return {dx * c + dy * s, -dx * s + dy * c, wrap_to_pi(phi - b.phi)};
    }
    /** operator-(b) with cos(b.phi), sin(b.phi) taken by the caller */
    TPose2D minusCS(const TPose2D& b, double c, double s) const
    {
        const double dx = x - b.x, dy = y - b.y;
        return {dx * c + dy * s, -dx * s + dy * c, wrap_to_pi(phi - b.phi)};
    }
    double   norm() const { return std::sqrt(x * x + y * y); }
    TPoint2D translation() const { return {x, y}; }
    bool     operator==(const TPose2D& o) const { return x == o.x && y == o.y && phi == o.phi; }
    bool     operator!=(const TPose2D& o) const { return !(*this == o); }
};

}  // namespace mpp
