// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product.
//
// Restatement of mrpt::math::solve_poly2/3/4 (module mrpt-math, file poly_roots.cpp,
// MRPT >= 2.12, version unpinned by the reference). MRPT is NOT vendored under
// /root/reference, so this is the *published* algorithm it wraps: S. Khashin's "poly34"
// (Descartes-Euler quartic through the resolvent cubic, trigonometric / Cardano cubic,
// one Newton polish step per real quartic root). Reference call sites that rely on it:
//   HolonomicBlend.cpp:565 (solve_poly2), :757-758 (solve_poly4), :765 (solve_poly3).
// Parity UNPINNED against MRPT itself; pinned here against numpy.roots
// (tests/test_oracle_known_answers.py).
#pragma once
#include <cmath>
#include <utility>

namespace orc
{
namespace poly
{
constexpr double kEps   = 1e-14;
constexpr double kTwoPi = 6.28318530717958648;

// cube root by range reduction to [1,8] and six Newton steps (poly34 "_root3")
inline double root3_pos(double x)
{
    double s = 1.;
    while (x < 1.)
    {
        x *= 8.;
        s *= 0.5;
    }
    while (x > 8.)
    {
        x *= 0.125;
        s *= 2.;
    }
    double r = 1.5;
    for (int i = 0; i < 6; i++) r -= 1. / 3. * (r - x / (r * r));
    return r * s;
}
inline double root3(double x)
{
    if (x > 0) return root3_pos(x);
    if (x < 0) return -root3_pos(-x);
    return 0.;
}

// x^3 + a x^2 + b x + c = 0.
// returns 3: x[0..2] real; 2: x[0], x[1] real (double root); 1: x[0] real, x[1] +- i x[2]
inline int solve_poly3(double* x, double a, double b, double c)
{
    const double a2 = a * a;
    double       q  = (a2 - 3 * b) / 9;
    const double r  = (a * (2 * a2 - 9 * b) + 27 * c) / 54;
    const double r2 = r * r;
    const double q3 = q * q * q;
    if (r2 <= (q3 + kEps))
    {
        double t = r / std::sqrt(q3);
        if (t < -1) t = -1;
        if (t > 1) t = 1;
        t = std::acos(t);
        a /= 3;
        q    = -2 * std::sqrt(q);
        x[0] = q * std::cos(t / 3) - a;
        x[1] = q * std::cos((t + kTwoPi) / 3) - a;
        x[2] = q * std::cos((t - kTwoPi) / 3) - a;
        return 3;
    }
    double A = -root3(std::fabs(r) + std::sqrt(r2 - q3));
    if (r < 0) A = -A;
    const double B = (A == 0 ? 0 : q / A);
    a /= 3;
    x[0] = (A + B) - a;
    x[1] = -0.5 * (A + B) - a;
    x[2] = 0.5 * std::sqrt(3.) * (A - B);
    if (std::fabs(x[2]) < kEps)
    {
        x[2] = x[1];
        return 2;
    }
    return 1;
}

// a + i b = sqrt(x + i y), a >= 0
inline void csqrt(double x, double y, double& a, double& b)
{
    double r = std::sqrt(x * x + y * y);
    if (y == 0)
    {
        r = std::sqrt(r);
        if (x >= 0)
        {
            a = r;
            b = 0;
        }
        else
        {
            a = 0;
            b = r;
        }
    }
    else
    {
        a = std::sqrt(0.5 * (x + r));
        b = 0.5 * y / a;
    }
}

// x^4 + b x^2 + d = 0
inline int solve_p4_biquadratic(double* x, double b, double d)
{
    const double D = b * b - 4 * d;
    if (D >= 0)
    {
        const double sD = std::sqrt(D);
        const double x1 = (-b + sD) / 2;
        const double x2 = (-b - sD) / 2;
        if (x2 >= 0)
        {
            const double sx1 = std::sqrt(x1), sx2 = std::sqrt(x2);
            x[0] = -sx1;
            x[1] = sx1;
            x[2] = -sx2;
            x[3] = sx2;
            return 4;
        }
        if (x1 < 0)
        {
            const double sx1 = std::sqrt(-x1), sx2 = std::sqrt(-x2);
            x[0] = 0;
            x[1] = sx1;
            x[2] = 0;
            x[3] = sx2;
            return 0;
        }
        const double sx1 = std::sqrt(x1), sx2 = std::sqrt(-x2);
        x[0] = -sx1;
        x[1] = sx1;
        x[2] = 0;
        x[3] = sx2;
        return 2;
    }
    const double sD2 = 0.5 * std::sqrt(-D);
    csqrt(-0.5 * b, sD2, x[0], x[1]);
    csqrt(-0.5 * b, -sD2, x[2], x[3]);
    return 0;
}

inline void sort3(double& a, double& b, double& c)
{
    if (a > b) std::swap(a, b);
    if (c < b)
    {
        std::swap(b, c);
        if (a > b) std::swap(a, b);
    }
}

// depressed quartic x^4 + b x^2 + c x + d = 0
inline int solve_p4_depressed(double* x, double b, double c, double d)
{
    if (std::fabs(c) < 1e-14 * (std::fabs(b) + std::fabs(d))) return solve_p4_biquadratic(x, b, d);
    const int res3 = solve_poly3(x, 2 * b, b * b - 4 * d, -c * c);
    if (res3 > 1)
    {
        sort3(x[0], x[1], x[2]);
        if (x[0] > 0)
        {
            const double sz1 = std::sqrt(x[0]), sz2 = std::sqrt(x[1]), sz3 = std::sqrt(x[2]);
            if (c > 0)
            {
                x[0] = (-sz1 - sz2 - sz3) / 2;
                x[1] = (-sz1 + sz2 + sz3) / 2;
                x[2] = (+sz1 - sz2 + sz3) / 2;
                x[3] = (+sz1 + sz2 - sz3) / 2;
                return 4;
            }
            x[0] = (-sz1 - sz2 + sz3) / 2;
            x[1] = (-sz1 + sz2 - sz3) / 2;
            x[2] = (+sz1 - sz2 - sz3) / 2;
            x[3] = (+sz1 + sz2 + sz3) / 2;
            return 4;
        }
        const double sz1 = std::sqrt(-x[0]), sz2 = std::sqrt(-x[1]), sz3 = std::sqrt(x[2]);
        if (c > 0)
        {
            x[0] = -sz3 / 2;
            x[1] = (sz1 - sz2) / 2;
            x[2] = sz3 / 2;
            x[3] = (-sz1 - sz2) / 2;
            return 0;
        }
        x[0] = sz3 / 2;
        x[1] = (-sz1 + sz2) / 2;
        x[2] = -sz3 / 2;
        x[3] = (sz1 + sz2) / 2;
        return 0;
    }
    if (x[0] < 0) x[0] = 0;
    const double sz1 = std::sqrt(x[0]);
    double       szr, szi;
    csqrt(x[1], x[2], szr, szi);
    if (c > 0)
    {
        x[0] = -sz1 / 2 - szr;
        x[1] = -sz1 / 2 + szr;
        x[2] = sz1 / 2;
        x[3] = szi;
        return 2;
    }
    x[0] = sz1 / 2 - szr;
    x[1] = sz1 / 2 + szr;
    x[2] = -sz1 / 2;
    x[3] = szi;
    return 2;
}

inline double newton4(double x, double a, double b, double c, double d)
{
    const double fxs = ((4 * x + 3 * a) * x + 2 * b) * x + c;
    if (fxs == 0) return x;
    const double fx = (((x + a) * x + b) * x + c) * x + d;
    return x - fx / fxs;
}

// x^4 + a x^3 + b x^2 + c x + d = 0. returns 4 / 2 / 0 = number of real roots (reals first)
inline int solve_poly4(double* x, double a, double b, double c, double d)
{
    const double d1  = d + 0.25 * a * (0.25 * b * a - 3. / 64 * a * a * a - c);
    const double c1  = c + 0.5 * a * (0.25 * a * a - b);
    const double b1  = b - 0.375 * a * a;
    const int    res = solve_p4_depressed(x, b1, c1, d1);
    if (res == 4)
    {
        x[0] -= a / 4;
        x[1] -= a / 4;
        x[2] -= a / 4;
        x[3] -= a / 4;
    }
    else if (res == 2)
    {
        x[0] -= a / 4;
        x[1] -= a / 4;
        x[2] -= a / 4;
    }
    else
    {
        x[0] -= a / 4;
        x[2] -= a / 4;
    }
    if (res > 0)
    {
        x[0] = newton4(x[0], a, b, c, d);
        x[1] = newton4(x[1], a, b, c, d);
    }
    if (res > 2)
    {
        x[2] = newton4(x[2], a, b, c, d);
        x[3] = newton4(x[3], a, b, c, d);
    }
    return res;
}

// a x^2 + b x + c = 0 ; r1 <= r2. returns 0/1/2
inline int solve_poly2(double a, double b, double c, double& r1, double& r2)
{
    if (std::fabs(a) < kEps)
    {
        if (std::fabs(b) < kEps) return 0;
        r1 = -c / b;
        r2 = 1e99;
        return 1;
    }
    double Di = b * b - 4 * a * c;
    if (Di < 0) return 0;
    Di = std::sqrt(Di);
    r1 = (-b + Di) / (2 * a);
    r2 = (-b - Di) / (2 * a);
    if (r2 < r1) std::swap(r1, r2);
    return 2;
}

}  // namespace poly
}  // namespace orc
