// Real-root solvers for monic cubics/quartics on the device (FP64).
//
// The reference calls mrpt::math::solve_poly3 / solve_poly4 (HolonomicBlend.cpp:757-765). MRPT
// is not vendored in the reference tree; those functions wrap S. Khashin's published "poly34"
// method (Descartes-Euler reduction of the quartic through its resolvent cubic, trigonometric
// or Cardano cubic, one Newton polish step per real quartic root). This header restates that
// method for sm_100a. Compiled with -fmad=false so that every +,-,*,/ and sqrt rounds exactly as
// on the reference's x86-64 build; only acos()/cos() (trigonometric branch of the cubic) come
// from the CUDA math library and may differ from glibc in the last bits.
#pragma once
#include <cuda_runtime.h>

namespace mppb
{
namespace p34
{
constexpr double kTiny  = 1e-14;
constexpr double kTwoPi = 6.28318530717958648;

__device__ __forceinline__ void swap_d(double& a, double& b)
{
    const double t = a;
    a              = b;
    b              = t;
}

// cube root through range reduction to [1,8] and six Newton iterations (as poly34 does;
// NOT cbrt(), whose rounding would differ from the reference)
__device__ __forceinline__ double cube_root(double v)
{
    if (v == 0) return 0.;
    const bool neg = v < 0;
    double     x   = neg ? -v : v;
    double     s   = 1.;
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
#pragma unroll
    for (int i = 0; i < 6; i++) r -= 1. / 3. * (r - x / (r * r));
    r *= s;
    return neg ? -r : r;
}

// x^3 + a x^2 + b x + c: returns 3 (x0,x1,x2 real), 2 (x0, x1 real), 1 (x0 real, x1 +- i x2)
__device__ __forceinline__ int cubic(double* x, double a, double b, double c)
{
    const double a2 = a * a;
    double       q  = (a2 - 3 * b) / 9;
    const double r  = (a * (2 * a2 - 9 * b) + 27 * c) / 54;
    const double r2 = r * r;
    const double q3 = q * q * q;
    if (r2 <= (q3 + kTiny))
    {
        double t = r / sqrt(q3);
        if (t < -1) t = -1;
        if (t > 1) t = 1;
        t = acos(t);
        a /= 3;
        q    = -2 * sqrt(q);
        x[0] = q * cos(t / 3) - a;
        x[1] = q * cos((t + kTwoPi) / 3) - a;
        x[2] = q * cos((t - kTwoPi) / 3) - a;
        return 3;
    }
    double A = -cube_root(fabs(r) + sqrt(r2 - q3));
    if (r < 0) A = -A;
    const double B = (A == 0 ? 0 : q / A);
    a /= 3;
    x[0] = (A + B) - a;
    x[1] = -0.5 * (A + B) - a;
    x[2] = 0.5 * sqrt(3.) * (A - B);
    if (fabs(x[2]) < kTiny)
    {
        x[2] = x[1];
        return 2;
    }
    return 1;
}

// principal square root of the complex number (x + i y): re >= 0
__device__ __forceinline__ void complex_sqrt(double x, double y, double& re, double& im)
{
    double r = sqrt(x * x + y * y);
    if (y == 0)
    {
        r = sqrt(r);
        if (x >= 0)
        {
            re = r;
            im = 0;
        }
        else
        {
            re = 0;
            im = r;
        }
    }
    else
    {
        re = sqrt(0.5 * (x + r));
        im = 0.5 * y / re;
    }
}

// x^4 + b x^2 + d
__device__ __forceinline__ int biquadratic(double* x, double b, double d)
{
    const double D = b * b - 4 * d;
    if (D >= 0)
    {
        const double sD = sqrt(D);
        const double x1 = (-b + sD) / 2, x2 = (-b - sD) / 2;
        if (x2 >= 0)
        {
            const double s1 = sqrt(x1), s2 = sqrt(x2);
            x[0] = -s1;
            x[1] = s1;
            x[2] = -s2;
            x[3] = s2;
            return 4;
        }
        if (x1 < 0)
        {
            x[0] = 0;
            x[1] = sqrt(-x1);
            x[2] = 0;
            x[3] = sqrt(-x2);
            return 0;
        }
        const double s1 = sqrt(x1);
        x[0]            = -s1;
        x[1]            = s1;
        x[2]            = 0;
        x[3]            = sqrt(-x2);
        return 2;
    }
    const double h = 0.5 * sqrt(-D);
    complex_sqrt(-0.5 * b, h, x[0], x[1]);
    complex_sqrt(-0.5 * b, -h, x[2], x[3]);
    return 0;
}

// depressed quartic x^4 + b x^2 + c x + d
__device__ __forceinline__ int depressed_quartic(double* x, double b, double c, double d)
{
    if (fabs(c) < 1e-14 * (fabs(b) + fabs(d))) return biquadratic(x, b, d);
    const int n3 = cubic(x, 2 * b, b * b - 4 * d, -c * c);
    if (n3 > 1)
    {
        // ascending order of the three real resolvent roots
        if (x[0] > x[1]) swap_d(x[0], x[1]);
        if (x[2] < x[1])
        {
            swap_d(x[1], x[2]);
            if (x[0] > x[1]) swap_d(x[0], x[1]);
        }
        if (x[0] > 0)
        {
            const double s1 = sqrt(x[0]), s2 = sqrt(x[1]), s3 = sqrt(x[2]);
            if (c > 0)
            {
                x[0] = (-s1 - s2 - s3) / 2;
                x[1] = (-s1 + s2 + s3) / 2;
                x[2] = (+s1 - s2 + s3) / 2;
                x[3] = (+s1 + s2 - s3) / 2;
            }
            else
            {
                x[0] = (-s1 - s2 + s3) / 2;
                x[1] = (-s1 + s2 - s3) / 2;
                x[2] = (+s1 - s2 - s3) / 2;
                x[3] = (+s1 + s2 + s3) / 2;
            }
            return 4;
        }
        const double s1 = sqrt(-x[0]), s2 = sqrt(-x[1]), s3 = sqrt(x[2]);
        if (c > 0)
        {
            x[0] = -s3 / 2;
            x[1] = (s1 - s2) / 2;
            x[2] = s3 / 2;
            x[3] = (-s1 - s2) / 2;
        }
        else
        {
            x[0] = s3 / 2;
            x[1] = (-s1 + s2) / 2;
            x[2] = -s3 / 2;
            x[3] = (s1 + s2) / 2;
        }
        return 0;
    }
    if (x[0] < 0) x[0] = 0;
    const double s1 = sqrt(x[0]);
    double       re, im;
    complex_sqrt(x[1], x[2], re, im);
    if (c > 0)
    {
        x[0] = -s1 / 2 - re;
        x[1] = -s1 / 2 + re;
        x[2] = s1 / 2;
    }
    else
    {
        x[0] = s1 / 2 - re;
        x[1] = s1 / 2 + re;
        x[2] = -s1 / 2;
    }
    x[3] = im;
    return 2;
}

__device__ __forceinline__ double polish(double x, double a, double b, double c, double d)
{
    const double df = ((4 * x + 3 * a) * x + 2 * b) * x + c;
    if (df == 0) return x;
    const double f = (((x + a) * x + b) * x + c) * x + d;
    return x - f / df;
}

// x^4 + a x^3 + b x^2 + c x + d: number of real roots (4, 2 or 0), stored first in x[]
__device__ __forceinline__ int quartic(double* x, double a, double b, double c, double d)
{
    const double d1 = d + 0.25 * a * (0.25 * b * a - 3. / 64 * a * a * a - c);
    const double c1 = c + 0.5 * a * (0.25 * a * a - b);
    const double b1 = b - 0.375 * a * a;
    const int    n  = depressed_quartic(x, b1, c1, d1);
    const double sh = a / 4;
    if (n == 4)
    {
        x[0] -= sh;
        x[1] -= sh;
        x[2] -= sh;
        x[3] -= sh;
    }
    else if (n == 2)
    {
        x[0] -= sh;
        x[1] -= sh;
        x[2] -= sh;
    }
    else
    {
        x[0] -= sh;
        x[2] -= sh;
    }
    if (n > 0)
    {
        x[0] = polish(x[0], a, b, c, d);
        x[1] = polish(x[1], a, b, c, d);
    }
    if (n > 2)
    {
        x[2] = polish(x[2], a, b, c, d);
        x[3] = polish(x[3], a, b, c, d);
    }
    return n;
}

}  // namespace p34
}  // namespace mppb
