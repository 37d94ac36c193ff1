// ORACLE / REFERENCE BUILD — TEST INFRASTRUCTURE ONLY. Not part of the shipped product.
//
// Minimal stand-in for the MRPT (>= 2.12) headers that the hot-path translation units of
// /root/reference/mrpt_path_planning include, so that THOSE SOURCE FILES — unmodified, compiled
// from where they lie — can be built into oracle/_ref/libmpp_ref.so (see oracle/Makefile, target
// `ref`). MRPT itself is not vendored under /root/reference and is not installed in this image.
//
// What this buys: every line of the reference's own code on the path (TPS_Astar.cpp,
// HolonomicBlend.cpp, DiffDriveCollisionGridBased.cpp, DiffDrive_C.cpp, the cost evaluators,
// transform_pc_square_clipping.cpp, tp_obstacles_single_path.cpp, edge_interpolated_path.cpp,
// refine_trajectory.cpp, trajectories.cpp ...) runs as written, and the line-by-line restatement in
// oracle/orc_*.hpp is checked against it (tests/test_oracle_vs_reference.py).
// What it does NOT buy: MRPT's own arithmetic (SE(2) composition, CDynamicGrid index math, PNPOLY,
// poly34 root solvers, exprtk, nanoflann distances, the PTG base class) is still the ASSUMED
// behaviour of SURVEY.md §8c, written here once more in MRPT's API shape. Every such piece is
// marked [MRPT-assumed].
//
// One header serves all <mrpt/...> include paths: the Makefile generates one-line forwarding
// headers under oracle/_ref/include/mrpt/ for each path the reference includes.
#pragma once

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <fstream>
#include <functional>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <optional>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <variant>
#include <vector>

#include "../orc_expr.hpp"  // stand-in for exprtk [MRPT-assumed]
#include "../orc_poly.hpp"  // poly34 restatement [MRPT-assumed]

// 2.14.1: the two-argument DEFINE_*_MRPT_OBJECT forms (MRPT_VERSION >= 0x020e00) without the
// TNavDynamicState::internalState member added in 2.14.2 (MoveEdgeSE2_TPS.h:54).
#define MRPT_VERSION 0x20e01

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

// ---------------------------------------------------------------------------------------------
// core: exceptions, asserts, macros
// ---------------------------------------------------------------------------------------------
namespace mrpt
{
inline std::string format(const char* fmt, ...)
{
    char    buf[4096];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return std::string(buf);
}
template <class EXC>
inline std::string exception_to_str(const EXC& e)
{
    return e.what();
}
}  // namespace mrpt

#define MRPT_SHIM_STR2(x) #x
#define MRPT_SHIM_STR(x) MRPT_SHIM_STR2(x)
#define THROW_EXCEPTION(msg) \
    throw std::runtime_error(std::string(msg) + std::string(" [" __FILE__ ":" MRPT_SHIM_STR(__LINE__) "]"))
#define THROW_EXCEPTION_FMT(fmt, ...) THROW_EXCEPTION(mrpt::format(fmt, __VA_ARGS__))
#define ASSERT_(c)                                                   \
    do {                                                             \
        if (!(c)) THROW_EXCEPTION("Assert condition failed: " #c);   \
    } while (0)
#define ASSERTMSG_(c, msg)                                           \
    do {                                                             \
        if (!(c)) THROW_EXCEPTION(std::string(msg));                 \
    } while (0)
#define ASSERT_GT_(a, b) ASSERT_((a) > (b))
#define ASSERT_GE_(a, b) ASSERT_((a) >= (b))
#define ASSERT_LT_(a, b) ASSERT_((a) < (b))
#define ASSERT_LE_(a, b) ASSERT_((a) <= (b))
#define ASSERT_EQUAL_(a, b) ASSERT_((a) == (b))
#define ASSERT_ABOVE_(a, b) ASSERT_((a) > (b))
#define ASSERT_BELOW_(a, b) ASSERT_((a) < (b))
#define MRPT_START
#define MRPT_END
#define MRPT_TODO(x)
#define MRPT_CHECK_NORMAL_NUMBER(v) ASSERT_(std::isfinite(v))
#define MRPT_THROW_UNKNOWN_SERIALIZATION_VERSION(v) \
    THROW_EXCEPTION_FMT("Cannot parse object: unknown serialization version number: '%i'", static_cast<int>(v))
#define MRPT_UNUSED_PARAM(x) (void)(x)

// logging: the planner's debug output is not part of the path; evaluate nothing.
#define MRPT_LOG_DEBUG(x) \
    do {                  \
    } while (0)
#define MRPT_LOG_DEBUG_STREAM(x) \
    do {                         \
    } while (0)
#define MRPT_LOG_DEBUG_FMT(...) \
    do {                        \
    } while (0)
#define MRPT_LOG_INFO_STREAM(x) \
    do {                        \
    } while (0)
#define MRPT_LOG_WARN_STREAM(x) \
    do {                        \
    } while (0)
#define MRPT_LOG_ERROR_STREAM(x) \
    do {                         \
    } while (0)

namespace mrpt
{
// [MRPT-assumed] bits_math.h / round.h
template <class T>
inline T square(const T x)
{
    return x * x;
}
inline double DEG2RAD(const double x) { return x * M_PI / 180.0; }
inline float  DEG2RAD(const float x) { return x * static_cast<float>(M_PI) / 180.0f; }
inline double DEG2RAD(const int x) { return x * M_PI / 180.0; }
inline double RAD2DEG(const double x) { return x * 180.0 / M_PI; }
template <typename T>
inline int sign(T x)
{
    return x < 0 ? -1 : 1;
}
template <typename T>
inline int signWithZero(T x)
{
    return (x == 0 || x == -0) ? 0 : sign(x);
}
template <typename T, typename K>
inline void keep_min(T& var, const K test_val)
{
    if (test_val < var) var = static_cast<T>(test_val);
}
template <typename T, typename K>
inline void keep_max(T& var, const K test_val)
{
    if (test_val > var) var = static_cast<T>(test_val);
}
// mrpt::round(): lrint (round-half-to-even in the default rounding mode)
template <typename T>
inline int round(const T value)
{
    return static_cast<int>(lrint(static_cast<double>(value)));
}
inline float d2f(const double d) { return static_cast<float>(d); }

namespace literals
{
constexpr inline double operator"" _deg(long double v)
{
    return static_cast<double>(v * 3.14159265358979323846264338327950288L / 180.0L);
}
}  // namespace literals
using namespace literals;

struct Clock
{
    using time_point = std::chrono::steady_clock::time_point;
    static time_point now() { return std::chrono::steady_clock::now(); }
    static double     nowDouble()
    {
        return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }
};

template <class MTX>
inline std::unique_lock<MTX> lockHelper(MTX& m)
{
    return std::unique_lock<MTX>(m);
}
template <class MTX>
inline std::unique_lock<MTX> lockHelper(const MTX& m)
{
    return std::unique_lock<MTX>(const_cast<MTX&>(m));
}

template <class T>
using optional_ref = std::optional<std::reference_wrapper<T>>;

class WorkerThreadsPool;  // only named by headers we do not compile
}  // namespace mrpt

// ---------------------------------------------------------------------------------------------
// mrpt::math — light-weight geometry [MRPT-assumed]
// ---------------------------------------------------------------------------------------------
namespace mrpt::math
{
template <class T>
inline void wrapTo2PiInPlace(T& a)
{
    const bool was_neg = a < 0;
    a                  = fmod(a, static_cast<T>(2.0 * M_PI));
    if (was_neg) a += static_cast<T>(2.0 * M_PI);
}
template <class T>
inline T wrapTo2Pi(T a)
{
    wrapTo2PiInPlace(a);
    return a;
}
template <class T>
inline T wrapToPi(T a)
{
    return wrapTo2Pi(a + static_cast<T>(M_PI)) - static_cast<T>(M_PI);
}
template <class T>
inline void wrapToPiInPlace(T& a)
{
    a = wrapToPi(a);
}
template <class T>
inline T angDistance(T from, T to)
{
    wrapToPiInPlace(from);
    wrapToPiInPlace(to);
    T d = to - from;
    if (d > M_PI)
        d -= 2. * M_PI;
    else if (d < -M_PI)
        d += 2. * M_PI;
    return d;
}

struct TPose2D;
template <typename T>
struct TPoint3D_;

template <typename T>
struct TPoint2D_
{
    T x{0}, y{0};
    constexpr TPoint2D_() = default;
    constexpr TPoint2D_(T x_, T y_) : x(x_), y(y_) {}
    explicit TPoint2D_(const TPose2D& p);
    double      norm() const { return std::sqrt(static_cast<double>(x) * x + static_cast<double>(y) * y); }
    TPoint2D_   operator-(const TPoint2D_& o) const { return {x - o.x, y - o.y}; }
    TPoint2D_   operator+(const TPoint2D_& o) const { return {x + o.x, y + o.y}; }
    TPoint2D_   operator*(T f) const { return {x * f, y * f}; }
    bool        operator==(const TPoint2D_& o) const { return x == o.x && y == o.y; }
    bool        operator!=(const TPoint2D_& o) const { return !(*this == o); }
    std::string asString() const { return mrpt::format("[%f %f]", static_cast<double>(x), static_cast<double>(y)); }
    template <typename U>
    TPoint2D_<U> cast() const
    {
        return TPoint2D_<U>(static_cast<U>(x), static_cast<U>(y));
    }
};
using TPoint2D  = TPoint2D_<double>;
using TPoint2Df = TPoint2D_<float>;

template <typename T>
struct TPoint3D_
{
    T x{0}, y{0}, z{0};
    constexpr TPoint3D_() = default;
    constexpr TPoint3D_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
    TPoint3D_& operator-=(const TPoint3D_& o)
    {
        x -= o.x;
        y -= o.y;
        z -= o.z;
        return *this;
    }
    TPoint3D_& operator+=(const TPoint3D_& o)
    {
        x += o.x;
        y += o.y;
        z += o.z;
        return *this;
    }
    template <typename U>
    TPoint3D_<U> cast() const
    {
        return TPoint3D_<U>(static_cast<U>(x), static_cast<U>(y), static_cast<U>(z));
    }
};
using TPoint3D  = TPoint3D_<double>;
using TPoint3Df = TPoint3D_<float>;

template <typename T>
struct TBoundingBox_
{
    TPoint3D_<T> min, max;
    TBoundingBox_() = default;
    TBoundingBox_(const TPoint3D_<T>& a, const TPoint3D_<T>& b) : min(a), max(b) {}
};
using TBoundingBox  = TBoundingBox_<double>;
using TBoundingBoxf = TBoundingBox_<float>;

struct TTwist2D
{
    double vx{0}, vy{0}, omega{0};
    constexpr TTwist2D() = default;
    constexpr TTwist2D(double vx_, double vy_, double w_) : vx(vx_), vy(vy_), omega(w_) {}
    // [MRPT-assumed] rotation of the linear part by `ang`
    void rotate(const double ang)
    {
        const double nvx = vx * cos(ang) - vy * sin(ang);
        const double nvy = vx * sin(ang) + vy * cos(ang);
        vx               = nvx;
        vy               = nvy;
    }
    bool        operator==(const TTwist2D& o) const { return vx == o.vx && vy == o.vy && omega == o.omega; }
    bool        operator!=(const TTwist2D& o) const { return !(*this == o); }
    std::string asString() const { return mrpt::format("[%f %f %f]", vx, vy, RAD2DEG(omega)); }
};

struct TPose2D
{
    double x{.0}, y{.0}, phi{.0};
    constexpr TPose2D() = default;
    constexpr TPose2D(double xx, double yy, double Phi) : x(xx), y(yy), phi(Phi) {}
    TPose2D(const TPoint2D& p) : x(p.x), y(p.y), phi(0) {}  // implicit, as in MRPT
    static constexpr TPose2D Identity() { return TPose2D(); }

    // [MRPT-assumed] SE(2) "oplus" / "ominus"
    TPose2D operator+(const TPose2D& b) const
    {
        const double A_cosphi = cos(this->phi), A_sinphi = sin(this->phi);
        const double new_x   = this->x + b.x * A_cosphi - b.y * A_sinphi;
        const double new_y   = this->y + b.x * A_sinphi + b.y * A_cosphi;
        const double new_phi = wrapToPi(this->phi + b.phi);
        return TPose2D(new_x, new_y, new_phi);
    }
    TPose2D operator-(const TPose2D& b) const
    {
        const double B_cosphi = cos(b.phi), B_sinphi = sin(b.phi);
        const double new_x   = (this->x - b.x) * B_cosphi + (this->y - b.y) * B_sinphi;
        const double new_y   = -(this->x - b.x) * B_sinphi + (this->y - b.y) * B_cosphi;
        const double new_phi = wrapToPi(this->phi - b.phi);
        return TPose2D(new_x, new_y, new_phi);
    }
    TPoint2D    translation() const { return {x, y}; }
    double      norm() const { return std::sqrt(x * x + y * y); }
    std::string asString() const { return mrpt::format("[%f %f %f]", x, y, RAD2DEG(phi)); }
};
inline bool operator==(const TPose2D& a, const TPose2D& b)
{
    return a.x == b.x && a.y == b.y && wrapTo2Pi(a.phi) == wrapTo2Pi(b.phi);
}
inline bool          operator!=(const TPose2D& a, const TPose2D& b) { return !(a == b); }
inline std::ostream& operator<<(std::ostream& o, const TPose2D& p) { return o << p.asString(); }
inline std::ostream& operator<<(std::ostream& o, const TPoint2D& p) { return o << p.asString(); }
inline std::ostream& operator<<(std::ostream& o, const TTwist2D& p) { return o << p.asString(); }

template <typename T>
inline TPoint2D_<T>::TPoint2D_(const TPose2D& p) : x(static_cast<T>(p.x)), y(static_cast<T>(p.y))
{
}

// TPS_Astar.cpp:501 subtracts a TPose2D from a TPoint2D. Which MRPT overload that resolves to could
// not be established without the MRPT sources; it is taken as the plain vector difference, the same
// reading the oracle restatement uses (DESIGN.md §6 lists it among the unverifiable assumptions).
inline TPoint2D operator-(const TPoint2D& a, const TPose2D& b) { return {a.x - b.x, a.y - b.y}; }

// [MRPT-assumed] TPolygon2D::contains = PNPOLY crossing test
struct TPolygon2D : public std::vector<TPoint2D>
{
    TPolygon2D() = default;
    explicit TPolygon2D(size_t n) : std::vector<TPoint2D>(n) {}
    TPolygon2D(const std::vector<TPoint2D>& v) : std::vector<TPoint2D>(v) {}
    bool contains(const TPoint2D& P) const
    {
        const size_t n = size();
        bool         c = false;
        for (size_t i = 0, j = n - 1; i < n; j = i++)
        {
            const auto& vi = (*this)[i];
            const auto& vj = (*this)[j];
            if (((vi.y > P.y) != (vj.y > P.y)) && (P.x < (vj.x - vi.x) * (P.y - vi.y) / (vj.y - vi.y) + vi.x)) c = !c;
        }
        return c;
    }
};

class CPolygon : public TPolygon2D
{
   public:
    void   AddVertex(double x, double y) { emplace_back(x, y); }
    double GetVertex_x(size_t i) const { return at(i).x; }
    double GetVertex_y(size_t i) const { return at(i).y; }
    size_t verticesCount() const { return size(); }
    bool   PointIntoPolygon(double x, double y) const { return contains(TPoint2D(x, y)); }
};

// [MRPT-assumed] poly_roots.h == Khashin's poly34
inline int solve_poly4(double* x, double a, double b, double c, double d) { return orc::poly::solve_poly4(x, a, b, c, d); }
inline int solve_poly3(double* x, double a, double b, double c) { return orc::poly::solve_poly3(x, a, b, c); }
inline int solve_poly2(double a, double b, double c, double& r1, double& r2)
{
    return orc::poly::solve_poly2(a, b, c, r1, r2);
}

template <typename T, size_t N>
struct CVectorFixed
{
    T             v[N] = {};
    T&            operator[](size_t i) { return v[i]; }
    const T&      operator[](size_t i) const { return v[i]; }
    CVectorFixed& operator-=(const CVectorFixed& o)
    {
        for (size_t i = 0; i < N; i++) v[i] -= o.v[i];
        return *this;
    }
    double norm() const
    {
        double s = 0;
        for (size_t i = 0; i < N; i++) s += v[i] * v[i];
        return std::sqrt(s);
    }
};

// [MRPT-assumed] CMatrixFixed::lu_solve = Eigen partial-pivot LU; restated as Gaussian
// elimination with row pivoting (HolonomicBlend.cpp:450).
template <typename T, size_t R, size_t C>
struct CMatrixFixed
{
    T        m[R][C] = {};
    T&       operator()(size_t r, size_t c) { return m[r][c]; }
    const T& operator()(size_t r, size_t c) const { return m[r][c]; }
    CVectorFixed<T, R> lu_solve(const CVectorFixed<T, R>& b) const
    {
        static_assert(R == C, "square only");
        T A[R][C + 1];
        for (size_t i = 0; i < R; i++)
        {
            for (size_t j = 0; j < C; j++) A[i][j] = m[i][j];
            A[i][C] = b[i];
        }
        for (size_t col = 0; col < C; col++)
        {
            size_t piv = col;
            for (size_t i = col + 1; i < R; i++)
                if (std::fabs(A[i][col]) > std::fabs(A[piv][col])) piv = i;
            if (piv != col)
                for (size_t j = 0; j < C + 1; j++) std::swap(A[col][j], A[piv][j]);
            for (size_t i = col + 1; i < R; i++)
            {
                const T f = A[i][col] / A[col][col];
                for (size_t j = col; j < C + 1; j++) A[i][j] -= f * A[col][j];
            }
        }
        CVectorFixed<T, R> x;
        for (int i = static_cast<int>(R) - 1; i >= 0; i--)
        {
            T s = A[i][C];
            for (size_t j = i + 1; j < C; j++) s -= A[i][j] * x[j];
            x[i] = s / A[i][i];
        }
        return x;
    }
};
using CMatrixDouble44 = CMatrixFixed<double, 4, 4>;
using CMatrixDouble33 = CMatrixFixed<double, 3, 3>;

// "[a b c]" row-vector parser (SE2_KinState.cpp:63-84)
class CMatrixDouble
{
   public:
    bool fromMatlabStringFormat(const std::string& s)
    {
        vals_.clear();
        const auto a = s.find('['), b = s.rfind(']');
        if (a == std::string::npos || b == std::string::npos || b < a) return false;
        std::string body = s.substr(a + 1, b - a - 1);
        for (auto& ch : body)
            if (ch == ',') ch = ' ';
        std::istringstream is(body);
        double             v;
        while (is >> v) vals_.push_back(v);
        return !vals_.empty();
    }
    int    rows() const { return vals_.empty() ? 0 : 1; }
    int    cols() const { return static_cast<int>(vals_.size()); }
    double operator()(int, int c) const { return vals_.at(c); }

   private:
    std::vector<double> vals_;
};
}  // namespace mrpt::math

// ---------------------------------------------------------------------------------------------
// mrpt::poses
// ---------------------------------------------------------------------------------------------
namespace mrpt::poses
{
// [MRPT-assumed] CPose2D with cached cos/sin; unary minus = inverse; composePoint
class CPose2D
{
   public:
    CPose2D() = default;
    CPose2D(double x, double y, double phi) : x_(x), y_(y), phi_(phi) {}
    explicit CPose2D(const mrpt::math::TPose2D& p) : x_(p.x), y_(p.y), phi_(p.phi) {}
    double x() const { return x_; }
    double y() const { return y_; }
    double phi() const { return phi_; }
    void   composePoint(double lx, double ly, double& gx, double& gy) const
    {
        update_cached_cos_sin();
        gx = x_ + lx * cosphi_ - ly * sinphi_;
        gy = y_ + lx * sinphi_ + ly * cosphi_;
    }
    void inverse()
    {
        update_cached_cos_sin();
        const double x = x_, y = y_;
        x_        = -x * cosphi_ - y * sinphi_;
        y_        = x * sinphi_ - y * cosphi_;
        phi_      = -phi_;
        uptodate_ = false;
    }
    mrpt::math::TPose2D asTPose() const { return {x_, y_, phi_}; }

   private:
    void update_cached_cos_sin() const
    {
        if (uptodate_) return;
        cosphi_   = ::cos(phi_);
        sinphi_   = ::sin(phi_);
        uptodate_ = true;
    }
    double         x_ = 0, y_ = 0, phi_ = 0;
    mutable double cosphi_ = 1, sinphi_ = 0;
    mutable bool   uptodate_ = false;
};
inline CPose2D operator-(const CPose2D& p)
{
    CPose2D r = p;
    r.inverse();
    return r;
}
class CPose3D
{
};
class CPose2DInterpolator
{
};
}  // namespace mrpt::poses

// ---------------------------------------------------------------------------------------------
// mrpt::rtti / serialization: just enough for the class declarations and a by-name PTG factory
// ---------------------------------------------------------------------------------------------
namespace mrpt::serialization
{
class CArchive
{
   public:
    virtual ~CArchive() = default;
    virtual void write(const void*, size_t) { THROW_EXCEPTION("serialization is not available in the shim"); }
    virtual void read(void*, size_t) { THROW_EXCEPTION("serialization is not available in the shim"); }
};
template <class T, std::enable_if_t<std::is_arithmetic_v<T>, int> = 0>
inline CArchive& operator<<(CArchive& a, const T& v)
{
    a.write(&v, sizeof(T));
    return a;
}
template <class T, std::enable_if_t<std::is_arithmetic_v<T>, int> = 0>
inline CArchive& operator>>(CArchive& a, T& v)
{
    a.read(&v, sizeof(T));
    return a;
}
inline CArchive& operator<<(CArchive& a, const std::string& s)
{
    const uint32_t n = static_cast<uint32_t>(s.size());
    a << n;
    if (n) a.write(s.data(), n);
    return a;
}
inline CArchive& operator>>(CArchive& a, std::string& s)
{
    uint32_t n = 0;
    a >> n;
    s.resize(n);
    if (n) a.read(&s[0], n);
    return a;
}
inline CArchive& operator<<(CArchive& a, const mrpt::math::CPolygon& p)
{
    const uint32_t n = static_cast<uint32_t>(p.size());
    a << n;
    for (const auto& v : p) a << v.x << v.y;
    return a;
}
inline CArchive& operator>>(CArchive& a, mrpt::math::CPolygon& p)
{
    uint32_t n = 0;
    a >> n;
    p.clear();
    for (uint32_t i = 0; i < n; i++)
    {
        double x, y;
        a >> x >> y;
        p.AddVertex(x, y);
    }
    return a;
}
template <class T>
inline CArchive& operator<<(CArchive& a, const std::vector<T>& v)
{
    const uint32_t n = static_cast<uint32_t>(v.size());
    a << n;
    for (const auto& e : v) a << e;
    return a;
}
template <class T>
inline CArchive& operator>>(CArchive& a, std::vector<T>& v)
{
    uint32_t n = 0;
    a >> n;
    v.resize(n);
    for (auto& e : v) a >> e;
    return a;
}
template <class STREAM>
inline CArchive archiveFrom(STREAM&)
{
    return CArchive();
}
class CSerializable;
}  // namespace mrpt::serialization

namespace mrpt::io
{
// The on-disk collision-grid cache (ptg_grid_NNN.dat.gz) is never found and never written:
// the grid is always recomputed (DiffDriveCollisionGridBased.cpp:664-759, "else" branch).
class CFileGZInputStream
{
   public:
    explicit CFileGZInputStream(const std::string&) {}
    bool fileOpenCorrectly() const { return false; }
};
class CFileGZOutputStream
{
   public:
    explicit CFileGZOutputStream(const std::string&) {}
    bool fileOpenCorrectly() const { return false; }
};
}  // namespace mrpt::io

namespace mrpt::rtti
{
class CObject
{
   public:
    using Ptr          = std::shared_ptr<CObject>;
    virtual ~CObject() = default;
};
struct TRuntimeClassId
{
    const char*                               className;
    std::function<std::shared_ptr<CObject>()> create;
};
inline std::map<std::string, const TRuntimeClassId*>& registry()
{
    static std::map<std::string, const TRuntimeClassId*> r;
    return r;
}
inline void registerClass(const TRuntimeClassId* id) { registry()[id->className] = id; }
inline void registerClassCustomName(const char* name, const TRuntimeClassId* id) { registry()[name] = id; }
inline const TRuntimeClassId* findRegisteredClass(const std::string& name)
{
    auto it = registry().find(name);
    return it == registry().end() ? nullptr : it->second;
}
struct AutoRegister
{
    explicit AutoRegister(const TRuntimeClassId* id) { registerClass(id); }
};
}  // namespace mrpt::rtti

#define MRPT_SHIM_PTR_CREATE(C)                             \
   public:                                                  \
    using Ptr      = std::shared_ptr<C>;                    \
    using ConstPtr = std::shared_ptr<const C>;              \
    template <typename... Args>                             \
    static Ptr Create(Args&&... args)                       \
    {                                                       \
        return std::make_shared<C>(std::forward<Args>(args)...); \
    }

#define DEFINE_MRPT_OBJECT(C, NS)                            \
    MRPT_SHIM_PTR_CREATE(C)                                  \
    static const mrpt::rtti::TRuntimeClassId& GetRuntimeClassIdStatic(); \
                                                             \
   private:

#define DEFINE_VIRTUAL_MRPT_OBJECT(C, NS)      \
   public:                                     \
    using Ptr      = std::shared_ptr<C>;       \
    using ConstPtr = std::shared_ptr<const C>; \
                                               \
   private:

#define IMPLEMENTS_MRPT_OBJECT(C, BASE, NS)                                              \
    const mrpt::rtti::TRuntimeClassId& NS::C::GetRuntimeClassIdStatic()                  \
    {                                                                                    \
        static const mrpt::rtti::TRuntimeClassId id{                                     \
            #NS "::" #C, []() -> std::shared_ptr<mrpt::rtti::CObject> {                  \
                return std::static_pointer_cast<mrpt::rtti::CObject>(std::make_shared<NS::C>()); \
            }};                                                                          \
        return id;                                                                       \
    }                                                                                    \
    static mrpt::rtti::AutoRegister mrpt_shim_autoreg_##C(&NS::C::GetRuntimeClassIdStatic());

#define IMPLEMENTS_VIRTUAL_MRPT_OBJECT(C, BASE, NS)

#define DEFINE_SERIALIZABLE(C, NS)                                         \
    DEFINE_MRPT_OBJECT(C, NS)                                              \
   protected:                                                              \
    uint8_t serializeGetVersion() const;                                   \
    void    serializeTo(mrpt::serialization::CArchive& out) const;         \
    void    serializeFrom(mrpt::serialization::CArchive& in, uint8_t version); \
                                                                           \
   private:

#define IMPLEMENTS_SERIALIZABLE(C, BASE, NS) IMPLEMENTS_MRPT_OBJECT(C, BASE, NS)
#define CLASS_ID(C) (&C::GetRuntimeClassIdStatic())

namespace mrpt::serialization
{
class CSerializable : public mrpt::rtti::CObject
{
};
}  // namespace mrpt::serialization

namespace mrpt::typemeta
{
}
#define MRPT_DECLARE_TTYPENAME_NAMESPACE(T, NS)
#define MRPT_ENUM_TYPE_BEGIN(T)
#define MRPT_ENUM_TYPE_END()
#define MRPT_FILL_ENUM_MEMBER(a, b)

// ---------------------------------------------------------------------------------------------
// mrpt::system — logger / profiler (the profiler only counts entries; tests read the counts)
// ---------------------------------------------------------------------------------------------
namespace mrpt::system
{
using TTimeStamp = mrpt::Clock::time_point;
inline double timeDifference(const TTimeStamp& a, const TTimeStamp& b)
{
    return std::chrono::duration<double>(b - a).count();
}
class COutputLogger
{
   public:
    COutputLogger() = default;
    explicit COutputLogger(const std::string&) {}
    virtual ~COutputLogger() = default;
};
using output_logger_callback_t = std::function<void()>;

class CTimeLogger
{
   public:
    CTimeLogger(bool enabled = true, const std::string& name = "") : name_(name) { (void)enabled; }
    void                          setName(const std::string& n) { name_ = n; }
    void                          enter(const char* n) { counts_[n]++; }
    std::map<std::string, size_t> counts_;
    void                          clear() { counts_.clear(); }

   private:
    std::string name_;
};
class CTimeLoggerEntry
{
   public:
    CTimeLoggerEntry(CTimeLogger& l, const char* name) { l.enter(name); }
    void stop() {}
};
class CTicTac
{
   public:
    void   Tic() { t0_ = mrpt::Clock::nowDouble(); }
    double Tac() const { return mrpt::Clock::nowDouble() - t0_; }

   private:
    double t0_ = 0;
};
inline std::string trim(const std::string& s)
{
    const auto a = s.find_first_not_of(" \t\r\n");
    if (a == std::string::npos) return "";
    const auto b = s.find_last_not_of(" \t\r\n");
    return s.substr(a, b - a + 1);
}
}  // namespace mrpt::system

// ---------------------------------------------------------------------------------------------
// mrpt::containers — CDynamicGrid [MRPT-assumed], yaml (tiny), map_as_vector (name only)
// ---------------------------------------------------------------------------------------------
namespace mrpt::containers
{
template <class T>
class CDynamicGrid
{
   protected:
    std::vector<T> m_map;
    double         m_x_min{0}, m_x_max{0}, m_y_min{0}, m_y_max{0}, m_resolution{0};
    size_t         m_size_x{0}, m_size_y{0};

   public:
    CDynamicGrid(double x_min = -10., double x_max = 10., double y_min = -10., double y_max = 10., double resolution = 0.10)
    {
        setSize(x_min, x_max, y_min, y_max, resolution);
    }
    virtual ~CDynamicGrid() = default;

    void setSize(
        const double x_min, const double x_max, const double y_min, const double y_max, const double resolution,
        const T* fill_value = nullptr)
    {
        m_x_min      = resolution * mrpt::round(x_min / resolution);
        m_y_min      = resolution * mrpt::round(y_min / resolution);
        m_x_max      = resolution * mrpt::round(x_max / resolution);
        m_y_max      = resolution * mrpt::round(y_max / resolution);
        m_resolution = resolution;
        m_size_x     = mrpt::round((m_x_max - m_x_min) / m_resolution);
        m_size_y     = mrpt::round((m_y_max - m_y_min) / m_resolution);
        if (fill_value)
            m_map.assign(m_size_x * m_size_y, *fill_value);
        else
            m_map.resize(m_size_x * m_size_y);
    }
    void   clear() { m_map.clear(); m_map.resize(m_size_x * m_size_y); }
    size_t getSizeX() const { return m_size_x; }
    size_t getSizeY() const { return m_size_y; }
    double getXMin() const { return m_x_min; }
    double getXMax() const { return m_x_max; }
    double getYMin() const { return m_y_min; }
    double getYMax() const { return m_y_max; }
    double getResolution() const { return m_resolution; }
    int    x2idx(double x) const { return static_cast<int>((x - m_x_min) / m_resolution); }
    int    y2idx(double y) const { return static_cast<int>((y - m_y_min) / m_resolution); }
    double idx2x(int cx) const { return m_x_min + (cx + 0.5) * m_resolution; }
    double idx2y(int cy) const { return m_y_min + (cy + 0.5) * m_resolution; }
    T*     cellByPos(double x, double y)
    {
        const int cx = x2idx(x), cy = y2idx(y);
        if (cx < 0 || cx >= static_cast<int>(m_size_x)) return nullptr;
        if (cy < 0 || cy >= static_cast<int>(m_size_y)) return nullptr;
        return &m_map[cx + cy * m_size_x];
    }
    const T* cellByPos(double x, double y) const { return const_cast<CDynamicGrid*>(this)->cellByPos(x, y); }
    T*       cellByIndex(unsigned int cx, unsigned int cy)
    {
        if (cx >= m_size_x || cy >= m_size_y) return nullptr;
        return &m_map[cx + cy * m_size_x];
    }
    const T* cellByIndex(unsigned int cx, unsigned int cy) const
    {
        return const_cast<CDynamicGrid*>(this)->cellByIndex(cx, cy);
    }
    const std::vector<T>& shim_cells() const { return m_map; }
};

template <class K, class V, class C = std::vector<std::pair<K, V>>>
class map_as_vector;

// Tiny YAML-like tree: scalars kept as text, maps and sequences. Enough for the *_yaml() members.
class yaml
{
   public:
    using sequence_t = std::vector<yaml>;
    using map_t      = std::map<std::string, yaml>;
    yaml()           = default;
    template <class T, std::enable_if_t<std::is_arithmetic_v<T>, int> = 0>
    yaml(T v)
    {
        *this = v;
    }
    yaml(const std::string& s) : kind_(SCALAR), scalar_(s) {}
    yaml(const char* s) : kind_(SCALAR), scalar_(s) {}
    static yaml Map()
    {
        yaml y;
        y.kind_ = MAP;
        return y;
    }
    static yaml Sequence()
    {
        yaml y;
        y.kind_ = SEQ;
        return y;
    }
    static yaml FromText(const std::string& text);
    static yaml FromFile(const std::string& fn)
    {
        std::ifstream f(fn);
        if (!f) THROW_EXCEPTION_FMT("cannot open yaml file '%s'", fn.c_str());
        std::stringstream ss;
        ss << f.rdbuf();
        return FromText(ss.str());
    }
    bool isMap() const { return kind_ == MAP; }
    bool isSequence() const { return kind_ == SEQ; }
    bool isScalar() const { return kind_ == SCALAR; }
    bool isNullNode() const { return kind_ == NUL; }
    bool empty() const { return kind_ == NUL || (kind_ == MAP && map_.empty()) || (kind_ == SEQ && seq_.empty()); }
    bool has(const std::string& k) const { return kind_ == MAP && map_.count(k) != 0; }
    yaml& operator[](const std::string& k)
    {
        if (kind_ == NUL) kind_ = MAP;
        ASSERT_(kind_ == MAP);
        return map_[k];
    }
    const yaml& operator[](const std::string& k) const
    {
        ASSERT_(kind_ == MAP);
        auto it = map_.find(k);
        if (it == map_.end()) THROW_EXCEPTION_FMT("yaml: missing key '%s'", k.c_str());
        return it->second;
    }
    yaml& operator()(int i) { return seq_.at(i); }
    const yaml& operator()(int i) const { return seq_.at(i); }
    sequence_t& asSequence()
    {
        ASSERT_(kind_ == SEQ);
        return seq_;
    }
    const sequence_t& asSequence() const
    {
        ASSERT_(kind_ == SEQ);
        return seq_;
    }
    map_t& asMap() { return map_; }
    const map_t& asMap() const { return map_; }
    template <class T>
    yaml& operator=(const T& v)
    {
        kind_ = SCALAR;
        if constexpr (std::is_same_v<T, bool>)
            scalar_ = v ? "true" : "false";
        else if constexpr (std::is_arithmetic_v<T>)
        {
            std::ostringstream os;
            os.precision(17);
            os << v;
            scalar_ = os.str();
        }
        else
            scalar_ = v;
        return *this;
    }
    template <class T>
    T as() const
    {
        ASSERT_(kind_ == SCALAR);
        if constexpr (std::is_same_v<T, bool>)
            return scalar_ == "true" || scalar_ == "True" || scalar_ == "1" || scalar_ == "yes";
        else if constexpr (std::is_same_v<T, std::string>)
            return scalar_;
        else if constexpr (std::is_floating_point_v<T>)
            return static_cast<T>(std::strtod(scalar_.c_str(), nullptr));
        else
            return static_cast<T>(std::strtoll(scalar_.c_str(), nullptr, 10));
    }
    template <class T>
    T getOrDefault(const std::string& k, const T& def) const
    {
        return has(k) ? (*this)[k].template as<T>() : def;
    }
    template <class T>
    std::vector<T> toStdVector() const
    {
        ASSERT_(kind_ == SEQ);
        std::vector<T> r;
        for (const auto& e : seq_) r.push_back(e.template as<T>());
        return r;
    }

   private:
    enum Kind
    {
        NUL,
        SCALAR,
        MAP,
        SEQ
    } kind_ = NUL;
    std::string scalar_;
    map_t       map_;
    sequence_t  seq_;
};
inline std::ostream& operator<<(std::ostream& o, const yaml&) { return o << "(yaml)"; }

// Flat "key: value" / "key: [a, b]" documents only (the share/*.yaml parameter files).
inline yaml yaml::FromText(const std::string& text)
{
    yaml               root = yaml::Map();
    std::istringstream is(text);
    std::string        line;
    while (std::getline(is, line))
    {
        const auto hash = line.find('#');
        if (hash != std::string::npos) line = line.substr(0, hash);
        if (line.empty() || line[0] == '%' || line.rfind("---", 0) == 0) continue;
        const auto colon = line.find(':');
        if (colon == std::string::npos) continue;
        const std::string key = mrpt::system::trim(line.substr(0, colon));
        std::string       val = mrpt::system::trim(line.substr(colon + 1));
        if (key.empty()) continue;
        if (!val.empty() && val.front() == '[')
        {
            yaml seq = yaml::Sequence();
            val      = val.substr(1, val.rfind(']') - 1);
            std::istringstream vs(val);
            std::string        item;
            while (std::getline(vs, item, ','))
            {
                item = mrpt::system::trim(item);
                if (!item.empty()) seq.asSequence().push_back(yaml(item));
            }
            root[key] = seq;
        }
        else
            root[key] = yaml(val);
    }
    return root;
}
}  // namespace mrpt::containers

#define MCP_LOAD_REQ(Y, V)                                                            \
    do {                                                                              \
        if (!(Y).has(#V)) THROW_EXCEPTION("Required parameter `" #V "` not an existing key in dictionary."); \
        V = (Y)[#V].as<decltype(V)>();                                                \
    } while (0)
#define MCP_LOAD_OPT(Y, V)                                    \
    do {                                                      \
        if ((Y).has(#V)) V = (Y)[#V].as<decltype(V)>();       \
    } while (0)
#define MCP_LOAD_REQ_DEG(Y, V)                   \
    do {                                         \
        MCP_LOAD_REQ(Y, V);                      \
        V = mrpt::DEG2RAD(V);                    \
    } while (0)
#define MCP_LOAD_OPT_DEG(Y, V)                   \
    do {                                         \
        if ((Y).has(#V))                         \
        {                                        \
            V = (Y)[#V].as<decltype(V)>();       \
            V = mrpt::DEG2RAD(V);                \
        }                                        \
    } while (0)
#define MCP_SAVE(Y, V) (Y)[#V] = V
#define MCP_SAVE_DEG(Y, V) (Y)[#V] = mrpt::RAD2DEG(V)

// ---------------------------------------------------------------------------------------------
// mrpt::config — INI files with `@define NAME value` and `${NAME}` substitution, `//` and `#`
// comments; CConfigFilePrefixer for the PTG<n>_ key prefixes. [MRPT-assumed]
// ---------------------------------------------------------------------------------------------
namespace mrpt::config
{
class CConfigFileBase
{
   public:
    virtual ~CConfigFileBase() = default;
    virtual bool        rawHas(const std::string& section, const std::string& key) const               = 0;
    virtual std::string rawGet(const std::string& section, const std::string& key) const               = 0;
    virtual void rawSet(const std::string& section, const std::string& key, const std::string& value) = 0;

    std::string read_string(const std::string& s, const std::string& k, const std::string& def, bool failIfNotFound = false) const
    {
        if (!rawHas(s, k))
        {
            if (failIfNotFound) THROW_EXCEPTION_FMT("Value '%s' not found in section '%s'", k.c_str(), s.c_str());
            return def;
        }
        return rawGet(s, k);
    }
    double read_double(const std::string& s, const std::string& k, double def, bool fail = false) const
    {
        if (!rawHas(s, k))
        {
            if (fail) THROW_EXCEPTION_FMT("Value '%s' not found in section '%s'", k.c_str(), s.c_str());
            return def;
        }
        return std::strtod(rawGet(s, k).c_str(), nullptr);
    }
    float read_float(const std::string& s, const std::string& k, float def, bool fail = false) const
    {
        return static_cast<float>(read_double(s, k, def, fail));
    }
    int read_int(const std::string& s, const std::string& k, int def, bool fail = false) const
    {
        if (!rawHas(s, k))
        {
            if (fail) THROW_EXCEPTION_FMT("Value '%s' not found in section '%s'", k.c_str(), s.c_str());
            return def;
        }
        return static_cast<int>(std::strtol(rawGet(s, k).c_str(), nullptr, 0));
    }
    uint64_t read_uint64_t(const std::string& s, const std::string& k, uint64_t def, bool fail = false) const
    {
        if (!rawHas(s, k))
        {
            if (fail) THROW_EXCEPTION_FMT("Value '%s' not found in section '%s'", k.c_str(), s.c_str());
            return def;
        }
        return std::strtoull(rawGet(s, k).c_str(), nullptr, 0);
    }
    bool read_bool(const std::string& s, const std::string& k, bool def, bool fail = false) const
    {
        if (!rawHas(s, k))
        {
            if (fail) THROW_EXCEPTION_FMT("Value '%s' not found in section '%s'", k.c_str(), s.c_str());
            return def;
        }
        const std::string v = rawGet(s, k);
        return v == "1" || v == "true" || v == "True" || v == "yes";
    }
    template <class VEC>
    void read_vector(const std::string& s, const std::string& k, const VEC& def, VEC& out, bool fail = false) const
    {
        if (!rawHas(s, k))
        {
            if (fail) THROW_EXCEPTION_FMT("Value '%s' not found in section '%s'", k.c_str(), s.c_str());
            out = def;
            return;
        }
        std::string v = rawGet(s, k);
        for (auto& ch : v)
            if (ch == ',' || ch == '[' || ch == ']') ch = ' ';
        std::istringstream                is(v);
        typename VEC::value_type          e;
        out.clear();
        // values are parsed as double and narrowed to the element type, as MRPT's tokenizer + atof does
        std::string tok;
        while (is >> tok) out.push_back(static_cast<typename VEC::value_type>(std::strtod(tok.c_str(), nullptr)));
        (void)e;
    }
    template <class T>
    void write(const std::string& s, const std::string& k, const T& v, int = -1, int = -1, const std::string& = "")
    {
        std::ostringstream os;
        os.precision(17);
        os << v;
        rawSet(s, k, os.str());
    }
};

class CConfigFileMemory : public CConfigFileBase
{
   public:
    CConfigFileMemory() = default;
    explicit CConfigFileMemory(const std::string& text) { setContent(text); }
    void setContent(const std::string& text)
    {
        data_.clear();
        std::map<std::string, std::string> defines;
        std::istringstream                 is(text);
        std::string                        line, section;
        while (std::getline(is, line))
        {
            line = mrpt::system::trim(line);
            if (line.empty() || line[0] == '#' || line[0] == ';' || line.rfind("//", 0) == 0) continue;
            if (line.rfind("@define", 0) == 0)
            {
                std::istringstream ds(line.substr(7));
                std::string        name, value;
                ds >> name;
                std::getline(ds, value);
                defines[name] = strip(value);
                continue;
            }
            if (line[0] == '[')
            {
                section = line.substr(1, line.find(']') - 1);
                continue;
            }
            const auto eq = line.find('=');
            if (eq == std::string::npos) continue;
            const std::string key = mrpt::system::trim(line.substr(0, eq));
            std::string       val = strip(line.substr(eq + 1));
            // ${NAME} substitution
            for (;;)
            {
                const auto a = val.find("${");
                if (a == std::string::npos) break;
                const auto b = val.find('}', a);
                if (b == std::string::npos) break;
                const std::string name = val.substr(a + 2, b - a - 2);
                auto              it   = defines.find(name);
                if (it == defines.end()) THROW_EXCEPTION_FMT("config: undefined variable '%s'", name.c_str());
                val = val.substr(0, a) + it->second + val.substr(b + 1);
            }
            data_[section][key] = val;
        }
    }
    bool rawHas(const std::string& s, const std::string& k) const override
    {
        auto it = data_.find(s);
        return it != data_.end() && it->second.count(k) != 0;
    }
    std::string rawGet(const std::string& s, const std::string& k) const override { return data_.at(s).at(k); }
    void        rawSet(const std::string& s, const std::string& k, const std::string& v) override { data_[s][k] = v; }

   private:
    static std::string strip(std::string v)
    {
        // trailing comments: " // ..." and " # ..."
        auto c = v.find("//");
        if (c != std::string::npos) v = v.substr(0, c);
        c = v.find('#');
        if (c != std::string::npos) v = v.substr(0, c);
        return mrpt::system::trim(v);
    }
    std::map<std::string, std::map<std::string, std::string>> data_;
};

class CConfigFile : public CConfigFileMemory
{
   public:
    explicit CConfigFile(const std::string& fn)
    {
        std::ifstream f(fn);
        if (!f) THROW_EXCEPTION_FMT("cannot open config file '%s'", fn.c_str());
        std::stringstream ss;
        ss << f.rdbuf();
        setContent(ss.str());
    }
};

class CConfigFilePrefixer : public CConfigFileBase
{
   public:
    CConfigFilePrefixer(const CConfigFileBase& o, const std::string& prefSec, const std::string& prefKey)
        : o_(const_cast<CConfigFileBase&>(o)), ps_(prefSec), pk_(prefKey)
    {
    }
    bool        rawHas(const std::string& s, const std::string& k) const override { return o_.rawHas(ps_ + s, pk_ + k); }
    std::string rawGet(const std::string& s, const std::string& k) const override { return o_.rawGet(ps_ + s, pk_ + k); }
    void rawSet(const std::string& s, const std::string& k, const std::string& v) override { o_.rawSet(ps_ + s, pk_ + k, v); }

   private:
    CConfigFileBase& o_;
    std::string      ps_, pk_;
};
}  // namespace mrpt::config

// MRPT's loader macros paste the type token into the reader's name (read_double, read_string ...).
namespace mrpt::config
{
using string = std::string;
}
#define MRPT_LOAD_CONFIG_VAR(V, T, CFG, SEC) \
    {                                        \
        V = (CFG).read_##T(SEC, #V, V);      \
    }
#define MRPT_LOAD_CONFIG_VAR_NO_DEFAULT(V, T, CFG, SEC) \
    {                                                   \
        V = (CFG).read_##T(SEC, #V, V, true);           \
    }
#define MRPT_LOAD_HERE_CONFIG_VAR(NAME, T, TARGET, CFG, SEC) TARGET = (CFG).read_##T(SEC, #NAME, TARGET, false);
#define MRPT_LOAD_HERE_CONFIG_VAR_NO_DEFAULT(NAME, T, TARGET, CFG, SEC) \
    {                                                                   \
        TARGET = (CFG).read_##T(SEC, #NAME, TARGET, true);              \
    }
#define MRPT_LOAD_HERE_CONFIG_VAR_DEGREES_NO_DEFAULT(NAME, T, TARGET, CFG, SEC) \
    {                                                                           \
        TARGET = mrpt::DEG2RAD((CFG).read_##T(SEC, #NAME, TARGET, true));       \
    }

// ---------------------------------------------------------------------------------------------
// mrpt::img / opengl / obs / kinematics — visualisation and command types: stubs
// ---------------------------------------------------------------------------------------------
namespace mrpt::img
{
struct TColor
{
    uint8_t R{0}, G{0}, B{0}, A{255};
    constexpr TColor() = default;
    constexpr TColor(uint8_t r, uint8_t g, uint8_t b, uint8_t a = 255) : R(r), G(g), B(b), A(a) {}
    constexpr TColor(unsigned int rgb, uint8_t a) : R((rgb >> 16) & 0xff), G((rgb >> 8) & 0xff), B(rgb & 0xff), A(a) {}
    static constexpr TColor black() { return TColor(0, 0, 0); }
};
enum TImageChannels
{
    CH_GRAY = 1,
    CH_RGB  = 3
};
enum TColormap
{
    cmJET = 0
};
inline TColor colormap(TColormap, double) { return TColor(); }
class CImage
{
   public:
    CImage(unsigned w, unsigned h, TImageChannels ch) : w_(w), ch_(ch), d_(static_cast<size_t>(w) * h * ch) {}
    void     filledRectangle(int, int, int, int, const TColor&) {}
    uint8_t* operator()(unsigned x, unsigned y) { return &d_[(static_cast<size_t>(y) * w_ + x) * ch_]; }

   private:
    unsigned             w_;
    int                  ch_;
    std::vector<uint8_t> d_;
};
}  // namespace mrpt::img

namespace mrpt::opengl
{
class CRenderizable
{
   public:
    using Ptr                = std::shared_ptr<CRenderizable>;
    virtual ~CRenderizable() = default;
    void setName(const std::string&) {}
    void setColor_u8(uint8_t, uint8_t, uint8_t, uint8_t = 255) {}
    void setLocation(double, double, double) {}
};
class CSetOfObjects : public CRenderizable
{
   public:
    using Ptr = std::shared_ptr<CSetOfObjects>;
    static Ptr Create() { return std::make_shared<CSetOfObjects>(); }
    template <class P>
    void insert(const P&)
    {
    }
};
class CTexturedPlane : public CRenderizable
{
   public:
    using Ptr = std::shared_ptr<CTexturedPlane>;
    static Ptr Create() { return std::make_shared<CTexturedPlane>(); }
    void       setPlaneCorners(float, float, float, float) {}
    void       assignImage(const mrpt::img::CImage&, const mrpt::img::CImage&) {}
};
class CDisk : public CRenderizable
{
   public:
    using Ptr = std::shared_ptr<CDisk>;
    static Ptr Create() { return std::make_shared<CDisk>(); }
    void       setDiskRadius(float, float) {}
};
class CSetOfLines;
class CArrow;
class COpenGLScene
{
   public:
    template <class P>
    void insert(const P&)
    {
    }
    bool saveToFile(const std::string&) const { return false; }
};
}  // namespace mrpt::opengl

namespace mrpt::obs
{
class CObservation
{
   public:
    using Ptr               = std::shared_ptr<CObservation>;
    virtual ~CObservation() = default;
};
}  // namespace mrpt::obs

namespace mrpt::kinematics
{
class CVehicleVelCmd
{
   public:
    using Ptr                 = std::shared_ptr<CVehicleVelCmd>;
    virtual ~CVehicleVelCmd() = default;
    struct TVelCmdParams
    {
    };
};
class CVehicleVelCmd_DiffDriven : public CVehicleVelCmd
{
   public:
    double lin_vel = 0, ang_vel = 0;
};
class CVehicleVelCmd_Holo : public CVehicleVelCmd
{
   public:
    using Ptr = std::shared_ptr<CVehicleVelCmd_Holo>;
    static Ptr Create() { return std::make_shared<CVehicleVelCmd_Holo>(); }
    double     vel = 0, dir_local = 0, ramp_time = 0, rot_speed = 0;
};
class CVehicleSimul_Holo;
}  // namespace mrpt::kinematics

// ---------------------------------------------------------------------------------------------
// mrpt::maps — float SoA point cloud; nanoflann semantics restated as brute force [MRPT-assumed]
// ---------------------------------------------------------------------------------------------
namespace mrpt::maps
{
class COccupancyGridMap2D
{
   public:
    using Ptr = std::shared_ptr<COccupancyGridMap2D>;
    static Ptr Create() { return std::make_shared<COccupancyGridMap2D>(); }
    void       setSize(float xmin, float xmax, float ymin, float ymax, float res)
    {
        nx_ = static_cast<unsigned>(mrpt::round((xmax - xmin) / res));
        ny_ = static_cast<unsigned>(mrpt::round((ymax - ymin) / res));
        d_.assign(static_cast<size_t>(nx_) * ny_, 0);
    }
    unsigned getSizeX() const { return nx_; }
    unsigned getSizeY() const { return ny_; }
    int8_t*  getRow(int cy) { return cy >= 0 && static_cast<unsigned>(cy) < ny_ ? &d_[static_cast<size_t>(cy) * nx_] : nullptr; }

   private:
    unsigned            nx_ = 0, ny_ = 0;
    std::vector<int8_t> d_;
};

class CPointsMap
{
   public:
    using Ptr             = std::shared_ptr<CPointsMap>;
    virtual ~CPointsMap() = default;

    void getPointsBuffer(size_t& n, const float*& xs, const float*& ys, const float*& zs) const
    {
        n  = x_.size();
        xs = x_.data();
        ys = y_.data();
        zs = z_.data();
    }
    const std::vector<float>& getPointsBufferRef_x() const { return x_; }
    const std::vector<float>& getPointsBufferRef_y() const { return y_; }
    void                      clear()
    {
        x_.clear();
        y_.clear();
        z_.clear();
    }
    void reserve(size_t n)
    {
        x_.reserve(n);
        y_.reserve(n);
        z_.reserve(n);
    }
    size_t size() const { return x_.size(); }
    bool   empty() const { return x_.empty(); }
    void   insertPointFast(float x, float y, float z = 0)
    {
        x_.push_back(x);
        y_.push_back(y);
        z_.push_back(z);
    }
    void insertPoint(float x, float y, float z = 0) { insertPointFast(x, y, z); }
    void insertPointFrom(const CPointsMap& o, size_t i) { insertPointFast(o.x_[i], o.y_[i], o.z_[i]); }
    void getPoint(size_t i, float& x, float& y) const
    {
        x = x_[i];
        y = y_[i];
    }
    mrpt::math::TBoundingBoxf boundingBox() const
    {
        mrpt::math::TBoundingBoxf bb;
        if (x_.empty()) return bb;
        bb.min = {x_[0], y_[0], z_[0]};
        bb.max = bb.min;
        for (size_t i = 1; i < x_.size(); i++)
        {
            bb.min.x = std::min(bb.min.x, x_[i]);
            bb.max.x = std::max(bb.max.x, x_[i]);
            bb.min.y = std::min(bb.min.y, y_[i]);
            bb.max.y = std::max(bb.max.y, y_[i]);
            bb.min.z = std::min(bb.min.z, z_[i]);
            bb.max.z = std::max(bb.max.z, z_[i]);
        }
        return bb;
    }
    void kdTreeEnsureIndexBuilt2D() const {}
    // nanoflann L2_Simple_Adaptor<float>: result += (q[i]-p[i])^2 for i = 0,1, in float
    float kdTreeClosestPoint2DsqrError(float qx, float qy) const
    {
        float best = std::numeric_limits<float>::max();
        for (size_t i = 0; i < x_.size(); i++)
        {
            const float dx = qx - x_[i], dy = qy - y_[i];
            float       d  = dx * dx;
            d += dy * dy;
            if (d < best) best = d;
        }
        return best;
    }
    // nanoflann radiusSearch: points with squared distance < squared radius, sorted by distance
    size_t kdTreeRadiusSearch2D(float qx, float qy, float maxRadiusSqr, std::vector<std::pair<size_t, float>>& out) const
    {
        out.clear();
        for (size_t i = 0; i < x_.size(); i++)
        {
            const float dx = qx - x_[i], dy = qy - y_[i];
            float       d  = dx * dx;
            d += dy * dy;
            if (d < maxRadiusSqr) out.emplace_back(i, d);
        }
        std::stable_sort(out.begin(), out.end(), [](const auto& a, const auto& b) { return a.second < b.second; });
        return out.size();
    }
    template <class OBS, class POSE>
    void insertObservation(const OBS&, const POSE&)
    {
        THROW_EXCEPTION("sensor observations are not available in the shim");
    }
    bool load2D_from_text_file(const std::string& fn)
    {
        std::ifstream f(fn);
        if (!f) return false;
        clear();
        std::string line;
        while (std::getline(f, line))
        {
            std::istringstream is(line);
            float              x, y;
            if (is >> x >> y) insertPointFast(x, y, 0);
        }
        return true;
    }

   protected:
    std::vector<float> x_, y_, z_;
};
class CSimplePointsMap : public CPointsMap
{
   public:
    using Ptr = std::shared_ptr<CSimplePointsMap>;
    static Ptr Create() { return std::make_shared<CSimplePointsMap>(); }
};
}  // namespace mrpt::maps
#define NANOFLANN_VERSION 0x142

// ---------------------------------------------------------------------------------------------
// mrpt::graphs
// ---------------------------------------------------------------------------------------------
namespace mrpt::graphs
{
using TNodeID = uint64_t;
#define INVALID_NODEID_SHIM static_cast<mrpt::graphs::TNodeID>(-1)
constexpr TNodeID INVALID_NODEID = static_cast<TNodeID>(-1);

template <class TYPE_EDGES>
class CDirectedTree
{
   public:
    struct TEdgeInfo
    {
        TNodeID    id;
        bool       reverse;
        TYPE_EDGES data;
        TEdgeInfo(TNodeID child_id_, bool direction_child_to_parent = false, const TYPE_EDGES& edge_data = TYPE_EDGES())
            : id(child_id_), reverse(direction_child_to_parent), data(edge_data)
        {
        }
    };
    using TListEdges         = std::list<TEdgeInfo>;
    using TMapNode2ListEdges = std::map<TNodeID, TListEdges>;
    TNodeID            root  = INVALID_NODEID;
    TMapNode2ListEdges edges_to_children;
    void               clear()
    {
        edges_to_children.clear();
        root = INVALID_NODEID;
    }
};
}  // namespace mrpt::graphs

// ---------------------------------------------------------------------------------------------
// mrpt::expr
// ---------------------------------------------------------------------------------------------
namespace mrpt::expr
{
class CRuntimeCompiledExpression
{
   public:
    void register_symbol_table(const std::map<std::string, double*>& vars) { e_.register_symbol_table(vars); }
    void compile(const std::string& expr, const std::map<std::string, double>& = {}, const std::string& = "")
    {
        e_.compile(expr);
    }
    double eval() const { return e_.eval(); }

   private:
    orc::Expression e_;
};
}  // namespace mrpt::expr

// ---------------------------------------------------------------------------------------------
// mrpt::nav — PTG base classes [MRPT-assumed; SURVEY.md §8c table]
// ---------------------------------------------------------------------------------------------
namespace mrpt::nav
{
enum PTG_collision_behavior_t
{
    COLL_BEH_BACK_AWAY = 0,
    COLL_BEH_STOP
};

class CParameterizedTrajectoryGenerator : public mrpt::serialization::CSerializable
{
   public:
    using Ptr = std::shared_ptr<CParameterizedTrajectoryGenerator>;

    struct TNavDynamicState
    {
        mrpt::math::TTwist2D curVelLocal{0, 0, 0};
        mrpt::math::TPose2D  relTarget{20.0, 0, 0};
        double               targetRelSpeed{0};
        bool                 operator==(const TNavDynamicState& o) const
        {
            return (curVelLocal == o.curVelLocal) && (relTarget == o.relTarget) && (targetRelSpeed == o.targetRelSpeed);
        }
        bool operator!=(const TNavDynamicState& o) const { return !(*this == o); }
    };

    static constexpr uint16_t INVALID_PTG_PATH_INDEX = static_cast<uint16_t>(-1);

    static PTG_collision_behavior_t& COLLISION_BEHAVIOR()
    {
        static PTG_collision_behavior_t b = COLL_BEH_BACK_AWAY;
        return b;
    }

    static Ptr CreatePTG(
        const std::string& ptgClassName_, const mrpt::config::CConfigFileBase& cfg, const std::string& sSection,
        const std::string& sKeyPrefix)
    {
        const std::string ptgClassName = mrpt::system::trim(ptgClassName_);
        const auto*       id           = mrpt::rtti::findRegisteredClass(ptgClassName);
        if (!id) THROW_EXCEPTION_FMT("[CreatePTG] No PTG named `%s` is registered!", ptgClassName.c_str());
        auto ptg = std::dynamic_pointer_cast<CParameterizedTrajectoryGenerator>(id->create());
        if (!ptg) THROW_EXCEPTION_FMT("[CreatePTG] Object of type `%s` seems not to be a PTG!", ptgClassName.c_str());
        mrpt::config::CConfigFilePrefixer cfp(cfg, "", sKeyPrefix);
        ptg->loadFromConfigFile(cfp, sSection);
        return ptg;
    }

    virtual std::string getDescription() const = 0;
    virtual bool        inverseMap_WS2TP(double x, double y, int& out_k, double& out_normalized_d, double tolerance_dist = 0.10) const = 0;
    virtual bool        PTG_IsIntoDomain(double x, double y) const
    {
        int    k;
        double d;
        return inverseMap_WS2TP(x, y, k, d);
    }
    virtual bool isBijectiveAt(uint16_t, uint32_t) const { return true; }
    virtual mrpt::kinematics::CVehicleVelCmd::Ptr directionToMotionCommand(uint16_t k) const = 0;
    virtual mrpt::kinematics::CVehicleVelCmd::Ptr getSupportedKinematicVelocityCommand() const = 0;

    void updateNavDynamicState(const TNavDynamicState& newState, const bool force_update = false)
    {
        if (force_update || m_nav_dyn_state != newState)
        {
            ASSERT_(newState.targetRelSpeed >= .0 && newState.targetRelSpeed <= 1.0);
            m_nav_dyn_state          = newState;
            m_nav_dyn_state_target_k = INVALID_PTG_PATH_INDEX;
            this->onNewNavDynamicState();
            if (this->supportSpeedAtTarget())
            {
                int    target_k      = -1;
                double target_norm_d = 0;
                this->inverseMap_WS2TP(m_nav_dyn_state.relTarget.x, m_nav_dyn_state.relTarget.y, target_k, target_norm_d, 1.0);
                if (target_norm_d > 0.01 && target_norm_d < 0.99 && target_k >= 0 && target_k < m_alphaValuesCount)
                {
                    m_nav_dyn_state_target_k = target_k;
                    this->onNewNavDynamicState();
                }
            }
        }
    }
    const TNavDynamicState& getCurrentNavDynamicState() const { return m_nav_dyn_state; }
    virtual void            onNewNavDynamicState() = 0;

    virtual void   setRefDistance(const double refDist) { refDistance = refDist; }
    virtual size_t getPathStepCount(uint16_t k) const                                    = 0;
    virtual mrpt::math::TPose2D getPathPose(uint16_t k, uint32_t step) const            = 0;
    virtual double              getPathDist(uint16_t k, uint32_t step) const            = 0;
    virtual double              getPathStepDuration() const                             = 0;
    virtual double              getMaxLinVel() const                                    = 0;
    virtual double              getMaxAngVel() const                                    = 0;
    virtual bool   getPathStepForDist(uint16_t k, double dist, uint32_t& out_step) const = 0;
    // default: finite differences between consecutive path poses
    virtual mrpt::math::TTwist2D getPathTwist(uint16_t k, uint32_t step) const
    {
        if (step > 0)
        {
            const auto   curPose  = getPathPose(k, step);
            const auto   prevPose = getPathPose(k, step - 1);
            const double dt       = getPathStepDuration();
            ASSERT_ABOVE_(dt, 0);
            return mrpt::math::TTwist2D(
                (curPose.x - prevPose.x) / dt, (curPose.y - prevPose.y) / dt, (curPose.phi - prevPose.phi) / dt);
        }
        return m_nav_dyn_state.curVelLocal;
    }
    virtual void updateTPObstacle(double ox, double oy, std::vector<double>& tp_obstacles) const         = 0;
    virtual void updateTPObstacleSingle(double ox, double oy, uint16_t k, double& tp_obstacle_k) const = 0;

    void initTPObstacles(std::vector<double>& TP_Obstacles) const
    {
        TP_Obstacles.resize(m_alphaValuesCount);
        for (size_t k = 0; k < m_alphaValuesCount; k++) initTPObstacleSingle(k, TP_Obstacles[k]);
    }
    void initTPObstacleSingle(uint16_t k, double& TP_Obstacle_k) const
    {
        TP_Obstacle_k = std::min(refDistance, this->getPathDist(k, this->getPathStepCount(k) - 1));
    }

    virtual void loadDefaultParams()
    {
        m_alphaValuesCount = 121;
        refDistance        = 6.0;
    }
    virtual bool   supportVelCmdNOP() const { return false; }
    virtual bool   supportSpeedAtTarget() const { return false; }
    virtual double maxTimeInVelCmdNOP(int) const { return .0; }
    virtual double getActualUnloopedPathLength(uint16_t k) const { return this->getPathDist(k, getPathStepCount(k) - 1); }
    virtual double evalPathRelativePriority(uint16_t, double) const { return 1.0; }
    virtual double getMaxRobotRadius() const                        = 0;
    virtual bool   isPointInsideRobotShape(const double x, const double y) const = 0;

    double index2alpha(uint16_t k) const { return index2alpha(k, m_alphaValuesCount); }
    static double index2alpha(uint16_t k, const unsigned int num_paths)
    {
        ASSERT_BELOW_(k, num_paths);
        return M_PI * (-1.0 + 2.0 * (k + 0.5) / num_paths);
    }
    uint16_t        alpha2index(double alpha) const { return alpha2index(alpha, m_alphaValuesCount); }
    static uint16_t alpha2index(double alpha, const unsigned int num_paths)
    {
        mrpt::math::wrapToPi(alpha);  // result discarded, as upstream does
        int k = mrpt::round(0.5 * (num_paths * (1.0 + alpha / M_PI) - 1.0));
        if (k < 0) k = 0;
        if (k >= static_cast<int>(num_paths)) k = num_paths - 1;
        return static_cast<uint16_t>(k);
    }
    uint16_t getAlphaValuesCount() const { return m_alphaValuesCount; }
    uint16_t getPathCount() const { return m_alphaValuesCount; }
    double   getRefDistance() const { return refDistance; }

    void initialize(const std::string& cacheFilename = std::string(), const bool verbose = true)
    {
        if (m_is_initialized) return;
        const std::string orig_expr_target_dir;
        this->internal_initialize(cacheFilename, verbose);
        m_is_initialized = true;
    }
    void deinitialize()
    {
        if (!m_is_initialized) return;
        this->internal_deinitialize();
        m_is_initialized = false;
    }
    bool isInitialized() const { return m_is_initialized; }

    virtual void loadFromConfigFile(const mrpt::config::CConfigFileBase& cfg, const std::string& sSection)
    {
        MRPT_LOAD_HERE_CONFIG_VAR_NO_DEFAULT(num_paths, uint64_t, m_alphaValuesCount, cfg, sSection);
        MRPT_LOAD_CONFIG_VAR_NO_DEFAULT(refDistance, double, cfg, sSection);
        MRPT_LOAD_HERE_CONFIG_VAR(score_priority, double, m_score_priority, cfg, sSection);
    }
    virtual void saveToConfigFile(mrpt::config::CConfigFileBase& cfg, const std::string& sSection) const
    {
        cfg.write(sSection, "num_paths", m_alphaValuesCount);
        cfg.write(sSection, "refDistance", refDistance);
        cfg.write(sSection, "score_priority", m_score_priority);
    }

   protected:
    double           refDistance{.0};
    uint16_t         m_alphaValuesCount{0};
    double           m_score_priority{1.0};
    bool             m_is_initialized{false};
    TNavDynamicState m_nav_dyn_state;
    uint16_t         m_nav_dyn_state_target_k{INVALID_PTG_PATH_INDEX};

    virtual void internal_initialize(const std::string& cacheFilename = std::string(), const bool verbose = true) = 0;
    virtual void internal_deinitialize()                                                                         = 0;
    virtual void internal_readFromStream(mrpt::serialization::CArchive&) {}
    virtual void internal_writeToStream(mrpt::serialization::CArchive&) const {}

    void internal_TPObsDistancePostprocess(const double ox, const double oy, const double new_tp_obs_dist, double& inout_tp_obs) const
    {
        const bool is_obs_inside_robot_shape = isPointInsideRobotShape(ox, oy);
        if (!is_obs_inside_robot_shape)
        {
            mrpt::keep_min(inout_tp_obs, new_tp_obs_dist);
            return;
        }
        switch (COLLISION_BEHAVIOR())
        {
            case COLL_BEH_STOP:
                inout_tp_obs = .0;
                break;
            case COLL_BEH_BACK_AWAY:
            {
                if (new_tp_obs_dist < getMaxRobotRadius())
                    return;  // moving away from the obstacle: ignore it
                else
                    inout_tp_obs = .0;
            }
            break;
        }
    }
};

class CPTG_RobotShape_Polygonal : public CParameterizedTrajectoryGenerator
{
   public:
    void setRobotShape(const mrpt::math::CPolygon& robotShape)
    {
        ASSERT_GE_(robotShape.size(), 3u);
        m_robotShape     = robotShape;
        m_robotMaxRadius = .0;
        for (const auto& v : m_robotShape) mrpt::keep_max(m_robotMaxRadius, v.norm());
        internal_processNewRobotShape();
    }
    const mrpt::math::CPolygon& getRobotShape() const { return m_robotShape; }
    double                      getMaxRobotRadius() const override { return m_robotMaxRadius; }
    bool isPointInsideRobotShape(const double x, const double y) const override { return m_robotShape.contains({x, y}); }

   protected:
    virtual void internal_processNewRobotShape() = 0;
    mrpt::math::CPolygon m_robotShape;
    double               m_robotMaxRadius{.01};
    void                 loadDefaultParams() override {}
    void                 loadShapeFromConfigFile(const mrpt::config::CConfigFileBase&, const std::string&) {}
    void                 saveToConfigFile(mrpt::config::CConfigFileBase&, const std::string&) const override {}
    void                 internal_shape_loadFromStream(mrpt::serialization::CArchive&) {}
    void                 internal_shape_saveToStream(mrpt::serialization::CArchive&) const {}
};

class CPTG_RobotShape_Circular : public CParameterizedTrajectoryGenerator
{
   public:
    void setRobotShapeRadius(const double robot_radius)
    {
        ASSERT_(robot_radius > 0);
        m_robotRadius = robot_radius;
        internal_processNewRobotShape();
    }
    double getRobotShapeRadius() const { return m_robotRadius; }
    double getMaxRobotRadius() const override { return m_robotRadius; }
    bool   isPointInsideRobotShape(const double x, const double y) const override { return ::hypot(x, y) < m_robotRadius; }

   protected:
    virtual void internal_processNewRobotShape() = 0;
    double       m_robotRadius{.0};
    void         loadDefaultParams() override { m_robotRadius = 0.2; }
    void         loadShapeFromConfigFile(const mrpt::config::CConfigFileBase& cfg, const std::string& sSection)
    {
        const double r = cfg.read_double(sSection, "robot_radius", m_robotRadius, false);
        if (r > 0) m_robotRadius = r;
    }
    void saveToConfigFile(mrpt::config::CConfigFileBase&, const std::string&) const override {}
    void internal_shape_loadFromStream(mrpt::serialization::CArchive&) {}
    void internal_shape_saveToStream(mrpt::serialization::CArchive&) const {}
};

class CAbstractNavigator;
}  // namespace mrpt::nav

namespace mrpt
{
// several reference .cpp files use these unqualified after `using namespace mrpt;`
using mrpt::system::trim;
}  // namespace mrpt
