// ============================================================================
// TEST INFRASTRUCTURE ONLY.  CPU restatement ("oracle") of the EDS event-to-
// model alignment hot path.  Nothing in the product (slam-eds_b200/) may link,
// import or call this file; only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py use it, as the checker or
// as the timed CPU baseline.
//
// PARITY UNPINNED: the reference (uzh-rpg/slam-eds) ships no tests, golden
// vectors or fixtures for this path and cannot be compiled here (Eigen, Ceres,
// OpenCV C++, Boost, Rock base-types are absent).  The solver
// semantics below restate ceres-solver 1.14..2.1 (un-vendored dependency of the
// reference, manifest.xml:12) from its published algorithm.  What pins this
// file instead: analytic known-answer tests, the in-repo double statements of
// the same maths, cv2 (real OpenCV) for blur/norm, and finite differences
// (tests/test_oracle_*.py).
//
// Every function cites the reference file:line it follows (paths relative to
// the reference root).  double precision throughout, like the reference.
// ============================================================================
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <atomic>
#include <functional>
#include <limits>
#include <thread>
#include <vector>

namespace {

inline int clampi(int v, int lo, int hi) { return std::max(lo, std::min(v, hi)); }

// ---------------------------------------------------------------------------
// Event frame
// ---------------------------------------------------------------------------

// src/utils/Utils.hpp:542-546 (expWeight), called at src/utils/Utils.cpp:72 as
// expWeight(idx / window_size, 1.0).
inline double exp_weight(double idx, double window_size) {
    double value = (idx - (window_size / 2)) / (window_size / 6.0);
    return std::exp(-0.5 * value * value);
}

// src/utils/Utils.cpp:50-122, int8 overload of drawValuesPoints.
// method: 0 = "nn", 1 = "bilinear".  sigma > 0 => cv::GaussianBlur(3x3, sigma).
void draw_values_points(const double* px, const double* py, const int8_t* val, int E, int H,
                        int W, int method, bool use_exp, double sigma, double* img) {
    std::fill(img, img + (size_t)H * W, 0.0);
    for (int i = 0; i < E; ++i) {
        double weight = use_exp ? exp_weight((double)i / (double)(uint32_t)E, 1.0) : 1.0;
        double v = (double)val[i];
        if (method == 0) {
            // cv::Point2i = cv::Point2d uses saturate_cast<int>(double) == cvRound
            // (round-half-to-even), Utils.cpp:75; then clip, Utils.cpp:77-78.
            int xi = (int)std::nearbyint(px[i]);
            int yi = (int)std::nearbyint(py[i]);
            xi = clampi(xi, 0, W - 1);
            yi = clampi(yi, 0, H - 1);
            img[(size_t)yi * W + xi] += weight * v;
        } else {
            // Utils.cpp:85-106
            int x0 = (int)std::floor(px[i]);
            int y0 = (int)std::floor(py[i]);
            int x1 = x0 + 1, y1 = y0 + 1;
            double x = px[i], y = py[i];
            auto in = [&](int xx, int yy) { return xx < W && yy < H && xx >= 0 && yy >= 0; };
            double wa = in(x0, y0) ? (x1 - x) * (y1 - y) : 0.0;
            double wb = in(x0, y1) ? (x1 - x) * (y - y0) : 0.0;
            double wc = in(x1, y0) ? (x - x0) * (y1 - y) : 0.0;
            double wd = in(x1, y1) ? (x - x0) * (y - y0) : 0.0;
            x0 = clampi(x0, 0, W - 1);
            x1 = clampi(x1, 0, W - 1);
            y0 = clampi(y0, 0, H - 1);
            y1 = clampi(y1, 0, H - 1);
            img[(size_t)y0 * W + x0] += weight * wa * v;
            img[(size_t)y1 * W + x0] += weight * wb * v;
            img[(size_t)y0 * W + x1] += weight * wc * v;
            img[(size_t)y1 * W + x1] += weight * wd * v;
        }
    }
    if (sigma > 0) {
        // Utils.cpp:113-119: ksize = (int(1.25*240/100), int(1.7*180/100)) = (3,3)
        // whatever the image size; cv::GaussianBlur default BORDER_REFLECT_101.
        // OpenCV getGaussianKernel(3, sigma): exp(-x^2/(2 sigma^2)), normalised.
        double k1 = 1.0, k0 = std::exp(-1.0 / (2.0 * sigma * sigma));
        double s = k0 + k1 + k0;
        k0 /= s;
        k1 /= s;
        auto r101 = [](int i, int n) {
            if (n == 1) return 0;
            if (i < 0) return -i;
            if (i >= n) return 2 * n - 2 - i;
            return i;
        };
        std::vector<double> tmp((size_t)H * W);
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const double* row = img + (size_t)y * W;
                // generic RowFilter order: left, centre, right
                tmp[(size_t)y * W + x] = row[r101(x - 1, W)] * k0 + row[x] * k1 + row[r101(x + 1, W)] * k0;
            }
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                // SymmColumnFilter order: centre, then k*(up+down)
                double c = tmp[(size_t)y * W + x];
                double u = tmp[(size_t)r101(y - 1, H) * W + x];
                double d = tmp[(size_t)r101(y + 1, H) * W + x];
                img[(size_t)y * W + x] = k1 * c + k0 * (u + d);
            }
    }
}

// ---------------------------------------------------------------------------
// Forward-mode dual numbers (stand-in for ceres::Jet<double,13>)
// ---------------------------------------------------------------------------
template <int N>
struct Jet {
    double a;
    double v[N];
    Jet() : a(0) { for (int i = 0; i < N; ++i) v[i] = 0; }
    explicit Jet(double s) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0; }
    Jet(double s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0; v[k] = 1.0; }
};
template <int N> inline Jet<N> operator+(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a + y.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] + y.v[i]; return r; }
template <int N> inline Jet<N> operator-(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a - y.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] - y.v[i]; return r; }
template <int N> inline Jet<N> operator-(const Jet<N>& x) { Jet<N> r; r.a = -x.a; for (int i = 0; i < N; ++i) r.v[i] = -x.v[i]; return r; }
template <int N> inline Jet<N> operator*(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a * y.a; for (int i = 0; i < N; ++i) r.v[i] = x.a * y.v[i] + x.v[i] * y.a; return r; }
template <int N> inline Jet<N> operator/(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; double iy = 1.0 / y.a; r.a = x.a * iy; for (int i = 0; i < N; ++i) r.v[i] = (x.v[i] - r.a * y.v[i]) * iy; return r; }
template <int N> inline Jet<N>& operator+=(Jet<N>& x, const Jet<N>& y) { x = x + y; return x; }
template <int N> inline Jet<N> jsqrt(const Jet<N>& x) { Jet<N> r; r.a = std::sqrt(x.a); double d = 0.5 / r.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * d; return r; }
inline double jsqrt(double x) { return std::sqrt(x); }
template <int N> inline double scalar_of(const Jet<N>& x) { return x.a; }
inline double scalar_of(double x) { return x; }

template <typename T> struct Lift;
template <> struct Lift<double> { static double c(double s) { return s; } };
template <int N> struct Lift<Jet<N>> { static Jet<N> c(double s) { return Jet<N>(s); } };

// ---------------------------------------------------------------------------
// ceres::Grid2D<double,1> + ceres::BiCubicInterpolator semantics
// (un-vendored; call site src/tracking/PhotometricError.hpp:110-111,172).
// ---------------------------------------------------------------------------
struct Grid {
    const double* data;
    int H, W;
    inline double at(int r, int c) const {
        r = clampi(r, 0, H - 1);
        c = clampi(c, 0, W - 1);
        return data[(size_t)r * W + c];
    }
};

inline void cubic_hermite(double p0, double p1, double p2, double p3, double x, double* f, double* dfdx) {
    const double a = 0.5 * (-p0 + 3.0 * p1 - 3.0 * p2 + p3);
    const double b = 0.5 * (2.0 * p0 - 5.0 * p1 + 4.0 * p2 - p3);
    const double c = 0.5 * (-p0 + p2);
    const double d = p1;
    if (f) *f = d + x * (c + x * (b + x * a));
    if (dfdx) *dfdx = c + x * (2.0 * b + 3.0 * a * x);
}

// Coordinates that would overflow int (or NaN) are clamped to a range where the
// clamped grid makes the interpolant constant anyway; identical to the
// reference for every finite in-range input.
inline double safe_coord(double v, int n) {
    double lo = -4.0, hi = (double)n + 4.0;
    if (!(v > lo)) return lo;  // also catches NaN
    if (v > hi) return hi;
    return v;
}

inline void bicubic(const Grid& g, double r, double c, double* f, double* dfdr, double* dfdc) {
    r = safe_coord(r, g.H);
    c = safe_coord(c, g.W);
    const int row = (int)std::floor(r);
    const int col = (int)std::floor(c);
    double fr[4], dfc[4];
    for (int k = 0; k < 4; ++k) {
        int rr = row - 1 + k;
        cubic_hermite(g.at(rr, col - 1), g.at(rr, col), g.at(rr, col + 1), g.at(rr, col + 2), c - col, &fr[k], &dfc[k]);
    }
    cubic_hermite(fr[0], fr[1], fr[2], fr[3], r - row, f, dfdr);
    if (dfdc) cubic_hermite(dfc[0], dfc[1], dfc[2], dfc[3], r - row, dfdc, nullptr);
}
inline void bicubic_eval(const Grid& g, double r, double c, double* f) { bicubic(g, r, c, f, nullptr, nullptr); }
template <int N>
inline void bicubic_eval(const Grid& g, const Jet<N>& r, const Jet<N>& c, Jet<N>* f) {
    double v, dr, dc;
    bicubic(g, r.a, c.a, &v, &dr, &dc);
    f->a = v;
    for (int i = 0; i < N; ++i) f->v[i] = dr * r.v[i] + dc * c.v[i];
}

// ---------------------------------------------------------------------------
// Tracking problem
// ---------------------------------------------------------------------------
struct Problem {
    int N, H, W, B;
    const double *grad, *norm_coord, *idp, *weights, *frame;  // grad/norm_coord interleaved xy
    double fx, fy, cx, cy;
    std::vector<double> kp;  // 3N, PhotometricError.hpp:94-105
    void block_range(int b, int* start, int* n) const {
        // src/tracking/Tracker.cpp:178-190
        int ne = N / B;
        *start = b * ne;
        *n = ne + ((b + 1 == B) ? (N - (b + 1) * ne) : 0);
    }
};

constexpr double kEps = 1e-05;  // PhotometricError.hpp:200

void make_kp(Problem& P) {
    P.kp.resize((size_t)3 * P.N);
    for (int i = 0; i < P.N; ++i) {
        double z = 1.0 / (P.idp[i] + kEps);
        P.kp[3 * i + 2] = z;
        P.kp[3 * i + 0] = P.norm_coord[2 * i] * z;
        P.kp[3 * i + 1] = P.norm_coord[2 * i + 1] * z;
    }
}

// Eigen::Quaternion::toRotationMatrix (no normalisation), PhotometricError.hpp:163.
template <typename T>
inline void quat_to_rot(const T* q /*x,y,z,w*/, T R[9]) {
    const T two = Lift<T>::c(2.0), one = Lift<T>::c(1.0);
    const T tx = two * q[0], ty = two * q[1], tz = two * q[2];
    const T twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
    const T txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
    const T tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    R[0] = one - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
    R[3] = txy + twz;         R[4] = one - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = one - (txx + tyy);
}

// PhotometricError.hpp:114-122
template <typename T>
inline void compute_flow(const T& xp, const T& yp, const T* vx, const T& idp, T* result) {
    const T one = Lift<T>::c(1.0);
    result[0] = (-idp * vx[0]) + (xp * idp * vx[2]) + (xp * yp * vx[3]) - (one + xp * xp) * vx[4] + (yp * vx[5]);
    result[1] = (-idp * vx[1]) + (yp * idp * vx[2]) + (one + yp * yp) * vx[3] - (xp * yp * vx[4]) - (xp * vx[5]);
}

// PhotometricError.hpp:124-182: one residual block (points [start, start+n)).
template <typename T>
void functor(const Problem& P, int start, int n, const T* px, const T* qx, const T* vx, T* residual) {
    T model_norm_sq = Lift<T>::c(1e-03);
    for (int i = 0; i < n; ++i) {
        int idx = start + i;
        T flow[2];
        compute_flow<T>(Lift<T>::c(P.norm_coord[2 * idx]), Lift<T>::c(P.norm_coord[2 * idx + 1]), vx, Lift<T>::c(P.idp[idx]), flow);
        residual[i] = -(Lift<T>::c(P.grad[2 * idx]) * flow[0] + Lift<T>::c(P.grad[2 * idx + 1]) * flow[1]);
        model_norm_sq += residual[i] * residual[i];
    }
    T model_norm = jsqrt(model_norm_sq);
    Grid g{P.frame, P.H, P.W};
    for (int i = 0; i < n; ++i) {
        int idx = start + i;
        T R[9];
        quat_to_rot<T>(qx, R);  // rebuilt per point, as the reference does (:163)
        T pt[3] = {Lift<T>::c(P.kp[3 * idx]), Lift<T>::c(P.kp[3 * idx + 1]), Lift<T>::c(P.kp[3 * idx + 2])};
        T p[3];
        for (int r = 0; r < 3; ++r) p[r] = R[3 * r] * pt[0] + R[3 * r + 1] * pt[1] + R[3 * r + 2] * pt[2] + px[r];
        T xp = Lift<T>::c(P.fx) * (p[0] / p[2]) + Lift<T>::c(P.cx);
        T yp = Lift<T>::c(P.fy) * (p[1] / p[2]) + Lift<T>::c(P.cy);
        T e;
        bicubic_eval(g, yp, xp, &e);
        residual[i] = Lift<T>::c(P.weights[idx]) * ((residual[i] / model_norm) - e);
    }
}

// Local parameterisations -------------------------------------------------
// ceres::EigenQuaternionParameterization::ComputeJacobian (4x3, xyzw storage),
// used at src/tracking/Tracker.cpp:111-112,197.
inline void quat_plus_jacobian(const double* x, double J[12]) {
    J[0] = x[3];  J[1] = x[2];   J[2] = -x[1];
    J[3] = -x[2]; J[4] = x[3];   J[5] = x[0];
    J[6] = x[1];  J[7] = -x[0];  J[8] = x[3];
    J[9] = -x[0]; J[10] = -x[1]; J[11] = -x[2];
}
// ceres::EigenQuaternionParameterization::Plus
inline void quat_plus(const double* x, const double* d, double* out) {
    const double nd = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (nd > 0.0) {
        const double s = std::sin(nd) / nd;
        const double dq[4] = {s * d[0], s * d[1], s * d[2], std::cos(nd)};  // x,y,z,w
        // out = dq * x (Hamilton product), xyzw storage
        const double ax = dq[0], ay = dq[1], az = dq[2], aw = dq[3];
        const double bx = x[0], by = x[1], bz = x[2], bw = x[3];
        out[3] = aw * bw - ax * bx - ay * by - az * bz;
        out[0] = aw * bx + ax * bw + ay * bz - az * by;
        out[1] = aw * by + ay * bw + az * bx - ax * bz;
        out[2] = aw * bz + az * bw + ax * by - ay * bx;
    } else {
        for (int i = 0; i < 4; ++i) out[i] = x[i];
    }
}
// UnitNormVectorAddition, PhotometricError.hpp:32-54
inline void unit_plus(const double* x, const double* d, double* out) {
    double sum = 0;
    for (int i = 0; i < 6; ++i) { double s = x[i] + d[i]; sum += s * s; out[i] = s; }
    sum = 1.0 / std::sqrt(sum);
    for (int i = 0; i < 6; ++i) out[i] *= sum;
}
// AutoDiffLocalParameterization<UnitNormVectorAddition,6,6>::ComputeJacobian:
// d Plus(x,delta) / d delta at delta = 0 (row-major 6x6).
inline void unit_plus_jacobian(const double* x, double J[36]) {
    double s = 0;
    for (int i = 0; i < 6; ++i) s += x[i] * x[i];
    double n = std::sqrt(s), n3 = n * s;
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) J[6 * i + j] = (i == j ? 1.0 / n : 0.0) - x[i] * x[j] / n3;
}
// Program-level Plus over [p(3) q(4) v(6)] <- delta(12)
inline void state_plus(const double* x, const double* d, double* out) {
    for (int i = 0; i < 3; ++i) out[i] = x[i] + d[i];
    quat_plus(x + 3, d + 3, out + 3);
    unit_plus(x + 7, d + 6, out + 7);
}

// Loss functions (ceres::HuberLoss / ceres::CauchyLoss), Tracker.cpp:146-161.
inline void loss_eval(int type, double a, double s, double rho[3]) {
    if (type == 1) {
        double b = a * a;
        if (s > b) {
            const double r = std::sqrt(s);
            rho[0] = 2.0 * a * r - b;
            rho[1] = std::max(std::numeric_limits<double>::min(), a / r);
            rho[2] = -rho[1] / (2.0 * s);
        } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
    } else if (type == 2) {
        double b = a * a, c = 1.0 / b;
        const double sum = 1.0 + s * c, inv = 1.0 / sum;
        rho[0] = b * std::log(sum);
        rho[1] = std::max(std::numeric_limits<double>::min(), inv);
        rho[2] = -c * (inv * inv);
    } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
}

// Per-block evaluation result in the 12-dim tangent space.
struct BlockEval {
    double cost;     // 0.5*rho(s)
    double sqnorm;   // s = ||r_b||^2 (uncorrected)
    double H[144];   // rho' * J^T J   (corrected, local)
    double g[12];    // rho' * J^T r
};

// Raw (uncorrected) residual and 13-column global Jacobian of one block.
// mode 0: analytic, mode 1: width-13 dual numbers through the
// functor (what ceres::AutoDiffCostFunction does, PhotometricError.hpp:197).
void block_raw(const Problem& P, int b, const double* x, int mode, bool want_jac, double* r, double* Jg /*n x 13 row-major*/) {
    int start, n;
    P.block_range(b, &start, &n);
    if (!want_jac) {
        functor<double>(P, start, n, x, x + 3, x + 7, r);
        return;
    }
    if (mode == 1) {
        typedef Jet<13> J13;
        J13 px[3], qx[4], vx[6];
        for (int i = 0; i < 3; ++i) px[i] = J13(x[i], i);
        for (int i = 0; i < 4; ++i) qx[i] = J13(x[3 + i], 3 + i);
        for (int i = 0; i < 6; ++i) vx[i] = J13(x[7 + i], 7 + i);
        std::vector<J13> res(n);
        functor<J13>(P, start, n, px, qx, vx, res.data());
        for (int i = 0; i < n; ++i) {
            r[i] = res[i].a;
            for (int k = 0; k < 13; ++k) Jg[(size_t)13 * i + k] = res[i].v[k];
        }
        return;
    }
    // analytic
    const double* t = x; const double* q = x + 3; const double* v = x + 7;
    std::vector<double> m(n), gm((size_t)6 * n);
    double S = 1e-03, c[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
        int idx = start + i;
        double X = P.norm_coord[2 * idx], Y = P.norm_coord[2 * idx + 1], d = P.idp[idx];
        double Gx = P.grad[2 * idx], Gy = P.grad[2 * idx + 1];
        double fxd[6] = {-d, 0, X * d, X * Y, -(1 + X * X), Y};
        double fyd[6] = {0, -d, Y * d, 1 + Y * Y, -X * Y, -X};
        double mi = 0;
        for (int k = 0; k < 6; ++k) { gm[6 * i + k] = -(Gx * fxd[k] + Gy * fyd[k]); mi += gm[6 * i + k] * v[k]; }
        m[i] = mi;
        S += mi * mi;
    }
    for (int i = 0; i < n; ++i) for (int k = 0; k < 6; ++k) c[k] += m[i] * gm[6 * i + k];
    double M = std::sqrt(S), M3 = M * S;
    double R[9];
    quat_to_rot<double>(q, R);
    Grid g{P.frame, P.H, P.W};
    // dR/dq (global, 4 columns) for Eigen's formula
    for (int i = 0; i < n; ++i) {
        int idx = start + i;
        const double* kp = &P.kp[3 * idx];
        double a[3], p[3];
        for (int rr = 0; rr < 3; ++rr) { a[rr] = R[3 * rr] * kp[0] + R[3 * rr + 1] * kp[1] + R[3 * rr + 2] * kp[2]; p[rr] = a[rr] + t[rr]; }
        double iz = 1.0 / p[2];
        double u = P.fx * (p[0] / p[2]) + P.cx, vv = P.fy * (p[1] / p[2]) + P.cy;
        double e, er, ec;
        bicubic(g, vv, u, &e, &er, &ec);
        double w = P.weights[idx];
        r[i] = w * (m[i] / M - e);
        double dP[3] = {-w * ec * P.fx * iz, -w * er * P.fy * iz, w * (ec * P.fx * p[0] + er * P.fy * p[1]) * iz * iz};
        double* J = &Jg[(size_t)13 * i];
        J[0] = dP[0]; J[1] = dP[1]; J[2] = dP[2];
        // d(R kp)/dq for R(q) = Eigen formula: derivative wrt x,y,z,w
        const double qx_ = q[0], qy = q[1], qz = q[2], qw = q[3];
        const double kx = kp[0], ky = kp[1], kz = kp[2];
        double dRk[3][4];
        // row 0: (1-2yy-2zz) kx + (2xy-2wz) ky + (2xz+2wy) kz
        dRk[0][0] = 2 * qy * ky + 2 * qz * kz;
        dRk[0][1] = -4 * qy * kx + 2 * qx_ * ky + 2 * qw * kz;
        dRk[0][2] = -4 * qz * kx - 2 * qw * ky + 2 * qx_ * kz;
        dRk[0][3] = -2 * qz * ky + 2 * qy * kz;
        // row 1: (2xy+2wz) kx + (1-2xx-2zz) ky + (2yz-2wx) kz
        dRk[1][0] = 2 * qy * kx - 4 * qx_ * ky - 2 * qw * kz;
        dRk[1][1] = 2 * qx_ * kx + 2 * qz * kz;
        dRk[1][2] = 2 * qw * kx - 4 * qz * ky + 2 * qy * kz;
        dRk[1][3] = 2 * qz * kx - 2 * qx_ * kz;
        // row 2: (2xz-2wy) kx + (2yz+2wx) ky + (1-2xx-2yy) kz
        dRk[2][0] = 2 * qz * kx + 2 * qw * ky - 4 * qx_ * kz;
        dRk[2][1] = -2 * qw * kx + 2 * qz * ky - 4 * qy * kz;
        dRk[2][2] = 2 * qx_ * kx + 2 * qy * ky;
        dRk[2][3] = -2 * qy * kx + 2 * qx_ * ky;
        for (int k = 0; k < 4; ++k) J[3 + k] = dP[0] * dRk[0][k] + dP[1] * dRk[1][k] + dP[2] * dRk[2][k];
        for (int k = 0; k < 6; ++k) J[7 + k] = w * (gm[6 * i + k] / M - m[i] * c[k] / M3);
    }
}

// One block: raw -> local parameterisation -> loss corrector -> normal equations.
// Follows ceres ResidualBlock::Evaluate: local Jacobians first, then Corrector
// (for Huber/Cauchy rho'' <= 0 so residual and Jacobian are both scaled by sqrt(rho')).
void block_eval(const Problem& P, int b, const double* x, int mode, int loss_type, double loss_a, bool want_jac,
                BlockEval* out, double* r_raw_out, double* Jlocal_out /*n x 12 or null*/) {
    int start, n;
    P.block_range(b, &start, &n);
    std::vector<double> r(n), Jg(want_jac ? (size_t)13 * n : 0);
    block_raw(P, b, x, mode, want_jac, r.data(), Jg.data());
    double s = 0;
    for (int i = 0; i < n; ++i) s += r[i] * r[i];
    double rho[3];
    loss_eval(loss_type, loss_a, s, rho);
    out->cost = 0.5 * rho[0];
    out->sqnorm = s;
    if (r_raw_out) std::memcpy(r_raw_out, r.data(), sizeof(double) * n);
    if (!want_jac) return;
    double Jq[12], Jv[36];
    quat_plus_jacobian(x + 3, Jq);
    unit_plus_jacobian(x + 7, Jv);
    std::fill(out->H, out->H + 144, 0.0);
    std::fill(out->g, out->g + 12, 0.0);
    double Jl[12];
    for (int i = 0; i < n; ++i) {
        const double* J = &Jg[(size_t)13 * i];
        for (int k = 0; k < 3; ++k) Jl[k] = J[k];
        for (int k = 0; k < 3; ++k) Jl[3 + k] = J[3] * Jq[k] + J[4] * Jq[3 + k] + J[5] * Jq[6 + k] + J[6] * Jq[9 + k];
        for (int k = 0; k < 6; ++k) { double acc = 0; for (int j = 0; j < 6; ++j) acc += J[7 + j] * Jv[6 * j + k]; Jl[6 + k] = acc; }
        if (Jlocal_out) std::memcpy(Jlocal_out + (size_t)12 * i, Jl, sizeof(Jl));
        for (int a = 0; a < 12; ++a) {
            out->g[a] += Jl[a] * r[i];
            for (int c2 = a; c2 < 12; ++c2) out->H[12 * a + c2] += Jl[a] * Jl[c2];
        }
    }
    for (int a = 0; a < 12; ++a) {
        out->g[a] *= rho[1];
        for (int c2 = a; c2 < 12; ++c2) { out->H[12 * a + c2] *= rho[1]; out->H[12 * c2 + a] = out->H[12 * a + c2]; }
    }
}

struct Eval { double cost; double H[144]; double g[12]; };

// Persistent workers for one solve: ceres keeps a thread pool alive for the whole ceres::Solve
// (options.num_threads, Tracker.cpp:138), so the timed CPU baseline must not pay a thread creation
// per evaluation either.  Workers spin on a generation counter between evaluations.
class WorkerPool {
  public:
    explicit WorkerPool(int n) : n_(std::max(1, n)) {
        for (int t = 1; t < n_; ++t) threads_.emplace_back([this, t]() { loop(t); });
    }
    ~WorkerPool() {
        stop_.store(true);
        gen_.fetch_add(1);
        for (auto& t : threads_) t.join();
    }
    int size() const { return n_; }
    template <typename F>
    void run(const F& f) {  // f(worker_index), worker 0 is the caller
        fn_ = [&f](int t) { f(t); };
        pending_.store(n_ - 1);
        gen_.fetch_add(1);
        f(0);
        while (pending_.load(std::memory_order_acquire) > 0) std::this_thread::yield();
    }

  private:
    void loop(int t) {
        unsigned seen = 0;
        for (;;) {
            while (gen_.load(std::memory_order_acquire) == seen) std::this_thread::yield();
            seen = gen_.load(std::memory_order_acquire);
            if (stop_.load()) return;
            fn_(t);
            pending_.fetch_sub(1, std::memory_order_release);
        }
    }
    int n_;
    std::vector<std::thread> threads_;
    std::function<void(int)> fn_;
    std::atomic<unsigned> gen_{0};
    std::atomic<int> pending_{0};
    std::atomic<bool> stop_{false};
};

void evaluate_all(const Problem& P, const double* x, int mode, int loss_type, double loss_a, bool want_jac, WorkerPool* pool, Eval* out) {
    std::vector<BlockEval> be(P.B);
    auto work = [&](int b) { block_eval(P, b, x, mode, loss_type, loss_a, want_jac, &be[b], nullptr, nullptr); };
    if (pool && pool->size() > 1) {
        // one worker per residual block, the way ceres parallelises with options.num_threads == number of
        // blocks (Tracker.cpp:138,178-195)
        const int nt = pool->size();
        pool->run([&](int t) { for (int b = t; b < P.B; b += nt) work(b); });
    } else {
        for (int b = 0; b < P.B; ++b) work(b);
    }
    out->cost = 0;
    if (want_jac) { std::fill(out->H, out->H + 144, 0.0); std::fill(out->g, out->g + 12, 0.0); }
    for (int b = 0; b < P.B; ++b) {
        out->cost += be[b].cost;
        if (want_jac) {
            for (int i = 0; i < 144; ++i) out->H[i] += be[b].H[i];
            for (int i = 0; i < 12; ++i) out->g[i] += be[b].g[i];
        }
    }
}

// Dense Cholesky solve of A y = rhs (12x12, SPD).  false if a pivot is <= 0.
bool chol_solve12(const double* A, const double* rhs, double* y) {
    double L[144];
    std::memset(L, 0, sizeof(L));
    for (int j = 0; j < 12; ++j) {
        double d = A[12 * j + j];
        for (int k = 0; k < j; ++k) d -= L[12 * j + k] * L[12 * j + k];
        if (!(d > 0.0) || !std::isfinite(d)) return false;
        L[12 * j + j] = std::sqrt(d);
        for (int i = j + 1; i < 12; ++i) {
            double s = A[12 * i + j];
            for (int k = 0; k < j; ++k) s -= L[12 * i + k] * L[12 * j + k];
            L[12 * i + j] = s / L[12 * j + j];
        }
    }
    double z[12];
    for (int i = 0; i < 12; ++i) { double s = rhs[i]; for (int k = 0; k < i; ++k) s -= L[12 * i + k] * z[k]; z[i] = s / L[12 * i + i]; }
    for (int i = 11; i >= 0; --i) { double s = z[i]; for (int k = i + 1; k < 12; ++k) s -= L[12 * k + i] * y[k]; y[i] = s / L[12 * i + i]; }
    for (int i = 0; i < 12; ++i) if (!std::isfinite(y[i])) return false;
    return true;
}

}  // namespace

// ===========================================================================
// C interface (ctypes)
// ===========================================================================
extern "C" {

struct eds_oracle_solver_config {
    int num_blocks;          // config.options.num_threads, Tracker.cpp:138,178
    int loss_type;           // 0 none, 1 Huber, 2 Cauchy (Tracker.cpp:146-161)
    double loss_param;       // config.loss_params[0]
    int max_iterations;      // config.options.max_num_iterations[id], Tracker.cpp:139
    double function_tolerance;   // Tracker.cpp:140
    double gradient_tolerance;   // 1e-8, Tracker.cpp:142
    double parameter_tolerance;  // 1e-6, Tracker.cpp:143
    int jacobian_mode;       // 0 analytic, 1 width-13 dual numbers (reference cost structure)
    int threads;             // worker threads (<= num_blocks)
};

struct eds_oracle_solver_info {
    int iterations;          // successful + unsuccessful steps (Tracker.cpp:211)
    int successful_steps, unsuccessful_steps;
    double initial_cost, final_cost;
    int usable;              // summary.IsSolutionUsable(), Tracker.cpp:213
    int termination;         // 0 CONVERGENCE, 1 NO_CONVERGENCE, 2 FAILURE
    double solve_time_us;    // wall clock of the solve only (Tracker.cpp:201-209)
    double final_radius;
};

// src/utils/Utils.cpp:50-122 on explicit double coordinates.
int eds_oracle_draw_values(const double* px, const double* py, const int8_t* val, int E, int H, int W,
                           int method, int use_exp, double sigma, double* img) {
    if (E < 0 || H <= 0 || W <= 0) return 1;
    draw_values_points(px, py, val, E, H, W, method, use_exp != 0, sigma, img);
    return 0;
}

// src/tracking/EventFrame.cpp:302-389 (level 0 only, out_scale == 1).
// mapx/mapy: forward undistortion LUT (CV_32F HxW, EventFrame.cpp:72-81); NULL = identity.
// returns 4 on first_ts > last_ts (the reference throws, EventFrame.cpp:325-329).
int eds_oracle_event_frame(const uint16_t* x, const uint16_t* y, const uint8_t* pol, const int64_t* ts_us, int E,
                           int H, int W, const float* mapx, const float* mapy, int method, int use_exp, double sigma,
                           double* img_out, double* frame_out, double* norm_out, int64_t* time_out, int64_t* delta_out) {
    if (E <= 0 || H <= 0 || W <= 0) return 1;
    std::vector<double> ux(E), uy(E);
    std::vector<int8_t> p(E);
    for (int i = 0; i < E; ++i) {
        if (x[i] >= W || y[i] >= H) return 1;
        if (mapx && mapy) {
            ux[i] = (double)mapx[(size_t)y[i] * W + x[i]];
            uy[i] = (double)mapy[(size_t)y[i] * W + x[i]];
        } else { ux[i] = x[i]; uy[i] = y[i]; }
        p[i] = pol[i] ? 1 : -1;  // EventFrame.cpp:318
    }
    if (ts_us) {
        int64_t first = ts_us[0], last = ts_us[E - 1];
        if (E == 1) last = first;
        if (first > last) return 4;
        if (time_out) *time_out = ts_us[E / 2];  // EventFrame.cpp:332-333
        if (delta_out) *delta_out = last - first;
    }
    std::vector<double> img((size_t)H * W);
    draw_values_points(ux.data(), uy.data(), p.data(), E, H, W, method, use_exp != 0, sigma, img.data());
    // cv::norm (L2), EventFrame.cpp:360-364; normalise :367-383
    double ss = 0;
    for (size_t i = 0; i < img.size(); ++i) ss += img[i] * img[i];
    double nrm = std::sqrt(ss);
    if (norm_out) *norm_out = nrm;
    if (img_out) std::memcpy(img_out, img.data(), sizeof(double) * img.size());
    if (frame_out) for (size_t i = 0; i < img.size(); ++i) frame_out[i] = img[i] / nrm;
    return 0;
}

// Bicubic sample (value + partials) of a row-major HxW grid at (row, col).
void eds_oracle_bicubic(const double* grid, int H, int W, int n, const double* rows, const double* cols, double* f, double* dfdr, double* dfdc) {
    Grid g{grid, H, W};
    for (int i = 0; i < n; ++i) bicubic(g, rows[i], cols[i], &f[i], &dfdr[i], &dfdc[i]);
}

static Problem make_problem(int N, const double* grad_xy, const double* norm_xy, const double* idp, const double* weights,
                            const double* frame, int H, int W, double fx, double fy, double cx, double cy, int B) {
    Problem P;
    P.N = N; P.H = H; P.W = W; P.B = B;
    P.grad = grad_xy; P.norm_coord = norm_xy; P.idp = idp; P.weights = weights; P.frame = frame;
    P.fx = fx; P.fy = fy; P.cx = cx; P.cy = cy;
    make_kp(P);
    return P;
}

// Residuals (no loss, what Tracker.cpp:223-230 writes into kf->residuals), the
// 12-column tangent-space Jacobian (uncorrected) and the robustified cost
// 0.5*sum_b rho(||r_b||^2) at state x = [p(3) q(xyzw) v(6)].
int eds_oracle_tracker_evaluate(int N, const double* grad_xy, const double* norm_xy, const double* idp, const double* weights,
                                const double* frame, int H, int W, double fx, double fy, double cx, double cy,
                                int num_blocks, int loss_type, double loss_param, int jacobian_mode, const double* x,
                                double* residuals_out, double* jac_out /*N x 12 or NULL*/, double* cost_out,
                                double* H_out /*144 or NULL*/, double* g_out /*12 or NULL*/, double* block_sqnorm_out /*B or NULL*/) {
    if (N <= 0 || num_blocks <= 0 || N < num_blocks) return 1;
    Problem P = make_problem(N, grad_xy, norm_xy, idp, weights, frame, H, W, fx, fy, cx, cy, num_blocks);
    double cost = 0;
    double Hs[144], gs[12];
    std::fill(Hs, Hs + 144, 0.0); std::fill(gs, gs + 12, 0.0);
    for (int b = 0; b < P.B; ++b) {
        int start, n;
        P.block_range(b, &start, &n);
        BlockEval be;
        block_eval(P, b, x, jacobian_mode, loss_type, loss_param, true, &be, residuals_out ? residuals_out + start : nullptr,
                   jac_out ? jac_out + (size_t)12 * start : nullptr);
        cost += be.cost;
        for (int i = 0; i < 144; ++i) Hs[i] += be.H[i];
        for (int i = 0; i < 12; ++i) gs[i] += be.g[i];
        if (block_sqnorm_out) block_sqnorm_out[b] = be.sqnorm;
    }
    if (cost_out) *cost_out = cost;
    if (H_out) std::memcpy(H_out, Hs, sizeof(Hs));
    if (g_out) std::memcpy(g_out, gs, sizeof(gs));
    return 0;
}

// MAD loss parameter, src/tracking/Tracker.cpp:292-305 + Utils.hpp:315-320
// (nth_element at index size/2 = upper median).  Reorders r in place like the reference.
double eds_oracle_mad_tau(double* r, int N) {
    std::nth_element(r, r + N / 2, r + N);
    double median = r[N / 2];
    std::vector<double> abs_med(N);
    for (int i = 0; i < N; ++i) abs_med[i] = std::abs(r[i] - median);
    std::nth_element(abs_med.begin(), abs_med.begin() + N / 2, abs_med.end());
    double mad = 1.4826 * abs_med[N / 2];
    return 1.345 * mad;
}

// Tracker::optimize (src/tracking/Tracker.cpp:104-241) with ceres::Solve restated:
// TrustRegionMinimizer + LevenbergMarquardtStrategy, jacobi_scaling on,
// monotonic steps, min_relative_decrease 1e-3, initial radius 1e4, max 1e16,
// min_lm_diagonal 1e-6, max_lm_diagonal 1e32, max 5 consecutive invalid steps.
int eds_oracle_tracker_solve(int N, const double* grad_xy, const double* norm_xy, const double* idp, const double* weights,
                             const double* frame, int H, int W, double fx, double fy, double cx, double cy,
                             const eds_oracle_solver_config* cfg, double* x /*13, in-out*/, double* residuals_out /*N or NULL*/,
                             double* next_loss_param_out, eds_oracle_solver_info* info, double* trace /*NULL or (max_it+1)*16*/) {
    if (N <= 0 || cfg->num_blocks <= 0 || N < cfg->num_blocks) return 1;
    Problem P = make_problem(N, grad_xy, norm_xy, idp, weights, frame, H, W, fx, fy, cx, cy, cfg->num_blocks);
    const int mode = cfg->jacobian_mode, lt = cfg->loss_type, th = cfg->threads;
    const double la = cfg->loss_param;
    WorkerPool pool(std::min(std::max(1, th), P.B));
    auto t0 = std::chrono::high_resolution_clock::now();

    double xc[13];
    std::memcpy(xc, x, sizeof(xc));
    Eval ev;
    evaluate_all(P, xc, mode, lt, la, true, &pool, &ev);
    double x_cost = ev.cost;
    double scale[12];
    for (int i = 0; i < 12; ++i) scale[i] = 1.0 / (1.0 + std::sqrt(ev.H[12 * i + i]));  // jacobi_scaling, iteration 0 only
    double Hs[144], gs[12];
    auto apply_scale = [&]() {
        for (int i = 0; i < 12; ++i) { gs[i] = ev.g[i] * scale[i]; for (int j = 0; j < 12; ++j) Hs[12 * i + j] = ev.H[12 * i + j] * scale[i] * scale[j]; }
    };
    apply_scale();
    auto grad_max_norm = [&](const double* xx, const double* g) {
        double ng[12], xp[13];
        for (int i = 0; i < 12; ++i) ng[i] = -g[i];
        state_plus(xx, ng, xp);
        double m = 0;
        for (int i = 0; i < 13; ++i) m = std::max(m, std::abs(xx[i] - xp[i]));
        return m;
    };
    auto norm13 = [](const double* a) { double s = 0; for (int i = 0; i < 13; ++i) s += a[i] * a[i]; return std::sqrt(s); };

    double radius = 1e4, decrease_factor = 2.0;
    bool reuse_diagonal = false;
    double diag[12];
    double gmax = grad_max_norm(xc, ev.g);
    double x_norm = norm13(xc);
    int iteration = 0, n_succ = 0, n_unsucc = 0, consecutive_invalid = 0;
    int termination = 1;  // NO_CONVERGENCE
    info->initial_cost = x_cost;
    if (trace) { trace[0] = x_cost; trace[1] = radius; trace[2] = gmax; trace[3] = 1; }
    bool is_valid_solution = std::isfinite(x_cost);
    if (!is_valid_solution) termination = 2;

    while (is_valid_solution) {
        if (iteration >= cfg->max_iterations) { termination = 1; break; }
        if (gmax <= cfg->gradient_tolerance) { termination = 0; break; }
        if (radius < 1e-32) { termination = 0; break; }
        ++iteration;
        // LevenbergMarquardtStrategy::ComputeStep
        if (!reuse_diagonal)
            for (int i = 0; i < 12; ++i) diag[i] = std::min(std::max(Hs[12 * i + i], 1e-6), 1e32);
        double A[144], y[12], step[12];
        std::memcpy(A, Hs, sizeof(A));
        for (int i = 0; i < 12; ++i) A[12 * i + i] += diag[i] / radius;
        bool solved = chol_solve12(A, gs, y);
        reuse_diagonal = true;
        bool step_valid = false;
        double model_cost_change = 0;
        if (solved) {
            for (int i = 0; i < 12; ++i) step[i] = -y[i];
            double sg = 0, sHs = 0;
            for (int i = 0; i < 12; ++i) { sg += step[i] * gs[i]; double hv = 0; for (int j = 0; j < 12; ++j) hv += Hs[12 * i + j] * step[j]; sHs += step[i] * hv; }
            model_cost_change = -sg - 0.5 * sHs;
            step_valid = model_cost_change > 0.0;
        }
        if (!step_valid) {
            // HandleInvalidStep
            if (++consecutive_invalid >= 5) { termination = 2; break; }
            radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
            ++n_unsucc;
            if (trace) { double* t = trace + 16 * iteration; t[0] = x_cost; t[1] = radius; t[2] = gmax; t[3] = -1; }
            continue;
        }
        consecutive_invalid = 0;
        double delta[12], cand[13];
        for (int i = 0; i < 12; ++i) delta[i] = step[i] * scale[i];
        state_plus(xc, delta, cand);
        Eval evc;
        evaluate_all(P, cand, mode, lt, la, false, &pool, &evc);
        double cand_cost = std::isfinite(evc.cost) ? evc.cost : std::numeric_limits<double>::max();
        // ParameterToleranceReached
        double sn = 0;
        for (int i = 0; i < 13; ++i) sn += (xc[i] - cand[i]) * (xc[i] - cand[i]);
        sn = std::sqrt(sn);
        if (sn <= cfg->parameter_tolerance * (x_norm + cfg->parameter_tolerance)) { termination = 0; break; }
        // FunctionToleranceReached
        double cost_change = x_cost - cand_cost;
        if (std::abs(cost_change) <= cfg->function_tolerance * x_cost) { termination = 0; break; }
        double rel_decrease = cost_change / model_cost_change;
        if (rel_decrease > 1e-3) {
            // HandleSuccessfulStep
            std::memcpy(xc, cand, sizeof(xc));
            x_norm = norm13(xc);
            evaluate_all(P, xc, mode, lt, la, true, &pool, &ev);
            x_cost = ev.cost;
            apply_scale();
            gmax = grad_max_norm(xc, ev.g);
            radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rel_decrease - 1.0, 3));
            radius = std::min(1e16, radius);
            decrease_factor = 2.0;
            reuse_diagonal = false;
            ++n_succ;
            if (trace) { double* t = trace + 16 * iteration; t[0] = x_cost; t[1] = radius; t[2] = gmax; t[3] = 1; t[4] = rel_decrease; }
        } else {
            radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
            ++n_unsucc;
            if (trace) { double* t = trace + 16 * iteration; t[0] = cand_cost; t[1] = radius; t[2] = gmax; t[3] = 0; t[4] = rel_decrease; }
        }
    }
    auto t1 = std::chrono::high_resolution_clock::now();
    info->solve_time_us = std::chrono::duration<double, std::micro>(t1 - t0).count();
    info->iterations = n_succ + n_unsucc;
    info->successful_steps = n_succ;
    info->unsuccessful_steps = n_unsucc;
    info->final_cost = x_cost;
    info->termination = termination;
    info->usable = (termination == 0 || termination == 1) ? 1 : 0;
    info->final_radius = radius;
    if (!info->usable) return 3;
    std::memcpy(x, xc, sizeof(xc));
    // residual write-back (no loss), Tracker.cpp:223-230, then MAD, :233
    std::vector<double> res(N);
    for (int b = 0; b < P.B; ++b) {
        int start, n;
        P.block_range(b, &start, &n);
        block_raw(P, b, xc, 0, false, res.data() + start, nullptr);
    }
    if (residuals_out) std::memcpy(residuals_out, res.data(), sizeof(double) * N);
    if (next_loss_param_out) *next_loss_param_out = eds_oracle_mad_tau(res.data(), N);
    return 0;
}

}  // extern "C"
