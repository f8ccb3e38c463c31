// CPU oracle for the per-point depth filter of uzh-rpg/slam-eds.
//
// TEST INFRASTRUCTURE ONLY: nothing under slam-eds_b200/ may include, link or call this file.
// PARITY UNPINNED: the reference ships no tests or golden vectors for this path and cannot be built
// here; this is a restatement of eds::mapping::DepthPoints::update (src/mapping/DepthPoints.cpp:93-228,
// :376-401; DepthPoints.hpp:151-191) and eds::utils::normPdf (src/utils/Utils.hpp:337-345) in double.
// cv::Mat::inv(DECOMP_SVD) of the invertible 3x3 blocks is restated as the exact inverse.
#include <algorithm>
#include <cmath>
#include <cstdint>

namespace {

void inv3(const double* M, double* I) {  // row-major
    const double a = M[0], b = M[1], c = M[2], d = M[3], e = M[4], f = M[5], g = M[6], h = M[7], i = M[8];
    const double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
    const double det = a * A + b * B + c * C;
    I[0] = A / det; I[1] = -(b * i - c * h) / det; I[2] = (b * f - c * e) / det;
    I[3] = B / det; I[4] = (a * i - c * g) / det;  I[5] = -(a * f - c * d) / det;
    I[6] = C / det; I[7] = -(a * h - b * g) / det; I[8] = (a * e - b * d) / det;
}
void mul3v(const double* M, const double* v, double* o) {
    for (int r = 0; r < 3; ++r) o[r] = M[3 * r] * v[0] + M[3 * r + 1] * v[1] + M[3 * r + 2] * v[2];
}
void cross(const double* a, const double* b, double* o) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
double dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

}  // namespace

extern "C" {

// DepthPoints::update for N points.  T_kf_ef: 4x4 row-major (the tracker's transform); kf_coord, ef_coord: N x 2 pixel
// coordinates (ef_coord are offsets "tracks" when coords_are_tracks != 0, DepthPoints.cpp:141-176); state: N x 4
// {mu (inverse depth), sigma2, a, b}, updated in place; ok_out: filterVogiatzis' return value per point (or null).
void eds_oracle_depth_update(int N, double fx, double fy, double cx, double cy, double mu_range, double px_error_angle, const double* T_kf_ef,
                             const double* kf_coord, const double* ef_coord, int coords_are_tracks, double* state, uint8_t* ok_out) {
    // T_ef_kf = T_kf_ef^-1 (rigid): R^T, -R^T t
    double R[9], t[3], Rt[9], tt[3];
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R[3 * r + c] = T_kf_ef[4 * r + c]; t[r] = T_kf_ef[4 * r + 3]; }
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Rt[3 * r + c] = R[3 * c + r];
    mul3v(Rt, t, tt);
    for (int r = 0; r < 3; ++r) tt[r] = -tt[r];
    const double K[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
    // P_kf = K [I|0], P_ef = K [R_ef_kf | t_ef_kf]  (:99-106)
    double M2[9], p2c3[3], InvM1[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) M2[3 * r + c] = K[3 * r] * Rt[c] + K[3 * r + 1] * Rt[3 + c] + K[3 * r + 2] * Rt[6 + c];
    mul3v(K, tt, p2c3);
    inv3(K, InvM1);
    // optical centre of camera 1 is the origin (P_kf's last column is zero): epipole = P_ef * [0,0,0,1] = 4th column of P_ef (:385-391)
    const double* epipole = p2c3;
    const double tnorm = std::sqrt(dot(t, t));
    for (int i = 0; i < N; ++i) {
        const double x_kf[3] = {kf_coord[2 * i], kf_coord[2 * i + 1], 1.0};
        double x_ef[3] = {ef_coord[2 * i], ef_coord[2 * i + 1], 1.0};
        if (coords_are_tracks) { x_ef[0] += x_kf[0]; x_ef[1] += x_kf[1]; }
        // invDepthTwoPointsEucl (:376-401)
        double ray[3], x1p[3], aux1[3], aux2[3];
        mul3v(InvM1, x_kf, ray);
        mul3v(M2, ray, x1p);
        cross(x1p, x_ef, aux1);
        cross(x_ef, epipole, aux2);
        const double inv_depth = dot(aux1, aux2) / dot(aux2, aux2);
        const double depth = 1.0 / inv_depth;
        // computeTau (DepthPoints.hpp:157-174)
        double bearing[3] = {(x_ef[0] - cx) / fx, (x_ef[1] - cy) / fy, 1.0};
        const double bn = std::sqrt(dot(bearing, bearing));
        for (double& v : bearing) v /= bn;
        const double a3[3] = {bearing[0] * depth - t[0], bearing[1] * depth - t[1], bearing[2] * depth - t[2]};
        const double a_norm = std::sqrt(dot(a3, a3));
        const double alpha = std::acos(dot(bearing, t) / tnorm);
        const double beta = std::acos(-dot(a3, t) / (tnorm * a_norm));
        const double beta_plus = beta + px_error_angle;
        const double gamma_plus = M_PI - alpha - beta_plus;
        const double z_plus = tnorm * std::sin(beta_plus) / std::sin(gamma_plus);
        const double depth_sigma = z_plus - depth;
        // getSigma2FromDepthSigma (:176-181)
        const double sg = 0.5 * (1.0 / std::max(0.000000000001, depth - depth_sigma) - 1.0 / (depth + depth_sigma));
        const double tau2 = sg * sg;
        // filterVogiatzis (:178-228)
        double& mu = state[4 * i]; double& sigma2 = state[4 * i + 1]; double& a = state[4 * i + 2]; double& b = state[4 * i + 3];
        const double z = inv_depth;
        const double norm_scale = std::sqrt(sigma2 + tau2);
        if (std::isnan(norm_scale)) { if (ok_out) ok_out[i] = 0; continue; }
        const double oldsigma2 = sigma2;
        const double s2 = 1.0 / (1.0 / sigma2 + 1.0 / tau2);
        const double m = s2 * (mu / sigma2 + z / tau2);
        const double uniform_x = 1.0 / mu_range;
        double exponent = z - mu;  // normPdf, Utils.hpp:337-345
        exponent *= -exponent;
        exponent /= 2 * norm_scale * norm_scale;
        double pdf = std::exp(exponent);
        pdf /= norm_scale * std::sqrt(2 * M_PI);
        double C1 = a / (a + b) * pdf;
        double C2 = b / (a + b) * uniform_x;
        const double nc = C1 + C2;
        C1 /= nc;
        C2 /= nc;
        const double f = C1 * (a + 1.0) / (a + b + 1.0) + C2 * a / (a + b + 1.0);
        const double e = C1 * (a + 1.0) * (a + 2.0) / ((a + b + 1.0) * (a + b + 2.0)) + C2 * a * (a + 1.0) / ((a + b + 1.0) * (a + b + 2.0));
        const double mu_new = C1 * m + C2 * mu;
        sigma2 = C1 * (s2 + m * m) + C2 * (sigma2 + mu * mu) - mu_new * mu_new;
        mu = mu_new;
        a = (e - f) / (f - e / f);
        b = a * (1.0 - f) / f;
        bool ok = true;
        if (sigma2 < 0.0) sigma2 = oldsigma2;
        if (mu < 0.0) { mu = 1.0; ok = false; }
        if (ok_out) ok_out[i] = ok ? 1 : 0;
    }
}

}  // extern "C"
