// Per-point depth filter: eds::mapping::DepthPoints::update
// (src/mapping/DepthPoints.cpp:93-228) -- two-view triangulation of every key-frame point against its
// position in the event frame (invDepthTwoPointsEucl, :376-401), depth uncertainty from a pixel of
// angular error (computeTau, DepthPoints.hpp:157-174), and the Vogiatzis / Hernandez Gaussian x Beta
// update of {mu, sigma2, a, b} (filterVogiatzis, :178-228).  Points are independent: one thread per
// point, fp64 like the reference, the filter state stays resident on the device between windows.
#include <vector>

#include "common.cuh"
#include "depth.cuh"

namespace {

struct DepthArgs {
    int N, tracks;
    double fx, fy, cx, cy, mu_range, px_error_angle;
    double M2[9];       // K R_ef_kf
    double epipole[3];  // K t_ef_kf
    double t[3];        // translation of T_kf_ef
    double tnorm;
    const double* kf_coord;
    const double* ef_coord;
    double* state;  // [N][4]
    unsigned char* ok;
    const unsigned char* skip;  // or null: points that left the event frame (Tracker::getCoord erases them before the update,
                                // Tracker.cpp:356-372; here they keep their state and report ok = 0)
};

__device__ __forceinline__ void cross3(const double* a, const double* b, double* o) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// Eigen::Quaterniond::toRotationMatrix for (x,y,z,w)
__device__ __forceinline__ void quat_rot(const double* q, double* R) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.0 - (txx + tyy);
}

__device__ __forceinline__ void depth_update_point(const DepthArgs& d, int i);

__global__ void __launch_bounds__(128) depth_update_kernel(DepthArgs d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.N) return;
    depth_update_point(d, i);
}

// the pose comes from a tracker's state on the device: (px, qx) is T_ef_kf (Tracker.cpp:338-341), the filter
// wants T_kf_ef = T_ef_kf^-1 (Tracker.cpp:220) and inverts it again (DepthPoints.cpp:102-104)
__global__ void __launch_bounds__(128) depth_update_tracked_kernel(DepthArgs d, const double* __restrict__ tracker_state) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.N) return;
    if (d.skip && d.skip[i]) {
        if (d.ok) d.ok[i] = 0;
        return;
    }
    double R[9];
    quat_rot(tracker_state + 3, R);
    const double px[3] = {tracker_state[0], tracker_state[1], tracker_state[2]};
    const double K[9] = {d.fx, 0, d.cx, 0, d.fy, d.cy, 0, 0, 1};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) d.M2[3 * r + c] = K[3 * r] * R[c] + K[3 * r + 1] * R[3 + c] + K[3 * r + 2] * R[6 + c];
        d.epipole[r] = K[3 * r] * px[0] + K[3 * r + 1] * px[1] + K[3 * r + 2] * px[2];
        d.t[r] = -(R[r] * px[0] + R[3 + r] * px[1] + R[6 + r] * px[2]);  // -R^T px
    }
    d.tnorm = sqrt(d.t[0] * d.t[0] + d.t[1] * d.t[1] + d.t[2] * d.t[2]);
    depth_update_point(d, i);
}

__device__ __forceinline__ void depth_update_point(const DepthArgs& d, int i) {
    const double2 kf = reinterpret_cast<const double2*>(d.kf_coord)[i];
    double2 ef = reinterpret_cast<const double2*>(d.ef_coord)[i];
    if (d.tracks) { ef.x += kf.x; ef.y += kf.y; }
    // invDepthTwoPointsEucl: ray = K^-1 x_kf (K upper triangular), x1p = K R_ef_kf ray
    const double ray[3] = {(kf.x - d.cx) / d.fx, (kf.y - d.cy) / d.fy, 1.0};
    double x1p[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) x1p[r] = d.M2[3 * r] * ray[0] + d.M2[3 * r + 1] * ray[1] + d.M2[3 * r + 2] * ray[2];
    const double x_ef[3] = {ef.x, ef.y, 1.0};
    double aux1[3], aux2[3];
    cross3(x1p, x_ef, aux1);
    cross3(x_ef, d.epipole, aux2);
    const double inv_depth = dot3(aux1, aux2) / dot3(aux2, aux2);
    const double depth = 1.0 / inv_depth;
    // computeTau
    double bearing[3] = {(ef.x - d.cx) / d.fx, (ef.y - d.cy) / d.fy, 1.0};
    const double bn = sqrt(dot3(bearing, bearing));
#pragma unroll
    for (int k = 0; k < 3; ++k) bearing[k] /= bn;
    const double a3[3] = {bearing[0] * depth - d.t[0], bearing[1] * depth - d.t[1], bearing[2] * depth - d.t[2]};
    const double a_norm = sqrt(dot3(a3, a3));
    const double alpha = acos(dot3(bearing, d.t) / d.tnorm);
    const double beta = acos(-dot3(a3, d.t) / (d.tnorm * a_norm));
    const double beta_plus = beta + d.px_error_angle;
    const double gamma_plus = 3.14159265358979323846 - alpha - beta_plus;
    const double z_plus = d.tnorm * sin(beta_plus) / sin(gamma_plus);
    const double depth_sigma = z_plus - depth;
    // getSigma2FromDepthSigma
    const double sg = 0.5 * (1.0 / fmax(0.000000000001, depth - depth_sigma) - 1.0 / (depth + depth_sigma));
    const double tau2 = sg * sg;
    // filterVogiatzis
    double4 st = reinterpret_cast<double4*>(d.state)[i];
    double mu = st.x, sigma2 = st.y, a = st.z, b = st.w;
    const double z = inv_depth;
    const double norm_scale = sqrt(sigma2 + tau2);
    if (isnan(norm_scale)) {
        if (d.ok) d.ok[i] = 0;
        return;
    }
    const double oldsigma2 = sigma2;
    const double s2 = 1.0 / (1.0 / sigma2 + 1.0 / tau2);
    const double m = s2 * (mu / sigma2 + z / tau2);
    const double uniform_x = 1.0 / d.mu_range;
    double exponent = z - mu;  // eds::utils::normPdf, Utils.hpp:337-345
    exponent *= -exponent;
    exponent /= 2 * norm_scale * norm_scale;
    double pdf = exp(exponent);
    pdf /= norm_scale * sqrt(2 * 3.14159265358979323846);
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
    reinterpret_cast<double4*>(d.state)[i] = make_double4(mu, sigma2, a, b);
    if (d.ok) d.ok[i] = ok ? 1 : 0;
}

}  // namespace


edsgpu_status edsgpu_depth_update_tracked(edsgpu_depth_points* d, const double* tracker_state_dev, const double* kf_coord_dev,
                                          const double* ef_coord_dev) {
    edsgpu_ctx* ctx = d->ctx;
    DepthArgs a{};
    a.N = d->N; a.tracks = 0;
    a.fx = d->fx; a.fy = d->fy; a.cx = d->cx; a.cy = d->cy; a.mu_range = d->mu_range; a.px_error_angle = d->px_error_angle;
    a.kf_coord = kf_coord_dev; a.ef_coord = ef_coord_dev;
    a.state = d->state; a.ok = d->ok; a.skip = d->outlier;
    depth_update_tracked_kernel<<<(d->N + 127) / 128, 128, 0, ctx->stream>>>(a, tracker_state_dev);
    ctx->launches++;
    EDS_CUDA(ctx, cudaGetLastError());
    return EDSGPU_OK;
}

extern "C" {

edsgpu_status edsgpu_depth_points_create(edsgpu_ctx* ctx, int num_points, double fx, double fy, double cx, double cy, double min_depth,
                                         double max_depth, const double* inv_depth, double init_a, double init_b, edsgpu_depth_points** out) {
    if (!ctx || !out) return EDSGPU_INVALID_ARGUMENT;
    EDS_REQUIRE(ctx, num_points > 0 && fx > 0 && fy > 0 && max_depth > min_depth, "depth_points_create: bad arguments");
    DeviceGuard g(ctx->device);
    edsgpu_depth_points* d = new edsgpu_depth_points();
    d->ctx = ctx; d->N = num_points; d->fx = fx; d->fy = fy; d->cx = cx; d->cy = cy;
    d->mu_range = max_depth - min_depth;                                   // DepthPoints.cpp:58,78
    const double px_noise = 3.0;                                           // DepthPoints.hpp:37
    d->px_error_angle = atan(px_noise / (2.0 * fx)) + atan(px_noise / (2.0 * fy));  // getAngleError, :151-154
    cudaError_t e = cudaMalloc(&d->state, sizeof(double) * 4 * (size_t)num_points);
    if (e == cudaSuccess) e = cudaMalloc(&d->coords, sizeof(double) * 4 * (size_t)num_points);
    if (e == cudaSuccess) e = cudaMalloc(&d->ok, (size_t)num_points);
    if (e == cudaSuccess) e = cudaMalloc(&d->outlier, (size_t)num_points);
    if (e == cudaSuccess) e = cudaMemsetAsync(d->outlier, 0, (size_t)num_points, ctx->stream);
    if (e != cudaSuccess) { edsgpu_depth_points_destroy(d); return edsgpu_fail(ctx, EDSGPU_CUDA_ERROR, cudaGetErrorString(e)); }
    // init (:52-91): without inverse depths every point starts at the mean depth with sigma2 = mu_range^2,
    // with inverse depths (from the global map) at them with sigma2 = mu_range^2 / 36
    std::vector<double> h(4 * (size_t)num_points);
    for (int i = 0; i < num_points; ++i) {
        h[4 * (size_t)i + 0] = inv_depth ? inv_depth[i] : 1.0 / ((max_depth - min_depth) / 2.0);
        h[4 * (size_t)i + 1] = inv_depth ? (d->mu_range * d->mu_range) / 36.0 : d->mu_range * d->mu_range;
        h[4 * (size_t)i + 2] = init_a;
        h[4 * (size_t)i + 3] = init_b;
    }
    EDS_CUDA(ctx, cudaMemcpyAsync(d->state, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, ctx->stream));
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = d;
    return EDSGPU_OK;
}

void edsgpu_depth_points_destroy(edsgpu_depth_points* d) {
    if (!d) return;
    DeviceGuard g(d->ctx->device);
    cudaStreamSynchronize(d->ctx->stream);
    if (d->state) cudaFree(d->state);
    if (d->coords) cudaFree(d->coords);
    if (d->ok) cudaFree(d->ok);
    if (d->outlier) cudaFree(d->outlier);
    delete d;
}

edsgpu_status edsgpu_depth_points_update(edsgpu_depth_points* d, const double T_kf_ef[16], const double* kf_coord, const double* ef_coord,
                                         int coords_are_tracks, uint8_t* ok_out) {
    if (!d) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = d->ctx;
    EDS_REQUIRE(ctx, T_kf_ef && kf_coord && ef_coord, "depth_points_update: null argument");
    DeviceGuard g(ctx->device);
    const size_t n = (size_t)d->N;
    EDS_CUDA(ctx, cudaMemcpyAsync(d->coords, kf_coord, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, ctx->stream));
    d->kf_coord_set = true;
    EDS_CUDA(ctx, cudaMemcpyAsync(d->coords + 2 * n, ef_coord, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, ctx->stream));
    DepthArgs a{};
    a.N = d->N; a.tracks = coords_are_tracks ? 1 : 0;
    a.fx = d->fx; a.fy = d->fy; a.cx = d->cx; a.cy = d->cy; a.mu_range = d->mu_range; a.px_error_angle = d->px_error_angle;
    // T_ef_kf = T_kf_ef^-1: R^T, -R^T t; P_ef = K [R_ef_kf | t_ef_kf]  (:99-106)
    double R[9], t[3], Rt[9], tt[3];
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R[3 * r + c] = T_kf_ef[4 * r + c]; t[r] = T_kf_ef[4 * r + 3]; }
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Rt[3 * r + c] = R[3 * c + r];
    for (int r = 0; r < 3; ++r) tt[r] = -(Rt[3 * r] * t[0] + Rt[3 * r + 1] * t[1] + Rt[3 * r + 2] * t[2]);
    const double K[9] = {d->fx, 0, d->cx, 0, d->fy, d->cy, 0, 0, 1};
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) a.M2[3 * r + c] = K[3 * r] * Rt[c] + K[3 * r + 1] * Rt[3 + c] + K[3 * r + 2] * Rt[6 + c];
        a.epipole[r] = K[3 * r] * tt[0] + K[3 * r + 1] * tt[1] + K[3 * r + 2] * tt[2];
        a.t[r] = t[r];
    }
    a.tnorm = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
    a.kf_coord = d->coords; a.ef_coord = d->coords + 2 * n;
    a.state = d->state; a.ok = d->ok;
    depth_update_kernel<<<(d->N + 127) / 128, 128, 0, ctx->stream>>>(a);
    ctx->launches++;
    EDS_CUDA(ctx, cudaGetLastError());
    if (ok_out) EDS_CUDA(ctx, cudaMemcpyAsync(ok_out, d->ok, n, cudaMemcpyDeviceToHost, ctx->stream));
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the caller's coordinate arrays may be pageable and reused
    return EDSGPU_OK;
}

edsgpu_status edsgpu_depth_points_get(edsgpu_depth_points* d, double* state_out) {
    if (!d || !state_out) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = d->ctx;
    DeviceGuard g(ctx->device);
    EDS_CUDA(ctx, cudaMemcpyAsync(state_out, d->state, sizeof(double) * 4 * (size_t)d->N, cudaMemcpyDeviceToHost, ctx->stream));
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EDSGPU_OK;
}

}  // extern "C"
