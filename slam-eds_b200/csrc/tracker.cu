// Event-to-model alignment on the device.
//
// Replaces Tracker::optimize (reference src/tracking/Tracker.cpp:104-241), the Ceres cost
// functor PhotometricError (src/tracking/PhotometricError.hpp:56-214) and ceres::Solve
// (un-vendored; trust-region LM restated from ceres-solver 1.14..2.1 semantics).
//
//   track_lm_kernel   thread-block CLUSTERS of 512-thread CTAs, one CTA per SM; a cluster keeps K <= 4
//                     tracking problems in flight.  Residual blocks (Tracker.cpp:178-195) are dealt
//                     round-robin to the CTAs of the cluster.  In every CTA producer warps evaluate
//                     32 points at a time (division-free fp64-exact projection, bicubic window as
//                     four texture gathers, analytic Jacobian in fp32; software-pipelined) into a
//                     shared-memory ring guarded by mbarriers, consumer warps own the 90 fp32
//                     outer-product accumulators, reduce a block with a halving warp butterfly,
//                     apply the per-block loss (rho') and store 92 doubles into the shared memory
//                     of the problem's leader CTA over DSMEM.  One leader warp per problem runs the
//                     whole Levenberg-Marquardt state machine (Jacobi scaling, damping, register-
//                     resident 12x12 Cholesky, retractions, accept/reject, tolerances) and publishes
//                     the next evaluation point to the cluster.  Evaluators and leaders hand over
//                     through ready/result mbarriers (no cluster-wide barrier in the loop), so the
//                     serial LM step of one problem overlaps the sweeps of the others.  The LM loop
//                     never returns to the host: one launch per batch of windows.
//   mad_kernel        next loss parameter (MAD / STD) from the written-back residuals
//                     (Tracker.cpp:281-317) by radix select.
//   kf_prepare_kernel keyframe upload: fp32 SoA gather streams, fp64 3-D points
//                     (PhotometricError.hpp:94-105) and the per-block 6x6 model Gram matrix
//                     A_b = sum g_i g_i^T (m_i = g_i . v is linear in v, so
//                     ||m||^2 = v^T A_b v and sum m_i g_i = A_b v need no sweep).
#include <cooperative_groups.h>
#include <float.h>
#include <stdlib.h>

#include "common.cuh"
#include "depth.cuh"
#include "frames.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int TRK_THREADS = 512;  // one CTA per SM: 13-14 producer warps, 2 consumer warps, (1 leader warp)
constexpr int TRK_WARPS = TRK_THREADS / 32;
constexpr int JLD = 20;         // floats per point row in shared memory (80 B: conflict-free 128-bit access)
#ifndef EDS_N_CONS
#define EDS_N_CONS 3
#endif
constexpr int N_CONS = EDS_N_CONS;           // consumer warps: own the outer-product accumulators (batches j = c mod N_CONS)
constexpr int N_PROD = TRK_WARPS - N_CONS;   // producer warps (one fewer in a CTA that hosts a leader warp)
constexpr int LEADER_WARP = TRK_WARPS - 1;
#ifndef EDS_RING_DEPTH
#define EDS_RING_DEPTH 3
#endif
constexpr int RING_DEPTH = EDS_RING_DEPTH;   // ring slots per producer warp (three absorb stragglers)
// capacity of the ring of 32-point batches; a CTA uses RING_DEPTH slots per producer warp it actually has
constexpr int N_SLOTS = RING_DEPTH * N_PROD;
constexpr int NSLOT = 92;       // 78 + 12 + cost + block squared norm
constexpr int MAX_BLOCKS = 16;  // residual blocks per problem (config.options.num_threads)
constexpr int MAX_CLUSTER = 8;
#ifndef EDS_MAX_K
#define EDS_MAX_K 4
#endif
constexpr int MAX_K = EDS_MAX_K;  // problems one cluster keeps in flight
constexpr double kEps = 1e-05;  // PhotometricError.hpp:200

enum { CMD_EVAL = 1, CMD_FINAL = 2, CMD_DONE = 3 };

#ifdef EDS_TIMING
__device__ unsigned long long g_timing[32];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#endif
enum { PHASE_INIT = 0, PHASE_CAND = 1 };

struct KfDev {
    const float4* gxy;   // {Gx, Gy, X, Y}
    const float2* dw;    // {idp, weight}
    const double* kpx;   // 3-D point (X,Y,1)/(idp+eps)
    const double* kpy;
    const double* kpz;
    const double* A;     // [B][21] upper triangle of sum g g^T per residual block
    int N, B, H, W;
    int ne;              // N / B: points per residual block (the last block also takes the remainder)
    double fx, fy, cx, cy;
};

struct ProblemDesc {
    KfDev kf;
    cudaTextureObject_t frame; // H x W un-normalised event frame (point-sampled, clamp-to-edge)
    const double* norms;       // {norm, 1/norm}
    double* state;             // 14 doubles: px(3) qx(4) vx(6) loss_param; in-out
    float* residuals;          // N, written by the final sweep
    edsgpu_tracker_info* info; // device
    int loss_type, max_iter, loss_param_method, eval_only;
    double ftol, gtol, ptol;
    float* jac_out;            // eval_only: N x 12 local Jacobian (no loss) or null
    double* eval_out;          // eval_only: cost, H(144), g(12)
};

// ------------------------------------------------------------------------------------------
// small fp64 helpers (device)
// ------------------------------------------------------------------------------------------
// Eigen::Quaternion::toRotationMatrix (no normalisation), PhotometricError.hpp:163
__device__ __forceinline__ void quat_to_rot(const double* q, double* R) {
    const double tx = 2.0 * q[0], ty = 2.0 * q[1], tz = 2.0 * q[2];
    const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
    const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
    const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
    R[3] = txy + twz;         R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.0 - (txx + tyy);
}

// Program-level Plus over [p(3) q(4) v(6)] <- delta(12):
// ceres::EigenQuaternionParameterization::Plus and UnitNormVectorAddition
// (PhotometricError.hpp:32-54), Tracker.cpp:111-114,197-198.
__device__ void state_plus(const double* x, const double* d, double* out) {
    for (int i = 0; i < 3; ++i) out[i] = x[i] + d[i];
    const double n2 = d[3] * d[3] + d[4] * d[4] + d[5] * d[5];
    if (n2 > 0.0) {
        double s, c;  // s = sin(|d|)/|d|, c = cos(|d|)
        if (n2 < 0.0625) {
            // |d| < 0.25: Taylor series in n2, truncation error < 1e-18 (faster than sin/cos + sqrt + div
            // on the serial LM path; agrees with libm to the last ulp or two)
            s = 1.0 + n2 * (-1.0 / 6 + n2 * (1.0 / 120 + n2 * (-1.0 / 5040 + n2 * (1.0 / 362880 + n2 * (-1.0 / 39916800 + n2 * (1.0 / 6227020800.0))))));
            c = 1.0 + n2 * (-0.5 + n2 * (1.0 / 24 + n2 * (-1.0 / 720 + n2 * (1.0 / 40320 + n2 * (-1.0 / 3628800 + n2 * (1.0 / 479001600.0 + n2 * (-1.0 / 87178291200.0)))))));
        } else {
            const double nd = sqrt(n2);
            s = sin(nd) / nd;
            c = cos(nd);
        }
        const double ax = s * d[3], ay = s * d[4], az = s * d[5], aw = c;
        const double bx = x[3], by = x[4], bz = x[5], bw = x[6];
        out[6] = aw * bw - ax * bx - ay * by - az * bz;
        out[3] = aw * bx + ax * bw + ay * bz - az * by;
        out[4] = aw * by + ay * bw + az * bx - ax * bz;
        out[5] = aw * bz + az * bw + ax * by - ay * bx;
    } else {
        for (int i = 3; i < 7; ++i) out[i] = x[i];
    }
    double sum = 0.0;
    for (int i = 0; i < 6; ++i) { double s = x[7 + i] + d[6 + i]; sum += s * s; out[7 + i] = s; }
    sum = rsqrt(sum);
    for (int i = 0; i < 6; ++i) out[7 + i] *= sum;
}

// ceres::HuberLoss / ceres::CauchyLoss (Tracker.cpp:146-161): rho(s), rho'(s)
__device__ __forceinline__ void loss_eval(int type, double a, double s, double* rho0, double* rho1) {
    if (type == EDSGPU_LOSS_HUBER) {
        const double b = a * a;
        if (s > b) {
            const double r = sqrt(s);
            *rho0 = 2.0 * a * r - b;
            *rho1 = fmax(DBL_MIN, a / r);
        } else { *rho0 = s; *rho1 = 1.0; }
    } else if (type == EDSGPU_LOSS_CAUCHY) {
        const double b = a * a, c = 1.0 / b;
        const double sum = 1.0 + s * c;
        *rho0 = b * log(sum);
        *rho1 = fmax(DBL_MIN, 1.0 / sum);
    } else { *rho0 = s; *rho1 = 1.0; }
}

__host__ __device__ __forceinline__ int tri_index(int a, int b) {  // a <= b < 12, row-major upper triangle
    return a * 12 - (a * (a - 1)) / 2 + (b - a);
}

// ------------------------------------------------------------------------------------------
// shared-memory layout of one CTA
// ------------------------------------------------------------------------------------------
struct LmState {
    double x[13], cand[13], delta[12];
    double Hs[78], gs[12];    // Jacobi-scaled normal equations at x
    double scale[12], diag[12];
    double x_cost, mcc, radius, dec, gmax, x_norm, initial_cost;
    int reuse_diag, iter, n_succ, n_unsucc, consec_invalid, termination, phase, n_eval;
};

// Everything a CTA needs for one evaluation; written by the leader into every CTA of the
// cluster (DSMEM) before sync (A).
struct EvalConst {
    double R[9], t[3];               // Eigen toRotationMatrix(q), translation
    float vf[6], inv_vs, inv_vn;     // velocity, 1/|v|^2, 1/|v|
    float blk[MAX_BLOCKS][8];        // per residual block: 1/M, alpha, beta[6] (see eval_point)
    int cmd;
};

// Per-problem state.  A cluster keeps up to MAX_K problems in flight (see track_lm_kernel): every CTA holds the
// published constants of all of them, the CTA that hosts a problem's leader additionally its reduction slots and LM state.
struct ProblemShared {
    EvalConst ec;
    double loss_a;
    // leader CTA only
    double slots[MAX_BLOCKS][NSLOT];  // one slot per residual block, filled over DSMEM
    double sum[NSLOT];
    double A[MAX_BLOCKS][21];         // per-block model Gram matrices
    double x_eval[13];
    LmState lm;
    ProblemDesc P;
    int valid;
};

struct CtaShared {
    ProblemShared prob[MAX_K];
    // ring of 32-point batches of rows [J(12) r pad], guarded by full/empty mbarriers
    alignas(16) float ring[N_SLOTS][32][JLD];
    alignas(8) unsigned long long full_bar[N_SLOTS];
    alignas(8) unsigned long long empty_bar[N_SLOTS];
    // block totals of consumer warps 1.., handed to consumer warp 0 (double-buffered by block)
    float cons_part[2][N_CONS - 1][96];
    double cons_s[2][N_CONS - 1];
    // dataflow between the evaluator warps of all CTAs and the leader warp of each problem
    alignas(8) unsigned long long ready_bar[MAX_K];   // every CTA: "evaluation constants of problem k have landed" (32 leader lanes)
    alignas(8) unsigned long long result_bar[MAX_K];  // leader CTA of k: "every evaluator warp of the cluster is done with problem k"
};

// ------------------------------------------------------------------------------------------
// per-point residual + analytic tangent-space Jacobian
// ------------------------------------------------------------------------------------------
// Catmull-Rom weights of ceres' CubicHermiteSpline (call site PhotometricError.hpp:172) in basis
// form: f = sum_k w_k(x) p_k, f' = sum_k dw_k(x) p_k  (same cubic as the Horner form a,b,c,d)
__device__ __forceinline__ void cr_weights(float x, float* w, float* dw) {
    const float x2 = x * x;
    w[0] = 0.5f * x * ((2.0f - x) * x - 1.0f);
    w[1] = 0.5f * (x2 * (3.0f * x - 5.0f) + 2.0f);
    w[2] = 0.5f * x * ((4.0f - 3.0f * x) * x + 1.0f);
    w[3] = 0.5f * x2 * (x - 1.0f);
    dw[0] = 0.5f * ((4.0f - 3.0f * x) * x - 1.0f);
    dw[1] = 0.5f * x * (9.0f * x - 10.0f);
    dw[2] = 0.5f * ((8.0f - 9.0f * x) * x + 1.0f);
    dw[3] = 0.5f * x * (3.0f * x - 2.0f);
}

// Stage 1 of a point: warp into the event camera and project (PhotometricError.hpp:157-168) ->
// interpolation cell, in-cell fractions and the camera-frame quantities the Jacobian needs.
struct PointGeo {
    int col, row;      // cell, clamped to [-4, W+3] x [-4, H+3]
    float tc, tr;      // fractions in the cell (0 where clamped)
    float iz, px, py;  // P = R kp + t: 1/Pz, Px, Py
    float ax, ay, az;  // R kp
};

// (u - cell, cell) for u = f p / pz + c without an fp64 division on the critical path: the cell comes
// from an fp32 estimate, the fraction from the exact fp64 numerator (f p + (c - cell) pz) times the
// fp32 reciprocal, i.e. absolute error ~1e-7 px where plain fp32 projection would carry ~2e-5 px
// (SURVEY.md section 7).  A cell guessed one off is repaired; at a cell boundary both choices give
// the same interpolant.  Outside [-4, limit + 3] the clamped Grid2D is constant: fraction dropped.
__device__ __forceinline__ void project_axis(double f, double c, double p, double pz, float pf, float izf, int limit, int& cell, float& frac) {
    const float est = fmaf((float)f * pf, izf, (float)c);
    int q = max(-6, min(__float2int_rd(est), limit + 5));  // saturating conversion, NaN -> 0
    float t = (float)fma(f, p, (c - (double)q) * pz) * izf;
    const int below = t < 0.f ? 1 : 0, above = t >= 1.f ? 1 : 0;  // at most one of them
    q += above - below;
    t += (float)(below - above);
    cell = max(-4, min(q, limit + 3));
    frac = (cell == q) ? t : 0.f;
}

struct Kp { double x, y, z; };  // 3-D key-frame point (X,Y,1)/(idp+eps)
__device__ __forceinline__ Kp load_kp(const KfDev& kf, int idx) { return Kp{__ldg(&kf.kpx[idx]), __ldg(&kf.kpy[idx]), __ldg(&kf.kpz[idx])}; }

__device__ __forceinline__ void point_geometry(const KfDev& kf, const EvalConst& K, const Kp& kp, PointGeo& G) {
    const double kx = kp.x, ky = kp.y, kz = kp.z;
    const double ax = K.R[0] * kx + K.R[1] * ky + K.R[2] * kz;
    const double ay = K.R[3] * kx + K.R[4] * ky + K.R[5] * kz;
    const double az = K.R[6] * kx + K.R[7] * ky + K.R[8] * kz;
    const double px = ax + K.t[0], py = ay + K.t[1], pz = az + K.t[2];
    G.ax = (float)ax; G.ay = (float)ay; G.az = (float)az;
    G.px = (float)px; G.py = (float)py;
    G.iz = __frcp_rn((float)pz);
    project_axis(kf.fx, kf.cx, px, pz, G.px, G.iz, kf.W, G.col, G.tc);
    project_axis(kf.fy, kf.cy, py, pz, G.py, G.iz, kf.H, G.row, G.tr);
}

// 4x4 taps as four 2x2 texture gathers. A gather at the corner shared by texels (i,j),(i+1,j),
// (i,j+1),(i+1,j+1) returns them as {w,z,x,y}; clamp-to-edge addressing clamps every texel index on
// its own, which is exactly the clamped ceres::Grid2D.
struct Taps { float4 q00, q10, q01, q11; };
__device__ __forceinline__ Taps fetch_taps(cudaTextureObject_t frame, int col, int row) {
    const float xc = (float)col, yr = (float)row;
    Taps t;
    t.q00 = tex2Dgather<float4>(frame, xc, yr, 0);
    t.q10 = tex2Dgather<float4>(frame, xc + 2.f, yr, 0);
    t.q01 = tex2Dgather<float4>(frame, xc, yr + 2.f, 0);
    t.q11 = tex2Dgather<float4>(frame, xc + 2.f, yr + 2.f, 0);
    return t;
}

// per-block constants published by the leader: bc = {1/M, alpha, beta[6]} with
// alpha = 1/(M |v|), beta_k = (c_k/M^3 + kappa v_k)/|v|, kappa = (1/M - c.v/M^3)/|v|^2, so that the
// tangent-space velocity Jacobian  (w (g/M - m c/M^3)) (I - v v^T/|v|^2)/|v|  =  w (alpha g - m beta)
//
// Stage 2 of a point: bicubic sample + residual + Jacobian row from the geometry, the fetched taps and
// the point's gradient record g4 = {Gx, Gy, X, Y}, dw = {inverse depth, weight}.
template <bool WANT_J>
__device__ __forceinline__ void point_finish(const KfDev& kf, const EvalConst& K, const float* __restrict__ bc, float inv_norm,
                                             const PointGeo& G, const Taps& T, float4 g4, float2 dw, float* __restrict__ J, float& r) {
    const float Gx = g4.x, Gy = g4.y, X = g4.z, Y = g4.w, d = dw.x, w = dw.y;
    // model term: m = g . v with g = -(Gx dflow_x/dv + Gy dflow_y/dv), PhotometricError.hpp:114-122,145
    float g[6];
    g[0] = Gx * d;
    g[1] = Gy * d;
    g[2] = -(Gx * X + Gy * Y) * d;
    g[3] = -(Gx * X * Y + Gy * (1.0f + Y * Y));
    g[4] = Gx * (1.0f + X * X) + Gy * X * Y;
    g[5] = Gy * X - Gx * Y;
    float m = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) m += g[k] * K.vf[k];
    float wc[4], dwc[4], wr[4], dwr[4];
    cr_weights(G.tc, wc, dwc);
    cr_weights(G.tr, wr, dwr);
    const float taps[4][4] = {{T.q00.w, T.q00.z, T.q10.w, T.q10.z}, {T.q00.x, T.q00.y, T.q10.x, T.q10.y},
                              {T.q01.w, T.q01.z, T.q11.w, T.q11.z}, {T.q01.x, T.q01.y, T.q11.x, T.q11.y}};
    float f = 0.f, dfdr = 0.f, dfdc = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float p0 = taps[k][0], p1 = taps[k][1], p2 = taps[k][2], p3 = taps[k][3];
        const float fr = wc[0] * p0 + wc[1] * p1 + wc[2] * p2 + wc[3] * p3;
        f += wr[k] * fr;
        dfdr += dwr[k] * fr;
        if (WANT_J) dfdc += wr[k] * (dwc[0] * p0 + dwc[1] * p1 + dwc[2] * p2 + dwc[3] * p3);
    }
    const float e = inv_norm * f;
    r = w * (m * bc[0] - e);  // PhotometricError.hpp:173
    if (!WANT_J) return;
    const float er = inv_norm * dfdr, ec = inv_norm * dfdc;
    // d r / d P  (P = R kp + t)
    const float fxf = (float)kf.fx, fyf = (float)kf.fy;
    const float wiz = w * G.iz;
    const float dPx = -wiz * ec * fxf;
    const float dPy = -wiz * er * fyf;
    const float dPz = -(dPx * G.px + dPy * G.py) * G.iz;
    J[0] = dPx; J[1] = dPy; J[2] = dPz;
    // quaternion tangent: q <- [sin|d| d/|d|, cos|d|] * q rotates by 2|d|: dP/dtheta = -2 [R kp]x
    const float axf = 2.0f * G.ax, ayf = 2.0f * G.ay, azf = 2.0f * G.az;
    J[3] = ayf * dPz - azf * dPy;
    J[4] = azf * dPx - axf * dPz;
    J[5] = axf * dPy - ayf * dPx;
    // velocity block in the tangent space of the unit-norm retraction
    const float wa = w * bc[1], wm = w * m;
#pragma unroll
    for (int k = 0; k < 6; ++k) J[6 + k] = wa * g[k] - wm * bc[2 + k];
}

// both stages back to back (residual-only sweeps, small paths)
template <bool WANT_J>
__device__ __forceinline__ void eval_point(const KfDev& kf, const EvalConst& K, const float* __restrict__ bc,
                                           cudaTextureObject_t frame, float inv_norm, int idx, float* __restrict__ J, float& r) {
    PointGeo G;
    point_geometry(kf, K, load_kp(kf, idx), G);
    const Taps T = fetch_taps(frame, G.col, G.row);
    point_finish<WANT_J>(kf, K, bc, inv_norm, G, T, __ldg(&kf.gxy[idx]), __ldg(&kf.dw[idx]), J, r);
}

// ---- mbarrier helpers (shared::cta) ---------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity, int tag = 0) {
    unsigned ok;
#ifdef EDS_WATCHDOG
    unsigned long long spins = 0;
#endif
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(4000u)  // suspend-time hint (ns): sleep instead of spinning
            : "memory");
#ifdef EDS_WATCHDOG
        if (!ok && ++spins > 400000ull) {  // debug build: report the wait that does not end (~1 s), then stop the kernel
            if ((threadIdx.x & 31) == 0)
                printf("[watchdog] block %d warp %d stuck in wait %d, barrier offset %u, parity %u\n", (int)blockIdx.x, (int)(threadIdx.x >> 5), tag,
                       smem_u32(bar), parity);
            __nanosleep(100000000);
            __trap();
        }
#endif
    } while (!ok);
}

// ---- cluster-scope variants: barrier in another CTA of the cluster, data written over DSMEM ----
__device__ __forceinline__ unsigned mapa_u32(unsigned local_addr, unsigned cta_rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(unsigned long long* local_bar, unsigned cta_rank) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa_u32(smem_u32(local_bar), cta_rank)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(unsigned long long* local_bar, unsigned cta_rank) {  // after fence_cluster()
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa_u32(smem_u32(local_bar), cta_rank)) : "memory");
}
__device__ __forceinline__ void fence_cluster();
__device__ __forceinline__ void mbar_wait_cluster(unsigned long long* bar, unsigned parity, int tag = 0) {
    unsigned ok;
#ifdef EDS_WATCHDOG
    unsigned long long spins = 0;
#endif
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(4000u)
            : "memory");
#ifdef EDS_WATCHDOG
        if (!ok && ++spins > 400000ull) {
            if ((threadIdx.x & 31) == 0)
                printf("[watchdog] block %d warp %d stuck in cluster wait %d, barrier offset %u, parity %u\n", (int)blockIdx.x, (int)(threadIdx.x >> 5),
                       tag, smem_u32(bar), parity);
            __nanosleep(100000000);
            __trap();
        }
#endif
    } while (!ok);
}
__device__ __forceinline__ void fence_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }

// halving butterfly: 2*HALF per-lane values -> HALF, lanes with bit OFFSET keep the upper half
template <int HALF, int OFFSET>
__device__ __forceinline__ void butterfly_step(float* a, unsigned lane) {
    const bool upper = (lane & OFFSET) != 0;
#pragma unroll
    for (int i = 0; i < HALF; ++i) {
        const float send = upper ? a[i] : a[i + HALF];
        const float keep = upper ? a[i + HALF] : a[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFFSET);
    }
}

// consumer: rows [R0,R1) of the upper triangle of [J r]^T [J r] (index 12 = r; the r*r entry is kept
// in fp64 by the producers).  Entry order: row-major over (a, c >= a).
template <int R0, int R1>
struct RowBlock {
    static constexpr int count() { int n = 0; for (int a = R0; a < R1; ++a) n += 13 - a; return n; }
    // slot of local entry i: 78 upper-triangle entries of J^T J, then 12 of J^T r
    static __device__ __forceinline__ int slot(int i) {
        int a = R0;
        while (i >= 13 - a) { i -= 13 - a; ++a; }
        const int c = a + i;
        return (c < 12) ? tri_index(a, c) : 78 + a;
    }
    static __device__ __forceinline__ void load(const float (*rows)[JLD], int lane, float* v) {
        const float4* row = reinterpret_cast<const float4*>(&rows[lane][0]);
#pragma unroll
        for (int q = R0 / 4; q < 4; ++q) {
            const float4 t = row[q];
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
    }
    static __device__ __forceinline__ float accumulate(const float* v, float* acc) {
        int e = 0;
#pragma unroll
        for (int a = R0; a < R1; ++a)
#pragma unroll
            for (int c = a; c < 13; ++c) acc[e++] += v[a] * v[c];
        return v[12];
    }
};
typedef RowBlock<0, 12> ConsRows;  // all 90 entries per consumer warp; the consumers split the batches, not the rows

// 96 per-lane partial sums -> warp totals; lane l ends with entries base(l)+{0,1,2}
__device__ __forceinline__ void reduce96(float* acc, unsigned lane) {
    butterfly_step<48, 16>(acc, lane);
    butterfly_step<24, 8>(acc, lane);
    butterfly_step<12, 4>(acc, lane);
    butterfly_step<6, 2>(acc, lane);
    butterfly_step<3, 1>(acc, lane);
}
__device__ __forceinline__ int reduce96_base(unsigned lane) {
    return 48 * ((lane >> 4) & 1) + 24 * ((lane >> 3) & 1) + 12 * ((lane >> 2) & 1) + 6 * ((lane >> 1) & 1) + 3 * (lane & 1);
}

// One CTA evaluates the residual blocks dealt to it with the constants in sh.ec and stores, per
// block, [rho' * JtJ (78) | rho' * Jtr (12) | 0.5 rho(s) | s] into slot_base[b] (leader smem, DSMEM).
//
// Warp-specialised: producer warps sweep the points 32 at a time (residual + analytic Jacobian row
// -> shared-memory ring slot, mbarrier "full"), consumer warps own the outer-product accumulators
// (90 fp32 registers per lane) and drain the ring (mbarrier "empty").  No CTA-wide barrier inside
// the sweep; 512 threads at <= 128 registers per thread = one CTA per SM, whose 16 warps hide each other's stalls
// (and the leader's serial LM step).  `batch_counter` numbers the batches of the whole kernel so that
// both sides derive slot and phase parity without talking to each other.
// Evaluator roles of one CTA: warps 0..N_PROD-2 produce, the next N_CONS warps consume, the last warp
// is the dedicated LM leader warp in a CTA that hosts a problem's leader and one more producer elsewhere.
struct Roles {
    int n_prod;      // producer warps of this CTA
    int pidx;        // producer index of this warp, -1 if not a producer
    int cidx;        // consumer index of this warp, -1 if not a consumer
    int n_eval_threads;  // threads taking part in the evaluation
    int etid;        // evaluator thread index
};
__device__ __forceinline__ Roles make_roles(bool hosts_leader) {
    const int warp = threadIdx.x >> 5;
    Roles r;
    r.n_prod = hosts_leader ? N_PROD - 1 : N_PROD;
    r.cidx = (warp >= N_PROD - 1 && warp < LEADER_WARP) ? warp - (N_PROD - 1) : -1;
    static_assert(N_CONS >= 2 && N_CONS <= 4, "consumer warps");
    r.pidx = (warp < N_PROD - 1) ? warp : ((warp == LEADER_WARP && !hosts_leader) ? N_PROD - 1 : -1);
    r.n_eval_threads = hosts_leader ? TRK_THREADS - 32 : TRK_THREADS;
    r.etid = threadIdx.x;  // the leader warp is the last one: evaluator threads keep their index
    return r;
}

template <bool RES_ONLY, bool HOSTS_LEADER>
__device__ void cta_evaluate(ProblemShared& ps, CtaShared& sh, double* slot_base /* [MAX_BLOCKS][NSLOT] on the leader */,
                             int rank, int csize, const Roles& role, bool write_residuals, unsigned& batch_counter, unsigned& block_counter) {
    const int tid = threadIdx.x, lane = tid & 31;
    const ProblemDesc& P = ps.P;
    const KfDev& kf = P.kf;
    const EvalConst& ec = ps.ec;
    const float inv_norm = (float)P.norms[1];
    const double loss_a = ps.loss_a;
    const int ne = kf.ne;
    // The ring of this CTA has a whole number of slots per producer warp, so that batch g and batch g + n_slots -- the
    // successive occupants of a slot -- always belong to the SAME producer.  That warp passes through every
    // "empty" wait of its slots in order and can never be a whole lap ahead of the barrier's phase (with slots shared
    // between producers a fast warp could test a parity that is two phases stale and overwrite an unconsumed batch).
    constexpr int CTA_PROD = HOSTS_LEADER ? N_PROD - 1 : N_PROD;  // compile-time: strides and wrap tests fold into immediates
    constexpr unsigned n_slots = (unsigned)(CTA_PROD * (N_SLOTS / CTA_PROD));
    // The mirror image on the "full" side: within a residual block the successive occupants of a slot must be drained by
    // the SAME consumer warp (it takes its batches in order, so it cannot test a stale parity either); batch j goes to
    // consumer j mod N_CONS, hence the ring size must be a multiple of N_CONS.  Across blocks the consumers' end-of-block
    // barrier keeps them within one block of each other.
    static_assert(n_slots % N_CONS == 0, "ring slots per CTA must be a multiple of the consumer warps");
    // The texture handle is read from shared memory, which the compiler cannot prove warp-uniform: it
    // would wrap every fetch in a loop over the distinct handles of the warp.  A warp-wide OR leaves
    // the value unchanged and lands in a uniform register.
    const cudaTextureObject_t frame = ((unsigned long long)__reduce_or_sync(0xffffffffu, (unsigned)(P.frame >> 32)) << 32) |
                                      (unsigned long long)__reduce_or_sync(0xffffffffu, (unsigned)P.frame);
    if constexpr (!RES_ONLY) {
        if (role.pidx >= 0) {
            // ---------------- producer ----------------
            // The CTA's batches of this visit are numbered t = 0 .. nb_tot-1 across its blocks (every block has
            // nb_reg batches except the problem's last one, which takes the remainder and belongs to one CTA as its
            // last block); batch t lives in ring slot (batch_counter + t) mod n_slots.  Warp p takes the batches
            // t = t0 + k n_prod, t0 chosen from the running number so that uneven shares even out over visits.
            // A batch is located from t alone (one multiply-high), so the three pipeline stages need no cursor state:
            //   A(k+2)  3-D points of the batch after next: loads in flight
            //   B(k+1)  fp64 geometry of the next batch, then its four texture gathers + gradient record in flight
            //   C(k)    finish the current batch (its gathers were issued a whole iteration ago) and hand it over
            const int B = kf.B;
            constexpr int n_prod = CTA_PROD;
            const int n_last = kf.N - (B - 1) * ne;  // Tracker.cpp:178-190
            const int nb_reg = (ne + 31) >> 5, nb_last = (n_last + 31) >> 5;
            const int n_blk = (B - rank + csize - 1) / csize;
            const bool has_last = ((B - 1 - rank) % csize) == 0;
            const int nb_tot = n_blk * nb_reg + (has_last ? nb_last - nb_reg : 0);
            // floor(t / nb_reg) = umulhi(t, ceil(2^32 / nb_reg)), exact while t nb_reg < 2^32 (keyframe_create: N <= 2^20);
            // nb_reg = 1 has no 32-bit magic number: the quotient is t itself
            const unsigned magic = nb_reg > 1 ? 0xFFFFFFFFu / (unsigned)nb_reg + 1u : 0u;
            auto locate = [&](int t, int& idx, int& b) -> bool {
                const int bl = min(nb_reg > 1 ? (int)__umulhi((unsigned)t, magic) : t, n_blk - 1);
                b = rank + bl * csize;
                const int i = ((t - bl * nb_reg) << 5) + lane;
                idx = b * ne + i;
                return (t < nb_tot) && (i < ((b + 1 == B) ? n_last : ne));
            };
            struct Stage { Taps T; float4 g4; float2 dw; PointGeo G; int idx, b; bool valid; };
            const PointGeo G0 = {0, 0, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const Kp kp0 = {0.0, 0.0, 1.0};
            // stage B: geometry + gathers in flight
            auto stage_b = [&](Stage& s, const Kp& kp) {
                s.G = G0;
                s.g4 = make_float4(0.f, 0.f, 0.f, 0.f);
                s.dw = make_float2(0.f, 0.f);
                if (s.valid) {
                    point_geometry(kf, ec, kp, s.G);
                    s.g4 = __ldg(&kf.gxy[s.idx]);
                    s.dw = __ldg(&kf.dw[s.idx]);
                }
                s.T = fetch_taps(frame, s.G.col, s.G.row);
            };
            int t = (role.pidx + n_prod - (int)(batch_counter % (unsigned)n_prod)) % n_prod;
            unsigned slot, phase;
            {
                const unsigned g0 = batch_counter + (unsigned)t;
                slot = g0 % n_slots;
                phase = (g0 / n_slots) & 1u;
            }
            // stage C: finish + hand over
            auto stage_c = [&](const Stage& s) {
                float J[12], r = 0.f;
                if (s.valid) {
                    point_finish<true>(kf, ec, ec.blk[s.b], inv_norm, s.G, s.T, s.g4, s.dw, J, r);
                    if (write_residuals) {
                        P.residuals[s.idx] = r;
                        if (P.jac_out) {
#pragma unroll
                            for (int k = 0; k < 12; ++k) P.jac_out[(size_t)12 * s.idx + k] = J[k];
                        }
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 12; ++k) J[k] = 0.f;
                }
#ifdef EDS_TIMING
                const long long te0 = clock64();
#endif
                mbar_wait(&sh.empty_bar[slot], phase ^ 1u, 1);
#ifdef EDS_TIMING
                if (lane == 0 && rank == csize - 1) { atomicAdd(&g_timing[18], (unsigned long long)(clock64() - te0)); atomicAdd(&g_timing[19], 1ull); }
#endif
                float4* dst = reinterpret_cast<float4*>(&sh.ring[slot][lane][0]);
                dst[0] = make_float4(J[0], J[1], J[2], J[3]);
                dst[1] = make_float4(J[4], J[5], J[6], J[7]);
                dst[2] = make_float4(J[8], J[9], J[10], J[11]);
                dst[3] = make_float4(r, 0.f, 0.f, 0.f);
                __syncwarp();  // all 32 rows are written: one elected arrival publishes the slot
                if (lane == 0) mbar_arrive(&sh.full_bar[slot]);
                slot += (unsigned)n_prod;
                if (slot >= n_slots) { slot -= n_slots; phase ^= 1u; }
            };
            // one pipeline step: `cur` holds batch t with its gathers in flight, `nxt` receives batch t + n_prod,
            // kp holds the 3-D points of batch t + n_prod and is refilled with those of batch t + 2 n_prod
            auto step = [&](Stage& cur, Stage& nxt, Kp& kp) {
                const Kp kpb = kp;
                int idx_a, b_a;
                const bool valid_a = locate(t + 2 * n_prod, idx_a, b_a);
                kp = kp0;
                if (valid_a) kp = load_kp(kf, idx_a);
                stage_b(nxt, kpb);
                stage_c(cur);
                // the batch after next becomes the next one
                cur.idx = idx_a; cur.b = b_a; cur.valid = valid_a;
                t += n_prod;
            };
            if (t < nb_tot) {
                Stage s0, s1;
                Kp kp;
                s0.valid = locate(t, s0.idx, s0.b);
                stage_b(s0, s0.valid ? load_kp(kf, s0.idx) : kp0);
                s1.valid = locate(t + n_prod, s1.idx, s1.b);
                kp = s1.valid ? load_kp(kf, s1.idx) : kp0;
                for (;;) {
                    step(s0, s1, kp);  // finishes s0, fills s1; s0's (idx, b, valid) now describe batch t + n_prod
                    if (t >= nb_tot) break;
                    step(s1, s0, kp);
                    if (t >= nb_tot) break;
                }
            }
            for (int bb = rank; bb < B; bb += csize) {  // the counters advance as they do for the consumers
                batch_counter += (unsigned)((bb + 1 == B) ? nb_last : nb_reg);
                block_counter++;
            }
            return;
        }
    }
    for (int b = rank; b < kf.B; b += csize) {
        const float* bc = ec.blk[b];
        const int start = b * ne;
        const int n = ne + ((b + 1 == kf.B) ? (kf.N - (b + 1) * ne) : 0);  // Tracker.cpp:178-190
        if constexpr (RES_ONLY) {
            // residual write-back only (Tracker.cpp:223-230): no Jacobian, no reduction, no DSMEM traffic
            for (int i = role.etid; i < n; i += role.n_eval_threads) {
                float r;
                eval_point<false>(kf, ec, bc, frame, inv_norm, start + i, nullptr, r);
                P.residuals[start + i] = r;
            }
            continue;
        }
        const int nb = (n + 31) >> 5;  // batches of this block
        if (role.cidx >= 0) {
            // ---------------- consumers: own the block's sums (warp c the batches j = c mod N_CONS, which
            // fixes the summation order), warp 0 applies the loss and publishes the slot ----------
            float acc[96];
#pragma unroll
            for (int i = 0; i < 96; ++i) acc[i] = 0.f;
            double s_acc = 0.0;  // sum of r^2 in fp64 (the cost decides accept / reject and the tolerances)
            unsigned slot, phase;
            {
                const unsigned g0 = batch_counter + (unsigned)role.cidx;
                slot = g0 % n_slots;
                phase = (g0 / n_slots) & 1u;
            }
            for (int j = role.cidx; j < nb; j += N_CONS, slot += N_CONS) {
                if (slot >= n_slots) { slot -= n_slots; phase ^= 1u; }
#ifdef EDS_TIMING
                const long long tw0 = clock64();
#endif
                mbar_wait(&sh.full_bar[slot], phase, 2);
#ifdef EDS_TIMING
                if (lane == 0 && rank == csize - 1 && role.cidx == 0) { atomicAdd(&g_timing[16], (unsigned long long)(clock64() - tw0)); atomicAdd(&g_timing[17], 1ull); }
#endif
                float v[16];
                ConsRows::load(sh.ring[slot], lane, v);
                __syncwarp();  // every lane has its row in registers: one elected arrival frees the slot
                if (lane == 0) mbar_arrive(&sh.empty_bar[slot]);
                const float rr = ConsRows::accumulate(v, acc);
                s_acc += (double)rr * (double)rr;
            }
            reduce96(acc, lane);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s_acc += __shfl_xor_sync(0xffffffffu, s_acc, o);
            const unsigned buf = block_counter & 1u;
            if (role.cidx != 0) {
#pragma unroll
                for (int i = 0; i < 3; ++i) sh.cons_part[buf][role.cidx - 1][3 * lane + i] = acc[i];
                if (lane == 0) sh.cons_s[buf][role.cidx - 1] = s_acc;
                asm volatile("bar.sync 2, %0;" ::"n"(32 * N_CONS) : "memory");
            } else {
                asm volatile("bar.sync 2, %0;" ::"n"(32 * N_CONS) : "memory");
#pragma unroll
                for (int c = 0; c < N_CONS - 1; ++c) {  // fixed order: the sums do not depend on timing
#pragma unroll
                    for (int i = 0; i < 3; ++i) acc[i] += sh.cons_part[buf][c][3 * lane + i];
                    s_acc += sh.cons_s[buf][c];
                }
                double rho0, rho1;
                loss_eval(P.loss_type, loss_a, s_acc, &rho0, &rho1);
                const int base = reduce96_base(lane);
                double* dst = slot_base + b * NSLOT;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const int e = base + i;
                    if (e < 90) dst[ConsRows::slot(e)] = rho1 * (double)acc[i];
                }
                if (lane == 0) { dst[90] = 0.5 * rho0; dst[91] = s_acc; }
            }
        }
        batch_counter += (unsigned)nb;
        block_counter++;
    }
}

// ------------------------------------------------------------------------------------------
// leader: Levenberg-Marquardt state machine (ceres TrustRegionMinimizer +
// LevenbergMarquardtStrategy semantics; options of Tracker.cpp:117-143)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Leader warp: publish the evaluation constants of point xe (13 doubles in leader smem) and the
// command to every CTA of the cluster.  Lane b derives the per-block model normalisation
// S_b = 1e-3 + v^T A_b v (PhotometricError.hpp:132,148) and c_b = A_b v.
__device__ void leader_publish(cg::cluster_group& cluster, CtaShared& sh, int which, const double* xe, int cmd, int B, int csize,
                               bool signal = true) {
    const int lane = threadIdx.x & 31;
    ProblemShared& ps = sh.prob[which];
    EvalConst& ec = ps.ec;  // build locally, then replicate
    // every lane computes the shared constants (no divergent serial section), lane 0 stores them
    double R[9];
    quat_to_rot(&xe[3], R);
    double v[6], vs = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) { v[k] = xe[7 + k]; vs += v[k] * v[k]; }
    const double ivn = rsqrt(vs), ivs = ivn * ivn;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 9; ++k) ec.R[k] = R[k];
        ec.t[0] = xe[0]; ec.t[1] = xe[1]; ec.t[2] = xe[2];
#pragma unroll
        for (int k = 0; k < 6; ++k) ec.vf[k] = (float)v[k];
        ec.inv_vs = (float)ivs;
        ec.inv_vn = (float)ivn;
        ec.cmd = cmd;
    }
    if (lane < B && cmd != CMD_DONE) {
        // S_b = 1e-3 + v^T A_b v (PhotometricError.hpp:132,148), c_b = A_b v
        const double* A = ps.A[lane];
        double c[6] = {0, 0, 0, 0, 0, 0};
        int k = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = i; j < 6; ++j, ++k) {
                c[i] += A[k] * v[j];
                if (j != i) c[j] += A[k] * v[i];
            }
        double cv = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) cv += v[i] * c[i];
        const double iM = rsqrt(1e-03 + cv);
        const double iM3 = iM * iM * iM;
        const double kappa = (iM - cv * iM3) * ivs;
        ec.blk[lane][0] = (float)iM;
        ec.blk[lane][1] = (float)(iM * ivn);
#pragma unroll
        for (int i = 0; i < 6; ++i) ec.blk[lane][2 + i] = (float)((c[i] * iM3 + kappa * v[i]) * ivn);
    }
    __syncwarp();
    // replicate the used prefix of EvalConst (R,t,v + B block records) and cmd
    const int nwords = (int)(offsetof(EvalConst, blk) / 4) + 8 * B;
    const int* src = reinterpret_cast<const int*>(&ec);
    const int self = (int)cluster.block_rank();
    for (int c = 0; c < csize; ++c) {
        if (c == self) continue;
        EvalConst* dst = &cluster.map_shared_rank(&sh, c)->prob[which].ec;
        int* d = reinterpret_cast<int*>(dst);
        for (int i = lane; i < nwords; i += 32) d[i] = src[i];
        if (lane == 0) dst->cmd = cmd;
    }
    if (signal) {
        // one fence orders all of this lane's stores, then every lane signals every CTA: the 32nd
        // arrival completes the phase there
        fence_cluster();
        for (int c = 0; c < csize; ++c) mbar_arrive_cluster_relaxed(&sh.ready_bar[which], (unsigned)c);
    }
}

// Leader warp (32 lanes, convergent): consume one evaluation and decide what to do next.
// ceres TrustRegionMinimizer + LevenbergMarquardtStrategy semantics (options of
// Tracker.cpp:117-143); returns the next command, lm.cand / lm.x hold the point to evaluate.
__device__ int lm_advance_warp(ProblemShared& sh) {
    const ProblemDesc& P = sh.P;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    LmState& lm = sh.lm;
    const int B = P.kf.B;
#ifdef EDS_TIMING
    unsigned long long tt0 = gtime();
#endif
    // fixed pairwise tree over the residual blocks: deterministic and independent of the cluster size
    bool fin = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int e = lane + 32 * k;
        if (e < 91) {
            // the tree is over MAX_BLOCKS leaves, missing blocks are zeros; with B <= 8 its first level only
            // adds zeros, so the 8-leaf tree gives the same bits with half the loads
            double v[MAX_BLOCKS];
            if (B <= MAX_BLOCKS / 2) {
#pragma unroll
                for (int b = 0; b < MAX_BLOCKS / 2; ++b) v[b] = (b < B) ? sh.slots[b][e] : 0.0;
            } else {
#pragma unroll
                for (int b = 0; b < MAX_BLOCKS; ++b) v[b] = (b < B) ? sh.slots[b][e] : 0.0;
#pragma unroll
                for (int b = 0; b < MAX_BLOCKS / 2; ++b) v[b] += v[b + MAX_BLOCKS / 2];
            }
#pragma unroll
            for (int w = MAX_BLOCKS / 4; w > 0; w >>= 1)
#pragma unroll
                for (int b = 0; b < w; ++b) v[b] += v[b + w];
            sh.sum[e] = v[0];
            fin = fin && isfinite(v[0]);
        }
    }
    fin = __all_sync(FULL, fin);
    __syncwarp();
#ifdef EDS_TIMING
    unsigned long long tt1 = gtime();
#endif
    const double cost = sh.sum[90];
    if (lane == 0) lm.n_eval++;
    double radius = lm.radius, dec = lm.dec;
    int reuse = lm.reuse_diag, n_succ = lm.n_succ, n_unsucc = lm.n_unsucc;
    bool take = false, first = false;
    int ret = 0;
    if (lm.phase == PHASE_INIT) {
        if (lane == 0) lm.initial_cost = cost;
        if (!fin) { if (lane == 0) { lm.termination = EDSGPU_TERM_FAILURE; lm.x_cost = cost; } ret = CMD_DONE; }
        else { take = true; first = true; }
    } else {
        const double cand_cost = isfinite(cost) ? cost : DBL_MAX;
        // ParameterToleranceReached: ||x - cand|| <= ptol (||x|| + ptol)
        double d = (lane < 13) ? lm.x[lane] - lm.cand[lane] : 0.0;
        const double sn = sqrt(warp_sum(d * d));
        const double cost_change = lm.x_cost - cand_cost;
        if (sn <= P.ptol * (lm.x_norm + P.ptol)) { if (lane == 0) lm.termination = EDSGPU_TERM_CONVERGENCE; ret = CMD_FINAL; }
        else if (fabs(cost_change) <= P.ftol * lm.x_cost) { if (lane == 0) lm.termination = EDSGPU_TERM_CONVERGENCE; ret = CMD_FINAL; }  // FunctionToleranceReached
        else {
            const double rel = cost_change / lm.mcc;
            if (rel > 1e-3) {  // min_relative_decrease: HandleSuccessfulStep
                if (!fin) { if (lane == 0) lm.termination = EDSGPU_TERM_FAILURE; ret = CMD_DONE; }
                else {
                    const double t = 2.0 * rel - 1.0;
                    radius = fmin(1e16, radius / fmax(1.0 / 3.0, 1.0 - t * t * t));
                    dec = 2.0; reuse = 0; n_succ++;
                    const double c = (lane < 13) ? lm.cand[lane] : 0.0;
                    __syncwarp();
                    if (lane < 13) lm.x[lane] = c;
                    take = true;
                }
            } else { radius = radius / dec; dec *= 2.0; reuse = 1; n_unsucc++; }
        }
    }
    __syncwarp();
    // row `lane` of the scaled system once a new point has been taken (kept in registers for the solve)
    double Hr[12];
#pragma unroll
    for (int q = 0; q < 12; ++q) Hr[q] = 0.0;
    if (take) {
        // EvaluateGradientAndJacobian: Jacobi scaling (iteration 0 only), scaled system,
        // gradient max norm || x - Plus(x, -g) ||_inf, ||x||
        // The unit-norm retraction makes n = [0(6), v/|v|] an exact null direction of J.
        // The fp32 sweep leaves ~1e-7 of noise along it, which a large trust-region radius
        // (D^2 -> 1e-12) would amplify into the step.  Project the reduced system in fp64:
        // H <- Pi H Pi, g <- Pi g, Pi = I - n n^T, so the device behaves like exact arithmetic.
        // Lane i < 12 holds row i of H and g_i in registers; vectors are exchanged by shuffles, so the
        // whole block needs no shared-memory round trip.
        const int la = lane < 12 ? lane : 0;
#pragma unroll
        for (int q = 0; q < 12; ++q) {
            const double v = sh.sum[la <= q ? tri_index(la, q) : tri_index(q, la)];
            Hr[q] = (lane < 12) ? v : 0.0;
        }
        double g = (lane < 12) ? sh.sum[78 + la] : 0.0;
        const double vq = (lane >= 6 && lane < 12) ? lm.x[1 + lane] : 0.0;
        const double inv_n = rsqrt(warp_sum(vq * vq));
        const double na = vq * inv_n;
        double wa = 0.0;  // (H n)_i
#pragma unroll
        for (int b = 6; b < 12; ++b) wa += Hr[b] * __shfl_sync(FULL, na, b);
        const double sw = warp_sum(na * wa);
        const double gn = warp_sum(na * g);
#pragma unroll
        for (int q = 0; q < 12; ++q) {
            const double nq = __shfl_sync(FULL, na, q), wq = __shfl_sync(FULL, wa, q);
            Hr[q] = Hr[q] - na * wq - wa * nq + na * nq * sw;
        }
        g = g - na * gn;
        double sc;
        if (first) {
            double diag = 0.0;
#pragma unroll
            for (int q = 0; q < 12; ++q) diag = (q == lane) ? Hr[q] : diag;
            sc = 1.0 / (1.0 + sqrt(fmax(diag, 0.0)));
            if (lane < 12) lm.scale[lane] = sc;
        } else {
            sc = lm.scale[la];
        }
#pragma unroll
        for (int q = 0; q < 12; ++q) Hr[q] = Hr[q] * sc * __shfl_sync(FULL, sc, q);
        if (lane < 12) {
            // kept for the re-solves after a rejected step and for the gradient tolerance test
#pragma unroll
            for (int q = 0; q < 12; ++q)
                if (q >= lane) lm.Hs[tri_index(lane, q)] = Hr[q];
            lm.gs[lane] = g * sc;
            sh.sum[78 + lane] = g;
        }
        double gt = (lane < 3) ? fabs(g) : 0.0;
        double gmax = warp_max(gt);  // translation part of x - Plus(x,-g) is exactly g
        if (!(gmax > P.gtol)) {
            // only now can the full projected-gradient norm decide the gradient tolerance test
            __syncwarp();
            double ng[12], xp[13];
            for (int i = 0; i < 12; ++i) ng[i] = -sh.sum[78 + i];
            state_plus(lm.x, ng, xp);
            for (int i = 0; i < 13; ++i) gmax = fmax(gmax, fabs(lm.x[i] - xp[i]));
        }
        const double xv = (lane < 13) ? lm.x[lane] : 0.0;
        const double xn = sqrt(warp_sum(xv * xv));
        if (lane == 0) { lm.x_cost = cost; lm.gmax = gmax; lm.x_norm = xn; }
    }
    __syncwarp();
#ifdef EDS_TIMING
    unsigned long long tt2 = gtime();
#endif
    const double gmax_now = lm.gmax;
    int iter = lm.iter, consec = lm.consec_invalid;
    double mcc = 0.0;
    while (ret == 0) {
        if (iter >= P.max_iter) { if (lane == 0) lm.termination = EDSGPU_TERM_NO_CONVERGENCE; ret = CMD_FINAL; break; }
        if (gmax_now <= P.gtol) { if (lane == 0) lm.termination = EDSGPU_TERM_CONVERGENCE; ret = CMD_FINAL; break; }
        if (radius < 1e-32) { if (lane == 0) lm.termination = EDSGPU_TERM_CONVERGENCE; ret = CMD_FINAL; break; }
        iter++;
        // LevenbergMarquardtStrategy::ComputeStep: D^2 = clamp(diag(JtJ)) / radius, (JtJ + D^2) y = Jt r
        if (!reuse && lane < 12) lm.diag[lane] = fmin(fmax(lm.Hs[tri_index(lane, lane)], 1e-6), 1e32);
        __syncwarp();
        reuse = 1;
        // lane i < 12 owns row i of the damped matrix (a) and of the un-damped one (h)
        double a[12], h[12];
        const int li = lane < 12 ? lane : 0;
        const double damping = lm.diag[li] / radius;  // D^2 of this lane's diagonal entry
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            if (take) {
                h[j] = Hr[j];  // this call has just formed the scaled system
            } else {
                const double v = lm.Hs[li <= j ? tri_index(li, j) : tri_index(j, li)];
                h[j] = (lane < 12) ? v : 0.0;
            }
            a[j] = (j == lane) ? h[j] + damping : h[j];
        }
        // Right-looking Cholesky in registers with the forward substitution L z = gs folded in.  After
        // step k lane i >= k holds L[i][k] in a[k]; lane k keeps its row j > k UNSCALED (L[k][k] L[j][k]),
        // the 1/L[k][k] it owes goes into the back substitution.  Straight-line code: lanes that must not
        // take part in an update multiply by zero instead of branching around it.
        bool ok = true;
        double myinv = 0.0;
        double z = (lane < 12) ? lm.gs[li] : 0.0;
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const double dk = __shfl_sync(FULL, a[k], k);
            const double zs = __shfl_sync(FULL, z, k);
            ok = ok && (dk > 0.0) && (dk < DBL_MAX);
            const double inv = rsqrt(dk);
            const double lik = a[k] * inv;
            const double zk = zs * inv;
            const double lm_ik = (lane > k) ? lik : 0.0;  // masked column: zero on and above the diagonal
            myinv = (lane == k) ? inv : myinv;
            z = (lane == k) ? zk : z - lm_ik * zk;
            a[k] = (lane >= k) ? lik : a[k];
#pragma unroll
            for (int j = k + 1; j < 12; ++j) a[j] -= lm_ik * __shfl_sync(FULL, lik, j);
        }
        // backward L^T y = z: lane i uses L[k][i] = a[k] / L[i][i] for k > i
#pragma unroll
        for (int k = 0; k < 12; ++k) a[k] = (lane < k) ? a[k] * myinv : 0.0;
#pragma unroll
        for (int k = 11; k >= 0; --k) {
            const double yk = __shfl_sync(FULL, z * myinv, k);
            z = (lane == k) ? yk : z - a[k] * yk;
        }
        const double step = (lane < 12) ? -z : 0.0;
        ok = ok && __all_sync(FULL, isfinite(step));
        if (ok) {
            // model_cost_change = -step^T (gs + 0.5 Hs step)
            double hv = 0.0;
#pragma unroll
            for (int j = 0; j < 12; ++j) hv += h[j] * __shfl_sync(FULL, step, j);
            const double gsl = (lane < 12) ? lm.gs[li] : 0.0;
            mcc = -warp_sum(step * (gsl + 0.5 * hv));
            ok = mcc > 0.0;
        }
        if (!ok) {  // HandleInvalidStep
            if (++consec >= 5) { if (lane == 0) lm.termination = EDSGPU_TERM_FAILURE; ret = CMD_DONE; break; }
            radius = radius / dec; dec *= 2.0; n_unsucc++;
            continue;
        }
        consec = 0;
        if (lane < 12) lm.delta[lane] = step * lm.scale[lane];
        __syncwarp();
#ifdef EDS_TIMING
        unsigned long long tt3 = gtime();
#endif
        if (lane == 0) state_plus(lm.x, lm.delta, lm.cand);
        ret = CMD_EVAL;
#ifdef EDS_TIMING
        if (lane == 0) { g_timing[6] += tt1 - tt0; g_timing[7] += tt2 - tt1; g_timing[8] += tt3 - tt2; g_timing[9] += gtime() - tt3; g_timing[10] += 1; }
#endif
    }
    if (lane == 0) {
        lm.radius = radius; lm.dec = dec; lm.reuse_diag = reuse; lm.n_succ = n_succ; lm.n_unsucc = n_unsucc;
        lm.iter = iter; lm.consec_invalid = consec; lm.mcc = mcc; lm.phase = PHASE_CAND;
    }
    __syncwarp();
    return ret;
}

// all threads of the CTA: descriptor, loss parameter and (leader CTA) the Gram matrices + state of one problem
__device__ __forceinline__ void load_problem(ProblemShared& ps, const ProblemDesc* problems, int pid, int count, bool leads) {
    const int tid = threadIdx.x;
    const bool valid = pid < count;
    if (tid == 0) ps.valid = valid;
    if (valid) {
        const int* src = reinterpret_cast<const int*>(&problems[pid]);
        int* dst = reinterpret_cast<int*>(&ps.P);
        for (int i = tid; i < (int)(sizeof(ProblemDesc) / sizeof(int)); i += TRK_THREADS) dst[i] = src[i];
    }
    __syncthreads();
    if (valid) {
        if (tid == 0) ps.loss_a = ps.P.state[13];
        if (leads) {
            for (int i = tid; i < 21 * ps.P.kf.B; i += TRK_THREADS) (&ps.A[0][0])[i] = ps.P.kf.A[i];
            if (tid < 13) ps.x_eval[tid] = ps.P.state[tid];
        }
    }
    if (tid == 0) ps.ec.cmd = valid ? 0 : CMD_DONE;
    __syncthreads();
}

__device__ __forceinline__ void init_barriers(CtaShared& sh, int evaluator_warps) {
    if (threadIdx.x < N_SLOTS) {
        mbar_init(&sh.full_bar[threadIdx.x], 1);   // the elected lane of the producer warp that filled the slot
        mbar_init(&sh.empty_bar[threadIdx.x], 1);  // the elected lane of the consumer warp that drained it
    }
    if (threadIdx.x < MAX_K) {
        mbar_init(&sh.ready_bar[threadIdx.x], 32);                          // the 32 lanes of the problem's leader warp
        mbar_init(&sh.result_bar[threadIdx.x], (unsigned)evaluator_warps);  // one arrival per evaluator warp of the cluster
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // ready/result are signalled from other CTAs
    __syncthreads();
}

// leader warp: reset the LM state of a problem and publish its first evaluation point
__device__ void leader_start(cg::cluster_group& cluster, CtaShared& sh, int which, int csize) {
    ProblemShared& ps = sh.prob[which];
    const int lane = threadIdx.x & 31;
    LmState& lm = ps.lm;
    if (lane < 13) lm.x[lane] = ps.x_eval[lane];
    if (lane == 0) {
        lm.radius = 1e4; lm.dec = 2.0; lm.reuse_diag = 0;
        lm.iter = 0; lm.n_succ = 0; lm.n_unsucc = 0; lm.consec_invalid = 0; lm.n_eval = 0;
        lm.termination = EDSGPU_TERM_NO_CONVERGENCE; lm.phase = PHASE_INIT;
        lm.x_cost = 0.0; lm.initial_cost = 0.0; lm.mcc = 0.0; lm.gmax = 0.0; lm.x_norm = 0.0;
    }
    __syncwarp();
    leader_publish(cluster, sh, which, lm.x, CMD_EVAL, ps.P.kf.B, csize);
}

// leader warp: consume the evaluation of a problem, advance its LM state, publish the next command
__device__ void leader_step(cg::cluster_group& cluster, CtaShared& sh, int which, int csize) {
    ProblemShared& ps = sh.prob[which];
    const int lane = threadIdx.x & 31;
    LmState& lm = ps.lm;
    const ProblemDesc& P = ps.P;
    const int next = lm_advance_warp(ps);
#ifdef EDS_TIMING
    const unsigned long long tp0 = gtime();
#endif
    leader_publish(cluster, sh, which, (next == CMD_EVAL) ? lm.cand : lm.x, next, P.kf.B, csize);
#ifdef EDS_TIMING
    if (lane == 0) g_timing[15] += gtime() - tp0;
#endif
    if (next != CMD_EVAL && lane == 0) {
        const bool usable = lm.termination != EDSGPU_TERM_FAILURE;
        if (usable) for (int i = 0; i < 13; ++i) P.state[i] = lm.x[i];  // Tracker.cpp:217-220
        edsgpu_tracker_info inf;
        inf.iterations = lm.n_succ + lm.n_unsucc;
        inf.successful_steps = lm.n_succ;
        inf.unsuccessful_steps = lm.n_unsucc;
        inf.termination = lm.termination;
        inf.usable = usable ? 1 : 0;
        inf.num_points = P.kf.N;
        inf.evaluations = lm.n_eval;
        inf.reserved = 0;
        inf.initial_cost = lm.initial_cost;
        inf.final_cost = lm.x_cost;
        inf.final_radius = lm.radius;
        *P.info = inf;
    }
}

// One cluster keeps K <= MAX_K independent tracking problems in flight (problems K*c .. K*c+K-1 for
// cluster c).  Problem k is led by the last warp of CTA rank k: a sequential Levenberg-Marquardt loop that
// waits for the reduced sums of an evaluation, takes its decision and publishes the next evaluation
// point to every CTA.  All other warps are evaluators: they visit the live problems round-robin,
// wait until the problem's constants have landed, sweep this CTA's residual blocks and report.
// There is no cluster-wide barrier in the loop, only two mbarrier hand-overs per evaluation
//     ready_bar[k]  (every CTA)    leader k   -> evaluators   constants + command are in your smem
//     result_bar[k] (leader's CTA) evaluators -> leader k     all block sums are in your smem
// so the serial LM step of one problem (~12 us of dependent fp64 latency) runs while the evaluators
// sweep the other K-1 problems.  With K = 1 sweep and LM step simply alternate.
__global__ void __launch_bounds__(TRK_THREADS, 1) track_lm_kernel(const ProblemDesc* __restrict__ problems, int count, int K) {
    cg::cluster_group cluster = cg::this_cluster();
    const int csize = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CtaShared& sh = *reinterpret_cast<CtaShared*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int first = (blockIdx.x / csize) * K;
    for (int k = 0; k < K; ++k) load_problem(sh.prob[k], problems, first + k, count, rank == k);
    const bool hosts_leader = rank < K;
    init_barriers(sh, TRK_WARPS * csize - K);
    const Roles role = make_roles(hosts_leader);
    const bool is_leader_warp = hosts_leader && ((tid >> 5) == LEADER_WARP);
    cluster.sync();  // barriers initialised and problems loaded everywhere before the first remote signal
#ifdef EDS_TIMING
    const unsigned long long t_kernel = gtime();
    unsigned long long t_wait = 0, t_work = 0, n_work = 0;
#endif
    if (is_leader_warp) {
        // ---------------- leader of problem `rank` ----------------
        if (sh.prob[rank].valid) {
            leader_start(cluster, sh, rank, csize);
            unsigned parity = 0;
            for (;;) {
#ifdef EDS_TIMING
                const unsigned long long t0 = gtime();
#endif
                mbar_wait_cluster(&sh.result_bar[rank], parity, 4);
                parity ^= 1u;
#ifdef EDS_TIMING
                const unsigned long long t1 = gtime();
#endif
                leader_step(cluster, sh, rank, csize);
#ifdef EDS_TIMING
                t_wait += t1 - t0; t_work += gtime() - t1; n_work++;
#endif
                if (sh.prob[rank].ec.cmd != CMD_EVAL) break;
            }
        }
#ifdef EDS_TIMING
        if (lane == 0) { atomicAdd(&g_timing[0], t_work); atomicAdd(&g_timing[1], t_wait); atomicAdd(&g_timing[2], n_work); }
#endif
    } else {
        // ---------------- evaluators ----------------
        unsigned batch_counter = 0, block_counter = 0;  // same in every evaluator thread of the CTA
        unsigned live = 0, parity = 0;
        for (int k = 0; k < K; ++k) live |= sh.prob[k].valid ? (1u << k) : 0u;
        while (live) {
            for (int k = 0; k < K; ++k) {
                if (!((live >> k) & 1u)) continue;
                ProblemShared& ps = sh.prob[k];
#ifdef EDS_TIMING
                const unsigned long long t0 = gtime();
#endif
                mbar_wait_cluster(&sh.ready_bar[k], (parity >> k) & 1u, 3);
                parity ^= 1u << k;
#ifdef EDS_TIMING
                const unsigned long long t1 = gtime();
#endif
                const int cmd = ps.ec.cmd;
                if (cmd == CMD_EVAL) {
                    double* slots = &cluster.map_shared_rank(&sh, k)->prob[k].slots[0][0];
                    if (hosts_leader) cta_evaluate<false, true>(ps, sh, slots, rank, csize, role, false, batch_counter, block_counter);
                    else cta_evaluate<false, false>(ps, sh, slots, rank, csize, role, false, batch_counter, block_counter);
                    // the consumer's block sums went over DSMEM: order them before the signal
                    if (role.cidx == 0) fence_cluster();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(&sh.result_bar[k], (unsigned)k);
                } else {
                    if (cmd == CMD_FINAL) cta_evaluate<true, false>(ps, sh, nullptr, rank, csize, role, true, batch_counter, block_counter);
                    live &= ~(1u << k);
                }
#ifdef EDS_TIMING
                t_wait += t1 - t0; t_work += gtime() - t1; n_work++;
#endif
            }
        }
#ifdef EDS_TIMING
        if (lane == 0 && rank == csize - 1) {
            if (role.cidx == 0) { atomicAdd(&g_timing[3], t_work); atomicAdd(&g_timing[4], t_wait); atomicAdd(&g_timing[5], n_work); }
            else if (tid == 0) { atomicAdd(&g_timing[13], t_work); atomicAdd(&g_timing[14], t_wait); }
            else if (role.pidx == role.n_prod - 1) { atomicAdd(&g_timing[20], t_work); }
            else if (role.pidx == 6) { atomicAdd(&g_timing[21], t_work); }
            else if (role.cidx == 1) { atomicAdd(&g_timing[22], t_work); }
        }
#endif
    }
    cluster.sync();  // nobody leaves while its shared memory may still be written or signalled
#ifdef EDS_TIMING
    if (rank == 0 && tid == 0) { atomicAdd(&g_timing[11], gtime() - t_kernel); atomicAdd(&g_timing[12], 1ull); }
#endif
}

// parity/debug entry (edsgpu_tracker_evaluate): one full evaluation at P.state, residuals and
// Jacobian rows written out, reduced normal equations returned.
__global__ void __launch_bounds__(TRK_THREADS, 1) track_eval_kernel(const ProblemDesc* __restrict__ problems, int count) {
    cg::cluster_group cluster = cg::this_cluster();
    const int csize = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CtaShared& sh = *reinterpret_cast<CtaShared*>(smem_raw);
    ProblemShared& ps = sh.prob[0];
    load_problem(ps, problems, blockIdx.x / csize, count, rank == 0);
    init_barriers(sh, 1);
    CtaShared* leader = cluster.map_shared_rank(&sh, 0);
    const Roles role = make_roles(rank == 0);
    const bool is_leader_warp = (rank == 0) && ((threadIdx.x >> 5) == LEADER_WARP);
    if (is_leader_warp) leader_publish(cluster, sh, 0, ps.x_eval, CMD_EVAL, ps.P.kf.B, csize, false);
    cluster.sync();
    unsigned batch_counter = 0, block_counter = 0;
    if (!is_leader_warp) {
        if (rank == 0) cta_evaluate<false, true>(ps, sh, &leader->prob[0].slots[0][0], rank, csize, role, true, batch_counter, block_counter);
        else cta_evaluate<false, false>(ps, sh, &leader->prob[0].slots[0][0], rank, csize, role, true, batch_counter, block_counter);
    }
    cluster.sync();
    const ProblemDesc& P = ps.P;
    if (rank == 0 && threadIdx.x == 0 && P.eval_out) {
        double cost = 0.0;
        for (int b = 0; b < P.kf.B; ++b) cost += ps.slots[b][90];
        P.eval_out[0] = cost;
        for (int a = 0; a < 12; ++a) {
            double gs = 0.0;
            for (int b = 0; b < P.kf.B; ++b) gs += ps.slots[b][78 + a];
            P.eval_out[1 + 144 + a] = gs;
            for (int c2 = a; c2 < 12; ++c2) {
                double h = 0.0;
                for (int b = 0; b < P.kf.B; ++b) h += ps.slots[b][tri_index(a, c2)];
                P.eval_out[1 + 12 * a + c2] = h;
                P.eval_out[1 + 12 * c2 + a] = h;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// next loss parameter (Tracker.cpp:281-317): MAD via exact radix select on fp32 keys
// ------------------------------------------------------------------------------------------
constexpr int MAD_THREADS = 1024;

__device__ __forceinline__ unsigned f2key(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// k-th smallest (0-based) of f(i), i < n.  All threads of the CTA must call.
template <typename F>
__device__ float select_kth(F f, int n, int k, unsigned* hist /*256*/, unsigned* bcast /*2*/) {
    unsigned prefix = 0, mask = 0;
    int kk = k;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        if (threadIdx.x < 256) hist[threadIdx.x] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned key = f2key(f(i));
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            // warp scan over 256 bins, 8 per lane
            unsigned local[8], sum = 0;
            for (int j = 0; j < 8; ++j) { local[j] = hist[threadIdx.x * 8 + j]; sum += local[j]; }
            unsigned incl = sum;
            for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)threadIdx.x >= o) incl += t; }
            unsigned excl = incl - sum;
            if ((unsigned)kk >= excl && (unsigned)kk < incl) {
                unsigned run = excl;
                for (int j = 0; j < 8; ++j) {
                    if ((unsigned)kk < run + local[j]) { bcast[0] = threadIdx.x * 8 + j; bcast[1] = run; break; }
                    run += local[j];
                }
            }
        }
        __syncthreads();
        prefix |= bcast[0] << shift;
        mask |= 255u << shift;
        kk -= (int)bcast[1];
        __syncthreads();
    }
    return key2f(prefix);
}

// Same selection with the values held in registers (n <= PER_THREAD * blockDim.x): the residuals are read once
// for both selections instead of once per radix pass.
constexpr int MAD_PER_THREAD = 16;
template <typename F>
__device__ float select_kth_reg(const float* vals, int cnt, F f, int k, unsigned* hist, unsigned* bcast) {
    unsigned prefix = 0, mask = 0;
    int kk = k;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        if (threadIdx.x < 256) hist[threadIdx.x] = 0;
        __syncthreads();
#pragma unroll
        for (int q = 0; q < MAD_PER_THREAD; ++q) {
            if (q < cnt) {
                const unsigned key = f2key(f(vals[q]));
                if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned local[8], sum = 0;
            for (int j = 0; j < 8; ++j) { local[j] = hist[threadIdx.x * 8 + j]; sum += local[j]; }
            unsigned incl = sum;
            for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)threadIdx.x >= o) incl += t; }
            unsigned excl = incl - sum;
            if ((unsigned)kk >= excl && (unsigned)kk < incl) {
                unsigned run = excl;
                for (int j = 0; j < 8; ++j) {
                    if ((unsigned)kk < run + local[j]) { bcast[0] = threadIdx.x * 8 + j; bcast[1] = run; break; }
                    run += local[j];
                }
            }
        }
        __syncthreads();
        prefix |= bcast[0] << shift;
        mask |= 255u << shift;
        kk -= (int)bcast[1];
        __syncthreads();
    }
    return key2f(prefix);
}

__global__ void __launch_bounds__(MAD_THREADS) mad_kernel(const ProblemDesc* __restrict__ problems) {
    const ProblemDesc& P = problems[blockIdx.x];
    __shared__ unsigned hist[256];
    __shared__ unsigned bcast[2];
    __shared__ double red[MAD_THREADS / 32];
    if (P.loss_param_method == EDSGPU_LOSS_PARAM_CONSTANT) return;
    if (!P.info->usable) return;  // Tracker.cpp:217: only after a usable solve
    const int n = P.kf.N;
    const float* r = P.residuals;
    if (P.loss_param_method == EDSGPU_LOSS_PARAM_MAD && n <= MAD_PER_THREAD * MAD_THREADS) {
        // n_quantile_vector(v, size/2) == element size/2 of the sorted vector (Utils.hpp:315-320); residuals in registers
        float vals[MAD_PER_THREAD];
        int cnt = 0;
#pragma unroll
        for (int q = 0; q < MAD_PER_THREAD; ++q) {
            const int i = threadIdx.x + q * MAD_THREADS;
            vals[q] = (i < n) ? r[i] : 0.f;
            cnt += (i < n) ? 1 : 0;
        }
        const float med = select_kth_reg(vals, cnt, [](float v) { return v; }, n / 2, hist, bcast);
        const float mad = select_kth_reg(vals, cnt, [med](float v) { return fabsf(v - med); }, n / 2, hist, bcast);
        if (threadIdx.x == 0) P.state[13] = 1.345 * (1.4826 * (double)mad);
    } else if (P.loss_param_method == EDSGPU_LOSS_PARAM_MAD) {
        const float med = select_kth([r](int i) { return r[i]; }, n, n / 2, hist, bcast);
        const float mad = select_kth([r, med](int i) { return fabsf(r[i] - med); }, n, n / 2, hist, bcast);
        if (threadIdx.x == 0) P.state[13] = 1.345 * (1.4826 * (double)mad);
    } else {
        // mean_std_vector (Utils.hpp:272-290): note the reference returns the VARIANCE as "std_dev"
        auto block_sum = [&](double v) {
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            __syncthreads();
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
            __syncthreads();
            double s = 0.0;
            for (int w = 0; w < MAD_THREADS / 32; ++w) s += red[w];
            return s;
        };
        double s = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)r[i];
        const double mu = block_sum(s) / (double)n;
        double q = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) { double d = (double)r[i] - mu; q += d * d / (double)(n - 1); }
        const double var = (n == 1) ? 0.0 : block_sum(q);
        if (threadIdx.x == 0) P.state[13] = 1.345 * var;
    }
}

// ------------------------------------------------------------------------------------------
// Tracker::getCoord (Tracker.cpp:319-376): the key-frame points at the filter's current inverse depths,
// warped with the tracker's pose into the event frame.  One thread per point, fp64.
// ------------------------------------------------------------------------------------------
__global__ void get_coord_kernel(int N, const double* __restrict__ norm_xy, const double* __restrict__ filter_state,
                                 const double* __restrict__ tracker_state, double fx, double fy, double cx, double cy, int W, int H,
                                 double* __restrict__ coord, unsigned char* __restrict__ outlier) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double R[9];
    quat_to_rot(tracker_state + 3, R);
    double p[3];
    p[2] = 1.0 / filter_state[4 * (size_t)i];  // eds::mapping::mu
    p[0] = norm_xy[2 * (size_t)i] * p[2];
    p[1] = norm_xy[2 * (size_t)i + 1] * p[2];
    double q[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) q[r] = R[3 * r] * p[0] + R[3 * r + 1] * p[1] + R[3 * r + 2] * p[2] + tracker_state[r];
    const double xp = fx * (q[0] / q[2]) + cx, yp = fy * (q[1] / q[2]) + cy;
    coord[2 * (size_t)i] = xp;
    coord[2 * (size_t)i + 1] = yp;
    if (outlier) outlier[i] = ((xp < 0.0 || xp > (double)W) || (yp < 0.0 || yp > (double)H)) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// keyframe upload
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kf_prepare_kernel(const double* __restrict__ grad_xy, const double* __restrict__ norm_xy,
                                                         const double* __restrict__ idp, int idp_stride, const double* __restrict__ weights, int N, int B,
                                                         float4* __restrict__ gxy, float2* __restrict__ dw, double* __restrict__ kpx,
                                                         double* __restrict__ kpy, double* __restrict__ kpz, double* __restrict__ A) {
    const int b = blockIdx.x;
    const int ne = N / B;
    const int start = b * ne;
    const int n = ne + ((b + 1 == B) ? (N - (b + 1) * ne) : 0);
    double a[21];
    for (int k = 0; k < 21; ++k) a[k] = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int idx = start + i;
        const double Gx = grad_xy[2 * idx], Gy = grad_xy[2 * idx + 1];
        const double X = norm_xy[2 * idx], Y = norm_xy[2 * idx + 1];
        const double d = idp[(size_t)idp_stride * idx], w = weights[idx];
        gxy[idx] = make_float4((float)Gx, (float)Gy, (float)X, (float)Y);
        dw[idx] = make_float2((float)d, (float)w);
        const double z = 1.0 / (d + kEps);  // PhotometricError.hpp:97-99
        kpz[idx] = z;
        kpx[idx] = X * z;
        kpy[idx] = Y * z;
        double g[6];
        g[0] = Gx * d;
        g[1] = Gy * d;
        g[2] = -(Gx * X + Gy * Y) * d;
        g[3] = -(Gx * X * Y + Gy * (1.0 + Y * Y));
        g[4] = Gx * (1.0 + X * X) + Gy * X * Y;
        g[5] = Gy * X - Gx * Y;
        int k = 0;
        for (int p = 0; p < 6; ++p)
            for (int q = p; q < 6; ++q, ++k) a[k] += g[p] * g[q];
    }
    __shared__ double red[8][21];
    for (int k = 0; k < 21; ++k) {
        double v = a[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 21) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
        A[21 * b + threadIdx.x] = s;
    }
}

// 14-double state record (px qx vx tau) of every problem of a batch, contiguous
__global__ void pack_states_kernel(const ProblemDesc* __restrict__ problems, int count, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 14 * count) out[i] = problems[i / 14].state[i % 14];
}

__global__ void float_to_double_kernel(const float* __restrict__ in, double* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)in[i];
}

}  // namespace

// ==========================================================================================
// host side
// ==========================================================================================
struct edsgpu_keyframe {
    edsgpu_ctx* ctx = nullptr;
    uint64_t uid = edsgpu_next_uid();
    KfDev dev{};
    void* block = nullptr;  // one allocation behind all device arrays
    double* src = nullptr;  // device copies of the caller's double arrays: grad_xy (2N) | norm_xy (2N) | idp (N) | weights (N)
};

struct edsgpu_tracker {
    edsgpu_ctx* ctx = nullptr;
    edsgpu_tracker_config cfg{};
    double* state = nullptr;            // device: 14 doubles
    edsgpu_tracker_info* info = nullptr; // device
    float* residuals = nullptr;         // device: kf->residuals of the last optimize (Tracker.cpp:223-230)
    int res_capacity = 0;
    uint64_t res_generation = 0;        // bumped when `residuals` is reallocated: batches holding the old pointer re-patch themselves
    // one-problem batch cached for repeated optimize() calls against the same keyframe / frame slot
    struct edsgpu_batch* cached = nullptr;
    uint64_t cached_kf = 0, cached_frames = 0;
    int cached_slot = -1;
};

struct LaunchShape { int csize, K, nclusters; };

struct edsgpu_batch {
    edsgpu_ctx* ctx = nullptr;
    int count = 0;
    LaunchShape shape{1, 1, 0};
    const edsgpu_frames* frames = nullptr;  // slots first_slot .. first_slot + count - 1
    int first_slot = 0;
    ProblemDesc* desc = nullptr;  // device
    std::vector<edsgpu_tracker*> trackers;
    std::vector<const edsgpu_keyframe*> keyframes;  // must outlive the batch (the descriptors hold their device arrays)
    std::vector<uint64_t> res_generation;           // trackers[i]->res_generation the descriptors were made with
};

namespace {

void cluster_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, cudaStream_t stream, int nclusters, int csize) {
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3(nclusters * csize);
    cfg.blockDim = dim3(TRK_THREADS);
    cfg.dynamicSmemBytes = sizeof(CtaShared);
    cfg.stream = stream;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
}

// clusters of c = 1 << k CTAs the device holds at once (cached per context)
int resident_clusters(edsgpu_ctx* ctx, int k) {
    if (ctx->lm_clusters[k] == 0) {
        const int c = 1 << k;
        cudaLaunchConfig_t cfg;
        cudaLaunchAttribute attr[1];
        cluster_config(cfg, attr, ctx->stream, 1, c);
        int n = 0;
        cudaFuncSetAttribute(track_lm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CtaShared));
        if (cudaOccupancyMaxActiveClusters(&n, track_lm_kernel, &cfg) != cudaSuccess || n <= 0) {
            cudaGetLastError();
            n = ctx->num_sms / c;  // one 512-thread CTA per SM
        }
        ctx->lm_clusters[k] = n;
        if (getenv("EDSGPU_VERBOSE")) fprintf(stderr, "[edsgpu] device holds %d clusters of %d tracker CTAs\n", n, c);
    }
    return ctx->lm_clusters[k];
}

// Shape of a batched launch: CTAs per cluster (<= residual blocks) and problems K kept in flight per
// cluster.  Wide clusters first; K grows only as far as needed to get down to about one CTA per SM
// (more problems per cluster hide the serial LM steps, more clusters use more SMs); everything must
// be resident at once, a second wave costs more than narrower clusters do.
LaunchShape pick_shape(edsgpu_ctx* ctx, int count, int B) {
    int force_c = 0, force_k = 0;
    if (const char* e = getenv("EDSGPU_CLUSTER")) {  // tuning/debug overrides
        const int v = atoi(e);
        if (v == 1 || v == 2 || v == 4 || v == 8) force_c = v;
    }
    if (const char* e = getenv("EDSGPU_INFLIGHT")) {
        const int v = atoi(e);
        if (v >= 1 && v <= MAX_K) force_k = v;
    }
    for (int k = 3; k >= 0; --k) {
        const int c = 1 << k;
        if (force_c ? (c != force_c) : (c > MAX_CLUSTER || c > B)) continue;
        int K = (int)(((size_t)count * c + ctx->num_sms - 1) / ctx->num_sms);
        K = std::max(1, std::min(K, std::min(MAX_K, c)));
        if (force_k) K = std::min(force_k, c);
        const int n = (count + K - 1) / K;
        if (force_c || k == 0 || n <= resident_clusters(ctx, k)) return LaunchShape{c, K, n};
    }
    return LaunchShape{1, 1, count};
}

template <typename K, typename... Args>
edsgpu_status launch_cluster(edsgpu_ctx* ctx, K kernel, int nclusters, int csize, Args... args) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    cluster_config(cfg, attr, ctx->stream, nclusters, csize);
    // per device, cheap: opt in to > 48 KB of dynamic shared memory
    EDS_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CtaShared)));
    if (getenv("EDSGPU_VERBOSE")) {  // debug aid: how many clusters of this shape the device holds at once
        int max_clusters = 0;
        cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &cfg);
        fprintf(stderr, "[edsgpu] launch: %d clusters of %d CTAs; device holds %d such clusters\n", nclusters, csize, max_clusters);
    }
    EDS_CUDA(ctx, cudaLaunchKernelEx(&cfg, kernel, args...));
    ctx->launches++;
    return EDSGPU_OK;
}

ProblemDesc make_desc(const edsgpu_tracker* tr, const edsgpu_keyframe* kf, const edsgpu_frames* frames, int slot) {
    ProblemDesc d{};
    d.kf = kf->dev;
    d.frame = frames->tex[slot];
    d.norms = frames->norms + 2 * slot;
    d.state = tr->state;
    d.residuals = tr->residuals;
    d.info = tr->info;
    d.loss_type = tr->cfg.loss_type;
    d.max_iter = tr->cfg.max_iterations;
    d.loss_param_method = tr->cfg.loss_param_method;
    d.eval_only = 0;
    d.ftol = tr->cfg.function_tolerance;
    d.gtol = tr->cfg.gradient_tolerance;
    d.ptol = tr->cfg.parameter_tolerance;
    d.jac_out = nullptr;
    d.eval_out = nullptr;
    return d;
}

// (re)build the device descriptors of a batch from its trackers' current buffers
edsgpu_status upload_descriptors(edsgpu_batch* b) {
    edsgpu_ctx* ctx = b->ctx;
    std::vector<ProblemDesc> hd(b->count);
    b->res_generation.resize(b->count);
    for (int i = 0; i < b->count; ++i) {
        EDS_REQUIRE(ctx, b->trackers[i]->res_capacity >= b->keyframes[i]->dev.N, "batch: a tracker's residual buffer is smaller than its key frame");
        hd[i] = make_desc(b->trackers[i], b->keyframes[i], b->frames, b->first_slot + i);
        b->res_generation[i] = b->trackers[i]->res_generation;
    }
    EDS_CUDA(ctx, cudaMemcpyAsync(b->desc, hd.data(), sizeof(ProblemDesc) * (size_t)b->count, cudaMemcpyHostToDevice, ctx->stream));
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // hd is a stack-owned source
    return EDSGPU_OK;
}

edsgpu_status check_pair(edsgpu_ctx* ctx, const edsgpu_tracker* tr, const edsgpu_keyframe* kf, const edsgpu_frames* frames, int slot) {
    EDS_REQUIRE(ctx, tr && kf && frames, "tracker: null handle");
    EDS_REQUIRE(ctx, tr->ctx == ctx && kf->ctx == ctx && frames->ctx == ctx, "tracker: handles belong to different contexts");
    EDS_REQUIRE(ctx, slot >= 0 && slot < frames->capacity, "tracker: frame slot out of range");
    EDS_REQUIRE(ctx, frames->H == kf->dev.H && frames->W == kf->dev.W, "tracker: keyframe and event frame sizes differ");
    EDS_REQUIRE(ctx, tr->cfg.num_blocks == kf->dev.B, "tracker: config.num_blocks differs from the keyframe's block partition");
    return EDSGPU_OK;
}

}  // namespace

extern "C" {

edsgpu_status edsgpu_keyframe_create(edsgpu_ctx* ctx, int num_points, const double* grad_xy, const double* norm_xy, const double* idp,
                                     const double* weights, int height, int width, double fx, double fy, double cx, double cy, int num_blocks,
                                     edsgpu_keyframe** out) {
    if (!ctx || !out) return EDSGPU_INVALID_ARGUMENT;
    EDS_REQUIRE(ctx, grad_xy && norm_xy && idp && weights, "keyframe_create: null array");
    EDS_REQUIRE(ctx, height > 0 && width > 0, "keyframe_create: bad image size");
    EDS_REQUIRE(ctx, num_blocks >= 1 && num_blocks <= MAX_BLOCKS, "keyframe_create: num_blocks must be in [1,16]");
    EDS_REQUIRE(ctx, num_points >= num_blocks, "keyframe_create: fewer points than residual blocks");
    EDS_REQUIRE(ctx, num_points <= (1 << 20), "keyframe_create: more than 2^20 points");
    DeviceGuard g(ctx->device);
    const size_t N = (size_t)num_points;
    edsgpu_keyframe* kf = new edsgpu_keyframe();
    kf->ctx = ctx;
    // layout: gxy | dw | kpx | kpy | kpz | A
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
    const size_t o_gxy = take(N * sizeof(float4)), o_dw = take(N * sizeof(float2)), o_kx = take(N * 8), o_ky = take(N * 8), o_kz = take(N * 8);
    const size_t o_A = take((size_t)num_blocks * 21 * 8);
    const size_t o_src = take(N * 6 * sizeof(double));  // kept: the inverse depths can be refreshed on the device
    cudaError_t e = cudaMalloc(&kf->block, off);
    if (e != cudaSuccess) { delete kf; return edsgpu_fail(ctx, EDSGPU_OUT_OF_MEMORY, cudaGetErrorString(e)); }
    char* base = (char*)kf->block;
    KfDev& d = kf->dev;
    d.gxy = (float4*)(base + o_gxy); d.dw = (float2*)(base + o_dw);
    d.kpx = (double*)(base + o_kx); d.kpy = (double*)(base + o_ky); d.kpz = (double*)(base + o_kz);
    d.A = (double*)(base + o_A);
    kf->src = (double*)(base + o_src);
    d.N = num_points; d.B = num_blocks; d.H = height; d.W = width;
    d.ne = num_points / num_blocks;
    d.fx = fx; d.fy = fy; d.cx = cx; d.cy = cy;
    // stage the double arrays: pinned -> device copy -> prepare kernel
    const size_t stage = N * 6 * sizeof(double);
    edsgpu_status st = edsgpu_ensure_pinned(ctx, stage);
    if (st != EDSGPU_OK) { edsgpu_keyframe_destroy(kf); return st; }
    double* hp = (double*)ctx->pinned;
    memcpy(hp, grad_xy, N * 16);
    memcpy(hp + 2 * N, norm_xy, N * 16);
    memcpy(hp + 4 * N, idp, N * 8);
    memcpy(hp + 5 * N, weights, N * 8);
    double* ds = kf->src;
    e = cudaMemcpyAsync(ds, hp, stage, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        kf_prepare_kernel<<<num_blocks, 256, 0, ctx->stream>>>(ds, ds + 2 * N, ds + 4 * N, 1, ds + 5 * N, num_points, num_blocks,
                                                                (float4*)d.gxy, (float2*)d.dw, (double*)d.kpx, (double*)d.kpy, (double*)d.kpz,
                                                                (double*)d.A);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // the pinned block is reused by later calls
    if (e != cudaSuccess) { edsgpu_keyframe_destroy(kf); return edsgpu_fail(ctx, EDSGPU_CUDA_ERROR, cudaGetErrorString(e)); }
    *out = kf;
    return EDSGPU_OK;
}

void edsgpu_keyframe_destroy(edsgpu_keyframe* kf) {
    if (!kf) return;
    DeviceGuard g(kf->ctx->device);
    cudaStreamSynchronize(kf->ctx->stream);
    if (kf->block) cudaFree(kf->block);
    delete kf;
}

edsgpu_status edsgpu_tracker_create(edsgpu_ctx* ctx, const edsgpu_tracker_config* config, double loss_param, edsgpu_tracker** out) {
    if (!ctx || !out || !config) return EDSGPU_INVALID_ARGUMENT;
    EDS_REQUIRE(ctx, config->num_blocks >= 1 && config->num_blocks <= MAX_BLOCKS, "tracker_create: num_blocks must be in [1,16]");
    EDS_REQUIRE(ctx, config->loss_type >= 0 && config->loss_type <= 2, "tracker_create: unknown loss type");
    EDS_REQUIRE(ctx, config->loss_param_method >= 0 && config->loss_param_method <= 2, "tracker_create: unknown loss-parameter method");
    EDS_REQUIRE(ctx, config->max_iterations >= 0, "tracker_create: negative max_iterations");
    DeviceGuard g(ctx->device);
    edsgpu_tracker* tr = new edsgpu_tracker();
    tr->ctx = ctx;
    tr->cfg = *config;
    cudaError_t e = cudaMalloc(&tr->state, 14 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&tr->info, sizeof(edsgpu_tracker_info));
    if (e == cudaSuccess) e = cudaMemsetAsync(tr->info, 0, sizeof(edsgpu_tracker_info), ctx->stream);
    if (e != cudaSuccess) { edsgpu_tracker_destroy(tr); return edsgpu_fail(ctx, EDSGPU_CUDA_ERROR, cudaGetErrorString(e)); }
    // Tracker::Tracker(config), Tracker.cpp:41-48
    const double v = 0.001 / sqrt(6.0 * 0.001 * 0.001);
    double init[14] = {0, 0, 0, 0, 0, 0, 1, v, v, v, v, v, v, loss_param};
    e = cudaMemcpyAsync(tr->state, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { edsgpu_tracker_destroy(tr); return edsgpu_fail(ctx, EDSGPU_CUDA_ERROR, cudaGetErrorString(e)); }
    *out = tr;
    return EDSGPU_OK;
}

void edsgpu_tracker_destroy(edsgpu_tracker* tr) {
    if (!tr) return;
    DeviceGuard g(tr->ctx->device);
    cudaStreamSynchronize(tr->ctx->stream);
    if (tr->state) cudaFree(tr->state);
    if (tr->info) cudaFree(tr->info);
    if (tr->cached) edsgpu_batch_destroy(tr->cached);
    if (tr->residuals) cudaFree(tr->residuals);
    delete tr;
}

edsgpu_status edsgpu_tracker_set_state(edsgpu_tracker* tr, const double px[3], const double qx[4], const double vx[6], const double* loss_param) {
    if (!tr) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = tr->ctx;
    DeviceGuard g(ctx->device);
    // small synchronous copies: the source arrays are the caller's stack/heap
    if (px) EDS_CUDA(ctx, cudaMemcpyAsync(tr->state, px, 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (qx) EDS_CUDA(ctx, cudaMemcpyAsync(tr->state + 3, qx, 4 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (vx) EDS_CUDA(ctx, cudaMemcpyAsync(tr->state + 7, vx, 6 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (loss_param) EDS_CUDA(ctx, cudaMemcpyAsync(tr->state + 13, loss_param, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EDSGPU_OK;
}

edsgpu_status edsgpu_tracker_get_state(edsgpu_tracker* tr, double px[3], double qx[4], double vx[6], double* loss_param, edsgpu_tracker_info* info) {
    if (!tr) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = tr->ctx;
    DeviceGuard g(ctx->device);
    double s[14];
    EDS_CUDA(ctx, cudaMemcpyAsync(s, tr->state, sizeof(s), cudaMemcpyDeviceToHost, ctx->stream));
    if (info) EDS_CUDA(ctx, cudaMemcpyAsync(info, tr->info, sizeof(*info), cudaMemcpyDeviceToHost, ctx->stream));
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (px) memcpy(px, s, 3 * sizeof(double));
    if (qx) memcpy(qx, s + 3, 4 * sizeof(double));
    if (vx) memcpy(vx, s + 7, 6 * sizeof(double));
    if (loss_param) *loss_param = s[13];
    return EDSGPU_OK;
}

void* edsgpu_tracker_state_dev(edsgpu_tracker* tr) { return tr ? (void*)tr->state : nullptr; }

#ifdef EDS_TIMING
void edsgpu_debug_timing(unsigned long long* out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_timing, sizeof(unsigned long long) * 32);
    if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(g_timing, z, sizeof(z)); }
}
#endif

edsgpu_status edsgpu_batch_create(edsgpu_ctx* ctx, edsgpu_tracker* const* trackers, const edsgpu_keyframe* const* keyframes, int count,
                                  const edsgpu_frames* frames, int first_slot, edsgpu_batch** out) {
    if (!ctx || !out) return EDSGPU_INVALID_ARGUMENT;
    EDS_REQUIRE(ctx, trackers && keyframes && count > 0, "batch_create: bad arguments");
    DeviceGuard g(ctx->device);
    int B = 0;
    for (int i = 0; i < count; ++i) {
        edsgpu_status st = check_pair(ctx, trackers[i], keyframes[i], frames, first_slot + i);
        if (st != EDSGPU_OK) return st;
        if (i == 0) B = keyframes[i]->dev.B;
        EDS_REQUIRE(ctx, keyframes[i]->dev.B == B, "batch_create: all problems of a batch must share num_blocks");
        for (int j = 0; j < i; ++j) EDS_REQUIRE(ctx, trackers[j] != trackers[i], "batch_create: a tracker appears twice");
        edsgpu_tracker* tr = trackers[i];
        if (tr->res_capacity < keyframes[i]->dev.N) {  // residuals live with the tracker: keyframes may be shared
            EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (tr->residuals) cudaFree(tr->residuals);
            tr->residuals = nullptr;
            tr->res_capacity = 0;
            EDS_CUDA(ctx, cudaMalloc(&tr->residuals, sizeof(float) * (size_t)keyframes[i]->dev.N));
            tr->res_capacity = keyframes[i]->dev.N;
            tr->res_generation++;  // older batches of this tracker hold the freed pointer: they re-patch before their next launch
        }
    }
    edsgpu_batch* b = new edsgpu_batch();
    b->ctx = ctx;
    b->count = count;
    b->shape = pick_shape(ctx, count, B);
    b->frames = frames;
    b->first_slot = first_slot;
    b->trackers.assign(trackers, trackers + count);
    b->keyframes.assign(keyframes, keyframes + count);
    cudaError_t e = cudaMalloc(&b->desc, sizeof(ProblemDesc) * (size_t)count);
    if (e != cudaSuccess) { edsgpu_batch_destroy(b); return edsgpu_fail(ctx, EDSGPU_CUDA_ERROR, cudaGetErrorString(e)); }
    edsgpu_status st = upload_descriptors(b);
    if (st != EDSGPU_OK) { edsgpu_batch_destroy(b); return st; }
    *out = b;
    return EDSGPU_OK;
}

void edsgpu_batch_destroy(edsgpu_batch* b) {
    if (!b) return;
    DeviceGuard g(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    if (b->desc) cudaFree(b->desc);
    delete b;
}

edsgpu_status edsgpu_batch_optimize(edsgpu_batch* b) {
    if (!b) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = b->ctx;
    DeviceGuard g(ctx->device);
    // a tracker of this batch was since paired with a larger key frame (its residual buffer moved): refresh the descriptors
    for (int i = 0; i < b->count; ++i)
        if (b->trackers[i]->res_generation != b->res_generation[i]) {
            edsgpu_status stp = upload_descriptors(b);
            if (stp != EDSGPU_OK) return stp;
            break;
        }
    // the event frames are built on their own stream: wait for the builds of our slots only
    edsgpu_status st = edsgpu_frames_wait_built(b->frames, b->first_slot, b->count, ctx->stream);
    if (st != EDSGPU_OK) return st;
    st = launch_cluster(ctx, track_lm_kernel, b->shape.nclusters, b->shape.csize, (const ProblemDesc*)b->desc, b->count, b->shape.K);
    if (st != EDSGPU_OK) return st;
    st = edsgpu_frames_mark_read(b->frames, b->first_slot, b->count, ctx->stream);
    if (st != EDSGPU_OK) return st;
    mad_kernel<<<b->count, MAD_THREADS, 0, ctx->stream>>>((const ProblemDesc*)b->desc);
    ctx->launches++;
    EDS_CUDA(ctx, cudaGetLastError());
    return EDSGPU_OK;
}

edsgpu_status edsgpu_batch_launch_shape(const edsgpu_batch* b, int* clusters, int* ctas_per_cluster, int* problems_in_flight) {
    if (!b) return EDSGPU_INVALID_ARGUMENT;
    if (clusters) *clusters = b->shape.nclusters;
    if (ctas_per_cluster) *ctas_per_cluster = b->shape.csize;
    if (problems_in_flight) *problems_in_flight = b->shape.K;
    return EDSGPU_OK;
}

edsgpu_status edsgpu_batch_pack_states_dev(edsgpu_batch* b, double* states_dev) {
    if (!b || !states_dev) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = b->ctx;
    DeviceGuard g(ctx->device);
    const int n = 14 * b->count;
    pack_states_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>((const ProblemDesc*)b->desc, b->count, states_dev);
    ctx->launches++;
    EDS_CUDA(ctx, cudaGetLastError());
    return EDSGPU_OK;
}

edsgpu_status edsgpu_trackers_optimize_batch(edsgpu_ctx* ctx, edsgpu_tracker* const* trackers, const edsgpu_keyframe* const* keyframes, int count,
                                             const edsgpu_frames* frames, int first_slot) {
    if (!ctx) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_batch* b = nullptr;
    edsgpu_status st = edsgpu_batch_create(ctx, trackers, keyframes, count, frames, first_slot, &b);
    if (st != EDSGPU_OK) return st;
    st = edsgpu_batch_optimize(b);
    edsgpu_batch_destroy(b);  // synchronises: the descriptors must outlive the launch
    return st;
}

edsgpu_status edsgpu_trackers_gather(edsgpu_ctx* ctx, edsgpu_tracker* const* trackers, int count, double* states_out, edsgpu_tracker_info* infos_out) {
    if (!ctx) return EDSGPU_INVALID_ARGUMENT;
    EDS_REQUIRE(ctx, trackers && count > 0 && states_out, "trackers_gather: bad arguments");
    DeviceGuard g(ctx->device);
    for (int i = 0; i < count; ++i) {
        EDS_CUDA(ctx, cudaMemcpyAsync(states_out + 14 * i, trackers[i]->state, 14 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (infos_out) EDS_CUDA(ctx, cudaMemcpyAsync(infos_out + i, trackers[i]->info, sizeof(edsgpu_tracker_info), cudaMemcpyDeviceToHost, ctx->stream));
    }
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EDSGPU_OK;
}

edsgpu_status edsgpu_tracker_optimize(edsgpu_tracker* tr, const edsgpu_keyframe* kf, const edsgpu_frames* frames, int slot, double px[3],
                                      double qx[4], double vx[6], double* residuals_out, double* next_loss_param_out, edsgpu_tracker_info* info) {
    if (!tr) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = tr->ctx;
    EDS_REQUIRE(ctx, kf && frames, "tracker_optimize: null handle");
    if (!tr->cached || tr->cached_kf != kf->uid || tr->cached_frames != frames->uid || tr->cached_slot != slot) {
        if (tr->cached) edsgpu_batch_destroy(tr->cached);
        tr->cached = nullptr;
        const edsgpu_keyframe* kfs[1] = {kf};
        edsgpu_tracker* trs[1] = {tr};
        edsgpu_status stc = edsgpu_batch_create(ctx, trs, kfs, 1, frames, slot, &tr->cached);
        if (stc != EDSGPU_OK) return stc;
        tr->cached_kf = kf->uid; tr->cached_frames = frames->uid; tr->cached_slot = slot;
    }
    edsgpu_status st = edsgpu_batch_optimize(tr->cached);
    if (st != EDSGPU_OK) return st;
    DeviceGuard g(ctx->device);
    edsgpu_tracker_info inf;
    double tau = 0.0;
    st = edsgpu_tracker_get_state(tr, px, qx, vx, &tau, &inf);
    if (st != EDSGPU_OK) return st;
    if (info) *info = inf;
    if (!inf.usable) return edsgpu_fail(ctx, EDSGPU_NOT_USABLE, "tracker: solution not usable");
    if (next_loss_param_out) *next_loss_param_out = tau;
    if (residuals_out) {
        // kf->residuals (Tracker.cpp:223-230): fp32 on the device, widened by a kernel
        const size_t N = (size_t)kf->dev.N;
        st = edsgpu_ensure_scratch(ctx, N * sizeof(double));
        if (st != EDSGPU_OK) return st;
        float_to_double_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(tr->residuals, (double*)ctx->scratch, N);
        ctx->launches++;
        EDS_CUDA(ctx, cudaGetLastError());
        EDS_CUDA(ctx, cudaMemcpyAsync(residuals_out, ctx->scratch, N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return EDSGPU_OK;
}

edsgpu_status edsgpu_tracker_evaluate(edsgpu_ctx* ctx, const edsgpu_keyframe* kf, const edsgpu_frames* frames, int slot, int loss_type,
                                      double loss_param, const double px[3], const double qx[4], const double vx[6], double* residuals_out,
                                      double* jacobian_out, double* cost_out, double* H_out, double* g_out) {
    if (!ctx) return EDSGPU_INVALID_ARGUMENT;
    EDS_REQUIRE(ctx, kf && frames && px && qx && vx, "tracker_evaluate: null argument");
    EDS_REQUIRE(ctx, kf->ctx == ctx && frames->ctx == ctx, "tracker_evaluate: handles belong to different contexts");
    EDS_REQUIRE(ctx, slot >= 0 && slot < frames->capacity, "tracker_evaluate: frame slot out of range");
    EDS_REQUIRE(ctx, frames->H == kf->dev.H && frames->W == kf->dev.W, "tracker_evaluate: keyframe and event frame sizes differ");
    DeviceGuard g(ctx->device);
    const size_t N = (size_t)kf->dev.N;
    // scratch: desc | state(14) | eval_out(157) | jac (N*12 float) | widened doubles (N*12)
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
    const size_t o_desc = take(sizeof(ProblemDesc)), o_state = take(14 * 8), o_eval = take(157 * 8), o_info = take(sizeof(edsgpu_tracker_info));
    const size_t o_jac = take(N * 12 * sizeof(float)), o_wide = take(N * 12 * sizeof(double)), o_res = take(N * sizeof(float));
    edsgpu_status st = edsgpu_ensure_scratch(ctx, off);
    if (st == EDSGPU_OK) st = edsgpu_ensure_pinned(ctx, sizeof(ProblemDesc) + 14 * 8 + 157 * 8);
    if (st != EDSGPU_OK) return st;
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    char* ds = (char*)ctx->scratch;
    char* hp = (char*)ctx->pinned;
    ProblemDesc* hd = (ProblemDesc*)hp;
    double* hstate = (double*)(hp + sizeof(ProblemDesc));
    memset(hd, 0, sizeof(ProblemDesc));
    hd->kf = kf->dev;
    hd->frame = frames->tex[slot];
    hd->norms = frames->norms + 2 * slot;
    hd->state = (double*)(ds + o_state);
    hd->residuals = (float*)(ds + o_res);
    hd->info = (edsgpu_tracker_info*)(ds + o_info);
    hd->loss_type = loss_type;
    hd->eval_only = 1;
    hd->jac_out = jacobian_out ? (float*)(ds + o_jac) : nullptr;
    hd->eval_out = (double*)(ds + o_eval);
    memcpy(hstate, px, 24); memcpy(hstate + 3, qx, 32); memcpy(hstate + 7, vx, 48);
    hstate[13] = loss_param;
    EDS_CUDA(ctx, cudaMemcpyAsync(ds + o_desc, hd, sizeof(ProblemDesc), cudaMemcpyHostToDevice, ctx->stream));
    EDS_CUDA(ctx, cudaMemcpyAsync(ds + o_state, hstate, 14 * 8, cudaMemcpyHostToDevice, ctx->stream));
    const int csize = pick_shape(ctx, 1, kf->dev.B).csize;
    st = edsgpu_frames_wait_built(frames, slot, 1, ctx->stream);
    if (st == EDSGPU_OK) st = launch_cluster(ctx, track_eval_kernel, 1, csize, (const ProblemDesc*)(ds + o_desc), 1);
    if (st == EDSGPU_OK) st = edsgpu_frames_mark_read(frames, slot, 1, ctx->stream);
    if (st != EDSGPU_OK) return st;
    double* hev = (double*)(hp + sizeof(ProblemDesc) + 14 * 8);
    EDS_CUDA(ctx, cudaMemcpyAsync(hev, ds + o_eval, 157 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (residuals_out) {
        float_to_double_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>((const float*)(ds + o_res), (double*)(ds + o_wide), N);
        ctx->launches++;
        EDS_CUDA(ctx, cudaMemcpyAsync(residuals_out, ds + o_wide, N * 8, cudaMemcpyDeviceToHost, ctx->stream));
        EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (jacobian_out) {
        float_to_double_kernel<<<(unsigned)((N * 12 + 255) / 256), 256, 0, ctx->stream>>>((const float*)(ds + o_jac), (double*)(ds + o_wide), N * 12);
        ctx->launches++;
        EDS_CUDA(ctx, cudaMemcpyAsync(jacobian_out, ds + o_wide, N * 12 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (cost_out) *cost_out = hev[0];
    if (H_out) memcpy(H_out, hev + 1, 144 * 8);
    if (g_out) memcpy(g_out, hev + 145, 12 * 8);
    return EDSGPU_OK;
}

// ---- tracker <-> depth filter on the device --------------------------
edsgpu_status edsgpu_tracker_get_coord(edsgpu_tracker* tr, const edsgpu_keyframe* kf, const edsgpu_depth_points* dp, double* coord_out,
                                       uint8_t* outlier_out) {
    if (!tr) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = tr->ctx;
    EDS_REQUIRE(ctx, kf && dp && kf->ctx == ctx && dp->ctx == ctx, "tracker_get_coord: handles belong to different contexts");
    EDS_REQUIRE(ctx, dp->N == kf->dev.N, "tracker_get_coord: the filter and the key frame have different point counts");
    DeviceGuard g(ctx->device);
    const size_t N = (size_t)kf->dev.N;
    double* coord_dev = dp->coords + 2 * N;  // the filter's event-frame coordinate buffer
    get_coord_kernel<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(kf->dev.N, kf->src + 2 * N, dp->state, tr->state, kf->dev.fx, kf->dev.fy,
                                                                           kf->dev.cx, kf->dev.cy, kf->dev.W, kf->dev.H, coord_dev, dp->ok);
    ctx->launches++;
    EDS_CUDA(ctx, cudaGetLastError());
    if (coord_out) EDS_CUDA(ctx, cudaMemcpyAsync(coord_out, coord_dev, sizeof(double) * 2 * N, cudaMemcpyDeviceToHost, ctx->stream));
    if (outlier_out) EDS_CUDA(ctx, cudaMemcpyAsync(outlier_out, dp->ok, N, cudaMemcpyDeviceToHost, ctx->stream));
    if (coord_out || outlier_out) EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EDSGPU_OK;
}

edsgpu_status edsgpu_keyframe_refresh_idepth(edsgpu_keyframe* kf, const edsgpu_depth_points* dp) {
    if (!kf) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = kf->ctx;
    EDS_REQUIRE(ctx, dp && dp->ctx == ctx && dp->N == kf->dev.N, "keyframe_refresh_idepth: filter of another context or size");
    DeviceGuard g(ctx->device);
    const size_t N = (size_t)kf->dev.N;
    KfDev& d = kf->dev;
    // same preparation as at upload, the inverse depths read from column 0 of the filter state (KeyFrame::inv_depth.getIDepth, Tracker.cpp:167)
    kf_prepare_kernel<<<d.B, 256, 0, ctx->stream>>>(kf->src, kf->src + 2 * N, dp->state, 4, kf->src + 5 * N, d.N, d.B, (float4*)d.gxy, (float2*)d.dw,
                                                   (double*)d.kpx, (double*)d.kpy, (double*)d.kpz, (double*)d.A);
    ctx->launches++;
    EDS_CUDA(ctx, cudaGetLastError());
    return EDSGPU_OK;
}

edsgpu_status edsgpu_depth_points_update_from_tracker(edsgpu_depth_points* dp, edsgpu_tracker* tr, edsgpu_keyframe* kf, const double* kf_coord,
                                                      int refresh_keyframe) {
    if (!dp || !tr || !kf) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = dp->ctx;
    EDS_REQUIRE(ctx, tr->ctx == ctx && kf->ctx == ctx && dp->N == kf->dev.N, "depth_points_update_from_tracker: mismatched handles");
    DeviceGuard g(ctx->device);
    const size_t N = (size_t)dp->N;
    if (kf_coord) {  // KeyFrame::coord, once per key frame
        EDS_CUDA(ctx, cudaMemcpyAsync(dp->coords, kf_coord, sizeof(double) * 2 * N, cudaMemcpyHostToDevice, ctx->stream));
        EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        dp->kf_coord_set = true;
    }
    EDS_REQUIRE(ctx, dp->kf_coord_set, "depth_points_update_from_tracker: the key-frame coordinates were never given");
    edsgpu_status st = edsgpu_tracker_get_coord(tr, kf, dp, nullptr, nullptr);
    if (st == EDSGPU_OK) st = edsgpu_depth_update_tracked(dp, tr->state, dp->coords, dp->coords + 2 * N);
    if (st == EDSGPU_OK && refresh_keyframe) st = edsgpu_keyframe_refresh_idepth(kf, dp);
    return st;
}

}  // extern "C"
