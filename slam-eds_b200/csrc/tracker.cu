// Event-to-model alignment on the device.
//
// Replaces Tracker::optimize (reference src/tracking/Tracker.cpp:104-241), the Ceres cost
// functor PhotometricError (src/tracking/PhotometricError.hpp:56-214) and ceres::Solve
// (un-vendored; trust-region LM restated from ceres-solver 1.14..2.1 semantics).
//
//   track_lm_kernel   persistent dataflow over global memory, 768-thread CTAs (24 warps at 80 registers), one per SM.  A few
//                     LEADER CTAs: each of their warps runs the whole Levenberg-Marquardt state machine of one problem (Jacobi
//                     scaling, damping, register-resident 12x12 Cholesky, retractions, accept/reject, tolerances) and turns
//                     every evaluation into one task per residual block (Tracker.cpp:178-195) in a global queue.
//                     All other CTAs are EVALUATORS: a control warp pops tasks (running ahead through a mailbox), 16
//                     producer warps evaluate 32 points at a time, start to finish (division-free fp64-exact projection,
//                     bicubic window as four texture gathers, analytic Jacobian in fp32) into a shared-memory ring guarded
//                     by mbarriers, three consumer PAIRS own the 90 fp32 outer-product accumulators (split by rows), reduce
//                     the block with a halving warp butterfly, and a combiner warp applies the per-block loss (rho'), stores
//                     92 doubles into the problem's slot and bumps the counter its leader polls.  Any evaluator serves any
//                     problem: every SM stays busy whatever the batch size, and the serial LM step of one problem overlaps
//                     the sweeps of all the others.  The LM loop never returns to the host: one launch per batch of windows.
//   mad_kernel        next loss parameter (MAD / STD) from the written-back residuals
//                     (Tracker.cpp:281-317) by radix select.
//   kf_prepare_kernel keyframe upload: the per-point model gradient g (fp32, two float4), fp64 3-D points
//                     (PhotometricError.hpp:94-105) and the per-block 6x6 model Gram matrix
//                     A_b = sum g_i g_i^T (m_i = g_i . v is linear in v, so
//                     ||m||^2 = v^T A_b v and sum m_i g_i = A_b v need no sweep).
#include <float.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "depth.cuh"
#include "frames.cuh"

namespace {

#ifndef EDS_TRK_THREADS
#define EDS_TRK_THREADS 768
#endif
constexpr int TRK_THREADS = EDS_TRK_THREADS;  // one CTA per SM; 24 warps = 6 per scheduler leave 80 registers per thread
constexpr int TRK_WARPS = TRK_THREADS / 32;
constexpr int JLD = 20;         // floats per point row in shared memory (80 B: conflict-free 128-bit access)
#ifndef EDS_N_CONS
#define EDS_N_CONS 3
#endif
constexpr int N_CONS = EDS_N_CONS;               // consumer PAIRS: pair c owns the batches j = c mod N_CONS of a block, its two warps split the rows
constexpr int N_CONS_WARPS = 2 * N_CONS;
constexpr int CTRL_WARP = TRK_WARPS - 1;         // evaluator CTA: the warp that fetches tasks from the global queue
// (warp TRK_WARPS - 2 of an evaluator CTA is the combiner: it adds the consumers' block totals and publishes the block)
constexpr int N_PROD = TRK_WARPS - 2 - N_CONS_WARPS;  // producer warps of an evaluator CTA
constexpr int N_EVAL_WARPS = TRK_WARPS - 1;      // producers + consumers + combiner
#ifndef EDS_RING_DEPTH
#define EDS_RING_DEPTH 3
#endif
constexpr int RING_DEPTH = EDS_RING_DEPTH;   // ring slots per producer warp
constexpr int N_SLOTS = RING_DEPTH * N_PROD; // ring of 32-point batches
constexpr int NSLOT = 92;       // 78 + 12 + cost + block squared norm
constexpr int MAX_BLOCKS = 16;  // residual blocks per problem (config.options.num_threads)
#ifndef EDS_MAILBOX
#define EDS_MAILBOX 2
#endif
constexpr int MAILBOX = EDS_MAILBOX;       // tasks an evaluator CTA holds: the running one and the prefetched ones
constexpr int LEAD_WARPS = 16;             // leader warps of a leader CTA (its other warps leave at once), one problem in flight each
constexpr int MAX_LEAD_CTAS = 8;
constexpr double kEps = 1e-05;  // PhotometricError.hpp:200

enum { CMD_EVAL = 1, CMD_FINAL = 2, CMD_DONE = 3, CMD_EXIT = 4, CMD_NOP = 5 };  // NOP: a reserved queue ticket that turned out not to be needed

#ifdef EDS_TIMING
__device__ unsigned long long g_timing[32];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define g_timing_cta ((int)gridDim.x - 1)  // the evaluator CTA whose warps report their clocks
#endif
enum { PHASE_INIT = 0, PHASE_CAND = 1 };

struct KfDev {
    const float4* ga;    // model gradient of the point, g = -(Gx dflow_x/dv + Gy dflow_y/dv) (PhotometricError.hpp:114-122): {g0, g1, g2, g3}
    const float4* gb;    // {g4, g5, weight, -}
    const double* kpx;   // 3-D point (X,Y,1)/(idp+eps)
    const double* kpy;
    const double* kpz;
    const double* A;     // [B][21] upper triangle of sum g g^T per residual block
    int N, B, H, W;
    int ne;              // N / B: points per residual block (the last block also takes the remainder)
    double fx, fy, cx, cy;
    float fxf, fyf;      // (float)fx, (float)fy
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
// dataflow state: global memory (leader <-> evaluators) and the shared memory of the two kinds of CTA
// ------------------------------------------------------------------------------------------
struct LmState {
    double x[13], cand[13], delta[12];
    double Hs[78], gs[12];    // Jacobi-scaled normal equations at x
    double scale[12], diag[12];
    double x_cost, mcc, radius, dec, gmax, x_norm, initial_cost;
    int reuse_diag, iter, n_succ, n_unsucc, consec_invalid, termination, phase, n_eval;
};

// Everything an evaluator needs for one evaluation; published by the problem's leader.
struct EvalConst {
    double R[9], t[3];               // Eigen toRotationMatrix(q), translation
    float vf[6], inv_vs, inv_vn;     // velocity, 1/|v|^2, 1/|v|
    float blk[MAX_BLOCKS][8];        // per residual block: 1/M, alpha, beta[6] (see eval_point)
    int cmd;
};

// One per problem of a batch, in global memory: what the leader publishes and what the evaluators report.
struct ProblemWork {
    EvalConst ec;                      // evaluation point of the tasks in the queue
    double loss_a;                     // loss parameter of this solve (state[13] when the solve started)
    double inv_norm;                   // 1 / L2 norm of the problem's event frame
    double slots[MAX_BLOCKS][NSLOT];   // block sums of the current evaluation, one writer (evaluator CTA) per block
    unsigned done;                     // +1 per finished EVAL task, +1 per evaluator warp of a finished FINAL task; never reset
    unsigned pad_[3];
};

// Task queue shared by all CTAs of a launch.  head / tail / finished only ever grow (also across launches of the same batch):
// ticket k lives in entries[k & mask] and is valid once its upper half equals k + 1.
struct QueueCtl { unsigned head, tail, finished, pad_; };
__device__ __forceinline__ unsigned pack_task(int problem, int block, int cmd) { return (unsigned)problem | ((unsigned)block << 20) | ((unsigned)cmd << 26); }

// A task as the evaluator warps see it (copied out of global memory by the CTA's control warp).
struct TaskShared {
    EvalConst ec;
    double loss_a;
    ProblemDesc P;
    double* slots;     // ProblemWork::slots of the problem
    unsigned* done;    // ProblemWork::done
    float inv_norm;    // 1 / L2 norm of the problem's event frame (EventFrame.cpp:360-383), applied at sample time
    int block, cmd;
};

// shared memory of an evaluator CTA
struct EvalShared {
    TaskShared task[MAILBOX];
    // ring of 32-point batches of rows [J(12) r pad], guarded by full/empty mbarriers
    alignas(16) float ring[N_SLOTS][32][JLD];
    alignas(8) unsigned long long full_bar[N_SLOTS];
    alignas(8) unsigned long long empty_bar[N_SLOTS];
    // block totals of the consumer warps (double-buffered by block), added up by the combiner warp
    float cons_part[2][N_CONS][96];
    double cons_s[2][N_CONS];
    alignas(8) unsigned long long part_full[2];    // consumers -> combiner: every consumer warp has left its totals of the block
    alignas(8) unsigned long long part_empty[2];   // combiner -> consumers: the buffer has been read
    // mailbox hand-over between the control warp and the evaluator warps
    alignas(8) unsigned long long task_full[MAILBOX];   // control warp -> evaluators: task[slot] is complete (1 arrival)
    alignas(8) unsigned long long task_empty[MAILBOX];  // evaluators -> control warp: every evaluator warp is done with task[slot]
};

// shared memory of a leader CTA: one problem in flight per warp
struct LeaderProblem {
    EvalConst ec;
    double loss_a;
    double sum[NSLOT];
    double A[MAX_BLOCKS][21];  // per-block model Gram matrices
    double x_eval[13];
    LmState lm;
    ProblemDesc P;
};
struct LeadShared { LeaderProblem prob[LEAD_WARPS]; };

// ---- global-memory hand-over primitives -----------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ------------------------------------------------------------------------------------------
// per-point residual + analytic tangent-space Jacobian
// ------------------------------------------------------------------------------------------
// Catmull-Rom weights of ceres' CubicHermiteSpline (call site PhotometricError.hpp:172) in basis
// form: f = sum_k w_k(x) p_k, f' = sum_k dw_k(x) p_k  (same cubic as the Horner form a,b,c,d).
// TWICE the weights: the factor 1/2 of every weight (1/4 per tap of the separable window) is exact in binary floating
// point, so it is applied once, to the sample, instead of sixteen times here -- same bits, fewer multiplications.
__device__ __forceinline__ void cr_weights2(float x, float* w, float* dw) {
    const float x2 = x * x;
    const float b = fmaf(-3.0f, x, 4.0f);
    w[0] = x * fmaf(2.0f - x, x, -1.0f);
    w[1] = fmaf(x2, fmaf(3.0f, x, -5.0f), 2.0f);
    w[2] = x * fmaf(b, x, 1.0f);
    w[3] = x2 * (x - 1.0f);
    dw[0] = fmaf(b, x, -1.0f);
    dw[1] = x * fmaf(9.0f, x, -10.0f);
    dw[2] = fmaf(fmaf(-9.0f, x, 8.0f), x, 1.0f);
    dw[3] = x * fmaf(3.0f, x, -2.0f);
}

// Stage 1 of a point: warp into the event camera and project (PhotometricError.hpp:157-168) ->
// interpolation cell, in-cell fractions and the camera-frame quantities the Jacobian needs.
struct PointGeo {
    int col, row;      // cell, clamped to [-4, W+3] x [-4, H+3]
    float tc, tr;      // fractions in the cell (0 where clamped)
    float iz, px, py;  // P = R kp + t: 1/Pz, Px, Py
    float ax, ay, az;  // R kp
};

// (cell, u - cell) for u = f p / pz + c without an fp64 division: 1/pz comes from the fp32 reciprocal and one Newton step
// in fp64 (relative error 2^-46), so u carries ~1e-11 px where plain fp32 projection would carry ~2e-5 px (SURVEY.md
// section 7).  At a cell boundary both choices of the cell give the same interpolant.  Outside [-4, limit + 3] the clamped
// Grid2D is constant: fraction dropped.  The conversion saturates (NaN -> 0).
__device__ __forceinline__ void project_axis(double fp, double c, double rz, int limit, int& cell, float& frac) {
    const double u = fma(fp, rz, c);
    const int q = __double2int_rd(u);
    cell = max(-4, min(q, limit + 3));
    const float t = (float)(u - (double)q);
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
    const double r0 = (double)G.iz;
    const double rz = fma(r0, fma(-pz, r0, 1.0), r0);
    project_axis(kf.fx * px, kf.cx, rz, kf.W, G.col, G.tc);
    project_axis(kf.fy * py, kf.cy, rz, kf.H, G.row, G.tr);
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
// the point's model record ga = {g0..g3}, gb = {g4, g5, weight, -} (prepared once per key frame, kf_prepare_kernel).
template <bool WANT_J>
__device__ __forceinline__ void point_finish(const KfDev& kf, const EvalConst& K, const float* __restrict__ bc, float inv_norm,
                                             const PointGeo& G, const Taps& T, float4 ga, float4 gb, float* __restrict__ J, float& r) {
    // model term: m = g . v with g = -(Gx dflow_x/dv + Gy dflow_y/dv), PhotometricError.hpp:114-122,145
    const float g[6] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y};
    const float w = gb.z;
    float m = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) m += g[k] * K.vf[k];
    float wc[4], dwc[4], wr[4], dwr[4];
    cr_weights2(G.tc, wc, dwc);
    cr_weights2(G.tr, wr, dwr);
    inv_norm *= 0.25f;  // the factor 1/2 of the column and of the row weights
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
    const float fxf = kf.fxf, fyf = kf.fyf;
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
    point_finish<WANT_J>(kf, K, bc, inv_norm, G, T, __ldg(&kf.ga[idx]), __ldg(&kf.gb[idx]), J, r);
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

// ------------------------------------------------------------------------------------------
// evaluator CTA: producer warps -> ring -> consumer warps
// ------------------------------------------------------------------------------------------
// A task is ONE residual block of one problem at one evaluation point.  The CTA's tasks arrive through the mailbox in
// an order every evaluator warp follows; the batches (32 points) of consecutive EVAL tasks carry one running number g,
// from which both sides derive ring slot and phase without talking to each other:
//   producers  sweep the points (residual + analytic Jacobian row -> ring slot g mod N_SLOTS, mbarrier "full").  Warp p
//              takes the batches g = p (mod N_PROD), one at a time and start to finish: latency is hidden by the NUMBER
//              of producer warps (18 at 80 registers), not by a software pipeline inside the warp -- the scoreboards a
//              warp has are shared by all its loads in flight, so fetches issued for a later batch made the finish of
//              the current one wait for them (measured in round 2: DEPBAR on the texture scoreboard).
//   consumers  own the outer-product accumulators.  Consumer PAIR c takes the batches j = c (mod N_CONS) of the block (a
//              fixed summation order); its two warps split the 90 entries by rows ([0,4): 46 entries, [4,12): 44), so a
//              consumer fits in 80 registers as well.  A block ends with a halving warp butterfly per warp, one warp (in
//              turn) adds the partial sums in index order, applies the per-block loss and stores
//              [rho' * JtJ (78) | rho' * Jtr (12) | 0.5 rho(s) | s] into the problem's slot of that block, then bumps
//              the counter the problem's leader polls.
// The ring has a whole number of slots per producer warp, so that the successive occupants of a slot always belong to
// the SAME producer: it passes through every "empty" wait of its slots in order and can never test a parity that is two
// phases stale.  The mirror image on the "full" side: the successive occupants of a slot within a block are drained by
// the same consumer pair (N_SLOTS is a multiple of N_CONS), and the consumers' end-of-block barrier keeps them within
// one block of each other.
struct Roles {
    int pidx;        // producer index of this warp, -1 if not a producer
    int cidx;        // consumer pair of this warp, -1 if not a consumer
    int half;        // which rows of the pair's batches this consumer warp accumulates
};
__device__ __forceinline__ Roles make_roles() {
    const int warp = threadIdx.x >> 5;
    Roles r;
    static_assert(MAILBOX >= 2, "the evaluate entry needs two mailbox slots");
    static_assert(N_CONS >= 2 && N_CONS <= 4, "consumer pairs");
    static_assert(N_SLOTS % N_CONS == 0, "ring slots per CTA must be a multiple of the consumer pairs");
    static_assert(N_PROD >= 1, "no producer warps");
    r.pidx = warp < N_PROD ? warp : -1;
    const bool cons = warp >= N_PROD && warp < N_PROD + N_CONS_WARPS;
    r.cidx = cons ? (warp - N_PROD) >> 1 : -1;
    r.half = cons ? (warp - N_PROD) & 1 : 0;
    return r;
}

// points [start, start + n) and batches of residual block b (the last block also takes the remainder, Tracker.cpp:178-190)
__device__ __forceinline__ void block_extent(const KfDev& kf, int b, int& start, int& n, int& nb) {
    start = b * kf.ne;
    n = kf.ne + ((b + 1 == kf.B) ? (kf.N - (b + 1) * kf.ne) : 0);
    nb = (n + 31) >> 5;
}

// the texture handle is read from shared memory, which the compiler cannot prove warp-uniform: it would wrap every fetch
// in a loop over the distinct handles of the warp.  A warp-wide OR leaves the value unchanged and lands in a uniform register.
__device__ __forceinline__ cudaTextureObject_t uniform_handle(cudaTextureObject_t h) {
    return ((unsigned long long)__reduce_or_sync(0xffffffffu, (unsigned)(h >> 32)) << 32) | (unsigned long long)__reduce_or_sync(0xffffffffu, (unsigned)h);
}

// residual write-back of a FINAL task (Tracker.cpp:223-230): no Jacobian, no reduction; every evaluator warp takes a
// slice and reports on its own (the leader waits for N_EVAL_WARPS arrivals per block)
__device__ void final_slice(const TaskShared& ts) {
    const ProblemDesc& P = ts.P;
    int start, n, nb;
    block_extent(P.kf, ts.block, start, n, nb);
    const cudaTextureObject_t frame = uniform_handle(P.frame);
    const float* bc = ts.ec.blk[ts.block];
    for (int i = threadIdx.x; i < n; i += 32 * N_EVAL_WARPS) {
        float r;
        eval_point<false>(P.kf, ts.ec, bc, frame, ts.inv_norm, start + i, nullptr, r);
        P.residuals[start + i] = r;
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) red_release_add(ts.done, 1u);
}

__device__ void producer_warp_main(EvalShared& sh, const int pidx) {
    const int lane = threadIdx.x & 31;
    unsigned base = 0;  // running batch number at the start of the current task
#ifdef EDS_TIMING
    unsigned long long t_wait = 0, t_work = 0;
#endif
    for (unsigned ord = 0;; ++ord) {
        const unsigned tslot = ord % MAILBOX;
#ifdef EDS_TIMING
        const unsigned long long t0 = gtime();
#endif
        mbar_wait(&sh.task_full[tslot], (ord / MAILBOX) & 1u, 3);
#ifdef EDS_TIMING
        const unsigned long long t1 = gtime();
        t_wait += t1 - t0;
#endif
        const TaskShared& ts = sh.task[tslot];
        const int cmd = ts.cmd;
        if (cmd == CMD_EXIT) {
#ifdef EDS_TIMING
            if (lane == 0 && g_timing_cta == (int)blockIdx.x && pidx == 0) { atomicAdd(&g_timing[13], t_work); atomicAdd(&g_timing[14], t_wait); }
#endif
            return;
        }
        if (cmd == CMD_FINAL) {
            final_slice(ts);
        } else if (cmd == CMD_EVAL) {
            const ProblemDesc& P = ts.P;
            const KfDev& kf = P.kf;
            int start, n, nb;
            block_extent(kf, ts.block, start, n, nb);
            const cudaTextureObject_t frame = uniform_handle(P.frame);
            const float* bc = ts.ec.blk[ts.block];
            const float inv_norm = ts.inv_norm;
            // this warp's batches have running number g = pidx (mod N_PROD)
            for (int t = (pidx + N_PROD - (int)(base % (unsigned)N_PROD)) % N_PROD; t < nb; t += N_PROD) {
                const int i = (t << 5) + lane;
                const bool live = i < n;
                const int idx = start + (live ? i : n - 1);  // idle lanes of a ragged batch repeat the block's last point, their row is zeroed
#ifdef EDS_TIMING
                const long long tp0 = clock64();
#endif
                const Kp kp = load_kp(kf, idx);
                const float4 ga = __ldg(&kf.ga[idx]);
                const float4 gb = __ldg(&kf.gb[idx]);
                PointGeo G;
                point_geometry(kf, ts.ec, kp, G);
#ifdef EDS_TIMING
                asm volatile("" ::"r"(G.col), "r"(G.row), "f"(G.tc), "f"(G.tr));
                const long long tp1 = clock64();
#endif
                const Taps T = fetch_taps(frame, G.col, G.row);
#ifdef EDS_TIMING
                asm volatile("" ::"f"(T.q00.x), "f"(T.q10.x), "f"(T.q01.x), "f"(T.q11.x), "f"(ga.x), "f"(gb.x));
                const long long tp2 = clock64();
#endif
                float J[12], r;
                point_finish<true>(kf, ts.ec, bc, inv_norm, G, T, ga, gb, J, r);
#ifdef EDS_TIMING
                asm volatile("" ::"f"(J[0]), "f"(J[5]), "f"(J[11]), "f"(r));
                const long long tp3 = clock64();
#endif
                if (P.eval_only && live) {  // parity/debug entry: residuals and Jacobian rows written out
                    P.residuals[idx] = r;
                    if (P.jac_out) {
#pragma unroll
                        for (int k = 0; k < 12; ++k) P.jac_out[(size_t)12 * idx + k] = J[k];
                    }
                }
                const float keep = live ? 1.f : 0.f;
#pragma unroll
                for (int k = 0; k < 12; ++k) J[k] *= keep;
                r *= keep;
                const unsigned g = base + (unsigned)t;
                const unsigned slot = g % (unsigned)N_SLOTS, phase = (g / (unsigned)N_SLOTS) & 1u;
#ifdef EDS_TIMING
                const long long te0 = clock64();
#endif
                mbar_wait(&sh.empty_bar[slot], phase ^ 1u, 1);
#ifdef EDS_TIMING
                if (lane == 0 && g_timing_cta == (int)blockIdx.x) { atomicAdd(&g_timing[18], (unsigned long long)(clock64() - te0)); atomicAdd(&g_timing[19], 1ull); }
#endif
                float4* dst = reinterpret_cast<float4*>(&sh.ring[slot][lane][0]);
                dst[0] = make_float4(J[0], J[1], J[2], J[3]);
                dst[1] = make_float4(J[4], J[5], J[6], J[7]);
                dst[2] = make_float4(J[8], J[9], J[10], J[11]);
                dst[3] = make_float4(r, 0.f, 0.f, 0.f);
                __syncwarp();  // all 32 rows are written: one elected arrival publishes the slot
                if (lane == 0) mbar_arrive(&sh.full_bar[slot]);
#ifdef EDS_TIMING
                if (lane == 0 && g_timing_cta == (int)blockIdx.x && pidx == 0) {
                    atomicAdd(&g_timing[27], (unsigned long long)(tp1 - tp0)); atomicAdd(&g_timing[28], (unsigned long long)(tp2 - tp1));
                    atomicAdd(&g_timing[29], (unsigned long long)(tp3 - tp2)); atomicAdd(&g_timing[30], (unsigned long long)(clock64() - tp3));
                    atomicAdd(&g_timing[31], 1ull);
                }
#endif
            }
            base += (unsigned)nb;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.task_empty[tslot]);
#ifdef EDS_TIMING
        t_work += gtime() - t1;
#endif
    }
}

// the two halves of a consumer pair: rows [0,4) of the upper triangle of [J r]^T [J r] (46 entries) and rows [4,12) (44)
typedef RowBlock<0, 4> ConsLo;
typedef RowBlock<4, 12> ConsHi;
constexpr int CONS_LO = 46, CONS_HI = 44;
static_assert(ConsLo::count() == CONS_LO && ConsHi::count() == CONS_HI, "row split");

// 48 per-lane partial sums -> warp totals; lane l ends with entries base48(l)+{0,1,2} (lanes 2k and 2k+1 hold the same three)
__device__ __forceinline__ void reduce48(float* acc, unsigned lane) {
    butterfly_step<24, 16>(acc, lane);
    butterfly_step<12, 8>(acc, lane);
    butterfly_step<6, 4>(acc, lane);
    butterfly_step<3, 2>(acc, lane);
#pragma unroll
    for (int i = 0; i < 3; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 1);
}
__device__ __forceinline__ int reduce48_base(unsigned lane) {
    return 24 * ((lane >> 4) & 1) + 12 * ((lane >> 3) & 1) + 6 * ((lane >> 2) & 1) + 3 * ((lane >> 1) & 1);
}

template <int HALF>
__device__ void consumer_warp_main(EvalShared& sh, const int cidx) {
    typedef typename std::conditional<HALF == 0, ConsLo, ConsHi>::type Rows;
    constexpr int COUNT = HALF == 0 ? CONS_LO : CONS_HI;   // entries of this half
    constexpr int FIRST = HALF == 0 ? 0 : CONS_LO;         // their position among the 90
    const int lane = threadIdx.x & 31;
#ifdef EDS_TIMING
    const int widx = 2 * cidx + HALF;  // consumer warp number
#endif
    unsigned base = 0, block_counter = 0;  // running batch number, EVAL tasks seen
#ifdef EDS_TIMING
    unsigned long long t_wait = 0, t_work = 0, n_work = 0;
#endif
    for (unsigned ord = 0;; ++ord) {
        const unsigned tslot = ord % MAILBOX;
#ifdef EDS_TIMING
        const unsigned long long t0 = gtime();
#endif
        mbar_wait(&sh.task_full[tslot], (ord / MAILBOX) & 1u, 3);
#ifdef EDS_TIMING
        const unsigned long long t1 = gtime();
#endif
        const TaskShared& ts = sh.task[tslot];
        const int cmd = ts.cmd;
        if (cmd == CMD_EXIT) break;
        if (cmd == CMD_FINAL) {
            final_slice(ts);
        } else if (cmd == CMD_EVAL) {
            const ProblemDesc& P = ts.P;
            int start, n, nb;
            block_extent(P.kf, ts.block, start, n, nb);
            float acc[48];
#pragma unroll
            for (int i = 0; i < 48; ++i) acc[i] = 0.f;
            double s_acc = 0.0;  // sum of r^2 in fp64 (the cost decides accept / reject and the tolerances), kept by the upper half
            unsigned slot, phase;
            {
                const unsigned g0 = base + (unsigned)cidx;
                slot = g0 % (unsigned)N_SLOTS;
                phase = (g0 / (unsigned)N_SLOTS) & 1u;
            }
            for (int j = cidx; j < nb; j += N_CONS, slot += N_CONS) {
                if (slot >= (unsigned)N_SLOTS) { slot -= (unsigned)N_SLOTS; phase ^= 1u; }
#ifdef EDS_TIMING
                const long long tw0 = clock64();
#endif
                mbar_wait(&sh.full_bar[slot], phase, 2);
#ifdef EDS_TIMING
                if (lane == 0 && g_timing_cta == (int)blockIdx.x && widx == 0) { atomicAdd(&g_timing[16], (unsigned long long)(clock64() - tw0)); atomicAdd(&g_timing[17], 1ull); }
#endif
                float v[16];
                Rows::load(sh.ring[slot], lane, v);
                __syncwarp();  // every lane has its row in registers: one elected arrival per half frees the slot
                if (lane == 0) mbar_arrive(&sh.empty_bar[slot]);
                const float rr = Rows::accumulate(v, acc);
                if (HALF == 1) s_acc += (double)rr * (double)rr;
            }
            reduce48(acc, lane);
            if (HALF == 1) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s_acc += __shfl_xor_sync(0xffffffffu, s_acc, o);
            }
            // every consumer leaves its block totals in shared memory (by entry number) for the combiner warp and moves on to
            // the next block; the buffer of block k is reused by block k + 2, once the combiner has read it
            const unsigned buf = block_counter & 1u, use = block_counter >> 1;
            if (use > 0) mbar_wait(&sh.part_empty[buf], (use - 1u) & 1u, 6);
            if ((lane & 1) == 0) {
                const int b0 = reduce48_base(lane);
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    if (b0 + i < COUNT) sh.cons_part[buf][cidx][FIRST + b0 + i] = acc[i];
            }
            if (HALF == 1 && lane == 0) sh.cons_s[buf][cidx] = s_acc;
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.part_full[buf]);
            base += (unsigned)nb;
            block_counter++;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.task_empty[tslot]);
#ifdef EDS_TIMING
        t_wait += t1 - t0; t_work += gtime() - t1; n_work++;
#endif
    }
#ifdef EDS_TIMING
    if (lane == 0 && g_timing_cta == (int)blockIdx.x && widx == 0) { atomicAdd(&g_timing[3], t_work); atomicAdd(&g_timing[4], t_wait); atomicAdd(&g_timing[5], n_work); }
#endif
}

// Combiner warp of an evaluator CTA: follows the task stream like the other evaluator warps; for every EVAL task it waits
// for the consumers' block totals, adds them in index order (a fixed summation order), applies the per-block loss and
// publishes the block to the problem's leader.  The consumers never wait for this: the loss (fp64 sqrt / division), the
// global stores and the release run beside the accumulation of the next block.
__device__ void combiner_warp_main(EvalShared& sh) {
    const int lane = threadIdx.x & 31;
    unsigned block_counter = 0;
    for (unsigned ord = 0;; ++ord) {
        const unsigned tslot = ord % MAILBOX;
        mbar_wait(&sh.task_full[tslot], (ord / MAILBOX) & 1u, 3);
        const TaskShared& ts = sh.task[tslot];
        const int cmd = ts.cmd;
        if (cmd == CMD_EXIT) break;
        if (cmd == CMD_FINAL) {
            final_slice(ts);
        } else if (cmd == CMD_EVAL) {
            const ProblemDesc& P = ts.P;
            const unsigned buf = block_counter & 1u, use = block_counter >> 1;
            mbar_wait(&sh.part_full[buf], use & 1u, 7);
            float tot[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) tot[i] = sh.cons_part[buf][0][3 * lane + i];
            double s_tot = sh.cons_s[buf][0];
#pragma unroll
            for (int c = 1; c < N_CONS; ++c) {
#pragma unroll
                for (int i = 0; i < 3; ++i) tot[i] += sh.cons_part[buf][c][3 * lane + i];
                s_tot += sh.cons_s[buf][c];
            }
            __syncwarp();  // every lane has read its entries: the buffer goes back to the consumers
            if (lane == 0) mbar_arrive(&sh.part_empty[buf]);
            double rho0, rho1;
            loss_eval(P.loss_type, ts.loss_a, s_tot, &rho0, &rho1);
            double* dst = ts.slots + ts.block * NSLOT;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int e = 3 * lane + i;
                if (e < 90) dst[ConsRows::slot(e)] = rho1 * (double)tot[i];
            }
            if (lane == 0) { dst[90] = 0.5 * rho0; dst[91] = s_tot; }
            // The lanes' stores happen before the warp barrier, the barrier before the elected lane's release: the
            // release is cumulative over what the lane has synchronised with, so the leader that acquires the count
            // sees every lane's sums (the CUTLASS semaphore pattern: barrier, then one thread's red.release.gpu).
            __syncwarp();
            if (!P.eval_only) {
                if (lane == 0) red_release_add(ts.done, 1u);
            } else {
                // parity/debug entry (track_eval_kernel): the block that finishes last sums the blocks
                unsigned before = 0;
                if (lane == 0) { __threadfence(); before = atomicAdd(ts.done, 1u); }
                before = __shfl_sync(0xffffffffu, before, 0);
                if (before == (unsigned)(P.kf.B - 1) && P.eval_out) {
                    __threadfence();
                    const int B = P.kf.B;
                    for (int e = lane; e < 91; e += 32) {
                        double v = 0.0;
                        for (int b = 0; b < B; ++b) v += __ldcg(&ts.slots[b * NSLOT + e]);
                        if (e == 90) P.eval_out[0] = v;
                        else if (e >= 78) P.eval_out[1 + 144 + (e - 78)] = v;
                        else {
                            int a = 0, rem = e;
                            while (rem >= 12 - a) { rem -= 12 - a; ++a; }
                            const int c2 = a + rem;
                            P.eval_out[1 + 12 * a + c2] = v;
                            P.eval_out[1 + 12 * c2 + a] = v;
                        }
                    }
                }
            }
            block_counter++;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.task_empty[tslot]);
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

// One warp: the evaluation constants of point xe (13 doubles) into ec.  Lane b derives the per-block model
// normalisation S_b = 1e-3 + v^T A_b v (PhotometricError.hpp:132,148) and c_b = A_b v from the Gram matrices A [B][21].
__device__ void compute_eval_const(EvalConst& ec, const double* xe, const double (*Ab)[21], int B, int cmd) {
    const int lane = threadIdx.x & 31;
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
        const double* A = Ab[lane];
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
}

// what a leader warp needs to talk to the evaluators
struct LeaderLink {
    ProblemWork* w;
    QueueCtl* ctl;
    unsigned long long* entries;
    unsigned mask;
    int pid;
};

// Leader warp: append `n` tasks (block 0..n-1 of problem pid, or n CMD_EXIT tasks) to the queue.  Everything the tasks
// refer to must have been written and fenced by the calling lanes before.
// `reserved` (lane 0): first of n tickets taken earlier with reserve_tickets, or RESERVE_NOW.
constexpr unsigned RESERVE_NOW = 0xffffffffu;
__device__ __forceinline__ unsigned reserve_tickets(const LeaderLink& L, int n) {  // lane 0's return value counts
    return ((threadIdx.x & 31) == 0) ? atomicAdd(&L.ctl->tail, (unsigned)n) : 0u;
}
__device__ void push_tasks(const LeaderLink& L, int n, int cmd, unsigned reserved = RESERVE_NOW) {
    const int lane = threadIdx.x & 31;
    unsigned t0 = reserved;
    if (__shfl_sync(0xffffffffu, reserved, 0) == RESERVE_NOW) t0 = reserve_tickets(L, n);
    t0 = __shfl_sync(0xffffffffu, t0, 0);
    for (int i = lane; i < n; i += 32) {
        const unsigned ticket = t0 + (unsigned)i;
        const unsigned payload = pack_task(L.pid, cmd == CMD_EXIT ? 0 : i, cmd);
        st_release_u64(&L.entries[ticket & L.mask], ((unsigned long long)(ticket + 1u) << 32) | payload);
    }
}

// Leader warp: publish the evaluation constants of point xe and queue one task per residual block.
__device__ void leader_publish(LeaderProblem& lp, const LeaderLink& L, const double* xe, int cmd) {
    const int lane = threadIdx.x & 31;
    const int B = lp.P.kf.B;
    // the round trip of the ticket counter runs beside the computation of the constants.  Not earlier: an evaluator that
    // draws a ticket waits for its entry, so a ticket must not stay unfilled for long
    const unsigned reserved = (cmd != CMD_DONE) ? reserve_tickets(L, B) : 0u;
    compute_eval_const(lp.ec, xe, lp.A, B, cmd);
    if (cmd == CMD_DONE) return;  // nothing left to evaluate
    // the used prefix of EvalConst (R, t, v + B block records) and the command go to global memory
    const int nwords = (int)(offsetof(EvalConst, blk) / 4) + 8 * B;
    const int* src = reinterpret_cast<const int*>(&lp.ec);
    int* dst = reinterpret_cast<int*>(&L.w->ec);
    for (int i = lane; i < nwords; i += 32) dst[i] = src[i];
    if (lane == 0) L.w->ec.cmd = cmd;
    // the lanes' stores are ordered before the queue entries ANY lane writes after this barrier: the entries are stored with
    // release semantics, which is cumulative over what the writing lane has synchronised with
    __syncwarp();
    push_tasks(L, B, cmd, reserved);
}

// Leader warp: wait until the problem's completion counter has reached `target`.
// Every lane polls (one broadcast request): a loop run by one lane only leaves the warp in diverged mode, in which
// every later shuffle of the LM step takes the slow collective path (measured: the step went from 8 to 28 us).
__device__ __forceinline__ void wait_done(const LeaderLink& L, unsigned target) {
    for (;;) {
        const unsigned v = __shfl_sync(0xffffffffu, ld_acquire_u32(&L.w->done), 0);
        if ((int)(v - target) >= 0) break;
        __nanosleep(20);
    }
}

// Leader warp (32 lanes, convergent): consume one evaluation and decide what to do next.
// ceres TrustRegionMinimizer + LevenbergMarquardtStrategy semantics (options of
// Tracker.cpp:117-143); returns the next command, lm.cand / lm.x hold the point to evaluate.
__device__ int lm_advance_warp(LeaderProblem& sh, const ProblemWork* w) {
    const ProblemDesc& P = sh.P;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    LmState& lm = sh.lm;
    const int B = P.kf.B;
#ifdef EDS_TIMING
    unsigned long long tt0 = gtime();
#endif
    // fixed pairwise tree over the residual blocks: deterministic and independent of the launch shape.  The tree is over
    // MAX_BLOCKS leaves, missing blocks are zeros; with B <= 8 its first level only adds zeros, so the 8-leaf tree gives the
    // same bits with half the loads.  All loads of the three entries a lane owns are issued before the first addition.
    bool fin = true;
    if (B <= MAX_BLOCKS / 2) {
        double v[3][MAX_BLOCKS / 2];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int e = min(lane + 32 * k, 90);
#pragma unroll
            for (int b = 0; b < MAX_BLOCKS / 2; ++b) v[k][b] = (b < B) ? __ldcg(&w->slots[b][e]) : 0.0;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
#pragma unroll
            for (int h = MAX_BLOCKS / 4; h > 0; h >>= 1)
#pragma unroll
                for (int b = 0; b < h; ++b) v[k][b] += v[k][b + h];
            sh.sum[min(lane + 32 * k, 90)] = v[k][0];  // the surplus lanes of the last pass repeat entry 90
            fin = fin && isfinite(v[k][0]);
        }
    } else {
#pragma unroll 1
        for (int k = 0; k < 3; ++k) {
            const int e = min(lane + 32 * k, 90);
            double v[MAX_BLOCKS];
#pragma unroll
            for (int b = 0; b < MAX_BLOCKS; ++b) v[b] = (b < B) ? __ldcg(&w->slots[b][e]) : 0.0;
#pragma unroll
            for (int h = MAX_BLOCKS / 2; h > 0; h >>= 1)
#pragma unroll
                for (int b = 0; b < h; ++b) v[b] += v[b + h];
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
#ifdef EDS_TIMING
    const unsigned long long td1 = gtime();
    unsigned long long td2 = td1, td3 = td1, td4 = td1;
#endif
    if (take) {
        double Hr[12];  // row `lane` of the system at the new point
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
#ifdef EDS_TIMING
        td2 = gtime();
#endif
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
#ifdef EDS_TIMING
        td3 = gtime();
#endif
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
#ifdef EDS_TIMING
        td4 = gtime();
#endif
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
        // lane i < 12 owns row i of the damped matrix; the un-damped scaled system stays in shared memory (lm.Hs), where this
        // call or an earlier one left it: nothing of it is carried in registers across the solve
        double a[12];
        const int li = lane < 12 ? lane : 0;
        const double damping = lm.diag[li] / radius;  // D^2 of this lane's diagonal entry
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            const double v = lm.Hs[li <= j ? tri_index(li, j) : tri_index(j, li)];
            const double hj = (lane < 12) ? v : 0.0;
            a[j] = (j == lane) ? hj + damping : hj;
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
            double hv = 0.0;  // lanes >= 12 have step = 0 and gsl = 0: whatever they read drops out of the sum
#pragma unroll
            for (int j = 0; j < 12; ++j) hv += lm.Hs[li <= j ? tri_index(li, j) : tri_index(j, li)] * __shfl_sync(FULL, step, j);
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
        if (lane == 0) { atomicAdd(&g_timing[22], td1 - tt1); atomicAdd(&g_timing[23], td2 - td1); atomicAdd(&g_timing[24], td3 - td2); atomicAdd(&g_timing[25], td4 - td3); atomicAdd(&g_timing[26], tt2 - td4); }
        if (lane == 0) { atomicAdd(&g_timing[6], tt1 - tt0); atomicAdd(&g_timing[7], tt2 - tt1); atomicAdd(&g_timing[8], tt3 - tt2); atomicAdd(&g_timing[9], gtime() - tt3); atomicAdd(&g_timing[10], 1ull); }
#endif
    }
    if (lane == 0) {
        lm.radius = radius; lm.dec = dec; lm.reuse_diag = reuse; lm.n_succ = n_succ; lm.n_unsucc = n_unsucc;
        lm.iter = iter; lm.consec_invalid = consec; lm.mcc = mcc; lm.phase = PHASE_CAND;
    }
    __syncwarp();
    return ret;
}

// ------------------------------------------------------------------------------------------
// leader CTA: one warp per problem in flight
// ------------------------------------------------------------------------------------------
// leader warp: reset the LM state of a problem and publish its first evaluation point
__device__ void leader_start(LeaderProblem& lp, const LeaderLink& L) {
    const int lane = threadIdx.x & 31;
    LmState& lm = lp.lm;
    if (lane < 13) lm.x[lane] = lp.x_eval[lane];
    if (lane == 0) {
        lm.radius = 1e4; lm.dec = 2.0; lm.reuse_diag = 0;
        lm.iter = 0; lm.n_succ = 0; lm.n_unsucc = 0; lm.consec_invalid = 0; lm.n_eval = 0;
        lm.termination = EDSGPU_TERM_NO_CONVERGENCE; lm.phase = PHASE_INIT;
        lm.x_cost = 0.0; lm.initial_cost = 0.0; lm.mcc = 0.0; lm.gmax = 0.0; lm.x_norm = 0.0;
        L.w->loss_a = lp.loss_a;
        L.w->inv_norm = lp.P.norms[1];
    }
    __syncwarp();
    leader_publish(lp, L, lm.x, CMD_EVAL);
}

// leader warp: consume the evaluation of a problem, advance its LM state, publish the next command; returns it
__device__ int leader_step(LeaderProblem& lp, const LeaderLink& L) {
    const int lane = threadIdx.x & 31;
    LmState& lm = lp.lm;
    const ProblemDesc& P = lp.P;
    const int next = lm_advance_warp(lp, L.w);
#ifdef EDS_TIMING
    const unsigned long long tp0 = gtime();
#endif
    leader_publish(lp, L, (next == CMD_EVAL) ? lm.cand : lm.x, next);
#ifdef EDS_TIMING
    if (lane == 0) atomicAdd(&g_timing[15], gtime() - tp0);
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
    return next;
}

// One leader warp: takes problems warp_id, warp_id + stride, ... to completion, one at a time.  A problem is a
// sequential Levenberg-Marquardt loop: wait for the block sums of the evaluation it has queued, take the decision,
// publish the next evaluation point and queue its tasks.  The serial step of one problem (~8 us of dependent fp64
// latency) overlaps the sweeps of all the others, which run on the evaluator CTAs.
__device__ void leader_warp_main(LeaderProblem& lp, const ProblemDesc* __restrict__ problems, ProblemWork* work, QueueCtl* ctl,
                                 unsigned long long* entries, unsigned mask, int count, int first, int stride, int n_eval_ctas) {
    const int lane = threadIdx.x & 31;
#ifdef EDS_TIMING
    unsigned long long t_wait = 0, t_work = 0, n_work = 0;
#endif
    for (int pid = first; pid < count; pid += stride) {
        LeaderLink L{work + pid, ctl, entries, mask, pid};
        {
            const int* src = reinterpret_cast<const int*>(&problems[pid]);
            int* dst = reinterpret_cast<int*>(&lp.P);
            for (int i = lane; i < (int)(sizeof(ProblemDesc) / sizeof(int)); i += 32) dst[i] = src[i];
        }
        __syncwarp();
        const ProblemDesc& P = lp.P;
        const int B = P.kf.B;
        if (lane == 0) lp.loss_a = P.state[13];
        for (int i = lane; i < 21 * B; i += 32) (&lp.A[0][0])[i] = P.kf.A[i];
        if (lane < 13) lp.x_eval[lane] = P.state[lane];
        // nothing of this problem is in flight: the counter's current value is the base of this solve
        unsigned target = __shfl_sync(0xffffffffu, ld_acquire_u32(&L.w->done), 0);
        __syncwarp();
        leader_start(lp, L);
        target += (unsigned)B;
        for (;;) {
#ifdef EDS_TIMING
            const unsigned long long t0 = gtime();
#endif
            wait_done(L, target);
#ifdef EDS_TIMING
            const unsigned long long t1 = gtime();
#endif
            const int next = leader_step(lp, L);
#ifdef EDS_TIMING
            t_wait += t1 - t0; t_work += gtime() - t1; n_work++;
#endif
            if (next == CMD_EVAL) { target += (unsigned)B; continue; }
            if (next == CMD_FINAL) {  // residual write-back (Tracker.cpp:223-230): every evaluator warp of every block reports
                target += (unsigned)(B * N_EVAL_WARPS);
                wait_done(L, target);
            }
            break;
        }
        // the last problem of the launch to finish sends every evaluator CTA home
        unsigned fin = 0;
        if (lane == 0) fin = atomicAdd(&ctl->finished, 1u) + 1u;
        fin = __shfl_sync(0xffffffffu, fin, 0);
        if (fin % (unsigned)count == 0u) push_tasks(L, n_eval_ctas, CMD_EXIT);
    }
#ifdef EDS_TIMING
    if (lane == 0) { atomicAdd(&g_timing[0], t_work); atomicAdd(&g_timing[1], t_wait); atomicAdd(&g_timing[2], n_work); }
#endif
}

// ------------------------------------------------------------------------------------------
// evaluator CTA
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void init_barriers(EvalShared& sh) {
    if (threadIdx.x < N_SLOTS) {
        mbar_init(&sh.full_bar[threadIdx.x], 1);   // the elected lane of the producer warp that filled the slot
        mbar_init(&sh.empty_bar[threadIdx.x], 2);  // the elected lanes of the two consumer warps (row halves) that drained it
    }
    if (threadIdx.x < MAILBOX) {
        mbar_init(&sh.task_full[threadIdx.x], 1);              // the elected lane of the control warp
        mbar_init(&sh.task_empty[threadIdx.x], N_EVAL_WARPS);  // one elected lane per evaluator warp
    }
    if (threadIdx.x < 2) {
        mbar_init(&sh.part_full[threadIdx.x], N_CONS_WARPS);   // one elected lane per consumer warp
        mbar_init(&sh.part_empty[threadIdx.x], 1);             // the elected lane of the combiner warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
}

// Control warp of an evaluator CTA: takes tickets from the global queue, waits for the ticket's entry, copies the task
// (problem descriptor + published constants) into the next mailbox slot and hands it to the evaluator warps.  It runs
// MAILBOX - 1 tasks ahead of them, so the three global round trips of a fetch are off the evaluators' path while the
// queue has work.
__device__ void control_warp_main(EvalShared& sh, const ProblemDesc* __restrict__ problems, ProblemWork* work, QueueCtl* ctl,
                                  const unsigned long long* entries, unsigned mask) {
    const int lane = threadIdx.x & 31;
    for (unsigned i = 0;; ++i) {
        const unsigned slot = i % MAILBOX, use = i / MAILBOX;
        if (use > 0) mbar_wait(&sh.task_empty[slot], (use - 1u) & 1u, 5);
#ifdef EDS_TIMING
        const unsigned long long tf0 = gtime();
#endif
        unsigned ticket = 0;
        if (lane == 0) ticket = atomicAdd(&ctl->head, 1u);
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        const unsigned long long* e = &entries[ticket & mask];
        unsigned payload;
        for (;;) {  // every lane polls (one broadcast request): the warp stays converged
            const unsigned long long v = ld_acquire_u64(e);
            const unsigned tag = __shfl_sync(0xffffffffu, (unsigned)(v >> 32), 0);
            payload = __shfl_sync(0xffffffffu, (unsigned)v, 0);
            if (tag == ticket + 1u) break;
            __nanosleep(20);
        }
        const int pid = (int)(payload & 0xfffffu), block = (int)((payload >> 20) & 63u), cmd = (int)(payload >> 26);
        TaskShared& ts = sh.task[slot];
        if (cmd != CMD_EXIT && cmd != CMD_NOP) {
            // one round trip: the descriptor (immutable during the launch) and the whole EvalConst the leader published
            // before the queue entry (acquired above; read past the non-coherent L1)
            ProblemWork* w = work + pid;
            const int* src = reinterpret_cast<const int*>(&problems[pid]);
            int* dst = reinterpret_cast<int*>(&ts.P);
            const int* es = reinterpret_cast<const int*>(&w->ec);
            int* ed = reinterpret_cast<int*>(&ts.ec);
            constexpr int ND = (int)(sizeof(ProblemDesc) / sizeof(int)), NE = (int)(sizeof(EvalConst) / sizeof(int));
            int dv[(ND + 31) / 32], evv[(NE + 31) / 32];
#pragma unroll
            for (int k = 0; k < (ND + 31) / 32; ++k) dv[k] = (lane + 32 * k < ND) ? __ldg(src + lane + 32 * k) : 0;
#pragma unroll
            for (int k = 0; k < (NE + 31) / 32; ++k) evv[k] = (lane + 32 * k < NE) ? __ldcg(es + lane + 32 * k) : 0;
            const double loss_a = __ldcg(&w->loss_a);
            const double inv_norm = __ldcg(&w->inv_norm);
#pragma unroll
            for (int k = 0; k < (ND + 31) / 32; ++k) if (lane + 32 * k < ND) dst[lane + 32 * k] = dv[k];
#pragma unroll
            for (int k = 0; k < (NE + 31) / 32; ++k) if (lane + 32 * k < NE) ed[lane + 32 * k] = evv[k];
            __syncwarp();
            if (lane == 0) {
                ts.loss_a = loss_a;
                ts.slots = &w->slots[0][0];
                ts.done = &w->done;
                ts.inv_norm = (float)inv_norm;
                ts.block = block;
            }
        }
        if (lane == 0) ts.cmd = cmd;
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.task_full[slot]);
#ifdef EDS_TIMING
        if (lane == 0 && g_timing_cta == (int)blockIdx.x) { atomicAdd(&g_timing[20], gtime() - tf0); atomicAdd(&g_timing[21], 1ull); }
#endif
        if (cmd == CMD_EXIT) break;
    }
}

// Batched Levenberg-Marquardt solve as a dataflow over global memory.  The first n_lead CTAs are LEADER CTAs: each of
// their warps runs the sequential LM state machine of one problem at a time (decision, 12x12 Cholesky, retraction) and
// turns every evaluation it needs into one task per residual block in a global queue.  All other CTAs are EVALUATORS:
// their control warp pops tasks, their producer / consumer warps sweep the block and report the block sums to the
// problem's slots in global memory.  Any evaluator serves any problem, so every SM stays busy whatever the number of
// problems, and the serial step of a problem (~8 us) is hidden behind the sweeps of all the others.  Summation orders are
// fixed per block and per problem: results do not depend on the launch shape or on the schedule.
__global__ void __launch_bounds__(TRK_THREADS, 1) track_lm_kernel(const ProblemDesc* __restrict__ problems, ProblemWork* work, QueueCtl* ctl,
                                                                  unsigned long long* entries, unsigned mask, int count, int n_lead) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5;
    if ((int)blockIdx.x < n_lead) {
        LeadShared& ls = *reinterpret_cast<LeadShared*>(smem_raw);
        if (warp >= LEAD_WARPS) return;
        // problems are dealt to the leader CTAs round-robin: a small batch spreads over all of them (fewer warps share an SM)
        leader_warp_main(ls.prob[warp], problems, work, ctl, entries, mask, count, warp * n_lead + (int)blockIdx.x, n_lead * LEAD_WARPS,
                         (int)gridDim.x - n_lead);
        return;
    }
    EvalShared& sh = *reinterpret_cast<EvalShared*>(smem_raw);
    init_barriers(sh);
    if (warp == CTRL_WARP) {
        control_warp_main(sh, problems, work, ctl, entries, mask);
        return;
    }
    const Roles role = make_roles();
    if (role.pidx >= 0) producer_warp_main(sh, role.pidx);
    else if (role.cidx < 0) combiner_warp_main(sh);
    else if (role.half == 0) consumer_warp_main<0>(sh, role.cidx);
    else consumer_warp_main<1>(sh, role.cidx);
}

// parity/debug entry (edsgpu_tracker_evaluate): one full evaluation at P.state, residuals and Jacobian rows written out,
// reduced normal equations returned.  One CTA per residual block; the CTA that finishes last sums the blocks.
__global__ void __launch_bounds__(TRK_THREADS, 1) track_eval_kernel(const ProblemDesc* __restrict__ problems, ProblemWork* work) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EvalShared& sh = *reinterpret_cast<EvalShared*>(smem_raw);
    __shared__ double A_sh[MAX_BLOCKS][21];
    __shared__ double x_sh[13];
    TaskShared& ts = sh.task[0];
    const int tid = threadIdx.x, warp = tid >> 5;
    {
        const int* src = reinterpret_cast<const int*>(&problems[0]);
        int* dst = reinterpret_cast<int*>(&ts.P);
        for (int i = tid; i < (int)(sizeof(ProblemDesc) / sizeof(int)); i += TRK_THREADS) dst[i] = src[i];
    }
    init_barriers(sh);  // ends with a CTA barrier
    const ProblemDesc& P = ts.P;
    const int B = P.kf.B;
    for (int i = tid; i < 21 * B; i += TRK_THREADS) (&A_sh[0][0])[i] = P.kf.A[i];
    if (tid < 13) x_sh[tid] = P.state[tid];
    if (tid == 0) {
        ts.loss_a = P.state[13];
        ts.slots = &work->slots[0][0];
        ts.done = &work->done;  // zeroed before the launch: the block that counts B - 1 before itself is the last one
        ts.inv_norm = (float)P.norms[1];
        ts.block = (int)blockIdx.x;
        ts.cmd = CMD_EVAL;
        sh.task[1].cmd = CMD_EXIT;  // the stream of this CTA: one evaluation, then the end
    }
    __syncthreads();
    if (warp == CTRL_WARP) {
        compute_eval_const(ts.ec, x_sh, A_sh, B, CMD_EVAL);
        if ((tid & 31) == 0) { mbar_arrive(&sh.task_full[0]); mbar_arrive(&sh.task_full[1]); }
        return;
    }
    const Roles role = make_roles();
    if (role.pidx >= 0) producer_warp_main(sh, role.pidx);
    else if (role.cidx < 0) combiner_warp_main(sh);
    else if (role.half == 0) consumer_warp_main<0>(sh, role.cidx);
    else consumer_warp_main<1>(sh, role.cidx);
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
                                                         float4* __restrict__ ga, float4* __restrict__ gb, double* __restrict__ kpx,
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
        ga[idx] = make_float4((float)g[0], (float)g[1], (float)g[2], (float)g[3]);
        gb[idx] = make_float4((float)g[4], (float)g[5], (float)w, 0.f);
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
    int cached_slot = -1, cached_level = -1;
    std::vector<int> level_iterations;  // config.options.max_num_iterations[id] (Tracker.cpp:139); empty = cfg.max_iterations at every level
};

struct LaunchShape { int n_lead, n_eval; };  // leader CTAs, evaluator CTAs

struct edsgpu_batch {
    edsgpu_ctx* ctx = nullptr;
    int count = 0;
    LaunchShape shape{1, 1};
    void* work_block = nullptr;  // ProblemWork[count] | QueueCtl | queue entries (device, zeroed once: the counters only grow)
    unsigned queue_mask = 0;
    const edsgpu_frames* frames = nullptr;  // slots first_slot .. first_slot + count - 1
    int first_slot = 0;
    int level = 0;                          // pyramid level of the frames the batch samples (Tracker::optimize(id, ...))
    ProblemDesc* desc = nullptr;  // device
    std::vector<edsgpu_tracker*> trackers;
    std::vector<const edsgpu_keyframe*> keyframes;  // must outlive the batch (the descriptors hold their device arrays)
    std::vector<uint64_t> res_generation;           // trackers[i]->res_generation the descriptors were made with
};

namespace {

constexpr size_t kLmSmem = sizeof(EvalShared) > sizeof(LeadShared) ? sizeof(EvalShared) : sizeof(LeadShared);

int env_int(const char* name, int lo, int hi) {  // tuning / debug overrides; -1 = not set
    const char* e = getenv(name);
    if (!e) return -1;
    const int v = atoi(e);
    return (v >= lo && v <= hi) ? v : -1;
}

// Shape of a batched launch.  Leader CTAs: one warp per problem in flight (up to MAX_LEAD_CTAS x 16 problems at once,
// further ones follow as warps finish).  Evaluator CTAs: all remaining SMs, but no more than there can be tasks in flight
// (one per residual block and problem).  EDSGPU_RESERVE_SMS leaves SMs to kernels of other streams (event-frame builds
// of the next window); EDSGPU_LEADER_CTAS / EDSGPU_EVAL_CTAS force the two counts.  Results do not depend on the shape.
LaunchShape pick_shape(edsgpu_ctx* ctx, int count, int B) {
    // four problems per leader CTA until all MAX_LEAD_CTAS are in use (leader warps that share an SM slow each other down;
    // measured on the 64-window batch: 4 CTAs 0.603 ms, 8 CTAs 0.582 ms, 12 CTAs 0.596 ms -- beyond 8 the evaluators miss the SMs)
    int n_lead = std::max(1, std::min((count + 3) / 4, MAX_LEAD_CTAS));
    const int force_lead = env_int("EDSGPU_LEADER_CTAS", 1, 64);
    if (force_lead > 0) n_lead = force_lead;
    const int reserve = std::max(0, env_int("EDSGPU_RESERVE_SMS", 0, 1024));
    const int in_flight = std::min(count, n_lead * LEAD_WARPS);
    int n_eval = std::max(1, std::min(ctx->num_sms - n_lead - reserve, in_flight * B));
    const int force_eval = env_int("EDSGPU_EVAL_CTAS", 1, 4096);
    if (force_eval > 0) n_eval = force_eval;
    return LaunchShape{n_lead, n_eval};
}

ProblemDesc make_desc(const edsgpu_tracker* tr, const edsgpu_keyframe* kf, const edsgpu_frames* frames, int slot, int level) {
    ProblemDesc d{};
    d.kf = kf->dev;
    const size_t img = (size_t)level * frames->capacity + slot;  // event_frame[id] of the slot
    d.frame = frames->tex[img];
    d.norms = frames->norms + 2 * img;
    d.state = tr->state;
    d.residuals = tr->residuals;
    d.info = tr->info;
    d.loss_type = tr->cfg.loss_type;
    d.max_iter = level < (int)tr->level_iterations.size() ? tr->level_iterations[level] : tr->cfg.max_iterations;
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
        hd[i] = make_desc(b->trackers[i], b->keyframes[i], b->frames, b->first_slot + i, b->level);
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
    EDS_RANGE("edsgpu_keyframe_create");
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
    // layout: ga | gb | kpx | kpy | kpz | A
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
    const size_t o_gxy = take(N * sizeof(float4)), o_dw = take(N * sizeof(float4)), o_kx = take(N * 8), o_ky = take(N * 8), o_kz = take(N * 8);
    const size_t o_A = take((size_t)num_blocks * 21 * 8);
    const size_t o_src = take(N * 6 * sizeof(double));  // kept: the inverse depths can be refreshed on the device
    cudaError_t e = cudaMalloc(&kf->block, off);
    if (e != cudaSuccess) { delete kf; return edsgpu_fail(ctx, EDSGPU_OUT_OF_MEMORY, cudaGetErrorString(e)); }
    char* base = (char*)kf->block;
    KfDev& d = kf->dev;
    d.ga = (float4*)(base + o_gxy); d.gb = (float4*)(base + o_dw);
    d.kpx = (double*)(base + o_kx); d.kpy = (double*)(base + o_ky); d.kpz = (double*)(base + o_kz);
    d.A = (double*)(base + o_A);
    kf->src = (double*)(base + o_src);
    d.N = num_points; d.B = num_blocks; d.H = height; d.W = width;
    d.ne = num_points / num_blocks;
    d.fx = fx; d.fy = fy; d.cx = cx; d.cy = cy;
    d.fxf = (float)fx; d.fyf = (float)fy;
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
                                                                (float4*)d.ga, (float4*)d.gb, (double*)d.kpx, (double*)d.kpy, (double*)d.kpz,
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
    return edsgpu_batch_create_level(ctx, trackers, keyframes, count, frames, first_slot, 0, out);
}

edsgpu_status edsgpu_batch_create_level(edsgpu_ctx* ctx, edsgpu_tracker* const* trackers, const edsgpu_keyframe* const* keyframes, int count,
                                        const edsgpu_frames* frames, int first_slot, int level, edsgpu_batch** out) {
    if (!ctx || !out) return EDSGPU_INVALID_ARGUMENT;
    EDS_REQUIRE(ctx, trackers && keyframes && count > 0, "batch_create: bad arguments");
    EDS_REQUIRE(ctx, frames && level >= 0 && level < frames->levels, "batch_create: the frames have no such pyramid level");
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
    b->level = level;
    b->trackers.assign(trackers, trackers + count);
    b->keyframes.assign(keyframes, keyframes + count);
    cudaError_t e = cudaMalloc(&b->desc, sizeof(ProblemDesc) * (size_t)count);
    if (e == cudaSuccess) {
        // queue capacity: every problem in flight has at most one task per block outstanding, plus the exit tasks; the
        // evaluators hold at most one unfilled ticket each
        const size_t in_flight = (size_t)std::min(count, b->shape.n_lead * LEAD_WARPS);
        size_t cap = 64;
        while (cap < in_flight * MAX_BLOCKS + 2 * (size_t)b->shape.n_eval + 64) cap <<= 1;
        b->queue_mask = (unsigned)(cap - 1);
        const size_t bytes = sizeof(ProblemWork) * (size_t)count + sizeof(QueueCtl) + sizeof(unsigned long long) * cap;
        e = cudaMalloc(&b->work_block, bytes);
        if (e == cudaSuccess) e = cudaMemsetAsync(b->work_block, 0, bytes, ctx->stream);
    }
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
    if (b->work_block) cudaFree(b->work_block);
    delete b;
}

edsgpu_status edsgpu_batch_optimize(edsgpu_batch* b) {
    EDS_RANGE("edsgpu_batch_optimize");
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
    {
        ProblemWork* work = (ProblemWork*)b->work_block;
        QueueCtl* ctl = (QueueCtl*)(work + b->count);
        unsigned long long* entries = (unsigned long long*)(ctl + 1);
        EDS_CUDA(ctx, cudaFuncSetAttribute(track_lm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLmSmem));
        track_lm_kernel<<<b->shape.n_lead + b->shape.n_eval, TRK_THREADS, kLmSmem, ctx->stream>>>((const ProblemDesc*)b->desc, work, ctl, entries,
                                                                                                b->queue_mask, b->count, b->shape.n_lead);
        EDS_CUDA(ctx, cudaGetLastError());
        ctx->launches++;
    }
    st = edsgpu_frames_mark_read(b->frames, b->first_slot, b->count, ctx->stream);
    if (st != EDSGPU_OK) return st;
    mad_kernel<<<b->count, MAD_THREADS, 0, ctx->stream>>>((const ProblemDesc*)b->desc);
    ctx->launches++;
    EDS_CUDA(ctx, cudaGetLastError());
    return EDSGPU_OK;
}

edsgpu_status edsgpu_batch_launch_shape(const edsgpu_batch* b, int* evaluator_ctas, int* leader_ctas, int* problems_in_flight) {
    if (!b) return EDSGPU_INVALID_ARGUMENT;
    if (evaluator_ctas) *evaluator_ctas = b->shape.n_eval;
    if (leader_ctas) *leader_ctas = b->shape.n_lead;
    if (problems_in_flight) *problems_in_flight = std::min(b->count, b->shape.n_lead * LEAD_WARPS);
    return EDSGPU_OK;
}

edsgpu_status edsgpu_batch_count(const edsgpu_batch* b, int* count) {
    if (!b || !count) return EDSGPU_INVALID_ARGUMENT;
    *count = b->count;
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
    return edsgpu_tracker_optimize_level(tr, kf, frames, slot, 0, px, qx, vx, residuals_out, next_loss_param_out, info);
}

edsgpu_status edsgpu_tracker_set_level_iterations(edsgpu_tracker* tr, const int* max_num_iterations, int num_levels) {
    if (!tr) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = tr->ctx;
    EDS_REQUIRE(ctx, num_levels >= 0 && (num_levels == 0 || max_num_iterations), "tracker_set_level_iterations: bad arguments");
    for (int i = 0; i < num_levels; ++i) EDS_REQUIRE(ctx, max_num_iterations[i] >= 0, "tracker_set_level_iterations: negative iteration cap");
    tr->level_iterations.assign(max_num_iterations, max_num_iterations + num_levels);
    if (tr->cached) {  // its descriptor carries the old cap
        edsgpu_batch_destroy(tr->cached);
        tr->cached = nullptr;
    }
    return EDSGPU_OK;
}

edsgpu_status edsgpu_tracker_optimize_level(edsgpu_tracker* tr, const edsgpu_keyframe* kf, const edsgpu_frames* frames, int slot, int level,
                                            double px[3], double qx[4], double vx[6], double* residuals_out, double* next_loss_param_out,
                                            edsgpu_tracker_info* info) {
    EDS_RANGE("edsgpu_tracker_optimize");
    if (!tr) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = tr->ctx;
    EDS_REQUIRE(ctx, kf && frames, "tracker_optimize: null handle");
    if (!tr->cached || tr->cached_kf != kf->uid || tr->cached_frames != frames->uid || tr->cached_slot != slot || tr->cached_level != level) {
        if (tr->cached) edsgpu_batch_destroy(tr->cached);
        tr->cached = nullptr;
        const edsgpu_keyframe* kfs[1] = {kf};
        edsgpu_tracker* trs[1] = {tr};
        edsgpu_status stc = edsgpu_batch_create_level(ctx, trs, kfs, 1, frames, slot, level, &tr->cached);
        if (stc != EDSGPU_OK) return stc;
        tr->cached_kf = kf->uid; tr->cached_frames = frames->uid; tr->cached_slot = slot; tr->cached_level = level;
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
    const size_t o_work = take(sizeof(ProblemWork));
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
    EDS_CUDA(ctx, cudaMemsetAsync(ds + o_work, 0, sizeof(ProblemWork), ctx->stream));  // the last block to finish sums: counter from 0
    st = edsgpu_frames_wait_built(frames, slot, 1, ctx->stream);
    if (st == EDSGPU_OK) {
        EDS_CUDA(ctx, cudaFuncSetAttribute(track_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EvalShared)));
        track_eval_kernel<<<kf->dev.B, TRK_THREADS, sizeof(EvalShared), ctx->stream>>>((const ProblemDesc*)(ds + o_desc), (ProblemWork*)(ds + o_work));
        EDS_CUDA(ctx, cudaGetLastError());
        ctx->launches++;
    }
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
                                                                           kf->dev.cx, kf->dev.cy, kf->dev.W, kf->dev.H, coord_dev, dp->outlier);
    ctx->launches++;
    EDS_CUDA(ctx, cudaGetLastError());
    if (coord_out) EDS_CUDA(ctx, cudaMemcpyAsync(coord_out, coord_dev, sizeof(double) * 2 * N, cudaMemcpyDeviceToHost, ctx->stream));
    if (outlier_out) EDS_CUDA(ctx, cudaMemcpyAsync(outlier_out, dp->outlier, N, cudaMemcpyDeviceToHost, ctx->stream));
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
    kf_prepare_kernel<<<d.B, 256, 0, ctx->stream>>>(kf->src, kf->src + 2 * N, dp->state, 4, kf->src + 5 * N, d.N, d.B, (float4*)d.ga, (float4*)d.gb,
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
