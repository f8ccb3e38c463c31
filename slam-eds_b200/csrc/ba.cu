// DSO-derived windowed-BA Hessian accumulation on the device.
//
// Replaces AccumulatedTopHessianSSE::addPoint<mode> + stitchDoubleMT (reference
// src/bundles/AccumulatedTopHessian.{h,cpp}), AccumulatedSCHessianSSE::addPoint + stitchDoubleMT
// (src/bundles/AccumulatedSCHessian.{h,cpp}), the AccumulatorApprox / AccumulatorXX float
// accumulators (src/bundles/MatrixAccumulators.h) and EFResidual::takeDataF
// (src/bundles/EnergyFunctionalStructs.cpp:38-48).
//
//   plan (host, once per graph)   residuals bucketed by (host,target) with a stable counting sort
//                                 and cut into single-key tiles: every CTA reduces into ONE 13x13
//                                 accumulator, no atomics, fixed summation order (the reference's
//                                 6-thread work stealing makes its float sums differ run to run).
//   ba_top_kernel<mode>           one thread per residual record (19 x 128-bit loads of the 304-byte
//                                 RawResidualJacobian), 91 fp32 accumulators per thread, halving warp
//                                 butterfly, fp64 across warps -> one partial per tile; per-residual
//                                 Hdd/bd/Hcd terms written for the point pass.
//   ba_top_finalize_kernel        ordered fp64 sum of the tile partials -> F*F x 13x13 doubles.
//   ba_point_sum_kernel           per point: Hdd_acc, bd_acc, Hcd_acc over its residuals (in order).
//   ba_sc_kernel                  per host frame the Schur terms are one weighted Gram matrix
//                                 sum_p HdiF_p [z_p; Hcd_p] [z_p; Hcd_p; bdSum_p]^T with z_p the
//                                 stacked JpJdF of the point's active residuals: register-tiled 4x4
//                                 through shared memory, chunked over points, ordered fp64 finalize.
//   ba_*_stitch_kernel            one CTA per 8x8 output block of H (adjoint sandwiches in fp64).
#include <algorithm>
#include <vector>

#include "common.cuh"
#include <cuda_pipeline.h>

namespace {

constexpr int REC = 76;  // floats per record: dso::RawResidualJacobian as laid out by Eigen (304 B)
constexpr int O_RES = 0, O_JPDXI0 = 8, O_JPDXI1 = 14, O_JPDC0 = 20, O_JPDC1 = 24, O_JPDD = 28, O_JIDX0 = 32, O_JIDX1 = 40,
              O_JAB0 = 48, O_JAB1 = 56, O_JIDX2 = 64, O_JABJIDX = 68, O_JAB2 = 72;
constexpr int CPARS = 4;
constexpr int TOP_THREADS = 256, NACC = 96;
constexpr int TOP_STAGE = 512;  // records staged in shared memory per round (512 x 304 B = 152 KB)
constexpr int MAXF = 8;
constexpr int SC_THREADS = 256, SC_CHUNK = 64, SC_BATCH = 32, SC_LD = 72;  // 8*MAXF + 5 padded to 72

struct BaDev {
    int F, P, R, num_tiles;
    const int32_t *host_idx, *target_idx, *point_of_res, *res_begin, *perm, *tile_key, *tile_start, *tile_count, *key_tile_begin;
    const float *recs, *res_toZero, *deltaF, *priorF, *adHTdeltaF, *cDeltaF;
    const uint8_t* flags;
    float* res_pt;        // R x 6: bd, Hdd, Hcd[4] of each residual (zero when filtered out)
    double* tile_partial; // num_tiles x 96
    int* tile_nres;       // num_tiles
    // second phase of the accumulation kernels (after a grid-wide barrier): ordered sums over tiles and over a point's residuals
    double* acc_out;      // F*F x 13x13 (AccumulatorApprox::finish layout)
    long long* num_out;   // F*F
    float *Hdd_out, *bd_out, *Hcd_out;  // P, P, P x 4
    unsigned* grid_bar;   // {arrivals, generation} of the grid barrier
};

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
__device__ __forceinline__ int butterfly_base(unsigned lane) {
    return 48 * ((lane >> 4) & 1) + 24 * ((lane >> 3) & 1) + 12 * ((lane >> 2) & 1) + 6 * ((lane >> 1) & 1) + 3 * (lane & 1);
}

// ------------------------------------------------------------------------------------------
// PointFrameResidual::linearize (src/tracking/Residuals.cpp:69-265), SURVEY.md 8 a12 / 8(f) rank 1:
// the feeder of the accumulators on the device, one thread per residual.  The records never cross
// PCIe: what is uploaded per linearisation is the point state (80 B/point) and the frame-frame
// precalc (112 B/pair) instead of 304 B/residual.  Arithmetic is float32 in the reference's
// operation order with explicit round-to-nearest mul/add (no FMA contraction), so the records are
// reproducible bit for bit by a plain float32 CPU evaluation.  Images are kept as float4 {I, dx, dy, 0} per pixel: one 128-bit
// load per bilinear tap.
// ------------------------------------------------------------------------------------------
constexpr int PRECALC = EDSGPU_PRECALC_FLOATS;
__constant__ int c_pattern[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};  // settings.cpp:276

__device__ __forceinline__ float fm(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fa(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fs(float a, float b) { return __fadd_rn(a, -b); }

struct LinDev {
    int F, P, R, H, W;
    const float4* images;   // [F][H*W]
    const float* precalc;   // [F*F][PRECALC]
    float fxl, fyl, cxl, cyl;
    const float* frame_energy_th;
    const float *pu, *pv, *idepth_zero, *idepth, *color, *weights;
    const int32_t *host_idx, *target_idx, *point_of_res;
    const uint8_t* state_in;    // or null
    const uint8_t* linearized;  // or null
    float* recs;
    uint8_t* flags;
    int32_t* state_new;
    float* energy_new;
};

// one residual: record (zero unless the residual is IN or an OUTLIER), new state (0 IN, 1 OOB, 2 OUTLIER) and energy
__device__ __forceinline__ void linearize_one(const LinDev& d, int r, float* __restrict__ rec, int& state, float& energy) {
#pragma unroll
    for (int i = 0; i < REC; ++i) rec[i] = 0.f;
    state = 1;  // OOB unless the residual survives
    energy = 0.f;
    const int p = d.point_of_res[r];
    const int h = d.host_idx[r], t = d.target_idx[r];
    const bool skip = d.state_in && d.state_in[r] == 1;  // :73-74
    if (!skip) {
        const float* pc = d.precalc + (size_t)PRECALC * (h + d.F * t);
        const float fxl = d.fxl, fyl = d.fyl, cxl = d.cxl, cyl = d.cyl;
        const float fxli = __fdiv_rn(1.0f, fxl), fyli = __fdiv_rn(1.0f, fyl);
        const float wM3G = (float)(d.W - 3), hM3G = (float)(d.H - 3);  // globalCalib.cpp:74-75
        const float u0 = d.pu[p], v0 = d.pv[p], idz = d.idepth_zero[p], ids = d.idepth[p];
        // centre point, ResidualProjections.h:60-86: ptp = R KliP + t idepth; 3x3 blocks are column-major
        const float K0 = fm(fs(u0, cxl), fxli), K1 = fm(fs(v0, cyl), fyli);
        float ptp[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) ptp[i] = fa(fa(fa(fm(__ldg(pc + i), K0), fm(__ldg(pc + 3 + i), K1)), __ldg(pc + 6 + i)), fm(__ldg(pc + 9 + i), idz));
        const float drescale = __fdiv_rn(1.0f, ptp[2]);
        const float new_idepth = fm(idz, drescale);
        const float u = fm(ptp[0], drescale), v = fm(ptp[1], drescale);
        const float Ku0 = fa(fm(u, fxl), cxl), Kv0 = fa(fm(v, fyl), cyl);
        bool ok = (drescale > 0) && Ku0 > 1.1f && Kv0 > 1.1f && Ku0 < wM3G && Kv0 < hM3G;
        if (ok) {
            const float R00 = __ldg(pc + 0), R10 = __ldg(pc + 1), R20 = __ldg(pc + 2), R01 = __ldg(pc + 3), R11 = __ldg(pc + 4), R21 = __ldg(pc + 5);
            const float t0 = __ldg(pc + 9), t1 = __ldg(pc + 10), t2 = __ldg(pc + 11);
            // :108-147 (SCALE_IDEPTH = SCALE_F = SCALE_C = 1)
            rec[O_JPDD] = fm(fm(drescale, fs(t0, fm(t2, u))), fxl);
            rec[O_JPDD + 1] = fm(fm(drescale, fs(t1, fm(t2, v))), fyl);
            const float dCx2 = fm(drescale, fs(fm(R20, u), R00));
            const float dCx3 = fm(fm(fm(fxl, drescale), fs(fm(R21, u), R01)), fyli);
            const float dCy2 = fm(fm(fm(fyl, drescale), fs(fm(R20, v), R10)), fxli);
            const float dCy3 = fm(drescale, fs(fm(R21, v), R11));
            rec[O_JPDC0 + 0] = fa(fm(K0, dCx2), u);
            rec[O_JPDC0 + 1] = fm(K1, dCx3);
            rec[O_JPDC0 + 2] = fa(dCx2, 1.0f);
            rec[O_JPDC0 + 3] = dCx3;
            rec[O_JPDC1 + 0] = fm(K0, dCy2);
            rec[O_JPDC1 + 1] = fa(fm(K1, dCy3), v);
            rec[O_JPDC1 + 2] = dCy2;
            rec[O_JPDC1 + 3] = fa(dCy3, 1.0f);
            rec[O_JPDXI0 + 0] = fm(new_idepth, fxl);
            rec[O_JPDXI0 + 2] = fm(fm(-new_idepth, u), fxl);
            rec[O_JPDXI0 + 3] = fm(fm(-u, v), fxl);
            rec[O_JPDXI0 + 4] = fm(fa(1.0f, fm(u, u)), fxl);
            rec[O_JPDXI0 + 5] = fm(-v, fxl);
            rec[O_JPDXI1 + 1] = fm(new_idepth, fyl);
            rec[O_JPDXI1 + 2] = fm(fm(-new_idepth, v), fyl);
            rec[O_JPDXI1 + 3] = fm(-fa(1.0f, fm(v, v)), fyl);
            rec[O_JPDXI1 + 4] = fm(fm(u, v), fyl);
            rec[O_JPDXI1 + 5] = fm(u, fyl);
            const float aff0 = __ldg(pc + 24), aff1 = __ldg(pc + 25), b0 = __ldg(pc + 26);
            const float4* img = d.images + (size_t)t * d.H * d.W;
            float J00 = 0, J11 = 0, J10 = 0, A00 = 0, A01 = 0, A10 = 0, A11 = 0, B00 = 0, B01 = 0, B11 = 0, wJI2 = 0;
            const float TH = 2500.0f, HUBER = 9.0f;  // settings.cpp:91,127
#pragma unroll
            for (int idx = 0; idx < 8; ++idx) {
                // ResidualProjections.h:46-56
                const float up = fa(u0, (float)c_pattern[idx][0]), vp = fa(v0, (float)c_pattern[idx][1]);
                float q[3];
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    q[i] = fa(fa(fa(fm(__ldg(pc + 12 + i), up), fm(__ldg(pc + 15 + i), vp)), __ldg(pc + 18 + i)), fm(__ldg(pc + 21 + i), ids));
                const float Ku = __fdiv_rn(q[0], q[2]), Kv = __fdiv_rn(q[1], q[2]);
                const bool inb = Ku > 1.1f && Kv > 1.1f && Ku < wM3G && Kv < hM3G;
                ok = ok && inb;
                // getInterpolatedElement33, globalFuncs.h:78-92 (a safe pixel once the residual is lost)
                const float Kus = inb ? Ku : 4.0f, Kvs = inb ? Kv : 4.0f;
                const int ix = (int)Kus, iy = (int)Kvs;
                const float dx = fs(Kus, (float)ix), dy = fs(Kvs, (float)iy), dxdy = fm(dx, dy);
                const float4* bp = img + (ix + iy * d.W);
                const float4 p00 = __ldg(bp), p10 = __ldg(bp + 1), p01 = __ldg(bp + d.W), p11 = __ldg(bp + 1 + d.W);
                const float w11 = dxdy, w01 = fs(dy, dxdy), w10 = fs(dx, dxdy), w00 = fa(fs(fs(1.0f, dx), dy), dxdy);
                const float hit0 = fa(fa(fa(fm(w11, p11.x), fm(w01, p01.x)), fm(w10, p10.x)), fm(w00, p00.x));
                float hit1 = fa(fa(fa(fm(w11, p11.y), fm(w01, p01.y)), fm(w10, p10.y)), fm(w00, p00.y));
                float hit2 = fa(fa(fa(fm(w11, p11.z), fm(w01, p01.z)), fm(w10, p10.z)), fm(w00, p00.z));
                const float col = d.color[8 * p + idx];
                const float residual = fs(hit0, fa(fm(aff0, col), aff1));
                const float drdA = fs(col, b0);
                ok = ok && isfinite(hit0);
                float w = __fsqrt_rn(__fdiv_rn(TH, fa(TH, fa(fm(hit1, hit1), fm(hit2, hit2)))));
                w = fm(0.5f, fa(w, d.weights[8 * p + idx]));
                const float ar = fabsf(residual);
                float hw = ar < HUBER ? 1.0f : __fdiv_rn(HUBER, ar);
                energy = fa(energy, fm(fm(fm(fm(fm(w, w), hw), residual), residual), fs(2.0f, hw)));
                if (hw < 1.0f) hw = __fsqrt_rn(hw);
                hw = fm(hw, w);
                hit1 = fm(hit1, hw);
                hit2 = fm(hit2, hw);
                rec[O_RES + idx] = fm(residual, hw);
                rec[O_JIDX0 + idx] = hit1;
                rec[O_JIDX1 + idx] = hit2;
                const float dh = fm(drdA, hw);
                rec[O_JAB0 + idx] = dh;
                rec[O_JAB1 + idx] = hw;
                J00 = fa(J00, fm(hit1, hit1)); J11 = fa(J11, fm(hit2, hit2)); J10 = fa(J10, fm(hit1, hit2));
                A00 = fa(A00, fm(dh, hit1)); A01 = fa(A01, fm(dh, hit2)); A10 = fa(A10, fm(hw, hit1)); A11 = fa(A11, fm(hw, hit2));
                B00 = fa(B00, fm(fm(fm(drdA, drdA), hw), hw)); B01 = fa(B01, fm(dh, hw)); B11 = fa(B11, fm(hw, hw));
                wJI2 = fa(wJI2, fm(fm(hw, hw), fa(fm(hit1, hit1), fm(hit2, hit2))));
            }
            if (ok) {
                // Mat22f members are column-major: (0,0) (1,0) (0,1) (1,1)
                rec[O_JIDX2 + 0] = J00; rec[O_JIDX2 + 1] = J10; rec[O_JIDX2 + 2] = J10; rec[O_JIDX2 + 3] = J11;
                rec[O_JABJIDX + 0] = A00; rec[O_JABJIDX + 1] = A10; rec[O_JABJIDX + 2] = A01; rec[O_JABJIDX + 3] = A11;
                rec[O_JAB2 + 0] = B00; rec[O_JAB2 + 1] = B01; rec[O_JAB2 + 2] = B01; rec[O_JAB2 + 3] = B11;
                const float th = fmaxf(d.frame_energy_th[h], d.frame_energy_th[t]);  // :253-261
                if (energy > th || wJI2 < 2.0f) { energy = th; state = 2; }
                else state = 0;
            }
        }
        if (state == 1) {  // lost on the way: no Jacobian (the reference leaves J stale and never reads it)
#pragma unroll
            for (int i = 0; i < REC; ++i) rec[i] = 0.f;
            energy = 0.f;
        }
    }
}

// EFResidual::takeDataF (EnergyFunctionalStructs.cpp:38-48) from a record in registers
__device__ __forceinline__ void jpjd_one(const float* __restrict__ J, float* __restrict__ out) {
    const float d0 = J[O_JPDD], d1 = J[O_JPDD + 1];
    const float* M = J + O_JIDX2;  // Mat22f column-major
    const float v0 = M[0] * d0 + M[2] * d1;
    const float v1 = M[1] * d0 + M[3] * d1;
#pragma unroll
    for (int i = 0; i < 6; ++i) out[i] = J[O_JPDXI0 + i] * v0 + J[O_JPDXI1 + i] * v1;
    const float* N = J + O_JABJIDX;
    out[6] = N[0] * d0 + N[2] * d1;
    out[7] = N[1] * d0 + N[3] * d1;
}

__global__ void ba_jpjd_kernel(const float* __restrict__ recs, int R, float* __restrict__ JpJdF) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const float* J = recs + (size_t)REC * r;
    float out[8];
    jpjd_one(J, out);
    float4* o = reinterpret_cast<float4*>(JpJdF + (size_t)8 * r);
    o[0] = make_float4(out[0], out[1], out[2], out[3]);
    o[1] = make_float4(out[4], out[5], out[6], out[7]);
}

__global__ void __launch_bounds__(128) ba_linearize_kernel(LinDev d) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= d.R) return;
    float rec[REC];
    int state;
    float energy;
    linearize_one(d, r, rec, state, energy);
    float4* out = reinterpret_cast<float4*>(d.recs + (size_t)REC * r);
#pragma unroll
    for (int i = 0; i < REC / 4; ++i) out[i] = make_float4(rec[4 * i], rec[4 * i + 1], rec[4 * i + 2], rec[4 * i + 3]);
    d.state_new[r] = state;
    d.energy_new[r] = energy;
    d.flags[r] = (uint8_t)((state == 0 ? EDSGPU_RES_ACTIVE : 0) | ((d.linearized && d.linearized[r]) ? EDSGPU_RES_LINEARIZED : 0));
}

// Vec3f image -> float4 {I, dx, dy, 0}
__global__ void ba_pad_image_kernel(const float* __restrict__ src, float4* __restrict__ dst, size_t npix) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npix) dst[i] = make_float4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 0.f);
}

// ------------------------------------------------------------------------------------------
// After the solve: back-substitution, linearised energy, fixLinearization.
// Everything they read (records, JpJdF, Hcd, bdSum, HdiF, flags) is already resident.
// ------------------------------------------------------------------------------------------
// EnergyFunctional::resubstituteFPt (EnergyFunctional.cpp:291-317): one thread per point.
__global__ void ba_resubstitute_kernel(int F, int P, const int32_t* __restrict__ res_begin, const int32_t* __restrict__ host_idx,
                                       const int32_t* __restrict__ target_idx, const uint8_t* __restrict__ flags,
                                       const float* __restrict__ JpJdF, const float* __restrict__ bdSum, const float* __restrict__ HcdA,
                                       const float* __restrict__ HcdL, const float* __restrict__ HdiF, const float* __restrict__ xAd,
                                       const float* __restrict__ cstep, float* __restrict__ step) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    int ngood = 0;
    float b = bdSum[p];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) dot += cstep[k] * (HcdA[4 * p + k] + HcdL[4 * p + k]);
    b -= dot;
    for (int r = res_begin[p]; r < res_begin[p + 1]; ++r) {
        if (!(flags[r] & EDSGPU_RES_ACTIVE)) continue;
        ngood++;
        const float* xa = xAd + 8 * (host_idx[r] * F + target_idx[r]);
        const float4 j0 = *reinterpret_cast<const float4*>(JpJdF + (size_t)8 * r), j1 = *reinterpret_cast<const float4*>(JpJdF + (size_t)8 * r + 4);
        float d = xa[0] * j0.x;
        d += xa[1] * j0.y; d += xa[2] * j0.z; d += xa[3] * j0.w;
        d += xa[4] * j1.x; d += xa[5] * j1.y; d += xa[6] * j1.z; d += xa[7] * j1.w;
        b -= d;
    }
    step[p] = ngood ? -b * HdiF[p] : 0.f;
}

// J * delta of one residual in the image plane (fixLinearizationF / calcLEnergyPt share it)
__device__ __forceinline__ void jp_delta(const float* __restrict__ J, const float* __restrict__ dp, const float* __restrict__ dc, float dd,
                                         float& jx, float& jy) {
    jx = 0.f; jy = 0.f;
    float cx = 0.f, cy = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) { jx += J[O_JPDXI0 + k] * dp[k]; jy += J[O_JPDXI1 + k] * dp[k]; }
#pragma unroll
    for (int k = 0; k < 4; ++k) { cx += J[O_JPDC0 + k] * dc[k]; cy += J[O_JPDC1 + k] * dc[k]; }
    jx = jx + cx + J[O_JPDD] * dd;
    jy = jy + cy + J[O_JPDD + 1] * dd;
}

// EFResidual::fixLinearizationF (EnergyFunctionalStructs.cpp:87-113): res_toZero = resF - J delta, isLinearized = true
__global__ void ba_fix_linearization_kernel(int F, int R, const float* __restrict__ recs, const int32_t* __restrict__ host_idx,
                                            const int32_t* __restrict__ target_idx, const int32_t* __restrict__ point_of_res,
                                            const float* __restrict__ deltaF, const float* __restrict__ adHTdeltaF,
                                            const float* __restrict__ cDeltaF, const uint8_t* __restrict__ select, float* __restrict__ res_toZero,
                                            uint8_t* __restrict__ flags) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    if (select ? !select[r] : !(flags[r] & EDSGPU_RES_ACTIVE)) return;
    const float* J = recs + (size_t)REC * r;
    const float* dp = adHTdeltaF + 8 * (host_idx[r] + F * target_idx[r]);
    float jx, jy;
    jp_delta(J, dp, cDeltaF, deltaF[point_of_res[r]], jx, jy);
    float out[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float v = J[O_RES + i];
        v -= J[O_JIDX0 + i] * jx;
        v -= J[O_JIDX1 + i] * jy;
        v -= J[O_JAB0 + i] * dp[6];
        v -= J[O_JAB1 + i] * dp[7];
        out[i] = v;
    }
    float4* o = reinterpret_cast<float4*>(res_toZero + (size_t)8 * r);
    o[0] = make_float4(out[0], out[1], out[2], out[3]);
    o[1] = make_float4(out[4], out[5], out[6], out[7]);
    flags[r] |= EDSGPU_RES_LINEARIZED;
}

// EnergyFunctional::calcLEnergyPt (EnergyFunctional.cpp:332-392): thread i takes residual i and point i; per-block
// partial sums in double, summed in block order by ba_sum_partials_kernel (deterministic).
constexpr int LE_THREADS = 256;
__global__ void __launch_bounds__(LE_THREADS) ba_lenergy_kernel(int F, int P, int R, const float* __restrict__ recs,
                                                                const int32_t* __restrict__ host_idx, const int32_t* __restrict__ target_idx,
                                                                const int32_t* __restrict__ point_of_res, const uint8_t* __restrict__ flags,
                                                                const float* __restrict__ res_toZero, const float* __restrict__ deltaF,
                                                                const float* __restrict__ priorF, const float* __restrict__ adHTdeltaF,
                                                                const float* __restrict__ cDeltaF, double* __restrict__ partial) {
    __shared__ double wsum[LE_THREADS / 32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double e = 0.0;
    if (i < R && (flags[i] & EDSGPU_RES_ACTIVE) && (flags[i] & EDSGPU_RES_LINEARIZED)) {
        const float* J = recs + (size_t)REC * i;
        const float* dp = adHTdeltaF + 8 * (host_idx[i] + F * target_idx[i]);
        float jx, jy;
        jp_delta(J, dp, cDeltaF, deltaF[point_of_res[i]], jx, jy);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float Jdelta = J[O_JIDX0 + k] * jx;
            Jdelta += J[O_JIDX1 + k] * jy;
            Jdelta += J[O_JAB0 + k] * dp[6];
            Jdelta += J[O_JAB1 + k] * dp[7];
            float r0 = res_toZero[(size_t)8 * i + k];
            r0 = r0 + r0;
            r0 = r0 + Jdelta;
            e += (double)(Jdelta * r0);
        }
    }
    if (i < P) e += (double)(deltaF[i] * deltaF[i] * priorF[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < LE_THREADS / 32; ++w) s += wsum[w];
        partial[blockIdx.x] = s;
    }
}
// (out / cdelta_out point into pinned host memory: the sum and the calibration delta reach the host without a copy)
__global__ void ba_sum_partials_kernel(const double* __restrict__ partial, int n, double* __restrict__ out, const float* __restrict__ cDeltaF,
                                       float* __restrict__ cdelta_out) {
    // one warp, fixed order: lane l sums l, l+32, ...; then a butterfly
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 32) s += partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) *out = s;
    if (threadIdx.x < 4) cdelta_out[threadIdx.x] = cDeltaF[threadIdx.x];
}

// ---- TMA bulk-copy + mbarrier helpers -------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// one 304-byte record: global -> shared through the TMA unit (cp.async.bulk, 16-byte aligned, size % 16 == 0)
__device__ __forceinline__ void tma_load_record(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// AccumulatedTopHessianSSE::addPoint<MODE> (AccumulatedTopHessian.cpp:39-159) over one single-key tile.
// Accumulator order (AccumulatorApprox, MatrixAccumulators.h:595-651): 55 upper-triangle entries of the
// 10x10 [C(4) xi(6)] block, 30 of the 10x3 top-right block [a b r], 6 of the 3x3 bottom-right block.
//
// The records a tile needs are gathered (perm) 304-byte PODs.  Index and flag loads are hoisted, then
// every thread issues one TMA bulk copy per record it will use into a shared-memory stage (up to
// TOP_STAGE records, ~152 KB in flight per SM) and all threads wait on one mbarrier: the gather runs
// at memory speed instead of one dependent round trip per record.  Shared reads are 128-bit with a
// 304-byte lane stride (conflict-free).
constexpr int TOP_KMAX = TOP_STAGE / TOP_THREADS;  // records per thread per stage

// AccumulatedTopHessianSSE::addPoint's accumulation of one residual (AccumulatedTopHessian.cpp:102-135): J = the record,
// res = the (possibly re-linearised) residual vector; acc = the 91 AccumulatorApprox entries, pt = {bd, Hdd, Hcd[4]} terms
__device__ __forceinline__ void top_add_point(const float* __restrict__ J, const float* __restrict__ res, float* __restrict__ acc, float* __restrict__ pt) {
    // :102-112
    float JIr0 = 0.f, JIr1 = 0.f, Jabr0 = 0.f, Jabr1 = 0.f, rsq = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        JIr0 += res[q] * J[O_JIDX0 + q];
        JIr1 += res[q] * J[O_JIDX1 + q];
        Jabr0 += res[q] * J[O_JAB0 + q];
        Jabr1 += res[q] * J[O_JAB1 + q];
        rsq += res[q] * res[q];
    }
    // x = [Jpdc[0] Jpdxi[0]], y = [Jpdc[1] Jpdxi[1]]  (:115-129)
    float x[10], y[10];
#pragma unroll
    for (int q = 0; q < 4; ++q) { x[q] = J[O_JPDC0 + q]; y[q] = J[O_JPDC1 + q]; }
#pragma unroll
    for (int q = 0; q < 6; ++q) { x[4 + q] = J[O_JPDXI0 + q]; y[4 + q] = J[O_JPDXI1 + q]; }
    const float a = J[O_JIDX2], b = J[O_JIDX2 + 2], c = J[O_JIDX2 + 3];  // (0,0) (0,1) (1,1), column-major
    // AccumulatorApprox::update: a x x^T + c y y^T + b (x y^T + y x^T) = x (a x + b y)^T + y (b x + c y)^T
    float px[10], py[10];
#pragma unroll
    for (int q = 0; q < 10; ++q) { px[q] = a * x[q] + b * y[q]; py[q] = b * x[q] + c * y[q]; }
    int e = 0;
#pragma unroll
    for (int p = 0; p < 10; ++p)
#pragma unroll
        for (int q = p; q < 10; ++q) acc[e++] += x[p] * px[q] + y[p] * py[q];
    // updateTopRight: TR00,TR10 = JabJIdx(0,0),(0,1); TR01,TR11 = JabJIdx(1,0),(1,1); TR02,TR12 = JI_r
    const float t00 = J[O_JABJIDX], t10 = J[O_JABJIDX + 2], t01 = J[O_JABJIDX + 1], t11 = J[O_JABJIDX + 3];
#pragma unroll
    for (int p = 0; p < 10; ++p) {
        acc[55 + 3 * p] += x[p] * t00 + y[p] * t10;
        acc[55 + 3 * p + 1] += x[p] * t01 + y[p] * t11;
        acc[55 + 3 * p + 2] += x[p] * JIr0 + y[p] * JIr1;
    }
    // updateBotRight
    acc[85] += J[O_JAB2]; acc[86] += J[O_JAB2 + 2]; acc[87] += Jabr0;
    acc[88] += J[O_JAB2 + 3]; acc[89] += Jabr1; acc[90] += rsq;
    // :132-135
    const float d0 = J[O_JPDD], d1 = J[O_JPDD + 1];
    const float q0 = a * d0 + b * d1, q1 = b * d0 + c * d1;
    pt[0] = JIr0 * d0 + JIr1 * d1;
    pt[1] = q0 * d0 + q1 * d1;
#pragma unroll
    for (int q = 0; q < 4; ++q) pt[2 + q] = J[O_JPDC0 + q] * q0 + J[O_JPDC1 + q] * q1;
}

// Grid-wide barrier of a cooperatively launched kernel (all CTAs resident), sense reversing: bar[0] counts arrivals, bar[1] is
// the generation.  No host-side state: a launch that fails leaves nothing behind that a later launch could wait for.
__device__ __forceinline__ void grid_barrier(unsigned* bar) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned gen;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(bar + 1) : "memory");
        __threadfence();
        if (atomicAdd(bar, 1u) == gridDim.x - 1u) {
            bar[0] = 0u;
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            unsigned g2;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g2) : "l"(bar + 1) : "memory");
            } while (g2 == gen);
        }
        __threadfence();
    }
    __syncthreads();
}

// Second phase of an accumulation kernel, by the whole grid once every tile partial and per-residual term is in memory:
//  * ordered fp64 sum of the tile partials of each (host,target) key -> 13x13 (AccumulatorApprox::finish layout),
//  * per point Hdd_acc, bd_acc, Hcd_acc (AccumulatedTopHessian.cpp:132-157), residuals in their stored order.
__device__ void top_second_phase(const BaDev& W) {
    const int nthreads = gridDim.x * blockDim.x, g = blockIdx.x * blockDim.x + threadIdx.x;
    for (int idx = g; idx < W.F * W.F * NACC; idx += nthreads) {
        const int key = idx / NACC, e = idx % NACC;
        const int t0 = W.key_tile_begin[key], t1 = W.key_tile_begin[key + 1];
        if (e < 91) {
            double s = 0.0;
            for (int t = t0; t < t1; ++t) s += __ldcg(&W.tile_partial[(size_t)NACC * t + e]);
            int r, c;
            if (e < 55) {
                int p = 0, rem = e;
                while (rem >= 10 - p) { rem -= 10 - p; ++p; }
                r = p; c = p + rem;
            } else if (e < 85) { r = (e - 55) / 3; c = 10 + (e - 55) % 3; }
            else { const int br[6][2] = {{10, 10}, {10, 11}, {10, 12}, {11, 11}, {11, 12}, {12, 12}}; r = br[e - 85][0]; c = br[e - 85][1]; }
            double* H = W.acc_out + (size_t)169 * key;
            H[13 * r + c] = s;
            H[13 * c + r] = s;
        } else if (e == 95) {
            long long n = 0;
            for (int t = t0; t < t1; ++t) n += __ldcg(&W.tile_nres[t]);
            W.num_out[key] = n;
        }
    }
    for (int p = g; p < W.P; p += nthreads) {
        float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int r = W.res_begin[p]; r < W.res_begin[p + 1]; ++r) {
            const float2* v = reinterpret_cast<const float2*>(W.res_pt + (size_t)6 * r);
            const float2 a = __ldcg(v), b = __ldcg(v + 1), c = __ldcg(v + 2);
            s[0] += a.x; s[1] += a.y; s[2] += b.x; s[3] += b.y; s[4] += c.x; s[5] += c.y;
        }
        W.bd_out[p] = s[0];
        W.Hdd_out[p] = s[1];
        reinterpret_cast<float4*>(W.Hcd_out)[p] = make_float4(s[2], s[3], s[4], s[5]);
    }
}

// warp butterfly + fp64 across warps -> this tile's partial (fixed order)
__device__ __forceinline__ void top_tile_partial(const BaDev& W, float* acc, int nres, double (*warp_part)[NACC], int* warp_n) {
    const int tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    butterfly_step<48, 16>(acc, lane);
    butterfly_step<24, 8>(acc, lane);
    butterfly_step<12, 4>(acc, lane);
    butterfly_step<6, 2>(acc, lane);
    butterfly_step<3, 1>(acc, lane);
    const int base = butterfly_base(lane);
    warp_part[warp][base] = (double)acc[0];
    warp_part[warp][base + 1] = (double)acc[1];
    warp_part[warp][base + 2] = (double)acc[2];
    for (int o = 16; o > 0; o >>= 1) nres += __shfl_xor_sync(0xffffffffu, nres, o);
    if (lane == 0) warp_n[warp] = nres;
    __syncthreads();
    if (tid < NACC) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < TOP_THREADS / 32; ++w) s += warp_part[w][tid];
        W.tile_partial[(size_t)NACC * tile + tid] = s;
    }
    if (tid == 0) {
        int n = 0;
        for (int w = 0; w < TOP_THREADS / 32; ++w) n += warp_n[w];
        W.tile_nres[tile] = n;
    }
}

template <int MODE>
__global__ void __launch_bounds__(TOP_THREADS) ba_top_kernel(BaDev W) {
    extern __shared__ __align__(16) unsigned char top_smem[];
    float* stage = reinterpret_cast<float*>(top_smem);  // [TOP_STAGE][REC]
    __shared__ __align__(8) unsigned long long bar;
    __shared__ double warp_part[TOP_THREADS / 32][NACC];
    __shared__ int warp_n[TOP_THREADS / 32];
    const int tile = blockIdx.x, tid = threadIdx.x;
    const int key = W.tile_key[tile], start = W.tile_start[tile], count = W.tile_count[tile];
    if (tid == 0) mbar_init(&bar, TOP_THREADS);
    float dp[8], dc[4];
    if (MODE == 1) {
#pragma unroll
        for (int k = 0; k < 8; ++k) dp[k] = W.adHTdeltaF[8 * key + k];  // ef->adHTdeltaF[htIDX], :72
#pragma unroll
        for (int k = 0; k < 4; ++k) dc[k] = W.cDeltaF[k];
    }
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
    int nres = 0;
    __syncthreads();  // barrier initialised
    unsigned parity = 0;
    for (int s0 = 0; s0 < count; s0 += TOP_STAGE, parity ^= 1u) {
        // hoisted, independent index / flag loads of this thread's records in the stage
        int rr[TOP_KMAX];
        bool use[TOP_KMAX];
#pragma unroll
        for (int k = 0; k < TOP_KMAX; ++k) {
            const int i = s0 + tid + k * TOP_THREADS;
            rr[k] = (i < count) ? W.perm[start + i] : -1;
        }
        unsigned bytes = 0;
#pragma unroll
        for (int k = 0; k < TOP_KMAX; ++k) {
            const unsigned fl = (rr[k] >= 0) ? W.flags[rr[k]] : 0u;
            const bool active = fl & 1u, lin = fl & 2u;
            if (MODE == 0) use[k] = active && !lin;        // :55-58
            else if (MODE == 1) use[k] = active && lin;    // :59-62
            else use[k] = active;                          // :63-67
            if (use[k]) bytes += 4u * REC;
        }
        mbar_arrive_expect_tx(&bar, bytes);
#pragma unroll
        for (int k = 0; k < TOP_KMAX; ++k)
            if (use[k]) tma_load_record(stage + (size_t)REC * (tid + k * TOP_THREADS), W.recs + (size_t)REC * rr[k], 4u * REC, &bar);
        mbar_wait(&bar, parity);
#pragma unroll 1
        for (int k = 0; k < TOP_KMAX; ++k) {
            const int r = rr[k];
            if (r < 0) continue;
            float pt[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (use[k]) {
                const float4* rec4 = reinterpret_cast<const float4*>(stage + (size_t)REC * (tid + k * TOP_THREADS));
                float J[REC];
#pragma unroll
                for (int q = 0; q < REC / 4; ++q) {
                    const float4 v = rec4[q];
                    J[4 * q] = v.x; J[4 * q + 1] = v.y; J[4 * q + 2] = v.z; J[4 * q + 3] = v.w;
                }
                float res[8];
                if (MODE == 0) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) res[q] = J[O_RES + q];
                } else {
                    const float4* rz = reinterpret_cast<const float4*>(W.res_toZero + (size_t)8 * r);
                    const float4 a = __ldg(rz), b = __ldg(rz + 1);
                    res[0] = a.x; res[1] = a.y; res[2] = a.z; res[3] = a.w; res[4] = b.x; res[5] = b.y; res[6] = b.z; res[7] = b.w;
                    if (MODE == 1) {  // :81-99: rtz + [JI*Jp Ja]*delta
                        const float dd = W.deltaF[W.point_of_res[r]];
                        float jx = 0.f, jy = 0.f, cxs = 0.f, cys = 0.f;
#pragma unroll
                        for (int q = 0; q < 6; ++q) { jx += J[O_JPDXI0 + q] * dp[q]; jy += J[O_JPDXI1 + q] * dp[q]; }
#pragma unroll
                        for (int q = 0; q < 4; ++q) { cxs += J[O_JPDC0 + q] * dc[q]; cys += J[O_JPDC1 + q] * dc[q]; }
                        jx = jx + cxs + J[O_JPDD] * dd;
                        jy = jy + cys + J[O_JPDD + 1] * dd;
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            res[q] = res[q] + J[O_JIDX0 + q] * jx + J[O_JIDX1 + q] * jy + J[O_JAB0 + q] * dp[6] + J[O_JAB1 + q] * dp[7];
                    }
                }
                top_add_point(J, res, acc, pt);
                nres++;
            }
            float2* o = reinterpret_cast<float2*>(W.res_pt + (size_t)6 * r);
            o[0] = make_float2(pt[0], pt[1]); o[1] = make_float2(pt[2], pt[3]); o[2] = make_float2(pt[4], pt[5]);
        }
        if (s0 + TOP_STAGE < count) __syncthreads();  // the stage is overwritten by the next round
    }
    top_tile_partial(W, acc, nres, warp_part, warp_n);
    grid_barrier(W.grid_bar);
    top_second_phase(W);
}

// PointFrameResidual::linearize FUSED with addPoint<0>: the thread that linearises a residual
// accumulates it straight from registers, in the tile structure and summation order of ba_top_kernel<0> (the results are
// bit-identical to linearize -> top_accumulate(0)).  The 304-byte record only goes to memory for the residuals that need
// it later (the linearized ones, which the mode-1 pass and the linearised energy read) unless the caller asks for all of
// them; state, energy, flags and JpJdF (takeDataF) are always written.
__global__ void __launch_bounds__(TOP_THREADS, 1) ba_lin_top_kernel(LinDev d, BaDev W, float* __restrict__ JpJdF, int write_all) {
    __shared__ double warp_part[TOP_THREADS / 32][NACC];
    __shared__ int warp_n[TOP_THREADS / 32];
    const int tile = blockIdx.x, tid = threadIdx.x;
    const int start = W.tile_start[tile], count = W.tile_count[tile];
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
    int nres = 0;
#pragma unroll 1
    for (int i = tid; i < count; i += TOP_THREADS) {
        const int r = W.perm[start + i];
        float rec[REC];
        int state;
        float energy;
        linearize_one(d, r, rec, state, energy);
        const bool lin = d.linearized && d.linearized[r];
        d.state_new[r] = state;
        d.energy_new[r] = energy;
        d.flags[r] = (uint8_t)((state == 0 ? EDSGPU_RES_ACTIVE : 0) | (lin ? EDSGPU_RES_LINEARIZED : 0));
        {
            float o8[8];
            jpjd_one(rec, o8);
            float4* o = reinterpret_cast<float4*>(JpJdF + (size_t)8 * r);
            o[0] = make_float4(o8[0], o8[1], o8[2], o8[3]);
            o[1] = make_float4(o8[4], o8[5], o8[6], o8[7]);
        }
        if (write_all || lin) {
            float4* out = reinterpret_cast<float4*>(d.recs + (size_t)REC * r);
#pragma unroll
            for (int q = 0; q < REC / 4; ++q) out[q] = make_float4(rec[4 * q], rec[4 * q + 1], rec[4 * q + 2], rec[4 * q + 3]);
        }
        float pt[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (state == 0 && !lin) {  // AccumulatedTopHessian.cpp:55-58
            top_add_point(rec, rec + O_RES, acc, pt);
            nres++;
        }
        float2* o = reinterpret_cast<float2*>(W.res_pt + (size_t)6 * r);
        o[0] = make_float2(pt[0], pt[1]); o[1] = make_float2(pt[2], pt[3]); o[2] = make_float2(pt[4], pt[5]);
    }
    top_tile_partial(W, acc, nres, warp_part, warp_n);
    grid_barrier(W.grid_bar);
    top_second_phase(W);
}

// ------------------------------------------------------------------------------------------
// Schur complement (AccumulatedSCHessianSSE::addPoint, AccumulatedSCHessian.cpp:34-77)
// ------------------------------------------------------------------------------------------
struct ScDev {
    int F, P, shift_prior;
    const int32_t *res_begin, *target_idx, *pt_perm, *chunk_host, *chunk_start, *chunk_count;
    const uint8_t* flags;
    const float *JpJdF, *HddA, *HddL, *bdA, *bdL, *HcdA, *HcdL, *priorF, *deltaF;
    float *HdiF, *bdSum;
    float* partial;  // num_chunks x M x N
};

__global__ void __launch_bounds__(SC_THREADS) ba_sc_kernel(ScDev S) {
    const int chunk = blockIdx.x, tid = threadIdx.x;
    const int F = S.F, M = 8 * F + 4, N = 8 * F + 5;
    const int tiles_n = (N + 3) / 4, tiles = ((M + 3) / 4) * tiles_n;
    const int start = S.chunk_start[chunk], count = S.chunk_count[chunk];
    __shared__ __align__(16) float z[SC_BATCH][SC_LD];   // column side: [z (8F) | Hcd (4) | bdSum]
    __shared__ __align__(16) float zs[SC_BATCH][SC_LD];  // row side, pre-scaled by HdiF
    float acc[2][16];
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[t][i] = 0.f;
    for (int b0 = 0; b0 < count; b0 += SC_BATCH) {
        const int nb = min(SC_BATCH, count - b0);
        __syncthreads();
        for (int i = tid; i < SC_BATCH * SC_LD; i += SC_THREADS) { (&z[0][0])[i] = 0.f; (&zs[0][0])[i] = 0.f; }
        __syncthreads();
        // eight lanes per point (AccumulatedSCHessian.cpp:36-76): flags and JpJdF of its residuals are
        // fetched with independent loads, the active count is a ballot over the lane group
        {
            const int q = tid >> 3, part = tid & 7, lane = tid & 31;
            const unsigned gmask = 0xffu << (lane & 24);
            const bool live = q < nb;
            const int p = live ? S.pt_perm[start + b0 + q] : 0;
            const int rb = live ? S.res_begin[p] : 0, re = live ? S.res_begin[p + 1] : 0;
            int ngood = 0;
            for (int r0 = rb; r0 < re; r0 += 8) {
                const int r = r0 + part;
                const bool act = (r < re) && (S.flags[r] & 1u);
                ngood += __popc(__ballot_sync(gmask, act));  // group-scoped: trip counts differ between groups
            }
            float HdiF = 0.f, bdSum = 0.f;
            if (live && ngood > 0) {
                float Hh = S.HddA[p] + S.HddL[p] + S.priorF[p];
                if (Hh < 1e-10f) Hh = 1e-10f;
                HdiF = (float)(1.0 / (double)Hh);
                bdSum = S.bdA[p] + S.bdL[p];
                if (S.shift_prior) bdSum += S.priorF[p] * S.deltaF[p];
                if (part < 4) {
                    const float hc = S.HcdA[4 * p + part] + S.HcdL[4 * p + part];
                    z[q][8 * F + part] = hc;
                    zs[q][8 * F + part] = HdiF * hc;
                } else if (part == 4) {
                    z[q][8 * F + 4] = bdSum;
                }
                for (int r = rb + part; r < re; r += 8) {
                    if (!(S.flags[r] & 1u)) continue;
                    const int t = S.target_idx[r];
                    const float4* j = reinterpret_cast<const float4*>(S.JpJdF + (size_t)8 * r);
                    const float4 a = __ldg(j), b = __ldg(j + 1);
                    *reinterpret_cast<float4*>(&z[q][8 * t]) = a;
                    *reinterpret_cast<float4*>(&z[q][8 * t + 4]) = b;
                    *reinterpret_cast<float4*>(&zs[q][8 * t]) = make_float4(HdiF * a.x, HdiF * a.y, HdiF * a.z, HdiF * a.w);
                    *reinterpret_cast<float4*>(&zs[q][8 * t + 4]) = make_float4(HdiF * b.x, HdiF * b.y, HdiF * b.z, HdiF * b.w);
                }
            }
            if (live && part == 0) { S.HdiF[p] = HdiF; S.bdSum[p] = bdSum; }
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int tl = tid + t * SC_THREADS;
            if (tl < tiles) {
                const int i0 = 4 * (tl / tiles_n), j0 = 4 * (tl % tiles_n);
                for (int q = 0; q < nb; ++q) {
                    const float4 rv = *reinterpret_cast<const float4*>(&zs[q][i0]);
                    const float4 cv = *reinterpret_cast<const float4*>(&z[q][j0]);
                    const float rr[4] = {rv.x, rv.y, rv.z, rv.w}, cc[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[t][4 * i + j] += rr[i] * cc[j];
                }
            }
        }
    }
    float* out = S.partial + (size_t)chunk * M * N;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int tl = tid + t * SC_THREADS;
        if (tl < tiles) {
            const int i0 = 4 * (tl / tiles_n), j0 = 4 * (tl % tiles_n);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (i0 + i < M && j0 + j < N) out[(size_t)(i0 + i) * N + j0 + j] = acc[t][4 * i + j];
        }
    }
}

// ordered fp64 sum over the chunks of each host -> accD / accE / accEB, per-host Hcc|bc
__global__ void ba_sc_finalize_kernel(int F, const float* __restrict__ partial, const int32_t* __restrict__ host_chunk_begin,
                                      double* __restrict__ accD, double* __restrict__ accE, double* __restrict__ accEB,
                                      double* __restrict__ hcc_host /*F x 20*/) {
    const int h = blockIdx.x, M = 8 * F + 4, N = 8 * F + 5, F2 = F * F;
    const int c0 = host_chunk_begin[h], c1 = host_chunk_begin[h + 1];
    const int e = blockIdx.y * blockDim.x + threadIdx.x;
    if (e >= M * N) return;
    double s = 0.0;
    for (int c = c0; c < c1; ++c) s += (double)partial[(size_t)c * M * N + e];
    const int i = e / N, j = e % N;
    if (i < 8 * F) {
        const int t1 = i / 8, a = i % 8;
        if (j < 8 * F) {           // accD[h + t1*F + t2*F2] (8x8)
            const int t2 = j / 8, b = j % 8;
            accD[(size_t)64 * (h + t1 * F + t2 * F2) + 8 * a + b] = s;
        } else if (j < 8 * F + 4) {  // accE[h + t1*F] (8x4)
            accE[(size_t)32 * (h + t1 * F) + 4 * a + (j - 8 * F)] = s;
        } else {                    // accEB[h + t1*F] (8)
            accEB[(size_t)8 * (h + t1 * F) + a] = s;
        }
    } else if (j >= 8 * F) {      // Hcd rows: [Hcc (4x4) | bc (4)]; the Hcd x z^T part is E^T, not needed
        hcc_host[20 * h + 5 * (i - 8 * F) + (j - 8 * F)] = s;
    }
}
__global__ void ba_sc_hcc_kernel(int F, const double* __restrict__ hcc_host, double* __restrict__ accHcc, double* __restrict__ accbc) {
    const int e = threadIdx.x;
    if (e >= 20) return;
    double s = 0.0;
    for (int h = 0; h < F; ++h) s += hcc_host[20 * h + e];
    const int i = e / 5, j = e % 5;
    if (j < 4) accHcc[4 * i + j] = s; else accbc[i] = s;
}

// ------------------------------------------------------------------------------------------
// stitches: one CTA (64 threads) per 8x8 output block (a,b) of H, fp64
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double adj(const double* A, int r, int c) { return A[c * 8 + r]; }  // Mat88 column-major

// A frame block of a stitched matrix is a sum of sandwiches  L C R^T  (adjoint, accumulator block, adjoint):
//   out(i,j) = sum_k x_k(i,j),  x_k(i,j) = sum_m L_k(i,m) T_k(m,j),  T_k(m,j) = sum_n C_k(m,n) R_k(j,n).
// A CTA of 256 threads = four groups of 64, thread e = 8 i + j of a group owns element (i, j).  Per chunk of terms:
//  (0) the three 8x8 blocks of every term are copied into shared memory with cp.async, one element per copy, all in flight
//      together: a chunk costs one memory round trip (a term-by-term evaluation from global memory spends a dependent round
//      trip of 24 loads per thread on every term: 49 us for the three stitches of solveSystemF);
//  (1) group g forms T_k, (2) then x_k, of the terms k = g, g + 4, ...;
//  (3) group 0 adds the x_k in their fixed order.  Same products and the same order of additions as term by term.
struct StitchTerm { const double *L, *C, *R; int ldc; };  // C(m,n) = C[ldc * m + n]; L, R: Mat88 column-major
constexpr int STITCH_CHUNK = 16, STITCH_THREADS = 256;
struct StitchShared { double L[STITCH_CHUNK][64], C[STITCH_CHUNK][64], R[STITCH_CHUNK][64], T[STITCH_CHUNK][64], X[STITCH_CHUNK][64]; };  // 40 KB

// returns the block's element (i, j) in the threads of group 0 (tid < 64)
template <typename TermFn>
__device__ __forceinline__ double stitch_terms(StitchShared& sh, int nterms, TermFn term) {
    const int tid = threadIdx.x, e = tid & 63, g = tid >> 6, i = e >> 3, j = e & 7;
    double s = 0.0;
    for (int k0 = 0; k0 < nterms; k0 += STITCH_CHUNK) {
        const int kn = min(STITCH_CHUNK, nterms - k0);
        __syncthreads();  // the buffers are free (previous chunk, or the caller's own use of them)
        for (int w = tid; w < kn * 192; w += STITCH_THREADS) {
            const int k = w / 192, r = w - 192 * k, which = r >> 6, el = r & 63;
            const StitchTerm q = term(k0 + k);
            if (which == 0) __pipeline_memcpy_async(&sh.L[k][el], q.L + el, 8);
            else if (which == 1) __pipeline_memcpy_async(&sh.R[k][el], q.R + el, 8);
            else __pipeline_memcpy_async(&sh.C[k][el], q.C + q.ldc * (el >> 3) + (el & 7), 8);  // row-major 8x8
        }
        __pipeline_commit();
        __pipeline_wait_prior(0);
        __syncthreads();
        for (int k = g; k < kn; k += 4) {
            double t = 0.0;
#pragma unroll
            for (int n = 0; n < 8; ++n) t += sh.C[k][8 * i + n] * adj(sh.R[k], j, n);
            sh.T[k][e] = t;
        }
        __syncthreads();
        for (int k = g; k < kn; k += 4) {
            double x = 0.0;
#pragma unroll
            for (int m = 0; m < 8; ++m) x += adj(sh.L[k], i, m) * sh.T[k][8 * m + j];
            sh.X[k][e] = x;
        }
        __syncthreads();
        if (g == 0)
            for (int k = 0; k < kn; ++k) s += sh.X[k][e];
    }
    return s;
}

// The calibration columns and the b segment of frame `a` (rows of the diagonal CTA): v(i, col) = sum over the 2F pairs frame
// a takes part in of  Ad_k(i, :) . E_k(:, col),  E_k an 8 x 5 block given element-wise by `e(k, m, col)`.  The adjoints
// and the E blocks go through shared memory like the sandwiches' operands; the sum keeps its order (pair by pair, m inside).
// Returns v(i, j) in the threads of group 0 with j < 5.
template <typename AdFn, typename EFn>
__device__ __forceinline__ double stitch_calibration(StitchShared& sh, int ncal, AdFn ad, EFn e) {
    const int tid = threadIdx.x, i = (tid & 63) >> 3, j = tid & 7;
    __syncthreads();
    for (int w = tid; w < ncal * 104; w += STITCH_THREADS) {
        const int k = w / 104, r = w - 104 * k;
        if (r < 64) __pipeline_memcpy_async(&sh.L[k][r], ad(k) + r, 8);
        else sh.C[k][r - 64] = e(k, (r - 64) / 5, (r - 64) % 5);
    }
    __pipeline_commit();
    __pipeline_wait_prior(0);
    __syncthreads();
    double v = 0.0;
    if (tid < 64 && j < CPARS + 1) {
        for (int k = 0; k < ncal; ++k)
#pragma unroll
            for (int m = 0; m < 8; ++m) v += adj(sh.L[k], i, m) * sh.C[k][5 * m + j];
    }
    return v;
}
static_assert(2 * MAXF <= STITCH_CHUNK, "the calibration pass keeps all 2F pairs of a frame in one chunk");

// AccumulatedTopHessianSSE::stitchDoubleInternal (AccumulatedTopHessian.cpp:241-303), by output block (a, b) = blk.
__device__ __forceinline__ void top_stitch_block(StitchShared& ssh, int blk, int F, const double* __restrict__ acc, const double* __restrict__ adHost,
                                                 const double* __restrict__ adTarget, int use_prior, const double* __restrict__ cPrior,
                                                 const float* __restrict__ cDeltaF, const double* __restrict__ frame_prior,
                                                 const double* __restrict__ frame_delta_prior, double* __restrict__ H, double* __restrict__ bvec) {
    const int a = blk % F, b = blk / F, n = CPARS + 8 * F;
    const int i = (threadIdx.x & 63) / 8, j = threadIdx.x % 8;
    const bool writer = threadIdx.x < 64;  // group 0 holds the results
    auto Hat = [&](int r, int c) -> double& { return H[(size_t)c * n + r]; };
    const int ndiag = a == b ? 2 * F : 0;
    const double s = stitch_terms(ssh, ndiag + 1, [&](int k) {
        StitchTerm q;
        q.ldc = 13;
        int kk;
        if (k < ndiag) {
            if (k < F) { kk = a + F * k; q.L = q.R = adHost + 64 * kk; }          // k = (h=a, t): adHost A88 adHost^T
            else { kk = (k - F) + F * a; q.L = q.R = adTarget + 64 * kk; }          // k = (h, t=a): adTarget A88 adTarget^T
        } else {
            kk = a + F * b; q.L = adHost + 64 * kk; q.R = adTarget + 64 * kk;       // H(hIdx,tIdx) += adHost A88 adTarget^T for (h=a, t=b)
        }
        q.C = acc + (size_t)169 * kk + 13 * CPARS + CPARS;
        return q;
    });
    if (writer) Hat(CPARS + 8 * a + i, CPARS + 8 * b + j) = s;
    if (a == b) {
        // calibration columns, b segment, priors for frame a; done by the diagonal CTA
        const double v = stitch_calibration(ssh, 2 * F,
            [&](int k) { return k < F ? adHost + 64 * (a + F * k) : adTarget + 64 * ((k - F) + F * a); },  // (h=a, t=k), then (h=k-F, t=a)
            [&](int k, int m, int c) {
                const int kk = k < F ? a + F * k : (k - F) + F * a;
                return acc[(size_t)169 * kk + 13 * (CPARS + m) + (c < CPARS ? c : CPARS + 8)];
            });
        if (writer && j < CPARS + 1) {
            if (j < CPARS) { Hat(CPARS + 8 * a + i, j) = v; Hat(j, CPARS + 8 * a + i) = v; }
            else bvec[CPARS + 8 * a + i] = v + (use_prior ? frame_prior[8 * a + i] * frame_delta_prior[8 * a + i] : 0.0);
        }
        if (writer && a == 0 && i < CPARS && j < CPARS + 1) {  // Hcc, bc
            double v = 0.0;
            const int col = (j < CPARS) ? j : CPARS + 8;
#pragma unroll 7
            for (int k = 0; k < F * F; ++k) v += acc[(size_t)169 * k + 13 * i + col];
            if (j < CPARS) Hat(i, j) = v + ((use_prior && i == j) ? cPrior[i] : 0.0);
            else bvec[i] = v + (use_prior ? cPrior[i] * (double)cDeltaF[i] : 0.0);
        }
    }
}

__global__ void __launch_bounds__(STITCH_THREADS) ba_top_stitch_kernel(int F, const double* __restrict__ acc, const double* __restrict__ adHost,
                                     const double* __restrict__ adTarget, int use_prior, const double* __restrict__ cPrior,
                                     const float* __restrict__ cDeltaF, const double* __restrict__ frame_prior,
                                     const double* __restrict__ frame_delta_prior, double* __restrict__ H, double* __restrict__ bvec) {
    __shared__ StitchShared ssh;
    top_stitch_block(ssh, blockIdx.x, F, acc, adHost, adTarget, use_prior, cPrior, cDeltaF, frame_prior, frame_delta_prior, H, bvec);
}
// "make diagonal by copying over parts" (AccumulatedTopHessian.h:125-137) + frame priors on the diagonal
__global__ void ba_top_symmetrise_kernel(int F, int use_prior, const double* __restrict__ frame_prior, double* __restrict__ H) {
    const int h = blockIdx.x % F, t = blockIdx.x / F, n = CPARS + 8 * F;
    const int i = threadIdx.x / 8, j = threadIdx.x % 8;
    auto Hat = [&](int r, int c) -> double& { return H[(size_t)c * n + r]; };
    if (t > h) {
        const double v = Hat(CPARS + 8 * h + i, CPARS + 8 * t + j) + Hat(CPARS + 8 * t + j, CPARS + 8 * h + i);
        Hat(CPARS + 8 * h + i, CPARS + 8 * t + j) = v;
        Hat(CPARS + 8 * t + j, CPARS + 8 * h + i) = v;
    } else if (t == h && use_prior && i == j) {
        Hat(CPARS + 8 * h + i, CPARS + 8 * h + i) += frame_prior[8 * h + i];
    }
}

// AccumulatedSCHessianSSE::stitchDoubleInternal (AccumulatedSCHessian.cpp:78-157), by output block (a, b) = blk.
__device__ __forceinline__ void sc_stitch_block(StitchShared& ssh, int blk, int F, const double* __restrict__ accD, const double* __restrict__ accE,
                                                const double* __restrict__ accEB, const double* __restrict__ accHcc, const double* __restrict__ accbc,
                                                const double* __restrict__ adHost, const double* __restrict__ adTarget, double* __restrict__ H,
                                                double* __restrict__ bvec) {
    const int a = blk % F, b = blk / F, n = CPARS + 8 * F, F2 = F * F;
    const int i = (threadIdx.x & 63) / 8, j = threadIdx.x % 8;
    const bool writer = threadIdx.x < 64;  // group 0 holds the results
    auto Hat = [&](int r, int c) -> double& { return H[(size_t)c * n + r]; };
    auto D = [&](int ii, int jj, int kk) { return accD + (size_t)64 * (ii + F * jj + F2 * kk); };
    const int ndiag = a == b ? F2 : 0;
    const double s = stitch_terms(ssh, ndiag + 3 * F, [&](int k) {
        StitchTerm q;
        q.ldc = 8;
        if (k < ndiag) {  // H(iIdx,iIdx) += adHost[ij] D[ijk] adHost[ik]^T over all j,k, i = a
            const int jj = k / F, kk = k % F;
            q.C = D(a, jj, kk); q.L = adHost + 64 * (a + F * jj); q.R = adHost + 64 * (a + F * kk);
            return q;
        }
        k -= ndiag;
        if (k < F) {  // H(jIdx,kIdx) += adTarget[ij] D[ijk] adTarget[ik]^T, j = a, k = b
            q.C = D(k, a, b); q.L = adTarget + 64 * (k + F * a); q.R = adTarget + 64 * (k + F * b);
        } else if (k < 2 * F) {  // H(jIdx,iIdx) += adTarget[ij] D[ijk] adHost[ik]^T, j = a, i = b
            const int kk = k - F;
            q.C = D(b, a, kk); q.L = adTarget + 64 * (b + F * a); q.R = adHost + 64 * (b + F * kk);
        } else {  // H(iIdx,kIdx) += adHost[ij] D[ijk] adTarget[ik]^T, i = a, k = b
            const int jj = k - 2 * F;
            q.C = D(a, jj, b); q.L = adHost + 64 * (a + F * jj); q.R = adTarget + 64 * (a + F * b);
        }
        return q;
    });
    if (writer) Hat(CPARS + 8 * a + i, CPARS + 8 * b + j) = s;
    if (a == b) {
        // rows of frame a as host i: adHost[ij] * E[ij], then as target j: adTarget[ij] * E[ij]
        const double v = stitch_calibration(ssh, 2 * F,
            [&](int k) { return k < F ? adHost + 64 * (a + F * k) : adTarget + 64 * ((k - F) + F * a); },
            [&](int k, int m, int c) {
                const int ij = k < F ? a + F * k : (k - F) + F * a;
                return c < CPARS ? accE[(size_t)32 * ij + 4 * m + c] : accEB[(size_t)8 * ij + m];
            });
        if (writer && j < CPARS + 1) {
            if (j < CPARS) { Hat(CPARS + 8 * a + i, j) = v; Hat(j, CPARS + 8 * a + i) = v; }
            else bvec[CPARS + 8 * a + i] = v;
        }
        if (writer && a == 0 && i < CPARS && j < CPARS + 1) {
            if (j < CPARS) Hat(i, j) = accHcc[4 * i + j];
            else bvec[i] = accbc[i];
        }
    }
}

__global__ void __launch_bounds__(STITCH_THREADS) ba_sc_stitch_kernel(int F, const double* __restrict__ accD, const double* __restrict__ accE, const double* __restrict__ accEB,
                                    const double* __restrict__ accHcc, const double* __restrict__ accbc, const double* __restrict__ adHost,
                                    const double* __restrict__ adTarget, double* __restrict__ H, double* __restrict__ bvec) {
    __shared__ StitchShared ssh;
    sc_stitch_block(ssh, blockIdx.x, F, accD, accE, accEB, accHcc, accbc, adHost, adTarget, H, bvec);
}

// solveSystemF's three stitches (active, linearized, Schur) in one launch: blockIdx.y picks the system
struct Stitch3Args {
    int F, use_prior;
    const double *acc0, *acc1, *adHost, *adTarget, *cPrior, *frame_prior, *frame_delta_prior;
    const float* cDeltaF;
    const double *accD, *accE, *accEB, *accHcc, *accbc;
    double *HA, *bA, *HL, *bL, *Hs, *bs;
};
__global__ void __launch_bounds__(STITCH_THREADS) ba_stitch3_kernel(Stitch3Args q) {
    __shared__ StitchShared ssh;
    if (blockIdx.y == 0)
        top_stitch_block(ssh, blockIdx.x, q.F, q.acc0, q.adHost, q.adTarget, 0, q.cPrior, q.cDeltaF, q.frame_prior, q.frame_delta_prior, q.HA, q.bA);
    else if (blockIdx.y == 1)
        top_stitch_block(ssh, blockIdx.x, q.F, q.acc1, q.adHost, q.adTarget, q.use_prior, q.cPrior, q.cDeltaF, q.frame_prior, q.frame_delta_prior, q.HL, q.bL);
    else
        sc_stitch_block(ssh, blockIdx.x, q.F, q.accD, q.accE, q.accEB, q.accHcc, q.accbc, q.adHost, q.adTarget, q.Hs, q.bs);
}
// the symmetrisation of the two top systems in one launch
__global__ void ba_top_symmetrise2_kernel(int F, int use_prior, const double* __restrict__ frame_prior, double* __restrict__ HA, double* __restrict__ HL) {
    const int h = blockIdx.x % F, t = blockIdx.x / F, n = CPARS + 8 * F;
    const int i = threadIdx.x / 8, j = threadIdx.x % 8;
    double* H = blockIdx.y == 0 ? HA : HL;
    const int prior = blockIdx.y == 0 ? 0 : use_prior;
    auto Hat = [&](int r, int c) -> double& { return H[(size_t)c * n + r]; };
    if (t > h) {
        const double v = Hat(CPARS + 8 * h + i, CPARS + 8 * t + j) + Hat(CPARS + 8 * t + j, CPARS + 8 * h + i);
        Hat(CPARS + 8 * h + i, CPARS + 8 * t + j) = v;
        Hat(CPARS + 8 * t + j, CPARS + 8 * h + i) = v;
    } else if (t == h && prior && i == j) {
        Hat(CPARS + 8 * h + i, CPARS + 8 * h + i) += frame_prior[8 * h + i];
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// EnergyFunctional::solveSystemF (EnergyFunctional.cpp:775-912) in its default solver mode (setting_solverMode =
// SOLVER_FIX_LAMBDA | SOLVER_ORTHOGONALIZE_X_LATER, settings.cpp:60): one CTA, the (4+8F)^2 system in shared memory.
//   HFinal = HL + HM + HA, bFinal = bL + (bM + HM delta) + bA - b_sc;  diag *= (1 + lambda);  HFinal -= H_sc / (1 + lambda)
//   SVecI = 1 / sqrt(diag + 10);  x = SVecI . (SVecI HFinal SVecI)^-1 (SVecI . bFinal)   (LDL^T: the scaled matrix is SPD)
//   optional orthogonalize(&x, 0): x -= N (N^T N)^-1 N^T x with the projector the host made from the frames' null spaces
// then the per-frame-pair row vectors xAd of resubstituteF_MT (:272-281, float like the reference) for the point pass.
// ------------------------------------------------------------------------------------------
constexpr int SOLVE_MAXN = CPARS + 8 * MAXF, SOLVE_LD = SOLVE_MAXN + 1, SOLVE_THREADS = 256;
constexpr int SOLVE_NB = (SOLVE_MAXN + 1 + 15) / 16;  // 16-strided entries per thread and dimension, border row included

__global__ void __launch_bounds__(SOLVE_THREADS) ba_solve_kernel(int F, double lambda, const double* __restrict__ HA, const double* __restrict__ bA,
                                                                  const double* __restrict__ HL, const double* __restrict__ bL,
                                                                  const double* __restrict__ Hsc, const double* __restrict__ bsc,
                                                                  const double* __restrict__ HM, const double* __restrict__ bM,
                                                                  const double* __restrict__ delta, const double* __restrict__ projector,
                                                                  const double* __restrict__ adHost, const double* __restrict__ adTarget,
                                                                  double* __restrict__ x_out, float* __restrict__ xAd, float* __restrict__ cstep) {
    __shared__ double A[SOLVE_MAXN + 1][SOLVE_LD];
    __shared__ double rhs[SOLVE_MAXN], sv[SOLVE_MAXN], xs[SOLVE_MAXN], col[SOLVE_MAXN];
    __shared__ double colbuf[2][SOLVE_NB * 16];
    __shared__ float xF[SOLVE_MAXN];
    const int n = CPARS + 8 * F, tid = threadIdx.x;
    const double inv1l = 1.0 / (1.0 + lambda);
    // entry (r, c) of the assembled system (column-major inputs)
    auto assembled = [&](int r, int c) {
        const int e = c * n + r;
        double v = HL[e] + (HM ? HM[e] : 0.0) + HA[e];
        if (r == c) v *= (1.0 + lambda);
        return v - Hsc[e] * inv1l;
    };
    for (int r = tid; r < n; r += SOLVE_THREADS) {
        double bm = bM ? bM[r] : 0.0;
        if (HM && delta)
            for (int c = 0; c < n; ++c) bm += HM[(size_t)c * n + r] * delta[c];
        const double s = 1.0 / sqrt(assembled(r, r) + 10.0);
        sv[r] = s;
        rhs[r] = (bL[r] + bm + bA[r] - bsc[r]) * s;
    }
    __syncthreads();
    // The scaled system, bordered by the right-hand side as row and column n, lives in REGISTERS: thread (ty, tx) of a
    // 16 x 16 arrangement owns the entries (ty + 16 a, tx + 16 b).  Right-looking LDL^T: at step k the owners of column k
    // publish it through shared memory (double buffered: one barrier per step), every thread then updates its own
    // entries, M(i,j) -= (M(i,k) / D_k) M(j,k), lower triangle only.  Eliminating the border row along with the rest makes
    // it the forward substitution: after the scaling it holds y = D^-1 L^-1 rhs.
    const int tx = tid & 15, ty = tid >> 4;
    double M[SOLVE_NB][SOLVE_NB];
#pragma unroll
    for (int a = 0; a < SOLVE_NB; ++a)
#pragma unroll
        for (int b = 0; b < SOLVE_NB; ++b) {
            const int i = ty + 16 * a, j = tx + 16 * b;
            double v = 0.0;
            if (j <= i && i <= n) {
                if (i < n) v = sv[i] * assembled(i, j) * sv[j];
                else if (j < n) v = rhs[j];
            }
            M[a][b] = v;
        }
    for (int k = 0; k < n; ++k) {
        double* cb = colbuf[k & 1];
        const int kb = k >> 4;
        if (tx == (k & 15)) {
#pragma unroll
            for (int b = 0; b < SOLVE_NB; ++b)
                if (b == kb) {
#pragma unroll
                    for (int a = 0; a < SOLVE_NB; ++a) cb[ty + 16 * a] = M[a][b];
                }
        }
        __syncthreads();
        const double rd = 1.0 / cb[k];
        double ci[SOLVE_NB], cj[SOLVE_NB];
#pragma unroll
        for (int a = 0; a < SOLVE_NB; ++a) { ci[a] = cb[ty + 16 * a] * rd; cj[a] = cb[tx + 16 * a]; }
#pragma unroll
        for (int a = 0; a < SOLVE_NB; ++a)
#pragma unroll
            for (int b = 0; b < SOLVE_NB; ++b) {
                const int i = ty + 16 * a, j = tx + 16 * b;
                if (j > k && j <= i && i <= n) M[a][b] -= ci[a] * cj[b];
            }
    }
    // L (unit lower, scaled columns) and D to shared memory for the back substitution
#pragma unroll
    for (int a = 0; a < SOLVE_NB; ++a)
#pragma unroll
        for (int b = 0; b < SOLVE_NB; ++b) {
            const int i = ty + 16 * a, j = tx + 16 * b;
            if (j <= i && i <= n && j < n) A[i][j] = M[a][b];
        }
    __syncthreads();
    for (int e = tid; e < (n + 1) * n; e += SOLVE_THREADS) {
        const int i = e / n, k = e - i * n;
        if (k < i) A[i][k] /= A[k][k];  // reads the diagonal, writes strictly below it
    }
    __syncthreads();
    // L^T w = y by the first warp: y is the scaled border row; a lane keeps the rows lane, lane + 32, lane + 64 in registers
    // and takes w_j from its owner by shuffle (the chain per step is a shuffle and a multiply-add)
    if (tid < 32) {
        double x0 = tid < n ? A[n][tid] : 0.0, x1 = tid + 32 < n ? A[n][tid + 32] : 0.0, x2 = tid + 64 < n ? A[n][tid + 64] : 0.0;
#pragma unroll 4
        for (int j = n - 1; j > 0; --j) {
            const int slot = j >> 5;
            const double mine = slot == 0 ? x0 : (slot == 1 ? x1 : x2);
            const double wj = __shfl_sync(0xffffffffu, mine, j & 31);
            if (tid < j) x0 -= A[j][tid] * wj;
            if (tid + 32 < j) x1 -= A[j][tid + 32] * wj;
            if (tid + 64 < j) x2 -= A[j][tid + 64] * wj;
        }
        if (tid < n) xs[tid] = x0 * sv[tid];
        if (tid + 32 < n) xs[tid + 32] = x1 * sv[tid + 32];
        if (tid + 64 < n) xs[tid + 64] = x2 * sv[tid + 64];
    }
    __syncthreads();
    if (projector) {  // orthogonalize(&x, 0)
        for (int r = tid; r < n; r += SOLVE_THREADS) {
            double p = 0.0;
            for (int c = 0; c < n; ++c) p += projector[(size_t)c * n + r] * xs[c];
            col[r] = xs[r] - p;
        }
        __syncthreads();
        for (int r = tid; r < n; r += SOLVE_THREADS) xs[r] = col[r];
        __syncthreads();
    }
    for (int r = tid; r < n; r += SOLVE_THREADS) { x_out[r] = xs[r]; xF[r] = (float)xs[r]; }
    __syncthreads();
    // xAd[F*h + t] = xF_h^T adHostF[h + F*t] + xF_t^T adTargetF[h + F*t], float in the reference's order (no FMA contraction)
    for (int e = tid; e < F * F * 8; e += SOLVE_THREADS) {
        const int c = e % 8, ht = e / 8, h = ht / F, t = ht % F;
        const double* AH = adHost + 64 * (size_t)(h + F * t);
        const double* AT = adTarget + 64 * (size_t)(h + F * t);
        float a = 0.f, b = 0.f;
        for (int k = 0; k < 8; ++k) a = __fadd_rn(a, __fmul_rn(xF[CPARS + 8 * h + k], (float)AH[c * 8 + k]));
        for (int k = 0; k < 8; ++k) b = __fadd_rn(b, __fmul_rn(xF[CPARS + 8 * t + k], (float)AT[c * 8 + k]));
        xAd[8 * (F * h + t) + c] = __fadd_rn(a, b);
    }
    if (tid < CPARS) cstep[tid] = xF[tid];
}

// ==========================================================================================
// host side
// ==========================================================================================
struct edsgpu_ba {
    edsgpu_ctx* ctx = nullptr;
    int F = 0, P = 0, R = 0, num_tiles = 0, num_chunks = 0;
    void* block = nullptr;  // one device allocation
    // plan
    int32_t *host_idx, *target_idx, *point_of_res, *res_begin, *perm, *tile_key, *tile_start, *tile_count, *key_tile_begin;
    int32_t *pt_perm, *chunk_host, *chunk_start, *chunk_count, *host_chunk_begin;
    // per-linearisation data
    float *recs, *res_toZero, *JpJdF, *deltaF, *priorF, *adHTdeltaF, *cDeltaF;
    uint8_t* flags;
    double *adHost, *adTarget;
    // results
    float *res_pt, *Hdd[2], *bd[2], *Hcd[2], *HdiF, *bdSum, *sc_partial;
    double *tile_partial, *acc[2], *accD, *accE, *accEB, *accHcc, *accbc, *hcc_host, *Hmat, *bvec, *prior_buf;
    long long* num[2];
    int* tile_nres;
    // device-side linearisation (edsgpu_ba_linearize): images, precalc, point state, results
    int H = 0, W = 0;
    float4* images = nullptr;    // [F][H*W], allocated by the first set_image
    void* lin_block = nullptr;   // precalc, thresholds, point arrays, state / energy / input flags
    float *precalc = nullptr, *frame_energy_th = nullptr, *pu = nullptr, *pv = nullptr, *idepth_zero = nullptr, *idepth = nullptr;
    float *color = nullptr, *weights = nullptr, *energy_new = nullptr;
    int32_t* state_new = nullptr;
    uint8_t *state_in = nullptr, *linearized = nullptr;
    float calib[4] = {0, 0, 0, 0};
    bool lin_inputs_set = false;
    unsigned images_set = 0;  // bit per frame
    double* solve_block = nullptr;  // edsgpu_ba_solve_system: three stitched systems, HM, bM, delta, projector, x (device)
    unsigned* grid_bar = nullptr;   // {arrivals, generation} of the accumulation kernels' grid barrier (device)
    bool have_top[2] = {false, false}, have_sc = false;  // which accumulations of the current linearisation are on the device
    std::vector<double> adHost_h, adTarget_h;  // host copies for xAd of the back-substitution
    void* post_block = nullptr;                // xAd (F*F*8 floats), cstep (4), step (P), energy partials
};

namespace {

BaDev ba_dev(const edsgpu_ba* w) {
    BaDev d{};
    d.F = w->F; d.P = w->P; d.R = w->R; d.num_tiles = w->num_tiles;
    d.host_idx = w->host_idx; d.target_idx = w->target_idx; d.point_of_res = w->point_of_res; d.res_begin = w->res_begin;
    d.perm = w->perm; d.tile_key = w->tile_key; d.tile_start = w->tile_start; d.tile_count = w->tile_count; d.key_tile_begin = w->key_tile_begin;
    d.recs = w->recs; d.res_toZero = w->res_toZero; d.deltaF = w->deltaF; d.priorF = w->priorF; d.adHTdeltaF = w->adHTdeltaF; d.cDeltaF = w->cDeltaF;
    d.flags = w->flags; d.res_pt = w->res_pt; d.tile_partial = w->tile_partial; d.tile_nres = w->tile_nres;
    d.grid_bar = w->grid_bar;
    return d;
}

// output slot of the next accumulation launch
void ba_dev_outputs(edsgpu_ba* w, BaDev& d, int slot) {
    d.acc_out = w->acc[slot]; d.num_out = w->num[slot];
    d.Hdd_out = w->Hdd[slot]; d.bd_out = w->bd[slot]; d.Hcd_out = w->Hcd[slot];
}

LinDev lin_dev(const edsgpu_ba* w, bool have_state_in, bool have_linearized) {
    LinDev d{};
    d.F = w->F; d.P = w->P; d.R = w->R; d.H = w->H; d.W = w->W;
    d.images = w->images; d.precalc = w->precalc;
    d.fxl = w->calib[0]; d.fyl = w->calib[1]; d.cxl = w->calib[2]; d.cyl = w->calib[3];
    d.frame_energy_th = w->frame_energy_th;
    d.pu = w->pu; d.pv = w->pv; d.idepth_zero = w->idepth_zero; d.idepth = w->idepth; d.color = w->color; d.weights = w->weights;
    d.host_idx = w->host_idx; d.target_idx = w->target_idx; d.point_of_res = w->point_of_res;
    d.state_in = have_state_in ? w->state_in : nullptr;
    d.linearized = have_linearized ? w->linearized : nullptr;
    d.recs = w->recs; d.flags = w->flags; d.state_new = w->state_new; d.energy_new = w->energy_new;
    return d;
}

// The accumulation kernels end with a grid-wide barrier: cooperative launch (every CTA resident, or the launch fails).
template <typename... Args>
cudaError_t launch_cooperative(void (*kernel)(Args...), int grid, int block, size_t smem, cudaStream_t stream, Args... args) {
    void* argv[] = {(void*)&args...};
    return cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid), dim3(block), argv, smem, stream);
}

edsgpu_status d2h(edsgpu_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!dst) return EDSGPU_OK;
    EDS_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return EDSGPU_OK;
}

}  // namespace

extern "C" {

edsgpu_status edsgpu_ba_create(edsgpu_ctx* ctx, int F, int P, int R, const int32_t* host_idx, const int32_t* target_idx,
                               const int32_t* res_begin, edsgpu_ba** out) {
    if (!ctx || !out) return EDSGPU_INVALID_ARGUMENT;
    EDS_REQUIRE(ctx, F >= 1 && F <= MAXF, "ba_create: F must be in [1,8]");
    EDS_REQUIRE(ctx, P >= 1 && R >= 1 && host_idx && target_idx && res_begin, "ba_create: bad arguments");
    EDS_REQUIRE(ctx, res_begin[0] == 0 && res_begin[P] == R, "ba_create: res_begin must span [0,R]");
    // ---- plan on the host: stable counting sort by (host,target), single-key tiles, points by host
    const int F2 = F * F;
    std::vector<int32_t> point_of_res(R), point_host(P, -1);
    for (int p = 0; p < P; ++p) {
        EDS_REQUIRE(ctx, res_begin[p] <= res_begin[p + 1], "ba_create: res_begin must be non-decreasing");
        for (int r = res_begin[p]; r < res_begin[p + 1]; ++r) {
            EDS_REQUIRE(ctx, host_idx[r] >= 0 && host_idx[r] < F && target_idx[r] >= 0 && target_idx[r] < F, "ba_create: frame index out of range");
            EDS_REQUIRE(ctx, point_host[p] < 0 || point_host[p] == host_idx[r], "ba_create: residuals of a point must share the host frame");
            point_host[p] = host_idx[r];
            point_of_res[r] = p;
        }
    }
    std::vector<int32_t> cnt(F2 + 1, 0), perm(R);
    for (int r = 0; r < R; ++r) cnt[host_idx[r] + F * target_idx[r] + 1]++;
    for (int k = 0; k < F2; ++k) cnt[k + 1] += cnt[k];
    {
        std::vector<int32_t> cur(cnt.begin(), cnt.end() - 1);
        for (int r = 0; r < R; ++r) perm[cur[host_idx[r] + F * target_idx[r]]++] = r;
    }
    // tile size: the smallest multiple of 64 (>= TOP_THREADS) whose tile count fits one wave of one CTA per SM
    int tile_sz = TOP_THREADS;
    for (;; tile_sz += 64) {
        long tiles = 0;
        for (int k = 0; k < F2; ++k) tiles += (cnt[k + 1] - cnt[k] + tile_sz - 1) / tile_sz;
        if (tiles <= ctx->num_sms || tile_sz >= 16384) break;
    }
    std::vector<int32_t> tile_key, tile_start, tile_count, key_tile_begin(F2 + 1, 0);
    for (int k = 0; k < F2; ++k) {
        key_tile_begin[k] = (int32_t)tile_key.size();
        for (int s = cnt[k]; s < cnt[k + 1]; s += tile_sz) {
            tile_key.push_back(k); tile_start.push_back(s); tile_count.push_back(std::min(tile_sz, cnt[k + 1] - s));
        }
    }
    key_tile_begin[F2] = (int32_t)tile_key.size();
    const int T = (int)tile_key.size();
    std::vector<int32_t> pt_perm, chunk_host, chunk_start, chunk_count, host_chunk_begin(F + 1, 0);
    for (int h = 0; h < F; ++h) {
        host_chunk_begin[h] = (int32_t)chunk_host.size();
        const int s0 = (int)pt_perm.size();
        for (int p = 0; p < P; ++p) if (point_host[p] == h) pt_perm.push_back(p);
        const int s1 = (int)pt_perm.size();
        for (int s = s0; s < s1; s += SC_CHUNK) { chunk_host.push_back(h); chunk_start.push_back(s); chunk_count.push_back(std::min(SC_CHUNK, s1 - s)); }
    }
    host_chunk_begin[F] = (int32_t)chunk_host.size();
    const int C = (int)chunk_host.size();
    const int Msc = 8 * F + 4, Nsc = 8 * F + 5, n = CPARS + 8 * F;

    DeviceGuard g(ctx->device);
    edsgpu_ba* w = new edsgpu_ba();
    w->ctx = ctx; w->F = F; w->P = P; w->R = R; w->num_tiles = T; w->num_chunks = C;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(std::max<size_t>(bytes, 16), 256); return o; };
    const size_t o_host = take(4 * (size_t)R), o_tgt = take(4 * (size_t)R), o_por = take(4 * (size_t)R), o_rb = take(4 * (size_t)(P + 1)), o_perm = take(4 * (size_t)R);
    const size_t o_tk = take(4 * (size_t)T), o_ts = take(4 * (size_t)T), o_tc = take(4 * (size_t)T), o_ktb = take(4 * (size_t)(F2 + 1));
    const size_t o_pp = take(4 * (size_t)pt_perm.size()), o_ch = take(4 * (size_t)C), o_cs = take(4 * (size_t)C), o_cc = take(4 * (size_t)C), o_hcb = take(4 * (size_t)(F + 1));
    const size_t o_recs = take(4 * (size_t)REC * R), o_rtz = take(32 * (size_t)R), o_jp = take(32 * (size_t)R), o_dF = take(4 * (size_t)P), o_pF = take(4 * (size_t)P);
    const size_t o_ad = take(32 * (size_t)F2), o_cd = take(16), o_fl = take((size_t)R), o_aH = take(512 * (size_t)F2), o_aT = take(512 * (size_t)F2);
    const size_t o_rpt = take(24 * (size_t)R);
    size_t o_Hdd[2], o_bd[2], o_Hcd[2], o_acc[2], o_num[2];
    for (int s = 0; s < 2; ++s) { o_Hdd[s] = take(4 * (size_t)P); o_bd[s] = take(4 * (size_t)P); o_Hcd[s] = take(16 * (size_t)P); o_acc[s] = take(8 * 169 * (size_t)F2); o_num[s] = take(8 * (size_t)F2); }
    const size_t o_hdi = take(4 * (size_t)P), o_bds = take(4 * (size_t)P), o_scp = take(4 * (size_t)C * Msc * Nsc), o_tp = take(8 * (size_t)NACC * T), o_tn = take(4 * (size_t)T);
    const size_t o_D = take(8 * 64 * (size_t)F2 * F), o_E = take(8 * 32 * (size_t)F2), o_EB = take(8 * 8 * (size_t)F2), o_Hcc = take(128), o_bc = take(32), o_hh = take(8 * 20 * (size_t)F);
    const size_t o_Hm = take(8 * (size_t)n * n), o_bv = take(8 * (size_t)n), o_pr = take(8 * (size_t)(4 + 16 * F)), o_gb = take(16);
    cudaError_t e = cudaMalloc(&w->block, off);
    if (e != cudaSuccess) { delete w; return edsgpu_fail(ctx, EDSGPU_OUT_OF_MEMORY, cudaGetErrorString(e)); }
    char* base = (char*)w->block;
    w->host_idx = (int32_t*)(base + o_host); w->target_idx = (int32_t*)(base + o_tgt); w->point_of_res = (int32_t*)(base + o_por);
    w->res_begin = (int32_t*)(base + o_rb); w->perm = (int32_t*)(base + o_perm); w->tile_key = (int32_t*)(base + o_tk);
    w->tile_start = (int32_t*)(base + o_ts); w->tile_count = (int32_t*)(base + o_tc); w->key_tile_begin = (int32_t*)(base + o_ktb);
    w->pt_perm = (int32_t*)(base + o_pp); w->chunk_host = (int32_t*)(base + o_ch); w->chunk_start = (int32_t*)(base + o_cs);
    w->chunk_count = (int32_t*)(base + o_cc); w->host_chunk_begin = (int32_t*)(base + o_hcb);
    w->recs = (float*)(base + o_recs); w->res_toZero = (float*)(base + o_rtz); w->JpJdF = (float*)(base + o_jp);
    w->deltaF = (float*)(base + o_dF); w->priorF = (float*)(base + o_pF); w->adHTdeltaF = (float*)(base + o_ad); w->cDeltaF = (float*)(base + o_cd);
    w->flags = (uint8_t*)(base + o_fl); w->adHost = (double*)(base + o_aH); w->adTarget = (double*)(base + o_aT);
    w->res_pt = (float*)(base + o_rpt);
    for (int s = 0; s < 2; ++s) {
        w->Hdd[s] = (float*)(base + o_Hdd[s]); w->bd[s] = (float*)(base + o_bd[s]); w->Hcd[s] = (float*)(base + o_Hcd[s]);
        w->acc[s] = (double*)(base + o_acc[s]); w->num[s] = (long long*)(base + o_num[s]);
    }
    w->HdiF = (float*)(base + o_hdi); w->bdSum = (float*)(base + o_bds); w->sc_partial = (float*)(base + o_scp);
    w->tile_partial = (double*)(base + o_tp); w->tile_nres = (int*)(base + o_tn);
    w->accD = (double*)(base + o_D); w->accE = (double*)(base + o_E); w->accEB = (double*)(base + o_EB); w->accHcc = (double*)(base + o_Hcc);
    w->accbc = (double*)(base + o_bc); w->hcc_host = (double*)(base + o_hh); w->Hmat = (double*)(base + o_Hm); w->bvec = (double*)(base + o_bv);
    w->prior_buf = (double*)(base + o_pr);
    w->grid_bar = (unsigned*)(base + o_gb);
    e = cudaMemsetAsync(w->block, 0, off, ctx->stream);
    auto up = [&](void* dst, const void* src, size_t bytes) { if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream); };
    up(w->host_idx, host_idx, 4 * (size_t)R); up(w->target_idx, target_idx, 4 * (size_t)R); up(w->point_of_res, point_of_res.data(), 4 * (size_t)R);
    up(w->res_begin, res_begin, 4 * (size_t)(P + 1)); up(w->perm, perm.data(), 4 * (size_t)R);
    up(w->tile_key, tile_key.data(), 4 * (size_t)T); up(w->tile_start, tile_start.data(), 4 * (size_t)T); up(w->tile_count, tile_count.data(), 4 * (size_t)T);
    up(w->key_tile_begin, key_tile_begin.data(), 4 * (size_t)(F2 + 1)); up(w->pt_perm, pt_perm.data(), 4 * pt_perm.size());
    up(w->chunk_host, chunk_host.data(), 4 * (size_t)C); up(w->chunk_start, chunk_start.data(), 4 * (size_t)C); up(w->chunk_count, chunk_count.data(), 4 * (size_t)C);
    up(w->host_chunk_begin, host_chunk_begin.data(), 4 * (size_t)(F + 1));
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // the plan vectors are stack-owned
    if (e != cudaSuccess) { edsgpu_ba_destroy(w); return edsgpu_fail(ctx, EDSGPU_CUDA_ERROR, cudaGetErrorString(e)); }
    *out = w;
    return EDSGPU_OK;
}

void edsgpu_ba_destroy(edsgpu_ba* w) {
    if (!w) return;
    DeviceGuard g(w->ctx->device);
    cudaStreamSynchronize(w->ctx->stream);
    if (w->block) cudaFree(w->block);
    if (w->images) cudaFree(w->images);
    if (w->lin_block) cudaFree(w->lin_block);
    if (w->post_block) cudaFree(w->post_block);
    if (w->solve_block) cudaFree(w->solve_block);
    delete w;
}

edsgpu_status edsgpu_ba_set_residuals(edsgpu_ba* w, const float* recs, const uint8_t* flags, const float* res_toZero) {
    if (!w) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = w->ctx;
    EDS_REQUIRE(ctx, recs && flags, "ba_set_residuals: null array");
    DeviceGuard g(ctx->device);
    w->have_top[0] = w->have_top[1] = w->have_sc = false;
    EDS_CUDA(ctx, cudaMemcpyAsync(w->recs, recs, 4 * (size_t)REC * w->R, cudaMemcpyHostToDevice, ctx->stream));
    EDS_CUDA(ctx, cudaMemcpyAsync(w->flags, flags, (size_t)w->R, cudaMemcpyHostToDevice, ctx->stream));
    if (res_toZero) EDS_CUDA(ctx, cudaMemcpyAsync(w->res_toZero, res_toZero, 32 * (size_t)w->R, cudaMemcpyHostToDevice, ctx->stream));
    ba_jpjd_kernel<<<(w->R + 255) / 256, 256, 0, ctx->stream>>>(w->recs, w->R, w->JpJdF);
    ctx->launches++;
    EDS_CUDA(ctx, cudaGetLastError());
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // caller buffers may be pageable and reused
    return EDSGPU_OK;
}

edsgpu_status edsgpu_ba_set_points(edsgpu_ba* w, const float* deltaF, const float* priorF) {
    if (!w) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = w->ctx;
    DeviceGuard g(ctx->device);
    w->have_top[1] = w->have_sc = false;  // deltaF feeds the re-linearisation of mode 1, deltaF / priorF the Schur terms
    if (deltaF) EDS_CUDA(ctx, cudaMemcpyAsync(w->deltaF, deltaF, 4 * (size_t)w->P, cudaMemcpyHostToDevice, ctx->stream));
    else EDS_CUDA(ctx, cudaMemsetAsync(w->deltaF, 0, 4 * (size_t)w->P, ctx->stream));
    if (priorF) EDS_CUDA(ctx, cudaMemcpyAsync(w->priorF, priorF, 4 * (size_t)w->P, cudaMemcpyHostToDevice, ctx->stream));
    else EDS_CUDA(ctx, cudaMemsetAsync(w->priorF, 0, 4 * (size_t)w->P, ctx->stream));
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EDSGPU_OK;
}

edsgpu_status edsgpu_ba_set_frames(edsgpu_ba* w, const float* adHTdeltaF, const float* cDeltaF, const double* adHost, const double* adTarget) {
    if (!w) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = w->ctx;
    DeviceGuard g(ctx->device);
    const size_t F2 = (size_t)w->F * w->F;
    if (adHTdeltaF || cDeltaF) w->have_top[1] = w->have_sc = false;  // inputs of the mode-1 re-linearisation
    if (adHTdeltaF) EDS_CUDA(ctx, cudaMemcpyAsync(w->adHTdeltaF, adHTdeltaF, 32 * F2, cudaMemcpyHostToDevice, ctx->stream));
    if (cDeltaF) EDS_CUDA(ctx, cudaMemcpyAsync(w->cDeltaF, cDeltaF, 16, cudaMemcpyHostToDevice, ctx->stream));
    if (adHost) EDS_CUDA(ctx, cudaMemcpyAsync(w->adHost, adHost, 512 * F2, cudaMemcpyHostToDevice, ctx->stream));
    if (adTarget) EDS_CUDA(ctx, cudaMemcpyAsync(w->adTarget, adTarget, 512 * F2, cudaMemcpyHostToDevice, ctx->stream));
    if (adHost) w->adHost_h.assign(adHost, adHost + 64 * F2);
    if (adTarget) w->adTarget_h.assign(adTarget, adTarget + 64 * F2);
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EDSGPU_OK;
}

edsgpu_status edsgpu_ba_top_accumulate(edsgpu_ba* w, int mode, double* acc_out, float* Hdd_out, float* bd_out, float* Hcd_out, int64_t* nres_out) {
    EDS_RANGE("edsgpu_ba_top_accumulate");
    if (!w) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = w->ctx;
    EDS_REQUIRE(ctx, mode >= 0 && mode <= 2, "ba_top_accumulate: mode must be 0 (active), 1 (linearized) or 2 (marginalize)");
    DeviceGuard g(ctx->device);
    const int slot = mode == 0 ? 0 : 1;
    // a new accumulation of either side makes the Schur terms that were built from the old one stale
    w->have_top[slot] = false;
    w->have_sc = false;
    if (mode == 2) w->have_top[0] = false;
    BaDev d = ba_dev(w);
    ba_dev_outputs(w, d, slot);
    const size_t stage_bytes = (size_t)TOP_STAGE * REC * sizeof(float);
    // one kernel: tile partials, grid barrier, ordered sums over tiles and over each point's residuals
    if (mode == 0) {
        EDS_CUDA(ctx, cudaFuncSetAttribute(ba_top_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes));
        EDS_CUDA(ctx, launch_cooperative(ba_top_kernel<0>, w->num_tiles, TOP_THREADS, stage_bytes, ctx->stream, d));
    } else if (mode == 1) {
        EDS_CUDA(ctx, cudaFuncSetAttribute(ba_top_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes));
        EDS_CUDA(ctx, launch_cooperative(ba_top_kernel<1>, w->num_tiles, TOP_THREADS, stage_bytes, ctx->stream, d));
    } else {
        EDS_CUDA(ctx, cudaFuncSetAttribute(ba_top_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes));
        EDS_CUDA(ctx, launch_cooperative(ba_top_kernel<2>, w->num_tiles, TOP_THREADS, stage_bytes, ctx->stream, d));
    }
    ctx->launches += 1;
    EDS_CUDA(ctx, cudaGetLastError());
    if (mode == 2) {  // AccumulatedTopHessian.cpp:152-157: marginalisation zeroes the active-side point terms
        EDS_CUDA(ctx, cudaMemsetAsync(w->Hdd[0], 0, 4 * (size_t)w->P, ctx->stream));
        EDS_CUDA(ctx, cudaMemsetAsync(w->bd[0], 0, 4 * (size_t)w->P, ctx->stream));
        EDS_CUDA(ctx, cudaMemsetAsync(w->Hcd[0], 0, 16 * (size_t)w->P, ctx->stream));
    }
    // set only once everything above was queued without error.  addPoint<2> defines BOTH point-term sides (the
    // linearized one and, as zeros, the active one), so marginalizePointsF's sequence addPoint<2> -> SC addPoint
    // (EnergyFunctional.cpp:538-560) needs no separate active pass.
    w->have_top[slot] = true;
    if (mode == 2) w->have_top[0] = true;
    if (acc_out || Hdd_out || bd_out || Hcd_out || nres_out) return edsgpu_ba_top_read(w, slot, acc_out, Hdd_out, bd_out, Hcd_out, nres_out);
    return EDSGPU_OK;
}

edsgpu_status edsgpu_ba_top_read(edsgpu_ba* w, int which, double* acc_out, float* Hdd_out, float* bd_out, float* Hcd_out, int64_t* nres_out) {
    if (!w) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = w->ctx;
    EDS_REQUIRE(ctx, which == 0 || which == 1, "ba_top_read: which must be 0 (active) or 1 (linearized)");
    EDS_REQUIRE(ctx, w->have_top[which], "ba_top_read: this side has not been accumulated for the current linearisation");
    DeviceGuard g(ctx->device);
    const int slot = which;
    const size_t F2 = (size_t)w->F * w->F;
    edsgpu_status st = edsgpu_ensure_pinned(ctx, 8 * F2);
    if (st != EDSGPU_OK) return st;
    if ((st = d2h(ctx, acc_out, w->acc[slot], 8 * 169 * F2)) != EDSGPU_OK) return st;
    if ((st = d2h(ctx, Hdd_out, w->Hdd[slot], 4 * (size_t)w->P)) != EDSGPU_OK) return st;
    if ((st = d2h(ctx, bd_out, w->bd[slot], 4 * (size_t)w->P)) != EDSGPU_OK) return st;
    if ((st = d2h(ctx, Hcd_out, w->Hcd[slot], 16 * (size_t)w->P)) != EDSGPU_OK) return st;
    if (nres_out) EDS_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, w->num[slot], 8 * F2, cudaMemcpyDeviceToHost, ctx->stream));
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (nres_out) { int64_t n = 0; for (size_t k = 0; k < F2; ++k) n += ((long long*)ctx->pinned)[k]; *nres_out = n; }
    return EDSGPU_OK;
}

edsgpu_status edsgpu_ba_top_stitch(edsgpu_ba* w, int which, int use_prior, const double* cPrior, const double* frame_prior,
                                   const double* frame_delta_prior, double* H, double* b) {
    if (!w) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = w->ctx;
    EDS_REQUIRE(ctx, which == 0 || which == 1, "ba_top_stitch: which must be 0 (active) or 1 (linearized)");
    EDS_REQUIRE(ctx, !use_prior || (cPrior && frame_prior && frame_delta_prior), "ba_top_stitch: priors requested but not given");
    DeviceGuard g(ctx->device);
    const int F = w->F, n = CPARS + 8 * F;
    double* pr = w->prior_buf;  // cPrior(4) | frame_prior(8F) | frame_delta_prior(8F)
    if (use_prior) {
        EDS_CUDA(ctx, cudaMemcpyAsync(pr, cPrior, 32, cudaMemcpyHostToDevice, ctx->stream));
        EDS_CUDA(ctx, cudaMemcpyAsync(pr + 4, frame_prior, 64 * (size_t)F, cudaMemcpyHostToDevice, ctx->stream));
        EDS_CUDA(ctx, cudaMemcpyAsync(pr + 4 + 8 * F, frame_delta_prior, 64 * (size_t)F, cudaMemcpyHostToDevice, ctx->stream));
    }
    ba_top_stitch_kernel<<<F * F, STITCH_THREADS, 0, ctx->stream>>>(F, w->acc[which], w->adHost, w->adTarget, use_prior, pr, w->cDeltaF, pr + 4, pr + 4 + 8 * F,
                                                        w->Hmat, w->bvec);
    ba_top_symmetrise_kernel<<<F * F, 64, 0, ctx->stream>>>(F, use_prior, pr + 4, w->Hmat);
    ctx->launches += 2;
    EDS_CUDA(ctx, cudaGetLastError());
    edsgpu_status st;
    if ((st = d2h(ctx, H, w->Hmat, 8 * (size_t)n * n)) != EDSGPU_OK) return st;
    if ((st = d2h(ctx, b, w->bvec, 8 * (size_t)n)) != EDSGPU_OK) return st;
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EDSGPU_OK;
}

edsgpu_status edsgpu_ba_sc_accumulate(edsgpu_ba* w, int shift_prior_to_zero, double* accD, double* accE, double* accEB, double* accHcc,
                                      double* accbc, float* HdiF_out, float* bdSum_out) {
    EDS_RANGE("edsgpu_ba_sc_accumulate");
    if (!w) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = w->ctx;
    DeviceGuard g(ctx->device);
    EDS_REQUIRE(ctx, w->have_top[0] && w->have_top[1], "ba_sc_accumulate: run top_accumulate for the active and the linearized residuals first");
    w->have_sc = false;
    ScDev s{};
    s.F = w->F; s.P = w->P; s.shift_prior = shift_prior_to_zero;
    s.res_begin = w->res_begin; s.target_idx = w->target_idx; s.pt_perm = w->pt_perm; s.chunk_host = w->chunk_host;
    s.chunk_start = w->chunk_start; s.chunk_count = w->chunk_count; s.flags = w->flags; s.JpJdF = w->JpJdF;
    s.HddA = w->Hdd[0]; s.HddL = w->Hdd[1]; s.bdA = w->bd[0]; s.bdL = w->bd[1]; s.HcdA = w->Hcd[0]; s.HcdL = w->Hcd[1];
    s.priorF = w->priorF; s.deltaF = w->deltaF; s.HdiF = w->HdiF; s.bdSum = w->bdSum; s.partial = w->sc_partial;
    ba_sc_kernel<<<w->num_chunks, SC_THREADS, 0, ctx->stream>>>(s);
    EDS_CUDA(ctx, cudaGetLastError());
    const size_t F2 = (size_t)w->F * w->F;
    {
        const int MN = (8 * w->F + 4) * (8 * w->F + 5);
        ba_sc_finalize_kernel<<<dim3(w->F, (MN + 127) / 128), 128, 0, ctx->stream>>>(w->F, w->sc_partial, w->host_chunk_begin, w->accD, w->accE, w->accEB, w->hcc_host);
    }
    ba_sc_hcc_kernel<<<1, 32, 0, ctx->stream>>>(w->F, w->hcc_host, w->accHcc, w->accbc);
    ctx->launches += 3;
    EDS_CUDA(ctx, cudaGetLastError());
    w->have_sc = true;
    if (accD || accE || accEB || accHcc || accbc || HdiF_out || bdSum_out) {
        edsgpu_status st;
        if ((st = d2h(ctx, accD, w->accD, 8 * 64 * F2 * w->F)) != EDSGPU_OK) return st;
        if ((st = d2h(ctx, accE, w->accE, 8 * 32 * F2)) != EDSGPU_OK) return st;
        if ((st = d2h(ctx, accEB, w->accEB, 8 * 8 * F2)) != EDSGPU_OK) return st;
        if ((st = d2h(ctx, accHcc, w->accHcc, 128)) != EDSGPU_OK) return st;
        if ((st = d2h(ctx, accbc, w->accbc, 32)) != EDSGPU_OK) return st;
        if ((st = d2h(ctx, HdiF_out, w->HdiF, 4 * (size_t)w->P)) != EDSGPU_OK) return st;
        if ((st = d2h(ctx, bdSum_out, w->bdSum, 4 * (size_t)w->P)) != EDSGPU_OK) return st;
        EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return EDSGPU_OK;
}

edsgpu_status edsgpu_ba_sc_stitch(edsgpu_ba* w, double* H, double* b) {
    if (!w) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = w->ctx;
    DeviceGuard g(ctx->device);
    const int F = w->F, n = CPARS + 8 * F;
    ba_sc_stitch_kernel<<<F * F, STITCH_THREADS, 0, ctx->stream>>>(F, w->accD, w->accE, w->accEB, w->accHcc, w->accbc, w->adHost, w->adTarget, w->Hmat, w->bvec);
    ctx->launches++;
    EDS_CUDA(ctx, cudaGetLastError());
    edsgpu_status st;
    if ((st = d2h(ctx, H, w->Hmat, 8 * (size_t)n * n)) != EDSGPU_OK) return st;
    if ((st = d2h(ctx, b, w->bvec, 8 * (size_t)n)) != EDSGPU_OK) return st;
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EDSGPU_OK;
}

edsgpu_status edsgpu_ba_get_jpjd(edsgpu_ba* w, float* JpJdF_out) {
    if (!w || !JpJdF_out) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = w->ctx;
    DeviceGuard g(ctx->device);
    EDS_CUDA(ctx, cudaMemcpyAsync(JpJdF_out, w->JpJdF, 32 * (size_t)w->R, cudaMemcpyDeviceToHost, ctx->stream));
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EDSGPU_OK;
}

// ---- device-side linearisation -------------------------------------------------------------
edsgpu_status edsgpu_ba_set_image(edsgpu_ba* w, int frame, int height, int width, const float* dI) {
    if (!w) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = w->ctx;
    EDS_REQUIRE(ctx, dI && frame >= 0 && frame < w->F && height > 4 && width > 4, "ba_set_image: bad arguments");
    EDS_REQUIRE(ctx, w->images == nullptr || (height == w->H && width == w->W), "ba_set_image: all frames of a window share one size");
    DeviceGuard g(ctx->device);
    const size_t npix = (size_t)height * width;
    if (!w->images) {
        cudaError_t e = cudaMalloc(&w->images, sizeof(float4) * npix * w->F);
        if (e != cudaSuccess) return edsgpu_fail(ctx, e == cudaErrorMemoryAllocation ? EDSGPU_OUT_OF_MEMORY : EDSGPU_CUDA_ERROR, cudaGetErrorString(e));
        w->H = height; w->W = width;
    }
    edsgpu_status st = edsgpu_ensure_scratch(ctx, npix * 3 * sizeof(float));
    if (st != EDSGPU_OK) return st;
    EDS_CUDA(ctx, cudaMemcpyAsync(ctx->scratch, dI, npix * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    ba_pad_image_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, ctx->stream>>>((const float*)ctx->scratch, w->images + npix * frame, npix);
    ctx->launches++;
    EDS_CUDA(ctx, cudaGetLastError());
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the caller's buffer may be pageable and reused; scratch is shared
    w->images_set |= 1u << frame;
    return EDSGPU_OK;
}

edsgpu_status edsgpu_ba_set_linearize_inputs(edsgpu_ba* w, const float* precalc, const float calib[4], const float* frame_energy_th,
                                             const float* u, const float* v, const float* idepth_zero_scaled, const float* idepth_scaled,
                                             const float* color, const float* weights) {
    if (!w) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = w->ctx;
    EDS_REQUIRE(ctx, precalc && calib && frame_energy_th && u && v && idepth_zero_scaled && idepth_scaled && color && weights,
                "ba_set_linearize_inputs: null array");
    DeviceGuard g(ctx->device);
    const size_t F = w->F, P = w->P, R = w->R;
    if (!w->lin_block) {
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
        const size_t o_pre = take(4 * PRECALC * F * F), o_th = take(4 * F), o_u = take(4 * P), o_v = take(4 * P), o_iz = take(4 * P), o_id = take(4 * P);
        const size_t o_c = take(32 * P), o_w = take(32 * P), o_e = take(4 * R), o_s = take(4 * R), o_si = take(R), o_l = take(R);
        cudaError_t e = cudaMalloc(&w->lin_block, off);
        if (e != cudaSuccess) return edsgpu_fail(ctx, e == cudaErrorMemoryAllocation ? EDSGPU_OUT_OF_MEMORY : EDSGPU_CUDA_ERROR, cudaGetErrorString(e));
        char* b = (char*)w->lin_block;
        w->precalc = (float*)(b + o_pre); w->frame_energy_th = (float*)(b + o_th); w->pu = (float*)(b + o_u); w->pv = (float*)(b + o_v);
        w->idepth_zero = (float*)(b + o_iz); w->idepth = (float*)(b + o_id); w->color = (float*)(b + o_c); w->weights = (float*)(b + o_w);
        w->energy_new = (float*)(b + o_e); w->state_new = (int32_t*)(b + o_s); w->state_in = (uint8_t*)(b + o_si); w->linearized = (uint8_t*)(b + o_l);
    }
    cudaStream_t s = ctx->stream;
    EDS_CUDA(ctx, cudaMemcpyAsync(w->precalc, precalc, 4 * PRECALC * F * F, cudaMemcpyHostToDevice, s));
    EDS_CUDA(ctx, cudaMemcpyAsync(w->frame_energy_th, frame_energy_th, 4 * F, cudaMemcpyHostToDevice, s));
    EDS_CUDA(ctx, cudaMemcpyAsync(w->pu, u, 4 * P, cudaMemcpyHostToDevice, s));
    EDS_CUDA(ctx, cudaMemcpyAsync(w->pv, v, 4 * P, cudaMemcpyHostToDevice, s));
    EDS_CUDA(ctx, cudaMemcpyAsync(w->idepth_zero, idepth_zero_scaled, 4 * P, cudaMemcpyHostToDevice, s));
    EDS_CUDA(ctx, cudaMemcpyAsync(w->idepth, idepth_scaled, 4 * P, cudaMemcpyHostToDevice, s));
    EDS_CUDA(ctx, cudaMemcpyAsync(w->color, color, 32 * P, cudaMemcpyHostToDevice, s));
    EDS_CUDA(ctx, cudaMemcpyAsync(w->weights, weights, 32 * P, cudaMemcpyHostToDevice, s));
    EDS_CUDA(ctx, cudaStreamSynchronize(s));
    memcpy(w->calib, calib, sizeof(w->calib));
    w->lin_inputs_set = true;
    return EDSGPU_OK;
}

edsgpu_status edsgpu_ba_linearize(edsgpu_ba* w, const uint8_t* state_in, const uint8_t* linearized, const float* res_toZero,
                                  int32_t* state_out, float* energy_out) {
    EDS_RANGE("edsgpu_ba_linearize");
    if (!w) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = w->ctx;
    EDS_REQUIRE(ctx, w->lin_inputs_set, "ba_linearize: call edsgpu_ba_set_linearize_inputs first");
    EDS_REQUIRE(ctx, w->images && w->images_set == (1u << w->F) - 1u, "ba_linearize: call edsgpu_ba_set_image for every frame first");
    DeviceGuard g(ctx->device);
    cudaStream_t s = ctx->stream;
    if (state_in) EDS_CUDA(ctx, cudaMemcpyAsync(w->state_in, state_in, (size_t)w->R, cudaMemcpyHostToDevice, s));
    if (linearized) EDS_CUDA(ctx, cudaMemcpyAsync(w->linearized, linearized, (size_t)w->R, cudaMemcpyHostToDevice, s));
    if (res_toZero) EDS_CUDA(ctx, cudaMemcpyAsync(w->res_toZero, res_toZero, 32 * (size_t)w->R, cudaMemcpyHostToDevice, s));
    w->have_top[0] = w->have_top[1] = w->have_sc = false;
    const LinDev d = lin_dev(w, state_in != nullptr, linearized != nullptr);
    ba_linearize_kernel<<<(w->R + 127) / 128, 128, 0, s>>>(d);
    ctx->launches++;
    EDS_CUDA(ctx, cudaGetLastError());
    ba_jpjd_kernel<<<(w->R + 255) / 256, 256, 0, s>>>(w->recs, w->R, w->JpJdF);
    ctx->launches++;
    EDS_CUDA(ctx, cudaGetLastError());
    if (state_out) EDS_CUDA(ctx, cudaMemcpyAsync(state_out, w->state_new, 4 * (size_t)w->R, cudaMemcpyDeviceToHost, s));
    if (energy_out) EDS_CUDA(ctx, cudaMemcpyAsync(energy_out, w->energy_new, 4 * (size_t)w->R, cudaMemcpyDeviceToHost, s));
    if (state_in || linearized || res_toZero || state_out || energy_out) EDS_CUDA(ctx, cudaStreamSynchronize(s));
    return EDSGPU_OK;
}

edsgpu_status edsgpu_ba_linearize_accumulate(edsgpu_ba* w, const uint8_t* state_in, const uint8_t* linearized, const float* res_toZero,
                                             int write_records, int32_t* state_out, float* energy_out) {
    EDS_RANGE("edsgpu_ba_linearize_accumulate");
    if (!w) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = w->ctx;
    EDS_REQUIRE(ctx, w->lin_inputs_set, "ba_linearize_accumulate: call edsgpu_ba_set_linearize_inputs first");
    EDS_REQUIRE(ctx, w->images && w->images_set == (1u << w->F) - 1u, "ba_linearize_accumulate: call edsgpu_ba_set_image for every frame first");
    DeviceGuard g(ctx->device);
    cudaStream_t s = ctx->stream;
    if (state_in) EDS_CUDA(ctx, cudaMemcpyAsync(w->state_in, state_in, (size_t)w->R, cudaMemcpyHostToDevice, s));
    if (linearized) EDS_CUDA(ctx, cudaMemcpyAsync(w->linearized, linearized, (size_t)w->R, cudaMemcpyHostToDevice, s));
    if (res_toZero) EDS_CUDA(ctx, cudaMemcpyAsync(w->res_toZero, res_toZero, 32 * (size_t)w->R, cudaMemcpyHostToDevice, s));
    w->have_top[0] = w->have_top[1] = w->have_sc = false;
    const LinDev d = lin_dev(w, state_in != nullptr, linearized != nullptr);
    BaDev t = ba_dev(w);
    ba_dev_outputs(w, t, 0);
    EDS_CUDA(ctx, launch_cooperative(ba_lin_top_kernel, w->num_tiles, TOP_THREADS, (size_t)0, s, d, t, w->JpJdF, write_records ? 1 : 0));
    ctx->launches++;
    EDS_CUDA(ctx, cudaGetLastError());
    w->have_top[0] = true;
    if (state_out) EDS_CUDA(ctx, cudaMemcpyAsync(state_out, w->state_new, 4 * (size_t)w->R, cudaMemcpyDeviceToHost, s));
    if (energy_out) EDS_CUDA(ctx, cudaMemcpyAsync(energy_out, w->energy_new, 4 * (size_t)w->R, cudaMemcpyDeviceToHost, s));
    if (state_in || linearized || res_toZero || state_out || energy_out) EDS_CUDA(ctx, cudaStreamSynchronize(s));
    return EDSGPU_OK;
}

// debug/parity: the records as they sit on the device (uploaded or linearised there)
edsgpu_status edsgpu_ba_get_residuals(edsgpu_ba* w, float* recs_out, uint8_t* flags_out) {
    if (!w) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = w->ctx;
    DeviceGuard g(ctx->device);
    if (recs_out) EDS_CUDA(ctx, cudaMemcpyAsync(recs_out, w->recs, 4 * (size_t)REC * w->R, cudaMemcpyDeviceToHost, ctx->stream));
    if (flags_out) EDS_CUDA(ctx, cudaMemcpyAsync(flags_out, w->flags, (size_t)w->R, cudaMemcpyDeviceToHost, ctx->stream));
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EDSGPU_OK;
}

// ---- after the solve ------------------------------------------------------------------------
namespace {
edsgpu_status ensure_post_block(edsgpu_ba* w) {
    if (w->post_block) return EDSGPU_OK;
    const size_t n_part = (size_t)(std::max(w->R, w->P) + LE_THREADS - 1) / LE_THREADS;
    const size_t bytes = align_up(32 * (size_t)w->F * w->F, 256) + 256 + align_up(4 * (size_t)w->P, 256) + align_up(8 * n_part, 256) + 256;
    cudaError_t e = cudaMalloc(&w->post_block, bytes);
    if (e != cudaSuccess) return edsgpu_fail(w->ctx, e == cudaErrorMemoryAllocation ? EDSGPU_OUT_OF_MEMORY : EDSGPU_CUDA_ERROR, cudaGetErrorString(e));
    return EDSGPU_OK;
}
}  // namespace

edsgpu_status edsgpu_ba_resubstitute(edsgpu_ba* w, const double* x, float* point_step_out) {
    if (!w) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = w->ctx;
    EDS_REQUIRE(ctx, x != nullptr, "ba_resubstitute: null x");
    const size_t F = w->F, F2 = F * F;
    EDS_REQUIRE(ctx, w->adHost_h.size() == 64 * F2 && w->adTarget_h.size() == 64 * F2, "ba_resubstitute: call edsgpu_ba_set_frames with the adjoints first");
    EDS_REQUIRE(ctx, w->have_top[0] && w->have_top[1] && w->have_sc, "ba_resubstitute: needs top_accumulate(0), top_accumulate(1) and sc_accumulate of this linearisation");
    DeviceGuard g(ctx->device);
    edsgpu_status st = ensure_post_block(w);
    if (st == EDSGPU_OK) st = edsgpu_ensure_pinned(ctx, 4 * (8 * F2 + 4));
    if (st != EDSGPU_OK) return st;
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the pinned block is shared
    // xAd[F*h + t] = xF_h^T adHostF[h + F*t] + xF_t^T adTargetF[h + F*t]  (EnergyFunctional.cpp:272-281), float
    float* xAd = (float*)ctx->pinned;
    float* cstep = xAd + 8 * F2;
    std::vector<float> xF(4 + 8 * F);
    for (size_t i = 0; i < xF.size(); ++i) xF[i] = (float)x[i];
    for (size_t h = 0; h < F; ++h)
        for (size_t t = 0; t < F; ++t) {
            const double* AH = &w->adHost_h[64 * (h + F * t)];
            const double* AT = &w->adTarget_h[64 * (h + F * t)];
            for (int c = 0; c < 8; ++c) {
                float a = 0.f, b = 0.f;
                for (int k = 0; k < 8; ++k) a += xF[4 + 8 * h + k] * (float)AH[c * 8 + k];
                for (int k = 0; k < 8; ++k) b += xF[4 + 8 * t + k] * (float)AT[c * 8 + k];
                xAd[8 * (F * h + t) + c] = a + b;
            }
        }
    for (int k = 0; k < 4; ++k) cstep[k] = xF[k];
    char* pb = (char*)w->post_block;
    float* d_xAd = (float*)pb;
    float* d_cstep = (float*)(pb + align_up(32 * F2, 256));
    float* d_step = (float*)(pb + align_up(32 * F2, 256) + 256);
    EDS_CUDA(ctx, cudaMemcpyAsync(d_xAd, xAd, 32 * F2, cudaMemcpyHostToDevice, ctx->stream));
    EDS_CUDA(ctx, cudaMemcpyAsync(d_cstep, cstep, 16, cudaMemcpyHostToDevice, ctx->stream));
    ba_resubstitute_kernel<<<(w->P + 127) / 128, 128, 0, ctx->stream>>>(w->F, w->P, w->res_begin, w->host_idx, w->target_idx, w->flags, w->JpJdF,
                                                                       w->bdSum, w->Hcd[0], w->Hcd[1], w->HdiF, d_xAd, d_cstep, d_step);
    ctx->launches++;
    EDS_CUDA(ctx, cudaGetLastError());
    if (point_step_out) EDS_CUDA(ctx, cudaMemcpyAsync(point_step_out, d_step, 4 * (size_t)w->P, cudaMemcpyDeviceToHost, ctx->stream));
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EDSGPU_OK;
}

edsgpu_status edsgpu_ba_fix_linearization(edsgpu_ba* w, const uint8_t* select, float* res_toZero_out) {
    if (!w) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = w->ctx;
    DeviceGuard g(ctx->device);
    uint8_t* d_select = nullptr;
    if (select) {
        edsgpu_status st = edsgpu_ensure_scratch(ctx, (size_t)w->R);
        if (st != EDSGPU_OK) return st;
        d_select = (uint8_t*)ctx->scratch;
        EDS_CUDA(ctx, cudaMemcpyAsync(d_select, select, (size_t)w->R, cudaMemcpyHostToDevice, ctx->stream));
    }
    ba_fix_linearization_kernel<<<(w->R + 255) / 256, 256, 0, ctx->stream>>>(w->F, w->R, w->recs, w->host_idx, w->target_idx, w->point_of_res,
                                                                            w->deltaF, w->adHTdeltaF, w->cDeltaF, d_select, w->res_toZero, w->flags);
    ctx->launches++;
    EDS_CUDA(ctx, cudaGetLastError());
    w->have_top[0] = w->have_top[1] = w->have_sc = false;  // flags (isLinearized) and res_toZero changed under every accumulation
    if (res_toZero_out) EDS_CUDA(ctx, cudaMemcpyAsync(res_toZero_out, w->res_toZero, 32 * (size_t)w->R, cudaMemcpyDeviceToHost, ctx->stream));
    if (select || res_toZero_out) EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EDSGPU_OK;
}

edsgpu_status edsgpu_ba_solve_system(edsgpu_ba* w, double lambda, const double* HM, const double* bM, const double* delta, const double* cPrior,
                                     const double* frame_prior, const double* frame_delta_prior, const double* nullspace_projector, double* x_out,
                                     float* point_step_out) {
    EDS_RANGE("edsgpu_ba_solve_system");
    if (!w) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = w->ctx;
    const int F = w->F, n = CPARS + 8 * F;
    const size_t F2 = (size_t)F * F, nn = (size_t)n * n;
    EDS_REQUIRE(ctx, x_out != nullptr, "ba_solve_system: null x_out");
    EDS_REQUIRE(ctx, (HM == nullptr) == (bM == nullptr), "ba_solve_system: give HM and bM or neither");
    EDS_REQUIRE(ctx, (cPrior == nullptr) == (frame_prior == nullptr) && (cPrior == nullptr) == (frame_delta_prior == nullptr),
                "ba_solve_system: give all three priors or none");
    EDS_REQUIRE(ctx, w->adHost_h.size() == 64 * F2 && w->adTarget_h.size() == 64 * F2, "ba_solve_system: call edsgpu_ba_set_frames with the adjoints first");
    EDS_REQUIRE(ctx, w->have_top[0] && w->have_top[1] && w->have_sc, "ba_solve_system: needs the active, the linearized and the Schur accumulation of this linearisation");
    DeviceGuard g(ctx->device);
    edsgpu_status st = ensure_post_block(w);
    if (st != EDSGPU_OK) return st;
    // device block: three stitched systems (H n*n, b n) | HM | bM | delta | projector | priors (cPrior 4, frame_prior 8F, frame_delta_prior 8F) | x
    const size_t sys = nn + (size_t)n, n_in = 2 * nn + 2 * (size_t)n + 4 + 16 * (size_t)F, total = 3 * sys + n_in + n;
    if (!w->solve_block) {
        cudaError_t e = cudaMalloc(&w->solve_block, sizeof(double) * total);
        if (e != cudaSuccess) return edsgpu_fail(ctx, e == cudaErrorMemoryAllocation ? EDSGPU_OUT_OF_MEMORY : EDSGPU_CUDA_ERROR, cudaGetErrorString(e));
    }
    double* S = w->solve_block;
    double *HAd = S, *bAd = S + nn, *HLd = S + sys, *bLd = S + sys + nn, *Hsd = S + 2 * sys, *bsd = S + 2 * sys + nn;
    double *HMd = S + 3 * sys, *bMd = HMd + nn, *dld = bMd + n, *prd = dld + n, *pr = prd + nn, *xd = pr + 4 + 16 * F;
    cudaStream_t s = ctx->stream;
    // the caller's inputs go through one pinned staging block and one copy (the caller's arrays are pageable: seven
    // separate copies of them cost more than the three stitches)
    if ((st = edsgpu_ensure_pinned(ctx, sizeof(double) * n_in)) != EDSGPU_OK) return st;
    {
        double* hst = (double*)ctx->pinned;
        size_t lo = n_in, hi = 0;  // the range of the block that carries something
        auto put = [&](size_t off, const double* src, size_t cnt) {
            memcpy(hst + off, src, sizeof(double) * cnt);
            lo = std::min(lo, off);
            hi = std::max(hi, off + cnt);
        };
        if (HM) {
            put(0, HM, nn);
            put(nn, bM, n);
            if (delta) put(nn + n, delta, n);
        }
        if (nullspace_projector) put(nn + 2 * (size_t)n, nullspace_projector, nn);
        if (cPrior) {
            const size_t o = 2 * nn + 2 * (size_t)n;
            put(o, cPrior, 4);
            put(o + 4, frame_prior, 8 * (size_t)F);
            put(o + 4 + 8 * F, frame_delta_prior, 8 * (size_t)F);
        }
        // (a gap between two present parts is copied along: it is never read on the device)
        if (hi > lo) EDS_CUDA(ctx, cudaMemcpyAsync(HMd + lo, hst + lo, sizeof(double) * (hi - lo), cudaMemcpyHostToDevice, s));
    }
    // accumulateAF_MT / accumulateLF_MT / accumulateSCF_MT's stitches (:790-800), results stay on the device
    {
        Stitch3Args q{};
        q.F = F; q.use_prior = cPrior ? 1 : 0;
        q.acc0 = w->acc[0]; q.acc1 = w->acc[1]; q.adHost = w->adHost; q.adTarget = w->adTarget;
        q.cPrior = pr; q.frame_prior = pr + 4; q.frame_delta_prior = pr + 4 + 8 * F; q.cDeltaF = w->cDeltaF;
        q.accD = w->accD; q.accE = w->accE; q.accEB = w->accEB; q.accHcc = w->accHcc; q.accbc = w->accbc;
        q.HA = HAd; q.bA = bAd; q.HL = HLd; q.bL = bLd; q.Hs = Hsd; q.bs = bsd;
        ba_stitch3_kernel<<<dim3(F * F, 3), STITCH_THREADS, 0, s>>>(q);
        ba_top_symmetrise2_kernel<<<dim3(F * F, 2), 64, 0, s>>>(F, cPrior ? 1 : 0, pr + 4, HAd, HLd);
    }
    char* pb = (char*)w->post_block;
    float* d_xAd = (float*)pb;
    float* d_cstep = (float*)(pb + align_up(32 * F2, 256));
    float* d_step = (float*)(pb + align_up(32 * F2, 256) + 256);
    ba_solve_kernel<<<1, SOLVE_THREADS, 0, s>>>(F, lambda, HAd, bAd, HLd, bLd, Hsd, bsd, HM ? HMd : nullptr, HM ? bMd : nullptr,
                                                (HM && delta) ? dld : nullptr, nullspace_projector ? prd : nullptr, w->adHost, w->adTarget, xd, d_xAd,
                                                d_cstep);
    // resubstituteF_MT(x) (:907-909): the point steps from the device-resident x
    ba_resubstitute_kernel<<<(w->P + 127) / 128, 128, 0, s>>>(w->F, w->P, w->res_begin, w->host_idx, w->target_idx, w->flags, w->JpJdF, w->bdSum,
                                                              w->Hcd[0], w->Hcd[1], w->HdiF, d_xAd, d_cstep, d_step);
    ctx->launches += 4;
    EDS_CUDA(ctx, cudaGetLastError());
    EDS_CUDA(ctx, cudaMemcpyAsync(x_out, xd, 8 * (size_t)n, cudaMemcpyDeviceToHost, s));
    if (point_step_out) EDS_CUDA(ctx, cudaMemcpyAsync(point_step_out, d_step, 4 * (size_t)w->P, cudaMemcpyDeviceToHost, s));
    EDS_CUDA(ctx, cudaStreamSynchronize(s));
    return EDSGPU_OK;
}

edsgpu_status edsgpu_ba_calc_l_energy(edsgpu_ba* w, const double* cPrior, const double* frame_prior, const double* frame_delta_prior,
                                      double* energy_out) {
    if (!w) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = w->ctx;
    EDS_REQUIRE(ctx, energy_out != nullptr, "ba_calc_l_energy: null output");
    DeviceGuard g(ctx->device);
    edsgpu_status st = ensure_post_block(w);
    if (st == EDSGPU_OK) st = edsgpu_ensure_pinned(ctx, 64);
    if (st != EDSGPU_OK) return st;
    const size_t F = w->F, F2 = F * F;
    const int n_part = (std::max(w->R, w->P) + LE_THREADS - 1) / LE_THREADS;
    char* pb = (char*)w->post_block;
    double* d_partial = (double*)(pb + align_up(32 * F2, 256) + 256 + align_up(4 * (size_t)w->P, 256));
    ba_lenergy_kernel<<<n_part, LE_THREADS, 0, ctx->stream>>>(w->F, w->P, w->R, w->recs, w->host_idx, w->target_idx, w->point_of_res, w->flags,
                                                            w->res_toZero, w->deltaF, w->priorF, w->adHTdeltaF, w->cDeltaF, d_partial);
    ctx->launches++;
    char* pin_dev = nullptr;  // the pinned block as the device sees it
    EDS_CUDA(ctx, cudaHostGetDevicePointer((void**)&pin_dev, ctx->pinned, 0));
    ba_sum_partials_kernel<<<1, 32, 0, ctx->stream>>>(d_partial, n_part, (double*)pin_dev, w->cDeltaF, (float*)(pin_dev + 16));
    ctx->launches++;
    EDS_CUDA(ctx, cudaGetLastError());
    float cd[4];
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double E = *(double*)ctx->pinned;
    memcpy(cd, (char*)ctx->pinned + 16, 16);
    // EnergyFunctional.cpp:403-409: frame and calibration prior terms
    if (frame_prior && frame_delta_prior)
        for (size_t i = 0; i < 8 * F; ++i) E += frame_delta_prior[i] * frame_prior[i] * frame_delta_prior[i];
    if (cPrior)
        for (int k = 0; k < 4; ++k) E += (double)cd[k] * (double)(float)cPrior[k] * (double)cd[k];
    *energy_out = E;
    return EDSGPU_OK;
}

}  // extern "C"
