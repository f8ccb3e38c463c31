// CPU oracle for the DSO coarse tracker of uzh-rpg/slam-eds.
//
// TEST INFRASTRUCTURE ONLY: nothing under slam-eds_b200/ may include, link or call this file.
// PARITY UNPINNED: the reference ships no tests or golden vectors for this path and cannot be built here;
// this restates CoarseTracker::calcRes / calcGSSSE / trackNewestCoarse (src/tracking/CoarseTracker.cpp:287-701),
// getInterpolatedElement33 (src/utils/globalFuncs.h:78-92), Accumulator9::updateSSE_eighted
// (src/bundles/MatrixAccumulators.h:1091-1150) and, un-vendored, Sophus' SE3::exp and Eigen's LDLT solve.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

extern "C" {

// CoarseTracker::calcRes + calcGSSSE, src/tracking/CoarseTracker.cpp:287-498: direct image
// alignment of a new frame against the reference point cloud at one pyramid level -- residual statistics and the
// 8x8 Gauss-Newton system (6 pose + 2 affine brightness).  float arithmetic in the reference's operation order
// (built with -ffp-contract=off); the sums E, H, b are taken in double where the reference uses float accumulators.
//   dINew [hl*wl*3] Vec3f {I,dx,dy};  Ki [9] column-major float (K[lvl].inverse());  R [9] row-major, t [3]: refToNew
//   affLL {a, b} = AffLight::fromToVecExposure(...), b0 = lastRef_aff_g2l.b
// out: rs[6] as calcRes returns it, H [64] row-major, b [8], counts {numTermsInE, numTermsInWarped (unpadded), numSaturated}
void eds_oracle_coarse_calc_res_gs(int lvl, int wl, int hl, const float* dINew, float fxl, float fyl, float cxl, float cyl, const float* Ki,
                                   const double* R, const double* t, const float* affLL, float b0, float cutoffTH, int n, const float* pc_u,
                                   const float* pc_v, const float* pc_idepth, const float* pc_color, double* rs, double* H_out, double* b_out,
                                   int64_t* counts) {
    const float setting_huberTH = 9.0f;  // settings.cpp:127
    float RKi[3][3], tf[3];
    for (int i = 0; i < 3; ++i) {
        tf[i] = (float)t[i];
        for (int j = 0; j < 3; ++j)
            RKi[i][j] = ((float)R[3 * i + 0] * Ki[3 * j + 0] + (float)R[3 * i + 1] * Ki[3 * j + 1]) + (float)R[3 * i + 2] * Ki[3 * j + 2];
    }
    auto KiM = [&](int i, int j) { return Ki[3 * j + i]; };
    double E = 0;
    int64_t numTermsInE = 0, numTermsInWarped = 0, numSaturated = 0;
    float sumSquaredShiftT = 0, sumSquaredShiftRT = 0, sumSquaredShiftNum = 0;
    const float maxEnergy = 2 * setting_huberTH * cutoffTH - setting_huberTH * setting_huberTH;  // :372
    double acc[9][9];
    for (auto& row : acc) for (double& v : row) v = 0.0;
    for (int i = 0; i < n; ++i) {
        const float id = pc_idepth[i], x = pc_u[i], y = pc_v[i];
        float pt[3];
        for (int k = 0; k < 3; ++k) pt[k] = ((RKi[k][0] * x + RKi[k][1] * y) + RKi[k][2] * 1.0f) + tf[k] * id;
        const float u = pt[0] / pt[2], v = pt[1] / pt[2];
        const float Ku = fxl * u + cxl, Kv = fyl * v + cyl;
        const float new_idepth = id / pt[2];
        if (lvl == 0 && i % 32 == 0) {  // :403-434: mean optical flow, translation only and rotation + translation
            float p1[3], p2[3], p3[3];
            for (int k = 0; k < 3; ++k) {
                const float kp = (KiM(k, 0) * x + KiM(k, 1) * y) + KiM(k, 2) * 1.0f;
                p1[k] = kp + tf[k] * id;
                p2[k] = kp - tf[k] * id;
                p3[k] = ((RKi[k][0] * x + RKi[k][1] * y) + RKi[k][2] * 1.0f) - tf[k] * id;
            }
            const float KuT = fxl * (p1[0] / p1[2]) + cxl, KvT = fyl * (p1[1] / p1[2]) + cyl;
            const float KuT2 = fxl * (p2[0] / p2[2]) + cxl, KvT2 = fyl * (p2[1] / p2[2]) + cyl;
            const float Ku3 = fxl * (p3[0] / p3[2]) + cxl, Kv3 = fyl * (p3[1] / p3[2]) + cyl;
            sumSquaredShiftT += (KuT - x) * (KuT - x) + (KvT - y) * (KvT - y);
            sumSquaredShiftT += (KuT2 - x) * (KuT2 - x) + (KvT2 - y) * (KvT2 - y);
            sumSquaredShiftRT += (Ku - x) * (Ku - x) + (Kv - y) * (Kv - y);
            sumSquaredShiftRT += (Ku3 - x) * (Ku3 - x) + (Kv3 - y) * (Kv3 - y);
            sumSquaredShiftNum += 2;
        }
        if (!(Ku > 2 && Kv > 2 && Ku < wl - 3 && Kv < hl - 3 && new_idepth > 0)) continue;
        const float refColor = pc_color[i];
        const int ix = (int)Ku, iy = (int)Kv;  // getInterpolatedElement33, globalFuncs.h:78-92
        const float dx = Ku - ix, dy = Kv - iy, dxdy = dx * dy;
        const float* bp = dINew + (size_t)3 * (ix + iy * wl);
        float hit[3];
        for (int c = 0; c < 3; ++c)
            hit[c] = ((dxdy * bp[3 * (1 + wl) + c] + (dy - dxdy) * bp[3 * wl + c]) + (dx - dxdy) * bp[3 + c]) + (1 - dx - dy + dxdy) * bp[c];
        if (!std::isfinite(hit[0])) continue;
        const float residual = hit[0] - (float)(affLL[0] * refColor + affLL[1]);
        const float hw = fabsf(residual) < setting_huberTH ? 1 : setting_huberTH / fabsf(residual);
        if (fabsf(residual) > cutoffTH) {
            E += maxEnergy;
            numTermsInE++;
            numSaturated++;
            continue;
        }
        E += hw * residual * residual * (2 - hw);
        numTermsInE++;
        numTermsInWarped++;
        // calcGSSSE :303-330 on the warped sample
        const float ddx = hit[1] * fxl, ddy = hit[2] * fyl;
        float J[9];
        J[0] = new_idepth * ddx;
        J[1] = new_idepth * ddy;
        J[2] = 0 - new_idepth * (u * ddx + v * ddy);
        J[3] = 0 - ((u * v) * ddx + ddy * (1 + v * v));
        J[4] = (u * v) * ddy + ddx * (1 + u * u);
        J[5] = u * ddy - v * ddx;
        J[6] = affLL[0] * (b0 - refColor);
        J[7] = -1;
        J[8] = residual;
        for (int a = 0; a < 9; ++a) {
            const float Jw = J[a] * hw;  // MatrixAccumulators.h:1091-1150
            for (int c = a; c < 9; ++c) acc[a][c] += (double)(Jw * J[c]);
        }
    }
    const int64_t npad = (numTermsInWarped + 3) / 4 * 4;  // :466-478: the SSE buffers are padded with zeros
    const float inv_n = 1.0f / (float)npad;
    for (int a = 0; a < 8; ++a) {
        for (int c = 0; c < 8; ++c) H_out[8 * a + c] = (double)(float)acc[a < c ? a : c][a < c ? c : a] * inv_n;
        b_out[a] = (double)(float)acc[a][8] * inv_n;
    }
    // SCALE_XI_ROT = SCALE_XI_TRANS = 1, SCALE_A = 10, SCALE_B = 1000 (HessianBlocks.h:59-65), :333-344
    const double sc[8] = {1, 1, 1, 1, 1, 1, 10.0f, 1000.0f};
    for (int a = 0; a < 8; ++a) {
        for (int c = 0; c < 8; ++c) H_out[8 * a + c] *= (sc[a] * sc[c]);
        b_out[a] *= sc[a];
    }
    rs[0] = E;
    rs[1] = (double)numTermsInE;
    rs[2] = sumSquaredShiftT / (sumSquaredShiftNum + 0.1);
    rs[3] = 0;
    rs[4] = sumSquaredShiftRT / (sumSquaredShiftNum + 0.1);
    rs[5] = numSaturated / (float)numTermsInE;
    counts[0] = numTermsInE; counts[1] = numTermsInWarped; counts[2] = numSaturated;
}

// ---- CoarseTracker::trackNewestCoarse (CoarseTracker.cpp:520-701): the coarse-to-fine Gauss-Newton loop around
// calcRes / calcGSSSE.  Sophus (un-vendored, version unpinned) is restated from its published formulas:
// SE3::exp(upsilon, omega) with the quaternion SO3::exp and V = I + (1-cos t)/t^2 W + (t - sin t)/t^3 W^2;
// Eigen's pivoted LDLT is restated as an un-pivoted LDL^T (the damped 8x8 systems are positive definite).
namespace coarse_loop {

void so3_exp(const double* w, double* R) {
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = std::sqrt(th2);
    double imag, real;
    if (th < 1e-10) {
        const double th4 = th2 * th2;
        imag = 0.5 - th2 / 48.0 + th4 / 3840.0;
        real = 1.0 - th2 / 8.0 + th4 / 384.0;
    } else {
        imag = std::sin(0.5 * th) / th;
        real = std::cos(0.5 * th);
    }
    const double x = imag * w[0], y = imag * w[1], z = imag * w[2], q = real;
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * q); R[2] = 2 * (x * z + y * q);
    R[3] = 2 * (x * y + z * q); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * q);
    R[6] = 2 * (x * z - y * q); R[7] = 2 * (y * z + x * q); R[8] = 1 - 2 * (x * x + y * y);
}

// new = exp(inc) * (R, t)
void se3_left_update(const double* inc6, double* R, double* t) {
    const double* u = inc6;
    const double* w = inc6 + 3;
    double Re[9];
    so3_exp(w, Re);
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = std::sqrt(th2);
    const double W[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double V[9];
    if (th < 1e-10) {
        for (int i = 0; i < 9; ++i) V[i] = Re[i];
    } else {
        const double A = (1.0 - std::cos(th)) / th2, B = (th - std::sin(th)) / (th2 * th);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                double w2 = 0;
                for (int k = 0; k < 3; ++k) w2 += W[3 * r + k] * W[3 * k + c];
                V[3 * r + c] = (r == c ? 1.0 : 0.0) + A * W[3 * r + c] + B * w2;
            }
    }
    double te[3], Rn[9], tn[3];
    for (int r = 0; r < 3; ++r) te[r] = V[3 * r] * u[0] + V[3 * r + 1] * u[1] + V[3 * r + 2] * u[2];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) Rn[3 * r + c] = Re[3 * r] * R[c] + Re[3 * r + 1] * R[3 + c] + Re[3 * r + 2] * R[6 + c];
        tn[r] = Re[3 * r] * t[0] + Re[3 * r + 1] * t[1] + Re[3 * r + 2] * t[2] + te[r];
    }
    for (int i = 0; i < 9; ++i) R[i] = Rn[i];
    for (int i = 0; i < 3; ++i) t[i] = tn[i];
}

// x = A^-1 rhs for a symmetric positive definite 8x8 (LDL^T)
void solve8(const double* A, const double* rhs, double* x) {
    double L[64] = {0}, D[8];
    for (int j = 0; j < 8; ++j) {
        double d = A[8 * j + j];
        for (int k = 0; k < j; ++k) d -= L[8 * j + k] * L[8 * j + k] * D[k];
        D[j] = d;
        L[8 * j + j] = 1.0;
        for (int i = j + 1; i < 8; ++i) {
            double v = A[8 * i + j];
            for (int k = 0; k < j; ++k) v -= L[8 * i + k] * L[8 * j + k] * D[k];
            L[8 * i + j] = v / d;
        }
    }
    double y[8];
    for (int i = 0; i < 8; ++i) { double v = rhs[i]; for (int k = 0; k < i; ++k) v -= L[8 * i + k] * y[k]; y[i] = v; }
    for (int i = 0; i < 8; ++i) y[i] /= D[i];
    for (int i = 7; i >= 0; --i) { double v = y[i]; for (int k = i + 1; k < 8; ++k) v -= L[8 * k + i] * x[k]; x[i] = v; }
}

}  // namespace coarse_loop

// levels: arrays of per-level inputs (index = pyramid level).  R/t (refToNew) and aff (a, b of the new frame) are in-out.
// Returns 1 when tracking succeeded (trackNewestCoarse's return value), 0 otherwise.
int eds_oracle_coarse_track(int coarsest_lvl, const int* wl, const int* hl, const float* const* dINew, const float* fxl, const float* fyl,
                            const float* cxl, const float* cyl, const float* const* Ki, const int* n, const float* const* pc_u,
                            const float* const* pc_v, const float* const* pc_idepth, const float* const* pc_color, double* R, double* t,
                            double* aff, const double* ref_aff, float ref_exposure, float new_exposure, const double* min_res_for_abort,
                            double* last_residuals, double* last_flow, int* evaluations) {
    using namespace coarse_loop;
    const float setting_coarseCutoffTH = 20.0f;  // settings.cpp:138
    const int maxIterations[5] = {10, 20, 100, 100, 100};
    const float lambdaExtrapolationLimit = 0.001f;
    auto aff_ll = [&](const double* g2T, float* out) {  // AffLight::fromToVecExposure, NumType.h:175-187
        float eF = ref_exposure, eT = new_exposure;
        if (eF == 0 || eT == 0) eF = eT = 1;
        const double a = std::exp(g2T[0] - ref_aff[0]) * eT / eF;
        out[0] = (float)a;
        out[1] = (float)(g2T[1] - a * ref_aff[1]);
    };
    auto eval = [&](int lvl, const double* Rc, const double* tc, const double* affc, float cutoff, double* rs, double* H, double* b) {
        float ll[2];
        aff_ll(affc, ll);
        int64_t counts[3];
        eds_oracle_coarse_calc_res_gs(lvl, wl[lvl], hl[lvl], dINew[lvl], fxl[lvl], fyl[lvl], cxl[lvl], cyl[lvl], Ki[lvl], Rc, tc, ll,
                                      (float)ref_aff[1], cutoff, n[lvl], pc_u[lvl], pc_v[lvl], pc_idepth[lvl], pc_color[lvl], rs, H, b, counts);
        ++*evaluations;
    };
    for (int i = 0; i < 5; ++i) last_residuals[i] = NAN;
    for (int i = 0; i < 3; ++i) last_flow[i] = 1000;
    *evaluations = 0;
    double Rc[9], tc[3], affc[2] = {aff[0], aff[1]};
    for (int i = 0; i < 9; ++i) Rc[i] = R[i];
    for (int i = 0; i < 3; ++i) tc[i] = t[i];
    bool haveRepeated = false;
    for (int lvl = coarsest_lvl; lvl >= 0; lvl--) {
        double H[64], b[8], resOld[6];
        float levelCutoffRepeat = 1;
        eval(lvl, Rc, tc, affc, setting_coarseCutoffTH * levelCutoffRepeat, resOld, H, b);
        while (resOld[5] > 0.6 && levelCutoffRepeat < 50) {
            levelCutoffRepeat *= 2;
            eval(lvl, Rc, tc, affc, setting_coarseCutoffTH * levelCutoffRepeat, resOld, H, b);
        }
        float lambda = 0.01f;
        for (int iteration = 0; iteration < maxIterations[lvl]; iteration++) {
            double Hl[64], nb[8], inc[8];
            for (int i = 0; i < 64; ++i) Hl[i] = H[i];
            for (int i = 0; i < 8; ++i) { Hl[9 * i] *= (1 + lambda); nb[i] = -b[i]; }
            solve8(Hl, nb, inc);
            float extrapFac = 1;
            if (lambda < lambdaExtrapolationLimit) extrapFac = std::sqrt(std::sqrt(lambdaExtrapolationLimit / lambda));
            for (int i = 0; i < 8; ++i) inc[i] *= extrapFac;
            double incScaled[8];
            for (int i = 0; i < 8; ++i) incScaled[i] = inc[i];
            incScaled[6] *= 10.0f;    // SCALE_A
            incScaled[7] *= 1000.0f;  // SCALE_B
            double sum = 0;
            for (int i = 0; i < 8; ++i) sum += incScaled[i];
            if (!std::isfinite(sum)) for (int i = 0; i < 8; ++i) incScaled[i] = 0;
            double Rn[9], tn[3], affn[2] = {affc[0] + incScaled[6], affc[1] + incScaled[7]};
            for (int i = 0; i < 9; ++i) Rn[i] = Rc[i];
            for (int i = 0; i < 3; ++i) tn[i] = tc[i];
            se3_left_update(incScaled, Rn, tn);
            double resNew[6], Hn[64], bn[8];
            eval(lvl, Rn, tn, affn, setting_coarseCutoffTH * levelCutoffRepeat, resNew, Hn, bn);
            const bool accept = (resNew[0] / resNew[1]) < (resOld[0] / resOld[1]);
            if (accept) {
                for (int i = 0; i < 64; ++i) H[i] = Hn[i];
                for (int i = 0; i < 8; ++i) b[i] = bn[i];
                for (int i = 0; i < 6; ++i) resOld[i] = resNew[i];
                affc[0] = affn[0]; affc[1] = affn[1];
                for (int i = 0; i < 9; ++i) Rc[i] = Rn[i];
                for (int i = 0; i < 3; ++i) tc[i] = tn[i];
                lambda *= 0.5f;
            } else {
                lambda *= 4;
                if (lambda < lambdaExtrapolationLimit) lambda = lambdaExtrapolationLimit;
            }
            double norm2 = 0;
            for (int i = 0; i < 8; ++i) norm2 += inc[i] * inc[i];
            if (!(std::sqrt(norm2) > 1e-3)) break;
        }
        last_residuals[lvl] = std::sqrt((float)(resOld[0] / resOld[1]));
        for (int i = 0; i < 3; ++i) last_flow[i] = resOld[2 + i];
        if (last_residuals[lvl] > 1.5 * min_res_for_abort[lvl]) return 0;
        if (levelCutoffRepeat > 1 && !haveRepeated) { lvl++; haveRepeated = true; }
    }
    for (int i = 0; i < 9; ++i) R[i] = Rc[i];
    for (int i = 0; i < 3; ++i) t[i] = tc[i];
    aff[0] = affc[0]; aff[1] = affc[1];
    if (std::fabs((float)aff[0]) > 1.2f || std::fabs((float)aff[1]) > 200.f) return 0;  // :683-685 (both modes != 0)
    return 1;
}


// CoarseTracker::makeCoarseDepthL0, src/tracking/CoarseTracker.cpp:127-283: the reference frame's inverse-depth map from the
// active points projected into it (weighted by their Hessian inverse), summed down the pyramid, dilated by one pixel where a
// pixel has no point (diagonal neighbours on levels 0-1, axis neighbours below), normalised, and compacted in scan-line order
// into the per-level point cloud pc_u / pc_v / pc_idepth / pc_color that calcRes reads.  float arithmetic in the reference's
// order.  One deviation: the dilation loops of the reference index one element before / past the image at its first and last
// pixel (i - 1 - wl = -1, i + 1 + wl = w*h); such reads count as "no point" here.
//   cu, cv, cid: PointFrameResidual::centerProjectedTo of the n points whose last residual is IN; HdiF: EFPoint::HdiF
//   dIRef[lvl]: lastRef->dIp[lvl], h*w Vec3f;  out per level: idepth, weightSums (h*w), pc_* (capacity h*w), pc_n
void eds_oracle_make_coarse_depth_l0(int L, const int* w, const int* h, int n, const float* cu, const float* cv, const float* cid, const float* HdiF,
                                     const float* const* dIRef, float* const* idepth, float* const* weightSums, float* const* pc_u,
                                     float* const* pc_v, float* const* pc_idepth, float* const* pc_color, int* pc_n) {
    memset(idepth[0], 0, sizeof(float) * w[0] * h[0]);
    memset(weightSums[0], 0, sizeof(float) * w[0] * h[0]);
    for (int p = 0; p < n; ++p) {
        const int u = (int)(cu[p] + 0.5f), v = (int)(cv[p] + 0.5f);
        const float weight = sqrtf((float)(1e-3 / ((double)HdiF[p] + 1e-12)));
        idepth[0][u + w[0] * v] += cid[p] * weight;
        weightSums[0][u + w[0] * v] += weight;
    }
    for (int lvl = 1; lvl < L; ++lvl) {
        const int wl = w[lvl], hl = h[lvl], wlm1 = w[lvl - 1];
        for (int y = 0; y < hl; ++y)
            for (int x = 0; x < wl; ++x) {
                const int b = 2 * x + 2 * y * wlm1;
                idepth[lvl][x + y * wl] = idepth[lvl - 1][b] + idepth[lvl - 1][b + 1] + idepth[lvl - 1][b + wlm1] + idepth[lvl - 1][b + wlm1 + 1];
                weightSums[lvl][x + y * wl] = weightSums[lvl - 1][b] + weightSums[lvl - 1][b + 1] + weightSums[lvl - 1][b + wlm1] + weightSums[lvl - 1][b + wlm1 + 1];
            }
    }
    for (int lvl = 0; lvl < L; ++lvl) {
        const int wl = w[lvl], total = w[lvl] * h[lvl], wh = total - wl;
        float* ws = weightSums[lvl];
        float* id = idepth[lvl];
        float* bak = new float[total];
        memcpy(bak, ws, sizeof(float) * total);
        const int off[2][4] = {{1 + wl, -1 - wl, wl - 1, -wl + 1}, {1, -1, wl, -wl}};
        const int* o = off[lvl < 2 ? 0 : 1];
        for (int i = wl; i < wh; ++i)
            if (bak[i] <= 0) {
                float sum = 0, num = 0, numn = 0;
                for (int k = 0; k < 4; ++k) {
                    const int j = i + o[k];
                    if (j >= 0 && j < total && bak[j] > 0) { sum += id[j]; num += bak[j]; numn++; }
                }
                if (numn > 0) { id[i] = sum / numn; ws[i] = num / numn; }
            }
        delete[] bak;
    }
    for (int lvl = 0; lvl < L; ++lvl) {
        const int wl = w[lvl], hl = h[lvl];
        float* ws = weightSums[lvl];
        float* id = idepth[lvl];
        int cnt = 0;
        for (int y = 2; y < hl - 2; ++y)
            for (int x = 2; x < wl - 2; ++x) {
                const int i = x + y * wl;
                if (ws[i] > 0) {
                    id[i] /= ws[i];
                    pc_u[lvl][cnt] = (float)x;
                    pc_v[lvl][cnt] = (float)y;
                    pc_idepth[lvl][cnt] = id[i];
                    pc_color[lvl][cnt] = dIRef[lvl][3 * i];
                    if (!std::isfinite(pc_color[lvl][cnt]) || !(id[i] > 0)) {
                        id[i] = -1;
                        continue;
                    }
                    cnt++;
                } else {
                    id[i] = -1;
                }
                ws[i] = 1;
            }
        pc_n[lvl] = cnt;
    }
}

}  // extern "C"
