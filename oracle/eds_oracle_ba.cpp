// ============================================================================
// TEST INFRASTRUCTURE ONLY.  CPU restatement ("oracle") of the DSO-derived
// windowed-BA Hessian accumulators of EDS (src/bundles).  See the header of
// eds_oracle_tracking.cpp for the rules that apply to this directory.
//
// PARITY UNPINNED (no reference tests/golden vectors exist; the reference
// cannot be compiled here).  Pinned instead by: dense J^T J of the stacked
// 8x13 rows (tests/test_oracle_ba.py), the two in-repo statements of each
// stitch, and the re-linearisation identity between
// AccumulatedTopHessian.cpp:84-98 and EnergyFunctionalStructs.cpp:101-110.
//
// The reference accumulates in float with a 3-level "shift-up"
// (MatrixAccumulators.h:937-971) and its chunk->thread mapping is dynamic, so
// its float sums differ run to run; this oracle reads the same float inputs,
// forms each per-residual term and accumulates in double.
// ============================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

// Record layout: dso::RawResidualJacobian (src/bundles/RawResidualJacobian.h:32-61)
// as Eigen lays it out with 16-byte alignment: 76 floats = 304 bytes.
constexpr int REC = 76;
constexpr int O_RES = 0, O_JPDXI0 = 8, O_JPDXI1 = 14, O_JPDC0 = 20, O_JPDC1 = 24, O_JPDD = 28,
              O_JIDX0 = 32, O_JIDX1 = 40, O_JAB0 = 48, O_JAB1 = 56, O_JIDX2 = 64, O_JABJIDX = 68, O_JAB2 = 72;
// Mat22f is column-major: [ (0,0) (1,0) (0,1) (1,1) ]
inline double m22(const float* m, int r, int c) { return (double)m[r + 2 * c]; }

constexpr int PATTERN = 8;  // patternNum, src/utils/settings.h:215
constexpr int CPARS = 4;    // src/utils/NumType.h:51

struct TopAcc {
    std::vector<double> H;  // F*F * 169, row-major 13x13, order [C(4) | xi(6) | a b | r]
    std::vector<long> num;
    long nres = 0;
    void init(int F) { H.assign((size_t)F * F * 169, 0.0); num.assign((size_t)F * F, 0); nres = 0; }
};

// AccumulatedTopHessianSSE::addPoint<mode>, src/bundles/AccumulatedTopHessian.cpp:39-159
void top_add_point(int mode, int F, int p, const float* recs, const int32_t* host_idx, const int32_t* target_idx,
                   const int32_t* res_begin, const uint8_t* flags, const float* res_toZero, const float* deltaF,
                   const float* adHTdeltaF, const float* cDeltaF, TopAcc& A, float* Hdd_out, float* bd_out, float* Hcd_out) {
    double dd = deltaF ? (double)deltaF[p] : 0.0;
    double bd_acc = 0, Hdd_acc = 0, Hcd_acc[4] = {0, 0, 0, 0};
    for (int r = res_begin[p]; r < res_begin[p + 1]; ++r) {
        bool active = flags[r] & 1, lin = flags[r] & 2;
        if (mode == 0 && (lin || !active)) continue;   // :55-58
        if (mode == 1 && (!lin || !active)) continue;  // :59-62
        if (mode == 2 && !active) continue;            // :63-67
        const float* J = recs + (size_t)REC * r;
        int ht = host_idx[r] + target_idx[r] * F;      // :71
        double res[PATTERN] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (mode == 0) for (int i = 0; i < PATTERN; ++i) res[i] = J[O_RES + i];
        if (mode == 2) for (int i = 0; i < PATTERN; ++i) res[i] = res_toZero[(size_t)8 * r + i];
        if (mode == 1) {
            // :81-99  rtz + JI*Jp*delta + Jab*delta_ab
            const float* dp = adHTdeltaF + (size_t)8 * ht;
            double jx = 0, jy = 0;
            for (int k = 0; k < 6; ++k) { jx += (double)J[O_JPDXI0 + k] * dp[k]; jy += (double)J[O_JPDXI1 + k] * dp[k]; }
            for (int k = 0; k < 4; ++k) { jx += (double)J[O_JPDC0 + k] * cDeltaF[k]; jy += (double)J[O_JPDC1 + k] * cDeltaF[k]; }
            jx += (double)J[O_JPDD] * dd;
            jy += (double)J[O_JPDD + 1] * dd;
            for (int i = 0; i < PATTERN; ++i)
                res[i] = (double)res_toZero[(size_t)8 * r + i] + (double)J[O_JIDX0 + i] * jx + (double)J[O_JIDX1 + i] * jy +
                         (double)J[O_JAB0 + i] * dp[6] + (double)J[O_JAB1 + i] * dp[7];
        }
        // :102-112
        double JI_r[2] = {0, 0}, Jab_r[2] = {0, 0}, rr = 0;
        for (int i = 0; i < PATTERN; ++i) {
            JI_r[0] += res[i] * J[O_JIDX0 + i];
            JI_r[1] += res[i] * J[O_JIDX1 + i];
            Jab_r[0] += res[i] * J[O_JAB0 + i];
            Jab_r[1] += res[i] * J[O_JAB1 + i];
            rr += res[i] * res[i];
        }
        // AccumulatorApprox::update / updateTopRight / updateBotRight
        // (src/bundles/MatrixAccumulators.h:754-915), x = [Jpdc[0] Jpdxi[0]], y = [Jpdc[1] Jpdxi[1]]
        double x[10], y[10];
        for (int k = 0; k < 4; ++k) { x[k] = J[O_JPDC0 + k]; y[k] = J[O_JPDC1 + k]; }
        for (int k = 0; k < 6; ++k) { x[4 + k] = J[O_JPDXI0 + k]; y[4 + k] = J[O_JPDXI1 + k]; }
        const double a = m22(J + O_JIDX2, 0, 0), bq = m22(J + O_JIDX2, 0, 1), c = m22(J + O_JIDX2, 1, 1);
        double* Hm = &A.H[(size_t)169 * ht];
        for (int i = 0; i < 10; ++i)
            for (int j = i; j < 10; ++j) {
                double v = a * x[i] * x[j] + c * y[i] * y[j] + bq * (x[i] * y[j] + y[i] * x[j]);
                Hm[13 * i + j] += v;
                if (j != i) Hm[13 * j + i] += v;
            }
        const double TR[3][2] = {{m22(J + O_JABJIDX, 0, 0), m22(J + O_JABJIDX, 0, 1)},
                                 {m22(J + O_JABJIDX, 1, 0), m22(J + O_JABJIDX, 1, 1)},
                                 {JI_r[0], JI_r[1]}};
        for (int i = 0; i < 10; ++i)
            for (int k = 0; k < 3; ++k) {
                double v = x[i] * TR[k][0] + y[i] * TR[k][1];
                Hm[13 * i + 10 + k] += v;
                Hm[13 * (10 + k) + i] += v;
            }
        const double br[6] = {m22(J + O_JAB2, 0, 0), m22(J + O_JAB2, 0, 1), Jab_r[0], m22(J + O_JAB2, 1, 1), Jab_r[1], rr};
        Hm[13 * 10 + 10] += br[0];
        Hm[13 * 10 + 11] += br[1]; Hm[13 * 11 + 10] += br[1];
        Hm[13 * 10 + 12] += br[2]; Hm[13 * 12 + 10] += br[2];
        Hm[13 * 11 + 11] += br[3];
        Hm[13 * 11 + 12] += br[4]; Hm[13 * 12 + 11] += br[4];
        Hm[13 * 12 + 12] += br[5];
        A.num[ht]++;
        // :132-135
        const double jd0 = J[O_JPDD], jd1 = J[O_JPDD + 1];
        const double Ji2Jd[2] = {a * jd0 + bq * jd1, bq * jd0 + c * jd1};
        bd_acc += JI_r[0] * jd0 + JI_r[1] * jd1;
        Hdd_acc += Ji2Jd[0] * jd0 + Ji2Jd[1] * jd1;
        for (int k = 0; k < 4; ++k) Hcd_acc[k] += (double)J[O_JPDC0 + k] * Ji2Jd[0] + (double)J[O_JPDC1 + k] * Ji2Jd[1];
        A.nres++;
    }
    // :140-157 (mode 2 additionally zeroes the *_accAF fields; the caller owns those)
    Hdd_out[p] = (float)Hdd_acc;
    bd_out[p] = (float)bd_acc;
    for (int k = 0; k < 4; ++k) Hcd_out[4 * p + k] = (float)Hcd_acc[k];
}

inline double& Hat(double* H, int n, int r, int c) { return H[(size_t)c * n + r]; }  // column-major like Eigen MatXX

// y(8x?) helpers on 8x8 column-major adjoint blocks (Mat88)
inline double ad(const double* A, int r, int c) { return A[c * 8 + r]; }

struct ScAcc {
    std::vector<double> D, E, EB;  // F^3*64 (row-major 8x8), F^2*32 (row-major 8x4), F^2*8
    double Hcc[16], bc[4];
    void init(int F) {
        D.assign((size_t)F * F * F * 64, 0.0); E.assign((size_t)F * F * 32, 0.0); EB.assign((size_t)F * F * 8, 0.0);
        std::fill(Hcc, Hcc + 16, 0.0); std::fill(bc, bc + 4, 0.0);
    }
};

// AccumulatedSCHessianSSE::addPoint, src/bundles/AccumulatedSCHessian.cpp:34-77
void sc_add_point(int F, int p, const int32_t* host_idx, const int32_t* target_idx, const int32_t* res_begin, const uint8_t* flags,
                  const float* JpJdF, const float* Hdd_A, const float* Hdd_L, const float* bd_A, const float* bd_L,
                  const float* Hcd_A, const float* Hcd_L, const float* priorF, const float* deltaF, bool shiftPriorToZero,
                  ScAcc& S, float* HdiF_out, float* bdSum_out) {
    int ngood = 0;
    for (int r = res_begin[p]; r < res_begin[p + 1]; ++r) if (flags[r] & 1) ngood++;
    if (ngood == 0) { HdiF_out[p] = 0; bdSum_out[p] = 0; return; }  // :38-45
    float Hf = Hdd_A[p] + Hdd_L[p] + priorF[p];  // float arithmetic like the reference (:47)
    if (Hf < 1e-10f) Hf = 1e-10f;
    float HdiF = (float)(1.0 / (double)Hf);  // reference: p->HdiF = 1.0 / H (double divide, stored to float)
    float bdSum = bd_A[p] + bd_L[p];
    if (shiftPriorToZero) bdSum += priorF[p] * deltaF[p];
    HdiF_out[p] = HdiF;
    bdSum_out[p] = bdSum;
    double Hcd[4];
    for (int k = 0; k < 4; ++k) Hcd[k] = (double)(float)(Hcd_A[4 * p + k] + Hcd_L[4 * p + k]);
    const double w = HdiF, wb = (double)HdiF * (double)bdSum;
    for (int i = 0; i < 4; ++i) { for (int j = 0; j < 4; ++j) S.Hcc[4 * i + j] += w * Hcd[i] * Hcd[j]; S.bc[i] += wb * Hcd[i]; }
    const int F2 = F * F;
    for (int r1 = res_begin[p]; r1 < res_begin[p + 1]; ++r1) {
        if (!(flags[r1] & 1)) continue;
        const int r1ht = host_idx[r1] + target_idx[r1] * F;
        const float* j1 = JpJdF + (size_t)8 * r1;
        for (int r2 = res_begin[p]; r2 < res_begin[p + 1]; ++r2) {
            if (!(flags[r2] & 1)) continue;
            const float* j2 = JpJdF + (size_t)8 * r2;
            double* D = &S.D[(size_t)64 * (r1ht + target_idx[r2] * F2)];
            for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) D[8 * i + j] += w * (double)j1[i] * (double)j2[j];
        }
        double* E = &S.E[(size_t)32 * r1ht];
        for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) E[4 * i + j] += w * (double)j1[i] * Hcd[j];
        double* EB = &S.EB[(size_t)8 * r1ht];
        for (int i = 0; i < 8; ++i) EB[i] += wb * (double)j1[i];
    }
}

template <typename Fn>
void parallel_chunks(int P, int threads, Fn fn) {
    // IndexThreadReduce hands out chunks of 50 points (EnergyFunctional.cpp:203);
    // static round-robin here so the result is reproducible.
    const int chunk = 50;
    int nchunks = (P + chunk - 1) / chunk;
    if (threads <= 1) { for (int c = 0; c < nchunks; ++c) fn(0, c * chunk, std::min(P, (c + 1) * chunk)); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t)
        th.emplace_back([=]() { for (int c = t; c < nchunks; c += threads) fn(t, c * chunk, std::min(P, (c + 1) * chunk)); });
    for (auto& t : th) t.join();
}

}  // namespace

extern "C" {

// EFResidual::takeDataF, src/bundles/EnergyFunctionalStructs.cpp:38-48 (float arithmetic).
void eds_oracle_ba_jpjd(int R, const float* recs, float* JpJdF) {
    for (int r = 0; r < R; ++r) {
        const float* J = recs + (size_t)REC * r;
        const float* M = J + O_JIDX2;
        float v0 = M[0] * J[O_JPDD] + M[2] * J[O_JPDD + 1];
        float v1 = M[1] * J[O_JPDD] + M[3] * J[O_JPDD + 1];
        for (int i = 0; i < 6; ++i) JpJdF[(size_t)8 * r + i] = J[O_JPDXI0 + i] * v0 + J[O_JPDXI1 + i] * v1;
        const float* N = J + O_JABJIDX;
        JpJdF[(size_t)8 * r + 6] = N[0] * J[O_JPDD] + N[2] * J[O_JPDD + 1];
        JpJdF[(size_t)8 * r + 7] = N[1] * J[O_JPDD] + N[3] * J[O_JPDD + 1];
    }
}

// EFResidual::fixLinearizationF, EnergyFunctionalStructs.cpp:87-113: res_toZero = resF - J*delta.
void eds_oracle_ba_fix_linearization(int F, int P, int R, const float* recs, const int32_t* host_idx, const int32_t* target_idx,
                                     const int32_t* res_begin, const float* deltaF, const float* adHTdeltaF, const float* cDeltaF,
                                     float* res_toZero) {
    for (int p = 0; p < P; ++p)
        for (int r = res_begin[p]; r < res_begin[p + 1]; ++r) {
            const float* J = recs + (size_t)REC * r;
            const float* dp = adHTdeltaF + (size_t)8 * (host_idx[r] + F * target_idx[r]);
            float jx = 0, jy = 0;
            for (int k = 0; k < 6; ++k) { jx += J[O_JPDXI0 + k] * dp[k]; jy += J[O_JPDXI1 + k] * dp[k]; }
            float cx = 0, cy = 0;
            for (int k = 0; k < 4; ++k) { cx += J[O_JPDC0 + k] * cDeltaF[k]; cy += J[O_JPDC1 + k] * cDeltaF[k]; }
            jx = jx + cx + J[O_JPDD] * deltaF[p];
            jy = jy + cy + J[O_JPDD + 1] * deltaF[p];
            for (int i = 0; i < PATTERN; ++i) {
                float v = J[O_RES + i];
                v -= J[O_JIDX0 + i] * jx;
                v -= J[O_JIDX1 + i] * jy;
                v -= J[O_JAB0 + i] * dp[6];
                v -= J[O_JAB1 + i] * dp[7];
                res_toZero[(size_t)8 * r + i] = v;
            }
        }
}

// accumulate{A,L}F_MT's addPointsInternal<mode> over all points
// (src/bundles/EnergyFunctional.cpp:197-238).  acc_out: F*F x 169 doubles
// (row-major 13x13 = AccumulatorApprox::finish()'s H cast to double).
int eds_oracle_ba_top_accumulate(int mode, int F, int P, int R, const float* recs, const int32_t* host_idx, const int32_t* target_idx,
                                 const int32_t* res_begin, const uint8_t* flags, const float* res_toZero, const float* deltaF,
                                 const float* adHTdeltaF, const float* cDeltaF, int threads, double* acc_out, int64_t* num_out /*F*F or NULL*/,
                                 float* Hdd_out, float* bd_out, float* Hcd_out, int64_t* nres_out) {
    if (mode < 0 || mode > 2 || F <= 0 || P < 0 || R < 0) return 1;
    int nt = std::max(1, threads);
    std::vector<TopAcc> acc(nt);
    for (auto& a : acc) a.init(F);
    parallel_chunks(P, nt, [&](int tid, int lo, int hi) {
        for (int p = lo; p < hi; ++p)
            top_add_point(mode, F, p, recs, host_idx, target_idx, res_begin, flags, res_toZero, deltaF, adHTdeltaF, cDeltaF, acc[tid],
                          Hdd_out, bd_out, Hcd_out);
    });
    std::fill(acc_out, acc_out + (size_t)F * F * 169, 0.0);
    int64_t nres = 0;
    if (num_out) std::fill(num_out, num_out + (size_t)F * F, 0);
    for (auto& a : acc) {  // per-thread partials summed at stitch time, AccumulatedTopHessian.cpp:263-268
        for (size_t i = 0; i < a.H.size(); ++i) acc_out[i] += a.H[i];
        if (num_out) for (size_t i = 0; i < a.num.size(); ++i) num_out[i] += a.num[i];
        nres += a.nres;
    }
    if (nres_out) *nres_out = nres;
    return 0;
}

// AccumulatedTopHessianSSE::stitchDoubleMT + stitchDoubleInternal
// (src/bundles/AccumulatedTopHessian.h:91-139, .cpp:241-303).  adHost/adTarget:
// F*F Mat88, column-major (Eigen default); H: (4+8F)^2 column-major; b: 4+8F.
void eds_oracle_ba_top_stitch(int F, const double* acc, const double* adHost, const double* adTarget, int usePrior,
                              const double* cPrior, const float* cDeltaF, const double* frame_prior, const double* frame_delta_prior,
                              double* H, double* b) {
    const int n = CPARS + 8 * F;
    std::fill(H, H + (size_t)n * n, 0.0);
    std::fill(b, b + n, 0.0);
    for (int k = 0; k < F * F; ++k) {
        const int h = k % F, t = k / F;
        const int hIdx = CPARS + h * 8, tIdx = CPARS + t * 8;
        const double* A = acc + (size_t)169 * k;  // row-major 13x13
        const double* AH = adHost + (size_t)64 * k;
        const double* AT = adTarget + (size_t)64 * k;
        auto A88 = [&](int r, int c) { return A[13 * (CPARS + r) + (CPARS + c)]; };
        // tmpH = AH * A88, tmpT = AT * A88
        double tH[64], tT[64];
        for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 8; ++j) {
                double sh = 0, st = 0;
                for (int m = 0; m < 8; ++m) { sh += ad(AH, i, m) * A88(m, j); st += ad(AT, i, m) * A88(m, j); }
                tH[8 * i + j] = sh; tT[8 * i + j] = st;
            }
        for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 8; ++j) {
                double hh = 0, tt = 0, ht = 0;
                for (int m = 0; m < 8; ++m) { hh += tH[8 * i + m] * ad(AH, j, m); tt += tT[8 * i + m] * ad(AT, j, m); ht += tH[8 * i + m] * ad(AT, j, m); }
                Hat(H, n, hIdx + i, hIdx + j) += hh;
                Hat(H, n, tIdx + i, tIdx + j) += tt;
                Hat(H, n, hIdx + i, tIdx + j) += ht;
            }
        for (int i = 0; i < 8; ++i) {
            for (int c = 0; c < CPARS; ++c) {
                double sh = 0, st = 0;
                for (int m = 0; m < 8; ++m) { sh += ad(AH, i, m) * A[13 * (CPARS + m) + c]; st += ad(AT, i, m) * A[13 * (CPARS + m) + c]; }
                Hat(H, n, hIdx + i, c) += sh;
                Hat(H, n, tIdx + i, c) += st;
            }
            double bh = 0, bt = 0;
            for (int m = 0; m < 8; ++m) { bh += ad(AH, i, m) * A[13 * (CPARS + m) + CPARS + 8]; bt += ad(AT, i, m) * A[13 * (CPARS + m) + CPARS + 8]; }
            b[hIdx + i] += bh;
            b[tIdx + i] += bt;
        }
        for (int r = 0; r < CPARS; ++r) {
            for (int c = 0; c < CPARS; ++c) Hat(H, n, r, c) += A[13 * r + c];
            b[r] += A[13 * r + CPARS + 8];
        }
    }
    if (usePrior) {  // .cpp:292-302
        for (int c = 0; c < CPARS; ++c) { Hat(H, n, c, c) += cPrior[c]; b[c] += cPrior[c] * (double)cDeltaF[c]; }
        for (int h = 0; h < F; ++h)
            for (int i = 0; i < 8; ++i) {
                Hat(H, n, CPARS + h * 8 + i, CPARS + h * 8 + i) += frame_prior[8 * h + i];
                b[CPARS + h * 8 + i] += frame_prior[8 * h + i] * frame_delta_prior[8 * h + i];
            }
    }
    // "make diagonal by copying over parts", .h:125-137
    for (int h = 0; h < F; ++h) {
        const int hIdx = CPARS + h * 8;
        for (int i = 0; i < 8; ++i) for (int c = 0; c < CPARS; ++c) Hat(H, n, c, hIdx + i) = Hat(H, n, hIdx + i, c);
        for (int t = h + 1; t < F; ++t) {
            const int tIdx = CPARS + t * 8;
            for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) Hat(H, n, hIdx + i, tIdx + j) += Hat(H, n, tIdx + j, hIdx + i);
            for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) Hat(H, n, tIdx + j, hIdx + i) = Hat(H, n, hIdx + i, tIdx + j);
        }
    }
}

// accumulateSCF_MT's addPointsInternal (EnergyFunctional.cpp:244-261).
int eds_oracle_ba_sc_accumulate(int F, int P, int R, const int32_t* host_idx, const int32_t* target_idx, const int32_t* res_begin,
                                const uint8_t* flags, const float* JpJdF, const float* Hdd_A, const float* Hdd_L, const float* bd_A,
                                const float* bd_L, const float* Hcd_A, const float* Hcd_L, const float* priorF, const float* deltaF,
                                int shiftPriorToZero, int threads, double* accD, double* accE, double* accEB, double* accHcc, double* accbc,
                                float* HdiF_out, float* bdSum_out) {
    if (F <= 0 || P < 0) return 1;
    int nt = std::max(1, threads);
    std::vector<ScAcc> acc(nt);
    for (auto& a : acc) a.init(F);
    parallel_chunks(P, nt, [&](int tid, int lo, int hi) {
        for (int p = lo; p < hi; ++p)
            sc_add_point(F, p, host_idx, target_idx, res_begin, flags, JpJdF, Hdd_A, Hdd_L, bd_A, bd_L, Hcd_A, Hcd_L, priorF, deltaF,
                         shiftPriorToZero != 0, acc[tid], HdiF_out, bdSum_out);
    });
    std::fill(accD, accD + (size_t)F * F * F * 64, 0.0);
    std::fill(accE, accE + (size_t)F * F * 32, 0.0);
    std::fill(accEB, accEB + (size_t)F * F * 8, 0.0);
    std::fill(accHcc, accHcc + 16, 0.0);
    std::fill(accbc, accbc + 4, 0.0);
    for (auto& a : acc) {
        for (size_t i = 0; i < a.D.size(); ++i) accD[i] += a.D[i];
        for (size_t i = 0; i < a.E.size(); ++i) accE[i] += a.E[i];
        for (size_t i = 0; i < a.EB.size(); ++i) accEB[i] += a.EB[i];
        for (int i = 0; i < 16; ++i) accHcc[i] += a.Hcc[i];
        for (int i = 0; i < 4; ++i) accbc[i] += a.bc[i];
    }
    return 0;
}

// AccumulatedSCHessianSSE::stitchDoubleMT + stitchDoubleInternal
// (src/bundles/AccumulatedSCHessian.h:93-133, .cpp:78-157).
void eds_oracle_ba_sc_stitch(int F, const double* accD, const double* accE, const double* accEB, const double* accHcc, const double* accbc,
                             const double* adHost, const double* adTarget, double* H, double* b) {
    const int n = CPARS + 8 * F, F2 = F * F;
    std::fill(H, H + (size_t)n * n, 0.0);
    std::fill(b, b + n, 0.0);
    for (int k = 0; k < F2; ++k) {
        const int i = k % F, j = k / F;
        const int iIdx = CPARS + i * 8, jIdx = CPARS + j * 8, ij = i + F * j;
        const double* E = accE + (size_t)32 * ij;    // 8x4 row-major
        const double* EB = accEB + (size_t)8 * ij;
        const double* AHij = adHost + (size_t)64 * ij;
        const double* ATij = adTarget + (size_t)64 * ij;
        for (int r = 0; r < 8; ++r) {
            for (int c = 0; c < CPARS; ++c) {
                double sh = 0, st = 0;
                for (int m = 0; m < 8; ++m) { sh += ad(AHij, r, m) * E[4 * m + c]; st += ad(ATij, r, m) * E[4 * m + c]; }
                Hat(H, n, iIdx + r, c) += sh;
                Hat(H, n, jIdx + r, c) += st;
            }
            double bh = 0, bt = 0;
            for (int m = 0; m < 8; ++m) { bh += ad(AHij, r, m) * EB[m]; bt += ad(ATij, r, m) * EB[m]; }
            b[iIdx + r] += bh;
            b[jIdx + r] += bt;
        }
        for (int kk = 0; kk < F; ++kk) {
            const int kIdx = CPARS + kk * 8, ijk = ij + kk * F2, ik = i + F * kk;
            const double* D = accD + (size_t)64 * ijk;  // 8x8 row-major
            const double* AHik = adHost + (size_t)64 * ik;
            const double* ATik = adTarget + (size_t)64 * ik;
            double tH[64], tT[64];
            for (int r = 0; r < 8; ++r)
                for (int c = 0; c < 8; ++c) {
                    double sh = 0, st = 0;
                    for (int m = 0; m < 8; ++m) { sh += ad(AHij, r, m) * D[8 * m + c]; st += ad(ATij, r, m) * D[8 * m + c]; }
                    tH[8 * r + c] = sh; tT[8 * r + c] = st;
                }
            for (int r = 0; r < 8; ++r)
                for (int c = 0; c < 8; ++c) {
                    double hh = 0, tt = 0, th = 0, ht = 0;
                    for (int m = 0; m < 8; ++m) {
                        hh += tH[8 * r + m] * ad(AHik, c, m);
                        tt += tT[8 * r + m] * ad(ATik, c, m);
                        th += tT[8 * r + m] * ad(AHik, c, m);
                        ht += tH[8 * r + m] * ad(ATik, c, m);
                    }
                    Hat(H, n, iIdx + r, iIdx + c) += hh;
                    Hat(H, n, jIdx + r, kIdx + c) += tt;
                    Hat(H, n, jIdx + r, iIdx + c) += th;
                    Hat(H, n, iIdx + r, kIdx + c) += ht;
                }
        }
    }
    for (int r = 0; r < CPARS; ++r) { for (int c = 0; c < CPARS; ++c) Hat(H, n, r, c) += accHcc[4 * r + c]; b[r] += accbc[r]; }
    for (int h = 0; h < F; ++h) {  // .h:127-131
        const int hIdx = CPARS + h * 8;
        for (int r = 0; r < 8; ++r) for (int c = 0; c < CPARS; ++c) Hat(H, n, c, hIdx + r) = Hat(H, n, hIdx + r, c);
    }
}

// PointFrameResidual::linearize, src/tracking/Residuals.cpp:69-265 (the feeder of the accumulators),
// for every residual of the window.  float arithmetic in the reference's operation order; built with
// -ffp-contract=off, so the GPU kernel (written with explicit round-to-nearest mul/add) can be compared bit for bit.
//   dI          [F][H*W*3]  FrameHessian::dI, Vec3f {I, dx, dy} per pixel
//   precalc     [F*F][28]   FrameFramePrecalc (HessianBlocks.h:77-104) of (host,target) at host + F*target:
//                           PRE_RTll_0 (9, column-major like Eigen), PRE_tTll_0 (3), PRE_KRKiTll (9, column-major),
//                           PRE_KtTll (3), PRE_aff_mode (2), PRE_b0_mode (1), pad (1)
//   calib       fxl, fyl, cxl, cyl
//   points      u, v, idepth_zero_scaled, idepth_scaled [P]; color, weights [P][8]
//   state_in    [R] ResState (0 IN, 1 OOB, 2 OUTLIER) or null = all IN; OOB residuals are skipped (:73-74)
//   frame_energy_th [F]     FrameHessian::frameEnergyTH
// out: recs [R][76] (zeroed where the residual leaves as OOB: the reference leaves J stale there and never reads it),
//      state_new [R], energy_new [R] (state_NewEnergy; 0 where OOB: the reference keeps the old state_energy)
void eds_oracle_ba_linearize(int F, int P, int R, int H, int W, const float* dI, const float* precalc, const float* calib,
                             const float* pu, const float* pv, const float* idepth_zero_scaled, const float* idepth_scaled,
                             const float* color, const float* weights, const int32_t* host_idx, const int32_t* target_idx,
                             const int32_t* res_begin, const uint8_t* state_in, const float* frame_energy_th, float* recs,
                             int32_t* state_new, float* energy_new) {
    static const int patternP[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};  // settings.cpp:276
    const float setting_huberTH = 9.0f;                      // settings.cpp:127
    const float setting_outlierTHSumComponent = 50.0f * 50.0f;  // settings.cpp:91
    const float fxl = calib[0], fyl = calib[1], cxl = calib[2], cyl = calib[3];
    const float fxli = 1.0f / fxl, fyli = 1.0f / fyl;
    const float wM3G = (float)(W - 3), hM3G = (float)(H - 3);  // globalCalib.cpp:74-75
    for (int p = 0; p < P; ++p)
        for (int r = res_begin[p]; r < res_begin[p + 1]; ++r) {
            float* J = recs + (size_t)REC * r;
            for (int i = 0; i < REC; ++i) J[i] = 0.f;
            energy_new[r] = 0.f;
            state_new[r] = 1;  // OOB unless the residual survives
            if (state_in && state_in[r] == 1) continue;  // :73-74
            const int h = host_idx[r], t = target_idx[r];
            const float* pc = precalc + (size_t)28 * (h + F * t);
            auto Rm = [&](int i, int j) { return pc[j * 3 + i]; };         // PRE_RTll_0(i,j)
            auto KRKi = [&](int i, int j) { return pc[12 + j * 3 + i]; };  // PRE_KRKiTll(i,j)
            const float* tt = pc + 9;
            const float* Kt = pc + 21;
            const float aff0 = pc[24], aff1 = pc[25], b0 = pc[26];
            const float* dIl = dI + (size_t)t * H * W * 3;
            // centre point, ResidualProjections.h:60-86 with dx = dy = 0
            const float KliP0 = (pu[p] + 0 - cxl) * fxli, KliP1 = (pv[p] + 0 - cyl) * fyli;
            float ptp[3];
            for (int i = 0; i < 3; ++i) ptp[i] = ((Rm(i, 0) * KliP0 + Rm(i, 1) * KliP1) + Rm(i, 2) * 1.0f) + tt[i] * idepth_zero_scaled[p];
            const float drescale = 1.0f / ptp[2];
            const float new_idepth = idepth_zero_scaled[p] * drescale;
            if (!(drescale > 0)) continue;
            const float u = ptp[0] * drescale, v = ptp[1] * drescale;
            const float Ku0 = u * fxl + cxl, Kv0 = v * fyl + cyl;
            if (!(Ku0 > 1.1f && Kv0 > 1.1f && Ku0 < wM3G && Kv0 < hM3G)) continue;
            // :108-147 (SCALE_IDEPTH = SCALE_F = SCALE_C = 1, HessianBlocks.h:58-62)
            float rec[REC];
            for (int i = 0; i < REC; ++i) rec[i] = 0.f;
            rec[O_JPDD] = drescale * (tt[0] - tt[2] * u) * 1.0f * fxl;
            rec[O_JPDD + 1] = drescale * (tt[1] - tt[2] * v) * 1.0f * fyl;
            float dCx2 = drescale * (Rm(2, 0) * u - Rm(0, 0));
            float dCx3 = fxl * drescale * (Rm(2, 1) * u - Rm(0, 1)) * fyli;
            float dCx0 = KliP0 * dCx2, dCx1 = KliP1 * dCx3;
            float dCy2 = fyl * drescale * (Rm(2, 0) * v - Rm(1, 0)) * fxli;
            float dCy3 = drescale * (Rm(2, 1) * v - Rm(1, 1));
            float dCy0 = KliP0 * dCy2, dCy1 = KliP1 * dCy3;
            rec[O_JPDC0 + 0] = (dCx0 + u) * 1.0f;
            rec[O_JPDC0 + 1] = dCx1 * 1.0f;
            rec[O_JPDC0 + 2] = (dCx2 + 1) * 1.0f;
            rec[O_JPDC0 + 3] = dCx3 * 1.0f;
            rec[O_JPDC1 + 0] = dCy0 * 1.0f;
            rec[O_JPDC1 + 1] = (dCy1 + v) * 1.0f;
            rec[O_JPDC1 + 2] = dCy2 * 1.0f;
            rec[O_JPDC1 + 3] = (dCy3 + 1) * 1.0f;
            rec[O_JPDXI0 + 0] = new_idepth * fxl;
            rec[O_JPDXI0 + 1] = 0;
            rec[O_JPDXI0 + 2] = -new_idepth * u * fxl;
            rec[O_JPDXI0 + 3] = -u * v * fxl;
            rec[O_JPDXI0 + 4] = (1 + u * u) * fxl;
            rec[O_JPDXI0 + 5] = -v * fxl;
            rec[O_JPDXI1 + 0] = 0;
            rec[O_JPDXI1 + 1] = new_idepth * fyl;
            rec[O_JPDXI1 + 2] = -new_idepth * v * fyl;
            rec[O_JPDXI1 + 3] = -(1 + v * v) * fyl;
            rec[O_JPDXI1 + 4] = u * v * fyl;
            rec[O_JPDXI1 + 5] = u * fyl;
            float J00 = 0, J11 = 0, J10 = 0, A00 = 0, A01 = 0, A10 = 0, A11 = 0, B00 = 0, B01 = 0, B11 = 0, wJI2 = 0, energyLeft = 0;
            bool oob = false;
            for (int idx = 0; idx < 8 && !oob; ++idx) {
                // ResidualProjections.h:46-56
                const float up = pu[p] + patternP[idx][0], vp = pv[p] + patternP[idx][1];
                float q[3];
                for (int i = 0; i < 3; ++i) q[i] = ((KRKi(i, 0) * up + KRKi(i, 1) * vp) + KRKi(i, 2) * 1.0f) + Kt[i] * idepth_scaled[p];
                const float Ku = q[0] / q[2], Kv = q[1] / q[2];
                if (!(Ku > 1.1f && Kv > 1.1f && Ku < wM3G && Kv < hM3G)) { oob = true; break; }
                // getInterpolatedElement33, globalFuncs.h:78-92
                const int ix = (int)Ku, iy = (int)Kv;
                const float dx = Ku - ix, dy = Kv - iy, dxdy = dx * dy;
                const float* bp = dIl + (size_t)3 * (ix + iy * W);
                float hit[3];
                for (int c = 0; c < 3; ++c)
                    hit[c] = ((dxdy * bp[3 * (1 + W) + c] + (dy - dxdy) * bp[3 * W + c]) + (dx - dxdy) * bp[3 + c]) + (1 - dx - dy + dxdy) * bp[c];
                const float residual = hit[0] - (float)(aff0 * color[8 * p + idx] + aff1);
                const float drdA = color[8 * p + idx] - b0;
                if (!std::isfinite(hit[0])) { oob = true; break; }
                float w = sqrtf(setting_outlierTHSumComponent / (setting_outlierTHSumComponent + (hit[1] * hit[1] + hit[2] * hit[2])));
                w = 0.5f * (w + weights[8 * p + idx]);
                float hw = fabsf(residual) < setting_huberTH ? 1 : setting_huberTH / fabsf(residual);
                energyLeft += w * w * hw * residual * residual * (2 - hw);
                if (hw < 1) hw = sqrtf(hw);
                hw = hw * w;
                hit[1] *= hw;
                hit[2] *= hw;
                rec[O_RES + idx] = residual * hw;
                rec[O_JIDX0 + idx] = hit[1];
                rec[O_JIDX1 + idx] = hit[2];
                rec[O_JAB0 + idx] = drdA * hw;
                rec[O_JAB1 + idx] = hw;
                J00 += hit[1] * hit[1]; J11 += hit[2] * hit[2]; J10 += hit[1] * hit[2];
                A00 += drdA * hw * hit[1]; A01 += drdA * hw * hit[2]; A10 += hw * hit[1]; A11 += hw * hit[2];
                B00 += drdA * drdA * hw * hw; B01 += drdA * hw * hw; B11 += hw * hw;
                wJI2 += hw * hw * (hit[1] * hit[1] + hit[2] * hit[2]);
            }
            if (oob) continue;
            // Mat22f members are column-major: (0,0) (1,0) (0,1) (1,1)
            rec[O_JIDX2 + 0] = J00; rec[O_JIDX2 + 1] = J10; rec[O_JIDX2 + 2] = J10; rec[O_JIDX2 + 3] = J11;
            rec[O_JABJIDX + 0] = A00; rec[O_JABJIDX + 1] = A10; rec[O_JABJIDX + 2] = A01; rec[O_JABJIDX + 3] = A11;
            rec[O_JAB2 + 0] = B00; rec[O_JAB2 + 1] = B01; rec[O_JAB2 + 2] = B01; rec[O_JAB2 + 3] = B11;
            for (int i = 0; i < REC; ++i) J[i] = rec[i];
            const float th = std::max<float>(frame_energy_th[h], frame_energy_th[t]);  // :253-261
            if (energyLeft > th || wJI2 < 2) { energyLeft = th; state_new[r] = 2; }
            else state_new[r] = 0;
            energy_new[r] = energyLeft;
        }
}

// EnergyFunctional::resubstituteF_MT + resubstituteFPt, src/bundles/EnergyFunctional.cpp:263-317 (float arithmetic):
// per-point inverse-depth step from the solved frame / calibration update x (4 + 8F doubles).
// adHost / adTarget: F*F column-major 8x8 (index host + F*target), cast to float like adHostF / adTargetF.
void eds_oracle_ba_resubstitute(int F, int P, int R, const double* x, const double* adHost, const double* adTarget,
                                const int32_t* host_idx, const int32_t* target_idx, const int32_t* res_begin, const uint8_t* flags,
                                const float* JpJdF, const float* bdSumF, const float* Hcd_accAF, const float* Hcd_accLF, const float* HdiF,
                                float* step_out) {
    std::vector<float> xF(CPARS + 8 * F);
    for (size_t i = 0; i < xF.size(); ++i) xF[i] = (float)x[i];
    std::vector<float> xAd((size_t)F * F * 8);  // xAd[F*h + t]
    for (int h = 0; h < F; ++h)
        for (int t = 0; t < F; ++t) {
            const double* AH = adHost + (size_t)64 * (h + F * t);
            const double* AT = adTarget + (size_t)64 * (h + F * t);
            for (int c = 0; c < 8; ++c) {
                float a = 0.f, b = 0.f;
                for (int k = 0; k < 8; ++k) a += xF[CPARS + 8 * h + k] * (float)AH[c * 8 + k];  // row vector times column c
                for (int k = 0; k < 8; ++k) b += xF[CPARS + 8 * t + k] * (float)AT[c * 8 + k];
                xAd[(size_t)8 * (F * h + t) + c] = a + b;
            }
        }
    for (int p = 0; p < P; ++p) {
        int ngoodres = 0;
        for (int r = res_begin[p]; r < res_begin[p + 1]; ++r) if (flags[r] & 1) ngoodres++;
        if (ngoodres == 0) { step_out[p] = 0.f; continue; }
        float b = bdSumF[p];
        float dot = 0.f;
        for (int k = 0; k < CPARS; ++k) dot += xF[k] * (Hcd_accAF[4 * p + k] + Hcd_accLF[4 * p + k]);
        b -= dot;
        for (int r = res_begin[p]; r < res_begin[p + 1]; ++r) {
            if (!(flags[r] & 1)) continue;
            const float* xa = &xAd[(size_t)8 * (host_idx[r] * F + target_idx[r])];
            float d = 0.f;
            for (int k = 0; k < 8; ++k) d += xa[k] * JpJdF[(size_t)8 * r + k];
            b -= d;
        }
        step_out[p] = -b * HdiF[p];
    }
}

// EnergyFunctional::calcLEnergyF_MT + calcLEnergyPt, EnergyFunctional.cpp:332-415: energy of the linearised part at
// the current delta, E = sum over linearised active residuals (2 res_toZero + J delta) . J delta + sum_p deltaF^2 priorF
// + the frame / calibration prior terms.  The reference sums in float (Accumulator11); the oracle in double.
double eds_oracle_ba_calc_l_energy(int F, int P, int R, const float* recs, const int32_t* host_idx, const int32_t* target_idx,
                                   const int32_t* res_begin, const uint8_t* flags, const float* res_toZero, const float* deltaF,
                                   const float* priorF, const float* adHTdeltaF, const float* cDeltaF, const double* cPrior,
                                   const double* frame_prior, const double* frame_delta_prior) {
    double E = 0.0;
    if (frame_prior && frame_delta_prior)
        for (int i = 0; i < 8 * F; ++i) E += frame_delta_prior[i] * frame_prior[i] * frame_delta_prior[i];
    if (cPrior)
        for (int k = 0; k < CPARS; ++k) E += (double)cDeltaF[k] * (double)(float)cPrior[k] * (double)cDeltaF[k];
    for (int p = 0; p < P; ++p) {
        const float dd = deltaF[p];
        for (int r = res_begin[p]; r < res_begin[p + 1]; ++r) {
            if (!(flags[r] & 2) || !(flags[r] & 1)) continue;
            const float* J = recs + (size_t)REC * r;
            const float* dp = adHTdeltaF + (size_t)8 * (host_idx[r] + F * target_idx[r]);
            float jx = 0, jy = 0, cx = 0, cy = 0;
            for (int k = 0; k < 6; ++k) { jx += J[O_JPDXI0 + k] * dp[k]; jy += J[O_JPDXI1 + k] * dp[k]; }
            for (int k = 0; k < 4; ++k) { cx += J[O_JPDC0 + k] * cDeltaF[k]; cy += J[O_JPDC1 + k] * cDeltaF[k]; }
            jx = jx + cx + J[O_JPDD] * dd;
            jy = jy + cy + J[O_JPDD + 1] * dd;
            for (int i = 0; i < PATTERN; ++i) {
                float Jdelta = J[O_JIDX0 + i] * jx;
                Jdelta += J[O_JIDX1 + i] * jy;
                Jdelta += J[O_JAB0 + i] * dp[6];
                Jdelta += J[O_JAB1 + i] * dp[7];
                float r0 = res_toZero[(size_t)8 * r + i];
                r0 = r0 + r0;
                r0 = r0 + Jdelta;
                E += (double)(Jdelta * r0);
            }
        }
        E += (double)(deltaF[p] * deltaF[p] * priorF[p]);
    }
    return E;
}

}  // extern "C"
