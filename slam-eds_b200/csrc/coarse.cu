// DSO-style coarse tracker evaluation: CoarseTracker::calcRes fused with
// CoarseTracker::calcGSSSE (src/tracking/CoarseTracker.cpp:287-498).  One launch per pyramid level
// and Gauss-Newton iteration: every reference point is warped into the new frame, its residual,
// Huber weight and the 8-DoF Jacobian (6 pose + 2 affine brightness) are formed in registers and
// reduced into the 9x9 system [J r]^T w [J r]; the buf_warped_* arrays of the reference are never
// materialised.  The same structure as the event tracker's sweep: gather -> residual -> weighted
// outer-product reduction; per-CTA partials are summed in a fixed order (deterministic).
//
// float32 in the reference's operation order with explicit round-to-nearest mul/add (no FMA
// contraction), so the in-bounds / cutoff decisions are those a plain float32 CPU evaluation takes;
// the sums E, H, b are carried in double across threads.
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"

namespace {

constexpr int CT_THREADS = 256;
constexpr int CT_NACC = 45;             // upper triangle of the 9x9 system
constexpr int CT_NPART = CT_NACC + 7;   // + E, numTermsInE, numTermsInWarped, numSaturated, shiftT, shiftRT, shiftNum

__device__ __forceinline__ float fm(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fa(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fs(float a, float b) { return __fadd_rn(a, -b); }

struct CoarseArgs {
    int lvl, wl, hl, n;
    float fxl, fyl, cxl, cyl;
    float RKi[9];  // row-major
    float Ki[9];   // row-major
    float t[3];
    float aff0, aff1, b0, cutoffTH;
    const float4* image;  // {I, dx, dy, 0}
    const float *pc_u, *pc_v, *pc_idepth, *pc_color;
    double* partials;     // [gridDim.x][CT_NPART]
};

__global__ void __launch_bounds__(CT_THREADS) coarse_res_gs_kernel(CoarseArgs a) {
    __shared__ double red[CT_THREADS / 32][CT_NPART];
    float acc[CT_NACC];
#pragma unroll
    for (int i = 0; i < CT_NACC; ++i) acc[i] = 0.f;
    double E = 0.0;
    int nE = 0, nW = 0, nSat = 0, nShift = 0;
    float shT = 0.f, shRT = 0.f;
    const float HUBER = 9.0f;  // settings.cpp:127
    const float maxEnergy = fs(fm(fm(2.0f, HUBER), a.cutoffTH), fm(HUBER, HUBER));  // :372
    const float wlim = (float)(a.wl - 3), hlim = (float)(a.hl - 3);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
        const float id = __ldg(a.pc_idepth + i), x = __ldg(a.pc_u + i), y = __ldg(a.pc_v + i);
        float pt[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) pt[k] = fa(fa(fa(fm(a.RKi[3 * k], x), fm(a.RKi[3 * k + 1], y)), a.RKi[3 * k + 2]), fm(a.t[k], id));
        const float u = __fdiv_rn(pt[0], pt[2]), v = __fdiv_rn(pt[1], pt[2]);
        const float Ku = fa(fm(a.fxl, u), a.cxl), Kv = fa(fm(a.fyl, v), a.cyl);
        const float new_idepth = __fdiv_rn(id, pt[2]);
        if (a.lvl == 0 && (i & 31) == 0) {  // :403-434
            float p1[3], p2[3], p3[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float kp = fa(fa(fm(a.Ki[3 * k], x), fm(a.Ki[3 * k + 1], y)), a.Ki[3 * k + 2]);
                const float ti = fm(a.t[k], id);
                p1[k] = fa(kp, ti);
                p2[k] = fs(kp, ti);
                p3[k] = fs(fa(fa(fm(a.RKi[3 * k], x), fm(a.RKi[3 * k + 1], y)), a.RKi[3 * k + 2]), ti);
            }
            const float KuT = fa(fm(a.fxl, __fdiv_rn(p1[0], p1[2])), a.cxl), KvT = fa(fm(a.fyl, __fdiv_rn(p1[1], p1[2])), a.cyl);
            const float KuT2 = fa(fm(a.fxl, __fdiv_rn(p2[0], p2[2])), a.cxl), KvT2 = fa(fm(a.fyl, __fdiv_rn(p2[1], p2[2])), a.cyl);
            const float Ku3 = fa(fm(a.fxl, __fdiv_rn(p3[0], p3[2])), a.cxl), Kv3 = fa(fm(a.fyl, __fdiv_rn(p3[1], p3[2])), a.cyl);
            auto sq = [](float d0, float d1) { return fa(fm(d0, d0), fm(d1, d1)); };
            shT = fa(shT, sq(fs(KuT, x), fs(KvT, y)));
            shT = fa(shT, sq(fs(KuT2, x), fs(KvT2, y)));
            shRT = fa(shRT, sq(fs(Ku, x), fs(Kv, y)));
            shRT = fa(shRT, sq(fs(Ku3, x), fs(Kv3, y)));
            nShift += 2;
        }
        if (!(Ku > 2.0f && Kv > 2.0f && Ku < wlim && Kv < hlim && new_idepth > 0.0f)) continue;
        const float refColor = __ldg(a.pc_color + i);
        const int ix = (int)Ku, iy = (int)Kv;  // getInterpolatedElement33, globalFuncs.h:78-92
        const float dx = fs(Ku, (float)ix), dy = fs(Kv, (float)iy), dxdy = fm(dx, dy);
        const float4* bp = a.image + (ix + iy * a.wl);
        const float4 p00 = __ldg(bp), p10 = __ldg(bp + 1), p01 = __ldg(bp + a.wl), p11 = __ldg(bp + 1 + a.wl);
        const float w11 = dxdy, w01 = fs(dy, dxdy), w10 = fs(dx, dxdy), w00 = fa(fs(fs(1.0f, dx), dy), dxdy);
        const float hit0 = fa(fa(fa(fm(w11, p11.x), fm(w01, p01.x)), fm(w10, p10.x)), fm(w00, p00.x));
        const float hit1 = fa(fa(fa(fm(w11, p11.y), fm(w01, p01.y)), fm(w10, p10.y)), fm(w00, p00.y));
        const float hit2 = fa(fa(fa(fm(w11, p11.z), fm(w01, p01.z)), fm(w10, p10.z)), fm(w00, p00.z));
        if (!isfinite(hit0)) continue;
        const float residual = fs(hit0, fa(fm(a.aff0, refColor), a.aff1));
        const float ar = fabsf(residual);
        const float hw = ar < HUBER ? 1.0f : __fdiv_rn(HUBER, ar);
        nE++;
        if (ar > a.cutoffTH) {
            E += (double)maxEnergy;
            nSat++;
            continue;
        }
        E += (double)fm(fm(fm(hw, residual), residual), fs(2.0f, hw));
        nW++;
        // calcGSSSE :303-330
        const float ddx = fm(hit1, a.fxl), ddy = fm(hit2, a.fyl);
        float J[9];
        J[0] = fm(new_idepth, ddx);
        J[1] = fm(new_idepth, ddy);
        J[2] = fs(0.0f, fm(new_idepth, fa(fm(u, ddx), fm(v, ddy))));
        J[3] = fs(0.0f, fa(fm(fm(u, v), ddx), fm(ddy, fa(1.0f, fm(v, v)))));
        J[4] = fa(fm(fm(u, v), ddy), fm(ddx, fa(1.0f, fm(u, u))));
        J[5] = fs(fm(u, ddy), fm(v, ddx));
        J[6] = fm(a.aff0, fs(a.b0, refColor));
        J[7] = -1.0f;
        J[8] = residual;
        int e = 0;
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            const float Jw = fm(J[r], hw);  // MatrixAccumulators.h:1091-1150
#pragma unroll
            for (int c = r; c < 9; ++c) acc[e++] += Jw * J[c];
        }
    }
    // warp -> CTA -> one partial per CTA, all in double
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double vals[CT_NPART];
#pragma unroll
    for (int i = 0; i < CT_NACC; ++i) vals[i] = (double)acc[i];
    vals[CT_NACC] = E; vals[CT_NACC + 1] = nE; vals[CT_NACC + 2] = nW; vals[CT_NACC + 3] = nSat;
    vals[CT_NACC + 4] = (double)shT; vals[CT_NACC + 5] = (double)shRT; vals[CT_NACC + 6] = nShift;
#pragma unroll
    for (int i = 0; i < CT_NPART; ++i) {
        double s = vals[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[warp][i] = s;
    }
    __syncthreads();
    if (threadIdx.x < CT_NPART) {
        double s = 0.0;
        for (int w = 0; w < CT_THREADS / 32; ++w) s += red[w][threadIdx.x];
        a.partials[(size_t)blockIdx.x * CT_NPART + threadIdx.x] = s;
    }
}

__global__ void coarse_finalize_kernel(const double* __restrict__ partials, int nblocks, double* __restrict__ out) {
    if (threadIdx.x < CT_NPART) {
        double s = 0.0;
        for (int b = 0; b < nblocks; ++b) s += partials[(size_t)b * CT_NPART + threadIdx.x];
        out[threadIdx.x] = s;
    }
}

__global__ void coarse_pad_image_kernel(const float* __restrict__ src, float4* __restrict__ dst, size_t npix) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npix) dst[i] = make_float4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 0.f);
}

struct Level {
    int w = 0, h = 0, n = 0, cap = 0;
    float fx = 0, fy = 0, cx = 0, cy = 0;
    float Ki[9] = {0};  // column-major as given
    float4* image = nullptr;
    float* pc = nullptr;  // [4][cap]: u, v, idepth, color
    bool has_image = false;
};

}  // namespace

struct edsgpu_coarse {
    edsgpu_ctx* ctx = nullptr;
    std::vector<Level> levels;
    double* partials = nullptr;  // [max grid][CT_NPART] + CT_NPART result
    int max_grid = 0;
};

extern "C" {

edsgpu_status edsgpu_coarse_create(edsgpu_ctx* ctx, int num_levels, edsgpu_coarse** out) {
    if (!ctx || !out) return EDSGPU_INVALID_ARGUMENT;
    EDS_REQUIRE(ctx, num_levels >= 1 && num_levels <= 8, "coarse_create: num_levels must be in [1,8]");
    DeviceGuard g(ctx->device);
    edsgpu_coarse* c = new edsgpu_coarse();
    c->ctx = ctx;
    c->levels.resize(num_levels);
    c->max_grid = 2 * ctx->num_sms;
    cudaError_t e = cudaMalloc(&c->partials, sizeof(double) * CT_NPART * ((size_t)c->max_grid + 1));
    if (e != cudaSuccess) { delete c; return edsgpu_fail(ctx, EDSGPU_CUDA_ERROR, cudaGetErrorString(e)); }
    *out = c;
    return EDSGPU_OK;
}

void edsgpu_coarse_destroy(edsgpu_coarse* c) {
    if (!c) return;
    DeviceGuard g(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    for (Level& l : c->levels) {
        if (l.image) cudaFree(l.image);
        if (l.pc) cudaFree(l.pc);
    }
    if (c->partials) cudaFree(c->partials);
    delete c;
}

edsgpu_status edsgpu_coarse_set_level(edsgpu_coarse* c, int lvl, int width, int height, float fx, float fy, float cx, float cy, const float Ki[9]) {
    if (!c) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = c->ctx;
    EDS_REQUIRE(ctx, lvl >= 0 && lvl < (int)c->levels.size() && width > 6 && height > 6 && Ki, "coarse_set_level: bad arguments");
    DeviceGuard g(ctx->device);
    Level& l = c->levels[lvl];
    if (l.image && (l.w != width || l.h != height)) {
        EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(l.image);
        l.image = nullptr;
        l.has_image = false;
    }
    l.w = width; l.h = height; l.fx = fx; l.fy = fy; l.cx = cx; l.cy = cy;
    memcpy(l.Ki, Ki, sizeof(l.Ki));
    return EDSGPU_OK;
}

edsgpu_status edsgpu_coarse_set_reference(edsgpu_coarse* c, int lvl, int n, const float* pc_u, const float* pc_v, const float* pc_idepth,
                                          const float* pc_color) {
    if (!c) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = c->ctx;
    EDS_REQUIRE(ctx, lvl >= 0 && lvl < (int)c->levels.size() && n >= 0, "coarse_set_reference: bad arguments");
    EDS_REQUIRE(ctx, n == 0 || (pc_u && pc_v && pc_idepth && pc_color), "coarse_set_reference: null array");
    DeviceGuard g(ctx->device);
    Level& l = c->levels[lvl];
    if (n > l.cap) {
        EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (l.pc) cudaFree(l.pc);
        l.pc = nullptr;
        l.cap = 0;
        EDS_CUDA(ctx, cudaMalloc(&l.pc, sizeof(float) * 4 * (size_t)n));
        l.cap = n;
    }
    l.n = n;
    const float* src[4] = {pc_u, pc_v, pc_idepth, pc_color};
    for (int k = 0; k < 4 && n > 0; ++k)
        EDS_CUDA(ctx, cudaMemcpyAsync(l.pc + (size_t)k * l.cap, src[k], sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EDSGPU_OK;
}

edsgpu_status edsgpu_coarse_set_new_frame(edsgpu_coarse* c, int lvl, const float* dI) {
    if (!c) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = c->ctx;
    EDS_REQUIRE(ctx, lvl >= 0 && lvl < (int)c->levels.size() && dI, "coarse_set_new_frame: bad arguments");
    Level& l = c->levels[lvl];
    EDS_REQUIRE(ctx, l.w > 0, "coarse_set_new_frame: call edsgpu_coarse_set_level first");
    DeviceGuard g(ctx->device);
    const size_t npix = (size_t)l.w * l.h;
    if (!l.image) EDS_CUDA(ctx, cudaMalloc(&l.image, sizeof(float4) * npix));
    edsgpu_status st = edsgpu_ensure_scratch(ctx, npix * 3 * sizeof(float));
    if (st != EDSGPU_OK) return st;
    EDS_CUDA(ctx, cudaMemcpyAsync(ctx->scratch, dI, npix * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    coarse_pad_image_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, ctx->stream>>>((const float*)ctx->scratch, l.image, npix);
    ctx->launches++;
    EDS_CUDA(ctx, cudaGetLastError());
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    l.has_image = true;
    return EDSGPU_OK;
}

edsgpu_status edsgpu_coarse_calc_res_gs(edsgpu_coarse* c, int lvl, const double R[9], const double t[3], const float affLL[2], float b0,
                                        float cutoffTH, double rs[6], double H[64], double b[8]) {
    if (!c) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = c->ctx;
    EDS_REQUIRE(ctx, lvl >= 0 && lvl < (int)c->levels.size() && R && t && affLL && rs, "coarse_calc_res_gs: bad arguments");
    Level& l = c->levels[lvl];
    EDS_REQUIRE(ctx, l.has_image && l.n > 0, "coarse_calc_res_gs: level needs a reference point cloud and a new frame");
    DeviceGuard g(ctx->device);
    edsgpu_status st = edsgpu_ensure_pinned(ctx, sizeof(double) * CT_NPART);
    if (st != EDSGPU_OK) return st;
    CoarseArgs a{};
    a.lvl = lvl; a.wl = l.w; a.hl = l.h; a.n = l.n;
    a.fxl = l.fx; a.fyl = l.fy; a.cxl = l.cx; a.cyl = l.cy;
    // RKi = refToNew.rotationMatrix().cast<float>() * Ki[lvl], t = translation().cast<float>()  (:363-364), float products
    for (int i = 0; i < 3; ++i) {
        a.t[i] = (float)t[i];
        for (int j = 0; j < 3; ++j) {
            a.Ki[3 * i + j] = l.Ki[3 * j + i];
            a.RKi[3 * i + j] = ((float)R[3 * i + 0] * l.Ki[3 * j + 0] + (float)R[3 * i + 1] * l.Ki[3 * j + 1]) + (float)R[3 * i + 2] * l.Ki[3 * j + 2];
        }
    }
    a.aff0 = affLL[0]; a.aff1 = affLL[1]; a.b0 = b0; a.cutoffTH = cutoffTH;
    a.image = l.image;
    a.pc_u = l.pc; a.pc_v = l.pc + l.cap; a.pc_idepth = l.pc + 2 * (size_t)l.cap; a.pc_color = l.pc + 3 * (size_t)l.cap;
    a.partials = c->partials;
    const int grid = std::max(1, std::min((l.n + CT_THREADS - 1) / CT_THREADS, c->max_grid));
    double* d_out = c->partials + (size_t)c->max_grid * CT_NPART;
    coarse_res_gs_kernel<<<grid, CT_THREADS, 0, ctx->stream>>>(a);
    ctx->launches++;
    coarse_finalize_kernel<<<1, 64, 0, ctx->stream>>>(c->partials, grid, d_out);
    ctx->launches++;
    EDS_CUDA(ctx, cudaGetLastError());
    EDS_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, d_out, sizeof(double) * CT_NPART, cudaMemcpyDeviceToHost, ctx->stream));
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const double* v = (const double*)ctx->pinned;
    const double E = v[CT_NACC], nE = v[CT_NACC + 1], nW = v[CT_NACC + 2], nSat = v[CT_NACC + 3];
    const float shT = (float)v[CT_NACC + 4], shRT = (float)v[CT_NACC + 5], shNum = (float)v[CT_NACC + 6];
    rs[0] = E;
    rs[1] = nE;
    rs[2] = shT / (shNum + 0.1);
    rs[3] = 0;
    rs[4] = shRT / (shNum + 0.1);
    rs[5] = (float)nSat / (float)nE;
    if (H && b) {
        // H_out = acc.H.topLeftCorner<8,8>().cast<double>() * (1.0f / n), n padded to a multiple of 4 (:466-478, :332-344)
        const long long npad = ((long long)nW + 3) / 4 * 4;
        const float inv_n = 1.0f / (float)npad;
        const double sc[8] = {1, 1, 1, 1, 1, 1, 10.0f, 1000.0f};  // SCALE_XI_ROT, SCALE_XI_TRANS, SCALE_A, SCALE_B
        double full[9][9];
        int e = 0;
        for (int r = 0; r < 9; ++r)
            for (int q = r; q < 9; ++q) { full[r][q] = full[q][r] = v[e++]; }
        for (int r = 0; r < 8; ++r) {
            for (int q = 0; q < 8; ++q) H[8 * r + q] = (double)(float)full[r][q] * inv_n * (sc[r] * sc[q]);
            b[r] = (double)(float)full[r][8] * inv_n * sc[r];
        }
    }
    return EDSGPU_OK;
}

// ---- trackNewestCoarse: the coarse-to-fine Gauss-Newton loop on the host around the device evaluation ----------
namespace {

// Sophus SO3::exp (quaternion form) -> rotation matrix, row-major
void so3_exp(const double* w, double* R) {
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2);
    double imag, real;
    if (th < 1e-10) {
        const double th4 = th2 * th2;
        imag = 0.5 - th2 / 48.0 + th4 / 3840.0;
        real = 1.0 - th2 / 8.0 + th4 / 384.0;
    } else {
        imag = sin(0.5 * th) / th;
        real = cos(0.5 * th);
    }
    const double x = imag * w[0], y = imag * w[1], z = imag * w[2], q = real;
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * q); R[2] = 2 * (x * z + y * q);
    R[3] = 2 * (x * y + z * q); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * q);
    R[6] = 2 * (x * z - y * q); R[7] = 2 * (y * z + x * q); R[8] = 1 - 2 * (x * x + y * y);
}

// (R, t) <- SE3::exp(inc) * (R, t), inc = [translation part, rotation part]
void se3_left_update(const double* inc6, double* R, double* t) {
    const double* u = inc6;
    const double* w = inc6 + 3;
    double Re[9], V[9];
    so3_exp(w, Re);
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2);
    const double W[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    if (th < 1e-10) {
        for (int i = 0; i < 9; ++i) V[i] = Re[i];
    } else {
        const double A = (1.0 - cos(th)) / th2, B = (th - sin(th)) / (th2 * th);
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
    memcpy(R, Rn, sizeof(Rn));
    memcpy(t, tn, sizeof(tn));
}

// x = A^-1 rhs, A symmetric positive definite 8x8 (LDL^T; the reference calls Eigen's ldlt().solve)
void solve8(const double* A, const double* rhs, double* x) {
    double L[64] = {0}, D[8], y[8];
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
    for (int i = 0; i < 8; ++i) { double v = rhs[i]; for (int k = 0; k < i; ++k) v -= L[8 * i + k] * y[k]; y[i] = v; }
    for (int i = 0; i < 8; ++i) y[i] /= D[i];
    for (int i = 7; i >= 0; --i) { double v = y[i]; for (int k = i + 1; k < 8; ++k) v -= L[8 * k + i] * x[k]; x[i] = v; }
}

}  // namespace

edsgpu_status edsgpu_coarse_track(edsgpu_coarse* c, int coarsest_lvl, double R[9], double t[3], double aff_g2l[2], const double ref_aff_g2l[2],
                                  float ref_exposure, float new_exposure, const double min_res_for_abort[5], double last_residuals[5],
                                  double last_flow[3], int* evaluations_out) {
    if (!c) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = c->ctx;
    EDS_REQUIRE(ctx, R && t && aff_g2l && ref_aff_g2l && last_residuals && last_flow, "coarse_track: null argument");
    EDS_REQUIRE(ctx, coarsest_lvl >= 0 && coarsest_lvl < 5 && coarsest_lvl < (int)c->levels.size(), "coarse_track: coarsest level out of range");
    const float setting_coarseCutoffTH = 20.0f;  // settings.cpp:138
    const int maxIterations[5] = {10, 20, 100, 100, 100};
    const float lambdaExtrapolationLimit = 0.001f;
    int evaluations = 0;
    edsgpu_status st = EDSGPU_OK;
    auto eval = [&](int lvl, const double* Rc, const double* tc, const double* affc, float cutoff, double* rs, double* H, double* b) {
        // AffLight::fromToVecExposure, NumType.h:175-187
        float eF = ref_exposure, eT = new_exposure;
        if (eF == 0 || eT == 0) eF = eT = 1;
        const double a = exp(affc[0] - ref_aff_g2l[0]) * eT / eF;
        const float ll[2] = {(float)a, (float)(affc[1] - a * ref_aff_g2l[1])};
        ++evaluations;
        return edsgpu_coarse_calc_res_gs(c, lvl, Rc, tc, ll, (float)ref_aff_g2l[1], cutoff, rs, H, b);
    };
    for (int i = 0; i < 5; ++i) last_residuals[i] = NAN;
    for (int i = 0; i < 3; ++i) last_flow[i] = 1000;
    double Rc[9], tc[3], affc[2] = {aff_g2l[0], aff_g2l[1]};
    memcpy(Rc, R, sizeof(Rc));
    memcpy(tc, t, sizeof(tc));
    bool haveRepeated = false, ok = true;
    for (int lvl = coarsest_lvl; lvl >= 0 && ok; lvl--) {
        double H[64], b[8], resOld[6];
        float levelCutoffRepeat = 1;
        if ((st = eval(lvl, Rc, tc, affc, setting_coarseCutoffTH * levelCutoffRepeat, resOld, H, b)) != EDSGPU_OK) return st;
        while (resOld[5] > 0.6 && levelCutoffRepeat < 50) {
            levelCutoffRepeat *= 2;
            if ((st = eval(lvl, Rc, tc, affc, setting_coarseCutoffTH * levelCutoffRepeat, resOld, H, b)) != EDSGPU_OK) return st;
        }
        float lambda = 0.01f;
        for (int iteration = 0; iteration < maxIterations[lvl]; iteration++) {
            double Hl[64], nb[8], inc[8];
            memcpy(Hl, H, sizeof(Hl));
            for (int i = 0; i < 8; ++i) { Hl[9 * i] *= (1 + lambda); nb[i] = -b[i]; }
            solve8(Hl, nb, inc);  // both affine parameters are optimised (setting_affineOptModeA/B >= 0, settings.cpp:119-120)
            float extrapFac = 1;
            if (lambda < lambdaExtrapolationLimit) extrapFac = sqrtf(sqrtf(lambdaExtrapolationLimit / lambda));
            for (int i = 0; i < 8; ++i) inc[i] *= extrapFac;
            double incScaled[8];
            memcpy(incScaled, inc, sizeof(inc));
            incScaled[6] *= 10.0f;    // SCALE_A (SCALE_XI_ROT = SCALE_XI_TRANS = 1)
            incScaled[7] *= 1000.0f;  // SCALE_B
            double sum = 0;
            for (int i = 0; i < 8; ++i) sum += incScaled[i];
            if (!std::isfinite(sum)) memset(incScaled, 0, sizeof(incScaled));
            double Rn[9], tn[3], affn[2] = {affc[0] + incScaled[6], affc[1] + incScaled[7]};
            memcpy(Rn, Rc, sizeof(Rn));
            memcpy(tn, tc, sizeof(tn));
            se3_left_update(incScaled, Rn, tn);
            double resNew[6], Hn[64], bn[8];  // the fused evaluation already has the system the reference recomputes on accept
            if ((st = eval(lvl, Rn, tn, affn, setting_coarseCutoffTH * levelCutoffRepeat, resNew, Hn, bn)) != EDSGPU_OK) return st;
            const bool accept = (resNew[0] / resNew[1]) < (resOld[0] / resOld[1]);
            if (accept) {
                memcpy(H, Hn, sizeof(H)); memcpy(b, bn, sizeof(b)); memcpy(resOld, resNew, sizeof(resOld));
                affc[0] = affn[0]; affc[1] = affn[1];
                memcpy(Rc, Rn, sizeof(Rc)); memcpy(tc, tn, sizeof(tc));
                lambda *= 0.5f;
            } else {
                lambda *= 4;
                if (lambda < lambdaExtrapolationLimit) lambda = lambdaExtrapolationLimit;
            }
            double norm2 = 0;
            for (int i = 0; i < 8; ++i) norm2 += inc[i] * inc[i];
            if (!(sqrt(norm2) > 1e-3)) break;
        }
        last_residuals[lvl] = sqrtf((float)(resOld[0] / resOld[1]));
        for (int i = 0; i < 3; ++i) last_flow[i] = resOld[2 + i];
        if (min_res_for_abort && last_residuals[lvl] > 1.5 * min_res_for_abort[lvl]) ok = false;
        if (ok && levelCutoffRepeat > 1 && !haveRepeated) { lvl++; haveRepeated = true; }
    }
    if (evaluations_out) *evaluations_out = evaluations;
    if (!ok) return edsgpu_fail(ctx, EDSGPU_NOT_USABLE, "coarse_track: residual above the abort threshold");
    memcpy(R, Rc, sizeof(Rc));
    memcpy(t, tc, sizeof(tc));
    aff_g2l[0] = affc[0]; aff_g2l[1] = affc[1];
    if (fabsf((float)affc[0]) > 1.2f || fabsf((float)affc[1]) > 200.f)  // :683-685
        return edsgpu_fail(ctx, EDSGPU_NOT_USABLE, "coarse_track: affine brightness parameters out of range");
    return EDSGPU_OK;
}

}  // extern "C"
