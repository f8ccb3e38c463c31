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

// halving butterfly: 2*HALF per-lane values -> HALF, lanes with bit OFFSET keep the upper half
template <int HALF, int OFFSET>
__device__ __forceinline__ void halving_step(float* a, unsigned lane) {
    const bool upper = (lane & OFFSET) != 0;
#pragma unroll
    for (int i = 0; i < HALF; ++i) {
        const float send = upper ? a[i] : a[i + HALF];
        const float keep = upper ? a[i + HALF] : a[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFFSET);
    }
}

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

// one CTA's share of an evaluation: points blockIdx.x * blockDim.x + threadIdx.x (+ grid stride) -> partial[CT_NPART]
template <int THREADS>
__device__ __forceinline__ void coarse_sweep(const CoarseArgs& a, double (*red)[CT_NPART], double* __restrict__ partial) {
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
    // warp -> CTA -> one partial per CTA.  The 45 fp32 accumulators go through a halving butterfly (48 shuffles instead of
    // 45 x 5: every step halves the values a lane holds), the seven statistics through plain butterflies; fp64 across warps.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {
        float a48[48];
#pragma unroll
        for (int i = 0; i < CT_NACC; ++i) a48[i] = acc[i];
        // the three free slots carry the counts (small integers: exact in fp32 in any order)
        a48[45] = (float)nE; a48[46] = (float)nW; a48[47] = (float)nSat;
        halving_step<24, 16>(a48, lane);
        halving_step<12, 8>(a48, lane);
        halving_step<6, 4>(a48, lane);
        halving_step<3, 2>(a48, lane);
#pragma unroll
        for (int i = 0; i < 3; ++i) a48[i] += __shfl_xor_sync(0xffffffffu, a48[i], 1);
        if ((lane & 1) == 0) {
            const int b0 = 24 * ((lane >> 4) & 1) + 12 * ((lane >> 3) & 1) + 6 * ((lane >> 2) & 1) + 3 * ((lane >> 1) & 1);
#pragma unroll
            for (int i = 0; i < 3; ++i) red[warp][b0 + i < CT_NACC ? b0 + i : b0 + i + 1] = (double)a48[i];  // slots 45..47 -> nE, nW, nSat behind E
        }
        double st[4] = {E, (double)shT, (double)shRT, (double)nShift};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double s = st[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) red[warp][i == 0 ? CT_NACC : CT_NACC + 3 + i] = s;
        }
    }
    __syncthreads();
    if (threadIdx.x < CT_NPART) {
        double s = 0.0;
        for (int w = 0; w < THREADS / 32; ++w) s += red[w][threadIdx.x];
        partial[threadIdx.x] = s;
    }
}

// One evaluation for the host-synchronous entry point: every CTA leaves its partial, the last one to finish (ticket) adds
// them in CTA order and writes the CT_NPART sums where `out` points -- pinned host memory: one launch, no copy back.
__global__ void __launch_bounds__(CT_THREADS) coarse_res_gs_kernel(CoarseArgs a, unsigned* __restrict__ ticket, double* __restrict__ out) {
    __shared__ double red[CT_THREADS / 32][CT_NPART];
    __shared__ bool last;
    coarse_sweep<CT_THREADS>(a, red, a.partials + (size_t)blockIdx.x * CT_NPART);
    __threadfence();  // this CTA's partial is visible before its ticket
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(ticket, 1u);
        last = (t == gridDim.x - 1u);
        if (last) *ticket = 0u;  // ready for the next launch
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (threadIdx.x < CT_NPART) {
        double s = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(a.partials + (size_t)b * CT_NPART + threadIdx.x);
        out[threadIdx.x] = s;
    }
}

__global__ void coarse_pad_image_kernel(const float* __restrict__ src, float4* __restrict__ dst, size_t npix) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npix) dst[i] = make_float4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 0.f);
}

// ---- CoarseTracker::makeCoarseDepthL0 (CoarseTracker.cpp:127-283) ---------------------------------------------------
__global__ void cd_scatter_kernel(int n, const float* __restrict__ cu, const float* __restrict__ cv, const float* __restrict__ cid,
                                  const float* __restrict__ HdiF, int w, int h, float* __restrict__ idepth, float* __restrict__ wsum) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int u = (int)fa(cu[p], 0.5f), v = (int)fa(cv[p], 0.5f);
    if (u < 0 || v < 0 || u >= w || v >= h) return;  // the reference trusts centerProjectedTo; out-of-image points are dropped here
    const float weight = sqrtf((float)(1e-3 / ((double)HdiF[p] + 1e-12)));
    // two points in one pixel are rare; their sum is taken in the order the atomics arrive (the reference: point order)
    atomicAdd(&idepth[u + w * v], fm(cid[p], weight));
    atomicAdd(&wsum[u + w * v], weight);
}

__global__ void cd_pyr_down_kernel(int wl, int hl, int wlm1, const float* __restrict__ id_in, const float* __restrict__ ws_in,
                                   float* __restrict__ id_out, float* __restrict__ ws_out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= wl || y >= hl) return;
    const int b = 2 * x + 2 * y * wlm1;
    id_out[x + y * wl] = fa(fa(fa(id_in[b], id_in[b + 1]), id_in[b + wlm1]), id_in[b + wlm1 + 1]);
    ws_out[x + y * wl] = fa(fa(fa(ws_in[b], ws_in[b + 1]), ws_in[b + wlm1]), ws_in[b + wlm1 + 1]);
}

// dilation by one pixel where a pixel has no point: diagonal neighbours on levels 0-1, axis neighbours below (:182-233).
// Reads the weights of before the pass (bak); inverse depths are only read where bak > 0 and only written where bak <= 0.
__global__ void cd_dilate_kernel(int wl, int hl, int diagonal, const float* __restrict__ bak, float* __restrict__ idepth, float* __restrict__ wsum) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, total = wl * hl;
    if (i < wl || i >= total - wl) return;
    if (bak[i] > 0) return;
    const int o[4] = {diagonal ? 1 + wl : 1, diagonal ? -1 - wl : -1, diagonal ? wl - 1 : wl, diagonal ? -wl + 1 : -wl};
    float sum = 0, num = 0, numn = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int j = i + o[k];
        if (j >= 0 && j < total && bak[j] > 0) { sum = fa(sum, idepth[j]); num = fa(num, bak[j]); numn = fa(numn, 1.0f); }
    }
    if (numn > 0) { idepth[i] = __fdiv_rn(sum, numn); wsum[i] = __fdiv_rn(num, numn); }
}

// normalisation + scan-line ordered compaction (:236-281), one warp per image row.  PASS 0 counts the row's points, PASS 1
// writes them at the row's offset (and finishes idepth / weightSums like the reference).
template <int PASS>
__global__ void cd_rows_kernel(int wl, int hl, float* __restrict__ idepth, float* __restrict__ wsum, const float* __restrict__ color, int* __restrict__ rows,
                               float* __restrict__ pc_u, float* __restrict__ pc_v, float* __restrict__ pc_id, float* __restrict__ pc_col) {
    const int y = 2 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (y >= hl - 2) return;
    int run = PASS ? rows[y] : 0;
    for (int x0 = 2; x0 < wl - 2; x0 += 32) {
        const int x = x0 + lane, i = x + y * wl;
        bool in = x < wl - 2, have = false, good = false;
        float q = 0.f, c = 0.f;
        if (in) {
            const float ws = wsum[i];
            have = ws > 0;
            if (have) {
                q = __fdiv_rn(idepth[i], ws);
                c = color[i];
                good = isfinite(c) && q > 0;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, good);
        if (PASS) {
            if (good) {
                const int k = run + __popc(m & ((1u << lane) - 1u));
                pc_u[k] = (float)x; pc_v[k] = (float)y; pc_id[k] = q; pc_col[k] = c;
            }
            if (in) {
                idepth[i] = good ? q : -1.0f;
                if (good || !have) wsum[i] = 1.0f;  // the reference's `continue` leaves the weight of a rejected pixel as it is
            }
        }
        run += __popc(m);
    }
    if (!PASS && lane == 0) rows[y] = run;
}

// exclusive scan of the row counts (a few hundred rows: one thread), rows[hl] = number of points of the level
__global__ void cd_row_scan_kernel(int hl, int* __restrict__ rows) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int acc = 0;
    for (int y = 2; y < hl - 2; ++y) { const int c = rows[y]; rows[y] = acc; acc += c; }
    rows[hl] = acc;
}

struct Level {
    int w = 0, h = 0, n = 0, cap = 0;
    float fx = 0, fy = 0, cx = 0, cy = 0;
    float Ki[9] = {0};  // column-major as given
    float4* image = nullptr;
    float* pc = nullptr;  // [4][cap]: u, v, idepth, color
    bool has_image = false;
    // makeCoarseDepthL0: the reference frame's intensity plane and the inverse-depth map under construction
    float* ref_color = nullptr;  // [h*w] lastRef->dIp[lvl][.][0]
    float* depth = nullptr;      // idepth | weightSums | weightSums_bak, h*w each
    int* rows = nullptr;         // per image row: points of the row, then their offset in the level's point cloud; [h] = total
};

}  // namespace

namespace {
__host__ __device__ void coarse_outputs(const double* v, double* rs, double* H, double* b);
}

struct edsgpu_coarse {
    edsgpu_ctx* ctx = nullptr;
    std::vector<Level> levels;
    double* partials = nullptr;  // [max grid][CT_NPART] + CT_NPART result
    int max_grid = 0;
    unsigned* track_bar = nullptr;   // [0] arrival counter of its grid barrier, [1] ticket of the single-evaluation kernel
    unsigned track_bar_count = 0;    // host mirror of the counter; track_bar_dirty: a launch failed, reset both first
    bool track_bar_dirty = false;
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
    if (e == cudaSuccess) e = cudaMalloc(&c->track_bar, 2 * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemsetAsync(c->track_bar, 0, 2 * sizeof(unsigned), ctx->stream);
    if (e != cudaSuccess) { edsgpu_coarse_destroy(c); return edsgpu_fail(ctx, EDSGPU_CUDA_ERROR, cudaGetErrorString(e)); }
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
        if (l.ref_color) cudaFree(l.ref_color);
        if (l.depth) cudaFree(l.depth);
        if (l.rows) cudaFree(l.rows);
    }
    if (c->partials) cudaFree(c->partials);
    if (c->track_bar) cudaFree(c->track_bar);
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
        if (l.ref_color) cudaFree(l.ref_color);
        if (l.depth) cudaFree(l.depth);
        if (l.rows) cudaFree(l.rows);
        l.ref_color = l.depth = nullptr;
        l.rows = nullptr;
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

// lastRef->dIp[lvl] of the reference frame (makeCoarseDepthL0 takes the point colours from it)
edsgpu_status edsgpu_coarse_set_reference_frame(edsgpu_coarse* c, int lvl, const float* dI) {
    if (!c) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = c->ctx;
    EDS_REQUIRE(ctx, lvl >= 0 && lvl < (int)c->levels.size() && dI, "coarse_set_reference_frame: bad arguments");
    Level& l = c->levels[lvl];
    EDS_REQUIRE(ctx, l.w > 0, "coarse_set_reference_frame: call edsgpu_coarse_set_level first");
    DeviceGuard g(ctx->device);
    const size_t npix = (size_t)l.w * l.h;
    if (!l.ref_color) EDS_CUDA(ctx, cudaMalloc(&l.ref_color, sizeof(float) * npix));
    // strided copy of channel 0 of the Vec3f image
    EDS_CUDA(ctx, cudaMemcpy2DAsync(l.ref_color, sizeof(float), dI, 3 * sizeof(float), sizeof(float), npix, cudaMemcpyHostToDevice, ctx->stream));
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EDSGPU_OK;
}

edsgpu_status edsgpu_coarse_make_depth_l0(edsgpu_coarse* c, int levels_used, int n, const float* proj_u, const float* proj_v, const float* proj_idepth,
                                          const float* HdiF, int* pc_n_out) {
    if (!c) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = c->ctx;
    EDS_RANGE("edsgpu_coarse_make_depth_l0");
    EDS_REQUIRE(ctx, levels_used >= 1 && levels_used <= (int)c->levels.size() && n >= 0, "coarse_make_depth_l0: bad arguments");
    EDS_REQUIRE(ctx, n == 0 || (proj_u && proj_v && proj_idepth && HdiF), "coarse_make_depth_l0: null array");
    DeviceGuard g(ctx->device);
    cudaStream_t s = ctx->stream;
    for (int lvl = 0; lvl < levels_used; ++lvl) {
        Level& l = c->levels[lvl];
        EDS_REQUIRE(ctx, l.w > 4 && l.h > 4 && l.ref_color, "coarse_make_depth_l0: every level needs edsgpu_coarse_set_level and edsgpu_coarse_set_reference_frame");
        if (lvl > 0) EDS_REQUIRE(ctx, l.w == c->levels[lvl - 1].w >> 1 && l.h == c->levels[lvl - 1].h >> 1, "coarse_make_depth_l0: levels must halve (makeK)");
        const size_t npix = (size_t)l.w * l.h;
        if (!l.depth) EDS_CUDA(ctx, cudaMalloc(&l.depth, sizeof(float) * 3 * npix));
        if (!l.rows) EDS_CUDA(ctx, cudaMalloc(&l.rows, sizeof(int) * ((size_t)l.h + 1)));
        if (l.cap < (int)npix) {  // the point cloud of a level can hold every pixel
            EDS_CUDA(ctx, cudaStreamSynchronize(s));
            if (l.pc) cudaFree(l.pc);
            l.pc = nullptr; l.cap = 0;
            EDS_CUDA(ctx, cudaMalloc(&l.pc, sizeof(float) * 4 * npix));
            l.cap = (int)npix;
        }
    }
    // the projected points: u | v | idepth | HdiF staged through the context's scratch
    edsgpu_status st = edsgpu_ensure_scratch(ctx, sizeof(float) * 4 * (size_t)std::max(n, 1));
    if (st != EDSGPU_OK) return st;
    float* d = (float*)ctx->scratch;
    const float* src[4] = {proj_u, proj_v, proj_idepth, HdiF};
    for (int k = 0; k < 4 && n > 0; ++k) EDS_CUDA(ctx, cudaMemcpyAsync(d + (size_t)k * n, src[k], sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, s));
    {
        Level& l0 = c->levels[0];
        const size_t npix = (size_t)l0.w * l0.h;
        EDS_CUDA(ctx, cudaMemsetAsync(l0.depth, 0, sizeof(float) * 2 * npix, s));
        if (n > 0) cd_scatter_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, d, d + n, d + 2 * (size_t)n, d + 3 * (size_t)n, l0.w, l0.h, l0.depth, l0.depth + npix);
    }
    for (int lvl = 1; lvl < levels_used; ++lvl) {
        Level& l = c->levels[lvl];
        Level& m = c->levels[lvl - 1];
        const size_t npix = (size_t)l.w * l.h, npm = (size_t)m.w * m.h;
        cd_pyr_down_kernel<<<dim3((l.w + 127) / 128, l.h), 128, 0, s>>>(l.w, l.h, m.w, m.depth, m.depth + npm, l.depth, l.depth + npix);
    }
    for (int lvl = 0; lvl < levels_used; ++lvl) {
        Level& l = c->levels[lvl];
        const size_t npix = (size_t)l.w * l.h;
        float *id = l.depth, *ws = l.depth + npix, *bak = l.depth + 2 * npix;
        EDS_CUDA(ctx, cudaMemcpyAsync(bak, ws, sizeof(float) * npix, cudaMemcpyDeviceToDevice, s));
        cd_dilate_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, s>>>(l.w, l.h, lvl < 2 ? 1 : 0, bak, id, ws);
        const int rows = l.h - 4, rpb = 8;  // eight warps (rows) per CTA
        float *pu = l.pc, *pv = l.pc + l.cap, *pi = l.pc + 2 * (size_t)l.cap, *pcol = l.pc + 3 * (size_t)l.cap;
        cd_rows_kernel<0><<<(rows + rpb - 1) / rpb, 32 * rpb, 0, s>>>(l.w, l.h, id, ws, l.ref_color, l.rows, pu, pv, pi, pcol);
        cd_row_scan_kernel<<<1, 32, 0, s>>>(l.h, l.rows);
        cd_rows_kernel<1><<<(rows + rpb - 1) / rpb, 32 * rpb, 0, s>>>(l.w, l.h, id, ws, l.ref_color, l.rows, pu, pv, pi, pcol);
        ctx->launches += 4;
    }
    EDS_CUDA(ctx, cudaGetLastError());
    st = edsgpu_ensure_pinned(ctx, sizeof(int) * (size_t)levels_used);
    if (st != EDSGPU_OK) return st;
    int* hn = (int*)ctx->pinned;
    for (int lvl = 0; lvl < levels_used; ++lvl)
        EDS_CUDA(ctx, cudaMemcpyAsync(hn + lvl, c->levels[lvl].rows + c->levels[lvl].h, sizeof(int), cudaMemcpyDeviceToHost, s));
    EDS_CUDA(ctx, cudaStreamSynchronize(s));
    for (int lvl = 0; lvl < levels_used; ++lvl) {
        c->levels[lvl].n = hn[lvl];  // pc_n[lvl]: the level is ready for calcRes / trackNewestCoarse
        if (pc_n_out) pc_n_out[lvl] = hn[lvl];
    }
    return EDSGPU_OK;
}

edsgpu_status edsgpu_coarse_get_reference(edsgpu_coarse* c, int lvl, int* n_out, float* pc_u, float* pc_v, float* pc_idepth, float* pc_color) {
    if (!c) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = c->ctx;
    EDS_REQUIRE(ctx, lvl >= 0 && lvl < (int)c->levels.size(), "coarse_get_reference: bad level");
    DeviceGuard g(ctx->device);
    const Level& l = c->levels[lvl];
    if (n_out) *n_out = l.n;
    float* dst[4] = {pc_u, pc_v, pc_idepth, pc_color};
    for (int k = 0; k < 4; ++k)
        if (dst[k] && l.n > 0) EDS_CUDA(ctx, cudaMemcpyAsync(dst[k], l.pc + (size_t)k * l.cap, sizeof(float) * (size_t)l.n, cudaMemcpyDeviceToHost, ctx->stream));
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EDSGPU_OK;
}

edsgpu_status edsgpu_coarse_calc_res_gs(edsgpu_coarse* c, int lvl, const double R[9], const double t[3], const float affLL[2], float b0,
                                        float cutoffTH, double rs[6], double H[64], double b[8]) {
    EDS_RANGE("edsgpu_coarse_calc_res_gs");
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
    double* out_dev = nullptr;  // the pinned block as the device sees it
    EDS_CUDA(ctx, cudaHostGetDevicePointer((void**)&out_dev, ctx->pinned, 0));
    coarse_res_gs_kernel<<<grid, CT_THREADS, 0, ctx->stream>>>(a, c->track_bar + 1, out_dev);
    ctx->launches++;
    EDS_CUDA(ctx, cudaGetLastError());
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const double* v = (const double*)ctx->pinned;
    coarse_outputs(v, rs, H, b);
    return EDSGPU_OK;
}

}  // extern "C"

// ---- trackNewestCoarse ---------------------------------------------------------------------------------------
namespace {

// (R, t) <- SE3::exp(inc) * (R, t), inc = [translation part, rotation part].  Sophus SO3::exp (quaternion form) for the
// rotation; sin / cos of the full angle in the translation's V matrix come from the half-angle pair by the double-angle
// identities (one sincos instead of four libm calls on the single thread that runs this).
__device__ void se3_left_update(const double* inc6, double* R, double* t) {
    const double* u = inc6;
    const double* w = inc6 + 3;
    double Re[9], V[9];
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2);
    const bool small = th < 1e-10;
    double imag, real, A = 0.0, B = 0.0;
    if (small) {
        const double th4 = th2 * th2;
        imag = 0.5 - th2 / 48.0 + th4 / 3840.0;
        real = 1.0 - th2 / 8.0 + th4 / 384.0;
    } else {
        double sh, ch;
        sincos(0.5 * th, &sh, &ch);
        const double ith = 1.0 / th;
        imag = sh * ith;
        real = ch;
        const double ith2 = ith * ith;
        A = 2.0 * sh * sh * ith2;                  // (1 - cos th) / th^2
        B = (th - 2.0 * sh * ch) * (ith2 * ith);  // (th - sin th) / th^3
    }
    {
        const double x = imag * w[0], y = imag * w[1], z = imag * w[2], q = real;
        Re[0] = 1 - 2 * (y * y + z * z); Re[1] = 2 * (x * y - z * q); Re[2] = 2 * (x * z + y * q);
        Re[3] = 2 * (x * y + z * q); Re[4] = 1 - 2 * (x * x + z * z); Re[5] = 2 * (y * z - x * q);
        Re[6] = 2 * (x * z - y * q); Re[7] = 2 * (y * z + x * q); Re[8] = 1 - 2 * (x * x + y * y);
    }
    const double W[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    if (small) {
        for (int i = 0; i < 9; ++i) V[i] = Re[i];
    } else {
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double w2 = 0;
#pragma unroll
                for (int k = 0; k < 3; ++k) w2 += W[3 * r + k] * W[3 * k + c];
                V[3 * r + c] = (r == c ? 1.0 : 0.0) + A * W[3 * r + c] + B * w2;
            }
    }
    double te[3], Rn[9], tn[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) te[r] = V[3 * r] * u[0] + V[3 * r + 1] * u[1] + V[3 * r + 2] * u[2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) Rn[3 * r + c] = Re[3 * r] * R[c] + Re[3 * r + 1] * R[3 + c] + Re[3 * r + 2] * R[6 + c];
        tn[r] = Re[3 * r] * t[0] + Re[3 * r + 1] * t[1] + Re[3 * r + 2] * t[2] + te[r];
    }
    for (int i = 0; i < 9; ++i) R[i] = Rn[i];
    for (int i = 0; i < 3; ++i) t[i] = tn[i];
}

// ------------------------------------------------------------------------------------------
// trackNewestCoarse on the device: ONE cooperative launch runs the whole coarse-to-fine loop.  Every evaluation is a
// sweep of the level's points by all CTAs, one grid barrier, then every CTA adds the per-CTA partials in order and its
// first warp advances the (identical) state machine: Levenberg damping, 8x8 LDL^T (a matrix row per lane), SE3::exp update,
// accept / reject, cutoff repeat, level change (the scalar decisions are lane 0's).  No host round trip inside the loop (it cost ~40 us per evaluation).
// ------------------------------------------------------------------------------------------
constexpr int CT_MAX_PER = 19;  // ceil(148 / 8): CTAs per eighth in the ordered sum of the partials
constexpr int CT_MAX_LEVELS = 5, CT_TRACK_THREADS = 512;  // wide CTAs: fewer of them at the grid barrier and in the ordered sum  // PYR_LEVELS the loop can visit (coarsest_lvl < 5, CoarseTracker.cpp:540)

struct LevelDev {
    int w, h, n;
    float fx, fy, cx, cy;
    float Ki[9];  // column-major as given
    const float4* image;
    const float *u, *v, *idepth, *color;
};

struct TrackIO {  // global memory, in-out
    double R[9], t[3], aff[2];
    double ref_aff[2], min_res[5];
    double last_residuals[5], last_flow[3];
    float ref_exposure, new_exposure;
    int has_min_res, coarsest;
    int evaluations, status;  // status: 0 = tracked, 1 = residual above the abort threshold, 2 = affine parameters out of range
};

struct TrackArgs {
    LevelDev lv[CT_MAX_LEVELS];
    TrackIO in;        // the request travels as a kernel parameter: no upload
    TrackIO* io;       // the result: pinned host memory the device writes directly, no copy back
    double* partials;  // [2][gridDim.x][CT_NPART], double-buffered by evaluation
    unsigned* bar;     // arrival counter of the grid barrier
    unsigned bar_base; // its value when this launch starts (the host keeps count: one barrier per evaluation)
};

struct TrackCtl {  // shared memory; written by the CTA's first warp only
    double Rc[9], tc[3], affc[2];   // accepted pose
    double Rn[9], tn[3], affn[2];   // proposal under evaluation
    double H[64], b[8], resOld[6], inc[8];
    double Hn[64], bn[8];           // the system of the evaluation just reduced
    float lambda, repeat;
    int lvl, iteration, state, haveRepeated, ok, evaluations, done, take, action;
    CoarseArgs a;                   // the evaluation every thread sweeps next
    double out[8];                  // lastResiduals (5), lastFlowIndicators (3): every CTA keeps its own copy, CTA 0 writes them out
};
enum { TS_FIRST = 0, TS_ITER = 1 };

// grid barrier (cooperative launch: every CTA is resident): one arrival counter that only grows, the k-th barrier of a
// launch is passed when it reaches base + (k + 1) * gridDim.x.  One release-add and one polled word per CTA; the writes
// of the CTA's other threads are ordered before the add by the __syncthreads (causality order), no separate fences.
__device__ __forceinline__ void track_grid_barrier(unsigned* bar, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        unsigned cur;
        do {  // spin: a handful of CTAs poll one word, the wait is a few microseconds at most
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(bar) : "memory");
        } while ((int)(cur - target) < 0);
    }
    __syncthreads();
}

// the evaluation request of (lvl, R, t, aff): what edsgpu_coarse_calc_res_gs sets up on the host
__device__ void track_request(TrackCtl& c, const TrackArgs& A, const TrackIO& io, int lvl, const double* R, const double* t, const double* aff) {
    const LevelDev& l = A.lv[lvl];
    CoarseArgs& a = c.a;
    a.lvl = lvl; a.wl = l.w; a.hl = l.h; a.n = l.n;
    a.fxl = l.fx; a.fyl = l.fy; a.cxl = l.cx; a.cyl = l.cy;
    for (int i = 0; i < 3; ++i) {
        a.t[i] = (float)t[i];
        for (int j = 0; j < 3; ++j) {
            a.Ki[3 * i + j] = l.Ki[3 * j + i];
            a.RKi[3 * i + j] = fa(fa(fm((float)R[3 * i + 0], l.Ki[3 * j + 0]), fm((float)R[3 * i + 1], l.Ki[3 * j + 1])), fm((float)R[3 * i + 2], l.Ki[3 * j + 2]));
        }
    }
    // AffLight::fromToVecExposure, NumType.h:175-187
    float eF = io.ref_exposure, eT = io.new_exposure;
    if (eF == 0 || eT == 0) eF = eT = 1;
    const double e = exp(aff[0] - io.ref_aff[0]) * eT / eF;
    a.aff0 = (float)e;
    a.aff1 = (float)(aff[1] - e * io.ref_aff[1]);
    a.b0 = (float)io.ref_aff[1];
    a.cutoffTH = 20.0f * c.repeat;  // setting_coarseCutoffTH, settings.cpp:138
    a.image = l.image;
    a.pc_u = l.u; a.pc_v = l.v; a.pc_idepth = l.idepth; a.pc_color = l.color;
    a.partials = nullptr;
    c.evaluations++;
}

// reduced sums -> calcRes' Vec6 and calcGSSSE's H, b (same arithmetic as the host entry point)
__host__ __device__ void coarse_outputs(const double* v, double* rs, double* H, double* b) {
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
}

// x = A^-1 rhs, A symmetric positive definite 8x8, by LDL^T (the reference calls Eigen's ldlt().solve): one matrix row per
// lane (lanes 0..7 of a warp; all 32 lanes must call).  A column costs one dependent chain instead of one per row, and one reciprocal instead of a division per entry.
__device__ __forceinline__ double solve8_warp(const double* row, double rhs, int lane) {
    double l[8], ut[8];  // l: this lane's row of L; ut: this lane's row of L^T (its column of L), collected for the back substitution
    double dinv_mine = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { l[i] = 0.0; ut[i] = 0.0; }
    double D[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        double v = row[j];
#pragma unroll
        for (int k = 0; k < j; ++k) {
            const double ljk = __shfl_sync(0xffffffffu, l[k], j);
            v -= l[k] * ljk * D[k];
        }
        const double d = __shfl_sync(0xffffffffu, v, j);
        D[j] = d;
        const double dinv = 1.0 / d;
        if (lane == j) dinv_mine = dinv;
        const double lij = lane > j ? v * dinv : (lane == j ? 1.0 : 0.0);
        l[j] = lij;
#pragma unroll
        for (int i = j + 1; i < 8; ++i) {
            const double t = __shfl_sync(0xffffffffu, lij, i);
            if (lane == j) ut[i] = t;
        }
    }
    double y = rhs;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const double yk = __shfl_sync(0xffffffffu, y, k);
        if (lane > k) y -= l[k] * yk;
    }
    double x = y * dinv_mine;
#pragma unroll
    for (int k = 7; k > 0; --k) {
        const double xk = __shfl_sync(0xffffffffu, x, k);
        if (lane < k) x -= ut[k] * xk;
    }
    return x;
}

// Warp 0 of every CTA: consume the evaluation that has just been reduced into `sums`, decide what happens next
// (CoarseTracker.cpp:540-701).  The scalar decisions are lane 0's; forming the 8x8 system from the sums, keeping it on an
// accepted step and the damped solve are spread over the lanes.
__device__ void track_advance(TrackCtl& c, const TrackArgs& A, const TrackIO& io, double* out, const double* sums, int lane) {
    const int maxIterations[5] = {10, 20, 100, 100, 100};
    const float lambdaExtrapolationLimit = 0.001f;
    enum { ACT_REPEAT = 0, ACT_PROPOSE = 1, ACT_LEVEL_END = 2 };
    // H_out = acc.H.topLeftCorner<8,8>().cast<double>() * (1.0f / n), n padded to a multiple of 4 (:466-478, :332-344): coarse_outputs, entry by entry
    {
        const long long npad = ((long long)sums[CT_NACC + 2] + 3) / 4 * 4;
        const float inv_n = 1.0f / (float)npad;
        for (int idx = lane; idx < 72; idx += 32) {
            const int r = idx < 64 ? idx >> 3 : idx - 64, q = idx < 64 ? idx & 7 : 8;
            const int lo = min(r, q), hi = max(r, q);
            const int e = lo * 9 - lo * (lo - 1) / 2 + (hi - lo);  // index of (lo, hi) in the row-major upper triangle of the 9x9
            const double sr = r == 6 ? 10.0f : (r == 7 ? 1000.0f : 1.0), sq = q == 6 ? 10.0f : (q == 7 ? 1000.0f : 1.0);
            const double v = (double)(float)sums[e] * inv_n;
            if (idx < 64) c.Hn[idx] = v * (sr * sq);
            else c.bn[r] = v * sr;
        }
    }
    if (lane == 0) {
        double rs[6];
        coarse_outputs(sums, rs, nullptr, nullptr);
        bool level_end = false;
        c.take = 0;
        if (c.state == TS_FIRST) {
            for (int i = 0; i < 6; ++i) c.resOld[i] = rs[i];
            c.take = 1;
            if (c.resOld[5] > 0.6 && c.repeat < 50) {
                c.repeat *= 2;
                c.action = ACT_REPEAT;
            } else {
                c.lambda = 0.01f;
                c.iteration = 0;
                c.action = ACT_PROPOSE;  // maxIterations >= 10
            }
        } else {
            const bool accept = (rs[0] / rs[1]) < (c.resOld[0] / c.resOld[1]);
            if (accept) {
                c.take = 1;
                for (int i = 0; i < 6; ++i) c.resOld[i] = rs[i];
                c.affc[0] = c.affn[0]; c.affc[1] = c.affn[1];
                for (int i = 0; i < 9; ++i) c.Rc[i] = c.Rn[i];
                for (int i = 0; i < 3; ++i) c.tc[i] = c.tn[i];
                c.lambda *= 0.5f;
            } else {
                c.lambda *= 4;
                if (c.lambda < lambdaExtrapolationLimit) c.lambda = lambdaExtrapolationLimit;
            }
            double norm2 = 0;
            for (int i = 0; i < 8; ++i) norm2 += c.inc[i] * c.inc[i];
            if (!(sqrt(norm2) > 1e-3)) level_end = true;
            c.iteration++;
            c.action = (!level_end && c.iteration < maxIterations[c.lvl]) ? ACT_PROPOSE : ACT_LEVEL_END;
        }
    }
    __syncwarp();
    if (c.take) {
        for (int idx = lane; idx < 72; idx += 32) {
            if (idx < 64) c.H[idx] = c.Hn[idx];
            else c.b[idx - 64] = c.bn[idx - 64];
        }
        __syncwarp();
    }
    const int action = c.action;
    if (action == ACT_PROPOSE) {
        // propose a step: both affine parameters are optimised (setting_affineOptModeA/B >= 0, settings.cpp:119-120)
        double row[8];
        const int r = lane & 7;
#pragma unroll
        for (int q = 0; q < 8; ++q) row[q] = c.H[8 * r + q] * (q == r ? (double)(1 + c.lambda) : 1.0);
        const double x = solve8_warp(row, -c.b[r], lane);
        float extrapFac = 1;
        if (c.lambda < lambdaExtrapolationLimit) extrapFac = sqrtf(sqrtf(lambdaExtrapolationLimit / c.lambda));
        if (lane < 8) c.inc[lane] = x * extrapFac;
        __syncwarp();
    }
    if (lane != 0) return;
    if (action == ACT_REPEAT) {
        track_request(c, A, io, c.lvl, c.Rc, c.tc, c.affc);
        return;
    }
    if (action == ACT_PROPOSE) {
        double incScaled[8];
        for (int i = 0; i < 8; ++i) incScaled[i] = c.inc[i];
        incScaled[6] *= 10.0f;    // SCALE_A (SCALE_XI_ROT = SCALE_XI_TRANS = 1)
        incScaled[7] *= 1000.0f;  // SCALE_B
        double sum = 0;
        for (int i = 0; i < 8; ++i) sum += incScaled[i];
        if (!isfinite(sum))
            for (int i = 0; i < 8; ++i) incScaled[i] = 0;
        c.affn[0] = c.affc[0] + incScaled[6];
        c.affn[1] = c.affc[1] + incScaled[7];
        for (int i = 0; i < 9; ++i) c.Rn[i] = c.Rc[i];
        for (int i = 0; i < 3; ++i) c.tn[i] = c.tc[i];
        se3_left_update(incScaled, c.Rn, c.tn);
        c.state = TS_ITER;
        track_request(c, A, io, c.lvl, c.Rn, c.tn, c.affn);  // the fused evaluation already has the system the reference recomputes on accept
        return;
    }
    // the level is finished
    out[c.lvl] = sqrtf((float)(c.resOld[0] / c.resOld[1]));
    for (int i = 0; i < 3; ++i) out[5 + i] = c.resOld[2 + i];
    if (io.has_min_res && out[c.lvl] > 1.5 * io.min_res[c.lvl]) c.ok = 0;
    if (c.ok && c.repeat > 1 && !c.haveRepeated) { c.lvl++; c.haveRepeated = 1; }
    c.lvl--;
    if (c.lvl < 0 || !c.ok) { c.done = 1; return; }
    c.repeat = 1;
    c.state = TS_FIRST;
    track_request(c, A, io, c.lvl, c.Rc, c.tc, c.affc);
}

__global__ void __launch_bounds__(CT_TRACK_THREADS) coarse_track_kernel(TrackArgs A) {
    __shared__ double red[CT_TRACK_THREADS / 32][CT_NPART];
    __shared__ double sums[CT_NPART];
    __shared__ TrackCtl c;
    const TrackIO& io = A.in;  // inputs are only read inside the loop; CTA 0 writes the outputs once, at the end
    if (threadIdx.x == 0) {
        // every CTA runs the same state machine on the same numbers
        for (int i = 0; i < 9; ++i) c.Rc[i] = io.R[i];
        for (int i = 0; i < 3; ++i) c.tc[i] = io.t[i];
        c.affc[0] = io.aff[0]; c.affc[1] = io.aff[1];
        c.lvl = io.coarsest; c.haveRepeated = 0; c.ok = 1; c.evaluations = 0; c.done = 0;
        c.repeat = 1; c.state = TS_FIRST; c.lambda = 0.01f; c.iteration = 0;
        for (int i = 0; i < 5; ++i) c.out[i] = io.last_residuals[i];
        for (int i = 0; i < 3; ++i) c.out[5 + i] = io.last_flow[i];
        track_request(c, A, io, c.lvl, c.Rc, c.tc, c.affc);
    }
    __syncthreads();
    for (unsigned k = 0;; ++k) {
        double* buf = A.partials + (size_t)(k & 1u) * gridDim.x * CT_NPART;
        coarse_sweep<CT_TRACK_THREADS>(c.a, red, buf + (size_t)blockIdx.x * CT_NPART);
        track_grid_barrier(A.bar, A.bar_base + (k + 1u) * gridDim.x);
        // ordered sum over the CTAs: eight threads per entry take an eighth of them each (fixed partition, loads issued
        // together), then the eighths are added in order within the eight lanes
        if (threadIdx.x < 8 * CT_NPART) {
            const int e = threadIdx.x >> 3, q = threadIdx.x & 7;
            const unsigned per = (gridDim.x + 7u) / 8u, b0 = q * per;
            double v[CT_MAX_PER];
#pragma unroll
            for (unsigned i = 0; i < CT_MAX_PER; ++i) v[i] = (i < per && b0 + i < gridDim.x) ? __ldcg(buf + (size_t)(b0 + i) * CT_NPART + e) : 0.0;
            double s = 0.0;
#pragma unroll
            for (unsigned i = 0; i < CT_MAX_PER; ++i)
                if (i < per) s += v[i];
            // lanes 8g .. 8g+7 hold the eighths of one entry: ((((((e0 + e1) + e2) + e3) + e4) + e5) + e6) + e7
            const unsigned base = threadIdx.x & 24u;
            double t = __shfl_sync(0xffffffffu, s, base);
#pragma unroll
            for (unsigned i = 1; i < 8; ++i) t += __shfl_sync(0xffffffffu, s, base + i);
            if (q == 0) sums[e] = t;
        }
        __syncthreads();
        if (threadIdx.x < 32) track_advance(c, A, io, c.out, sums, threadIdx.x);
        __syncthreads();
        if (c.done) break;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        TrackIO& o = *A.io;
        o.evaluations = c.evaluations;
        for (int i = 0; i < 5; ++i) o.last_residuals[i] = c.out[i];
        for (int i = 0; i < 3; ++i) o.last_flow[i] = c.out[5 + i];
        int status = 0;
        if (!c.ok) status = 1;
        else {
            for (int i = 0; i < 9; ++i) o.R[i] = c.Rc[i];
            for (int i = 0; i < 3; ++i) o.t[i] = c.tc[i];
            o.aff[0] = c.affc[0]; o.aff[1] = c.affc[1];
            if (fabsf((float)c.affc[0]) > 1.2f || fabsf((float)c.affc[1]) > 200.f) status = 2;  // :683-685
        }
        o.status = status;
    }
}

}  // namespace

extern "C" {

edsgpu_status edsgpu_coarse_track(edsgpu_coarse* c, int coarsest_lvl, double R[9], double t[3], double aff_g2l[2], const double ref_aff_g2l[2],
                                  float ref_exposure, float new_exposure, const double min_res_for_abort[5], double last_residuals[5],
                                  double last_flow[3], int* evaluations_out) {
    EDS_RANGE("edsgpu_coarse_track");
    if (!c) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = c->ctx;
    EDS_REQUIRE(ctx, R && t && aff_g2l && ref_aff_g2l && last_residuals && last_flow, "coarse_track: null argument");
    EDS_REQUIRE(ctx, coarsest_lvl >= 0 && coarsest_lvl < CT_MAX_LEVELS && coarsest_lvl < (int)c->levels.size(), "coarse_track: coarsest level out of range");
    DeviceGuard g(ctx->device);
    TrackArgs A{};
    int nmax = 1;
    for (int lvl = 0; lvl <= coarsest_lvl; ++lvl) {
        const Level& l = c->levels[lvl];
        EDS_REQUIRE(ctx, l.has_image && l.n > 0, "coarse_track: every level needs a reference point cloud and a new frame");
        LevelDev& d = A.lv[lvl];
        d.w = l.w; d.h = l.h; d.n = l.n; d.fx = l.fx; d.fy = l.fy; d.cx = l.cx; d.cy = l.cy;
        memcpy(d.Ki, l.Ki, sizeof(d.Ki));
        d.image = l.image;
        d.u = l.pc; d.v = l.pc + l.cap; d.idepth = l.pc + 2 * (size_t)l.cap; d.color = l.pc + 3 * (size_t)l.cap;
        nmax = std::max(nmax, l.n);
    }
    edsgpu_status st = edsgpu_ensure_pinned(ctx, sizeof(TrackIO));
    if (st != EDSGPU_OK) return st;
    TrackIO* h = (TrackIO*)ctx->pinned;
    TrackIO& in = A.in;
    memcpy(in.R, R, sizeof(in.R)); memcpy(in.t, t, sizeof(in.t));
    in.aff[0] = aff_g2l[0]; in.aff[1] = aff_g2l[1];
    in.ref_aff[0] = ref_aff_g2l[0]; in.ref_aff[1] = ref_aff_g2l[1];
    in.ref_exposure = ref_exposure; in.new_exposure = new_exposure;
    in.has_min_res = min_res_for_abort ? 1 : 0;
    if (min_res_for_abort) memcpy(in.min_res, min_res_for_abort, sizeof(in.min_res));
    in.coarsest = coarsest_lvl;
    for (int i = 0; i < 5; ++i) in.last_residuals[i] = NAN;
    for (int i = 0; i < 3; ++i) in.last_flow[i] = 1000;
    *h = in;  // a failed track leaves R, t, aff as they came
    EDS_CUDA(ctx, cudaHostGetDevicePointer((void**)&A.io, h, 0));
    A.partials = c->partials;
    A.bar = c->track_bar;
    // one CTA per 512 points of the largest level, at most one per SM (cooperative launch: all resident)
    const int grid = std::max(1, std::min((nmax + CT_TRACK_THREADS - 1) / CT_TRACK_THREADS, std::min(ctx->num_sms, c->max_grid / 2)));
    if (c->track_bar_dirty) {
        EDS_CUDA(ctx, cudaMemsetAsync(c->track_bar, 0, 2 * sizeof(unsigned), ctx->stream));
        c->track_bar_count = 0;
        c->track_bar_dirty = false;
    }
    A.bar_base = c->track_bar_count;
    void* argv[] = {(void*)&A};
    c->track_bar_dirty = true;  // until the launch is known to have run to its end
    EDS_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)coarse_track_kernel, dim3(grid), dim3(CT_TRACK_THREADS), argv, 0, ctx->stream));
    ctx->launches++;
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    c->track_bar_count += (unsigned)h->evaluations * (unsigned)grid;  // one barrier per evaluation
    c->track_bar_dirty = false;
    for (int i = 0; i < 5; ++i) last_residuals[i] = h->last_residuals[i];
    for (int i = 0; i < 3; ++i) last_flow[i] = h->last_flow[i];
    if (evaluations_out) *evaluations_out = h->evaluations;
    if (h->status == 1) return edsgpu_fail(ctx, EDSGPU_NOT_USABLE, "coarse_track: residual above the abort threshold");
    memcpy(R, h->R, sizeof(h->R));
    memcpy(t, h->t, sizeof(h->t));
    aff_g2l[0] = h->aff[0]; aff_g2l[1] = h->aff[1];
    if (h->status == 2) return edsgpu_fail(ctx, EDSGPU_NOT_USABLE, "coarse_track: affine brightness parameters out of range");
    return EDSGPU_OK;
}

}  // extern "C"
