// Event window -> brightness-increment event frame on the device.
//
// Replaces EventFrame::create (reference src/tracking/EventFrame.cpp:302-389) and
// utils::drawValuesPoints (src/utils/Utils.cpp:50-122) for pyramid level 0:
//   k1  scatter_events_kernel   LUT gather + bilinear/nn vote x exp time weight, accumulated
//                               with warp-aggregated 64-bit fixed-point atomics (value * 2^40):
//                               deterministic, exact for the integer (nn, unweighted) mode,
//                               2^-41 absolute quantisation otherwise.
//   k2  blur_norm_kernel        3x3 separable Gaussian (BORDER_REFLECT_101) in fp64 through a
//                               shared-memory tile, fp32 frame out, fused sum of squares;
//                               the last CTA of each frame reduces the per-tile partials in a
//                               fixed order and publishes norm and 1/norm.
// The frame is stored un-normalised in fp32; consumers multiply by 1/norm when sampling.
#include <stdlib.h>

#include "common.cuh"
#include "frames.cuh"

namespace {

constexpr double kQ = 1099511627776.0;  // 2^40
// blur: a warp owns a strip of BLUR_COLS output columns (its two outer lanes only carry the halo columns) by BLUR_ROWS rows
constexpr int BLUR_THREADS = 256, BLUR_WARPS = BLUR_THREADS / 32, BLUR_COLS = 30, BLUR_ROWS = 64;
__host__ __device__ constexpr int blur_strips_x(int W) { return (W + BLUR_COLS - 1) / BLUR_COLS; }
__host__ __device__ constexpr int blur_strips_y(int H) { return (H + BLUR_ROWS - 1) / BLUR_ROWS; }
__host__ __device__ constexpr int blur_ctas(int H, int W) { return (blur_strips_x(W) * blur_strips_y(H) + BLUR_WARPS - 1) / BLUR_WARPS; }

// Utils.hpp:542-546 with idx = i / E, window_size = 1 (Utils.cpp:72).
__device__ __forceinline__ double exp_weight(int i, int E) {
    double value = ((double)i / (double)(unsigned)E - 0.5) / (1.0 / 6.0);
    return exp(-0.5 * value * value);
}

// One 64-bit atomic per distinct key in the warp: lanes with equal `key` are merged.
// `val` is the fixed-point vote; `scaled` optional extra per-corner weights are applied by
// the caller (all lanes of a group share them, see scatter_events_kernel).
__device__ __forceinline__ long long warp_aggregate(unsigned active, unsigned key, long long val, bool* is_leader) {
    const unsigned lane = threadIdx.x & 31;
    unsigned peers = __match_any_sync(active, key);
    int leader = __ffs(peers) - 1;
    *is_leader = ((int)lane == leader);
    // as many rounds as the largest group of the warp has members (usually 1-3): in round k every lane fetches the value of
    // its k-th peer, in ascending lane order (the order of the sum is fixed)
    const int rounds = __reduce_max_sync(active, __popc(peers));
    if (rounds <= 1) return val;
    long long sum = 0;
    unsigned m = peers;
    for (int k = 0; k < rounds; ++k) {
        const int src = m ? __ffs(m) - 1 : (int)lane;
        const long long o = __shfl_sync(active, val, src);
        if (m) { sum += o; m &= m - 1u; }
    }
    return sum;
}

// exp_weight(i, E) for i < E: depends on the window length only, so it is tabulated once per E instead of evaluated (a double
// division and a double exp) for every event of every window
__global__ void exp_table_kernel(int E, double* __restrict__ tab) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < E) tab[i] = exp_weight(i, E);
}

__global__ void __launch_bounds__(256) scatter_events_kernel(const uint16_t* __restrict__ ex, const uint16_t* __restrict__ ey,
                                                             const uint8_t* __restrict__ epol, int E, int H, int W,
                                                             const float* __restrict__ mapx, const float* __restrict__ mapy,
                                                             int mode, const double* __restrict__ exp_tab, long long* __restrict__ acc) {
    const int win = blockIdx.y;
    const size_t ebase = (size_t)win * E;
    long long* img = acc + (size_t)win * H * W;
    const int stride = gridDim.x * blockDim.x;
    // warp-uniform trip count so that every lane reaches the warp collectives
    const int first = blockIdx.x * blockDim.x + threadIdx.x;
    const int iters = (E + stride - 1) / stride;
    for (int it = 0; it < iters; ++it) {
        const int i = first + it * stride;
        const bool live = i < E;
        const unsigned active = __ballot_sync(0xffffffffu, live);
        if (!live) continue;
        const int x = __ldcs(ex + ebase + i), y = __ldcs(ey + ebase + i);  // streaming: every event is read once
        const bool inside = (x < W) && (y < H);
        const double pol = __ldcs(epol + ebase + i) ? 1.0 : -1.0;  // EventFrame.cpp:318
        const double wt = exp_tab ? __ldg(exp_tab + i) : 1.0;  // Utils.hpp:542-546
        double ux = x, uy = y;
        if (mapx != nullptr && inside) {  // EventFrame.cpp:316-317
            ux = (double)mapx[(size_t)y * W + x];
            uy = (double)mapy[(size_t)y * W + x];
        }
        // events from the same raw pixel share coordinates, hence all weights: merge them
        const unsigned key = inside ? (unsigned)(y * W + x) : 0xffffffffu;
        bool leader;
        if (mode == EDSGPU_DRAW_NN) {
            // cv::Point2i(Point2d) rounds half to even (cvRound), then clip: Utils.cpp:75-78
            int xi = __double2int_rn(ux), yi = __double2int_rn(uy);
            xi = max(0, min(xi, W - 1));
            yi = max(0, min(yi, H - 1));
            long long v = inside ? __double2ll_rn(wt * pol * kQ) : 0;
            v = warp_aggregate(active, key, v, &leader);
            if (leader && inside && v != 0) atomicAdd((unsigned long long*)&img[(size_t)yi * W + xi], (unsigned long long)v);
        } else {
            // Utils.cpp:85-106
            const double fx0 = floor(ux), fy0 = floor(uy);
            int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
            const double ax = ux - fx0, ay = uy - fy0;  // == ux - x0, exact
            const double bx = (double)x1 - ux, by = (double)y1 - uy;
            const bool x0in = x0 >= 0 && x0 < W, x1in = x1 >= 0 && x1 < W;
            const bool y0in = y0 >= 0 && y0 < H, y1in = y1 >= 0 && y1 < H;
            const double wa = (x0in && y0in) ? bx * by : 0.0;
            const double wb = (x0in && y1in) ? bx * ay : 0.0;
            const double wc = (x1in && y0in) ? ax * by : 0.0;
            const double wd = (x1in && y1in) ? ax * ay : 0.0;
            x0 = max(0, min(x0, W - 1)); x1 = max(0, min(x1, W - 1));
            y0 = max(0, min(y0, H - 1)); y1 = max(0, min(y1, H - 1));
            // merge the time-weight * polarity of same-pixel events, then apply the 4 corner weights
            const double s = inside ? wt * pol : 0.0;
            unsigned peers = __match_any_sync(active, key);
            leader = ((int)(threadIdx.x & 31) == __ffs(peers) - 1);
            double sum = s;
            const int rounds = __reduce_max_sync(active, __popc(peers));
            if (rounds > 1) {  // see warp_aggregate
                sum = 0.0;
                unsigned m = peers;
                for (int k = 0; k < rounds; ++k) {
                    const int src = m ? __ffs(m) - 1 : (int)(threadIdx.x & 31);
                    const double o = __shfl_sync(active, s, src);
                    if (m) { sum += o; m &= m - 1u; }
                }
            }
            if (leader && inside) {
                long long va = __double2ll_rn(sum * wa * kQ), vb = __double2ll_rn(sum * wb * kQ);
                long long vc = __double2ll_rn(sum * wc * kQ), vd = __double2ll_rn(sum * wd * kQ);
                if (va) atomicAdd((unsigned long long*)&img[(size_t)y0 * W + x0], (unsigned long long)va);
                if (vb) atomicAdd((unsigned long long*)&img[(size_t)y1 * W + x0], (unsigned long long)vb);
                if (vc) atomicAdd((unsigned long long*)&img[(size_t)y0 * W + x1], (unsigned long long)vc);
                if (vd) atomicAdd((unsigned long long*)&img[(size_t)y1 * W + x1], (unsigned long long)vd);
            }
        }
    }
}

// accumulator clear with streaming stores (evict-first in L2): the zeros are consumed by the scatter and the blur of the same
// build and never again
__global__ void clear_acc_kernel(ulonglong2* __restrict__ p, size_t n16, long long* __restrict__ tail) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) __stcs(p + i, make_ulonglong2(0ull, 0ull));
    if (tail && blockIdx.x == 0 && threadIdx.x == 0) *tail = 0;
}

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    if (i < 0) return -i;
    if (i >= n) return 2 * n - 2 - i;
    return i;
}

// 3x3 separable Gaussian (cv::GaussianBlur, BORDER_REFLECT_101, Utils.cpp:113-119) + sum of squares (cv::norm L2,
// EventFrame.cpp:360-364).  One thread per column of a strip, walking down the rows: the accumulator is read ONCE per
// pixel (coalesced 64-bit loads), the row pass takes its neighbours from the adjacent lanes by shuffle, the column pass
// keeps the last three row-filtered values in registers -- no shared-memory tile, no CTA barrier in the sweep.  fp64 in
// OpenCV's operation order (explicit mul / add, no FMA contraction): bit-exact against cv::GaussianBlur (CV_64F) on exactly
// representable input.
// OUT64 == false: write the fp32 frame + per-CTA sum of squares, the last CTA of a window publishes norm and 1/norm.
// OUT64 == true : write out64 = blurred * scale (debug / host read-back path).
template <bool OUT64>
__global__ void __launch_bounds__(BLUR_THREADS) blur_norm_kernel(const long long* __restrict__ acc, int H, int W, double k0, double k1,
                                                                 const cudaSurfaceObject_t* __restrict__ surfs, double* __restrict__ partials,
                                                                 unsigned* __restrict__ tickets, double* __restrict__ norms,
                                                                 int first_slot, int first_acc, double* __restrict__ out64,
                                                                 const double* __restrict__ scale_ptr) {
    __shared__ double wsum[BLUR_WARPS];
    __shared__ bool is_last;
    const int win = blockIdx.z;
    const int slot = first_slot + win;
    const long long* img = acc + (size_t)(first_acc + win) * H * W;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler too: the loops below get uniform trip counts
    const int sx = blur_strips_x(W), strip = blockIdx.x * BLUR_WARPS + warp;
    double sq = 0.0;
    {
        // a warp beyond the last strip (tail of the last CTA) repeats the last strip and stores nothing: no branch around the
        // shuffles, so they need no divergence handling
        const int nstrips = sx * blur_strips_y(H);
        const bool live = strip < nstrips;
        const int st = live ? strip : nstrips - 1;
        const int x = (st % sx) * BLUR_COLS + lane - 1;             // this lane's column; lanes 0 and 31 are halo columns
        const int y0 = (st / sx) * BLUR_ROWS, y1 = min(y0 + BLUR_ROWS, H);
        const int xs = reflect101(min(max(x, -1), W), W);           // source column (reflected at the border, clamped for idle lanes)
        const bool out_col = live && lane >= 1 && lane <= BLUR_COLS && x < W;
        const double scale = OUT64 ? (scale_ptr ? scale_ptr[2 * slot + 1] : 1.0) : 1.0;
        cudaSurfaceObject_t surf = 0;
        if (!OUT64) surf = surfs[slot];
        double up = 0.0, mid = 0.0;  // row-filtered values of rows y - 2 and y - 1
        constexpr int RB = 8;        // rows in flight: their loads are issued together, ahead of the arithmetic
        for (int yb = y0 - 1; yb <= y1; yb += RB) {
            double cs[RB];
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                const int ys = reflect101(min(yb + j, min(y1, H)), H);  // warp-uniform; rows past the strip repeat its last one (unused)
                cs[j] = (double)__ldcs(img + (size_t)ys * W + xs);  // streaming: read once, must not push the tracker's working set out of L2
            }
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                // (rows past the strip, in its last chunk only, run through the arithmetic too and store nothing: a uniform
                // trip count keeps the shuffles free of divergence handling)
                const int y = yb + j;
                const double c = cs[j] * (1.0 / kQ);
                const double l = __shfl_up_sync(0xffffffffu, c, 1), r = __shfl_down_sync(0xffffffffu, c, 1);
                // row pass, generic cv::RowFilter order: left, centre, right
                const double down = __dadd_rn(__dadd_rn(__dmul_rn(l, k0), __dmul_rn(c, k1)), __dmul_rn(r, k0));
                if (y > y0 && y <= y1 && out_col) {
                    // cv::SymmColumnFilter order: centre, then k*(up+down); the value of output row y - 1
                    const double v = __dadd_rn(__dmul_rn(k1, mid), __dmul_rn(k0, __dadd_rn(up, down)));
                    if (OUT64) {
                        out64[(size_t)win * H * W + (size_t)(y - 1) * W + x] = v * scale;
                    } else {
                        surf2Dwrite((float)v, surf, x * (int)sizeof(float), y - 1);
                        sq += v * v;
                    }
                }
                up = mid;
                mid = down;
            }
        }
    }
    if constexpr (!OUT64) {
        // fused sum of squares: warp -> CTA -> ordered final pass over the CTAs of the window
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (lane == 0) wsum[warp] = sq;
        __syncthreads();
        const int nctas = gridDim.x;
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < BLUR_WARPS; ++w) s += wsum[w];
            partials[(size_t)win * nctas + blockIdx.x] = s;
            __threadfence();
            unsigned t = atomicAdd(&tickets[win], 1u);
            is_last = (t == (unsigned)nctas - 1);
        }
        __syncthreads();
        if (is_last && tid < 32) {
            __threadfence();
            const volatile double* p = partials + (size_t)win * nctas;
            double s = 0.0;
            for (int i = tid; i < nctas; i += 32) s += p[i];
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (tid == 0) {
                double nrm = sqrt(s);
                norms[2 * slot] = nrm;
                norms[2 * slot + 1] = 1.0 / nrm;
                tickets[win] = 0;
            }
        }
    }
}

// Pyramid level i >= 1 of EventFrame::create (EventFrame.cpp:349-357): cv::dilate + cv::erode of the level-0 frame with a
// (2i+1) x (2i+1) rectangle, same resolution, and the level's own L2 norm (:360-364).  The default border of cv::dilate /
// cv::erode leaves pixels outside the image out of the maximum / minimum; clamp-to-edge sampling does the same (the
// replicated edge texel is already inside every clipped window).  Separable: row maxima / minima of a tile with halo in
// shared memory, then columns.  The level is stored like level 0: fp32, un-normalised, {norm, 1/norm} beside it.
constexpr int MORPH_TILE = 32, MORPH_THREADS = 256, MORPH_MAX_R = 4;
__host__ __device__ constexpr int morph_ctas_x(int W) { return (W + MORPH_TILE - 1) / MORPH_TILE; }
__host__ __device__ constexpr int morph_ctas_y(int H) { return (H + MORPH_TILE - 1) / MORPH_TILE; }

__global__ void __launch_bounds__(MORPH_THREADS) morph_level_kernel(const cudaTextureObject_t* __restrict__ tex0, const cudaSurfaceObject_t* __restrict__ surfs,
                                                                    int H, int W, int R, int first_slot, double* __restrict__ partials,
                                                                    unsigned* __restrict__ tickets, double* __restrict__ norms) {
    constexpr int TS = MORPH_TILE + 2 * MORPH_MAX_R;
    __shared__ float tile[TS][TS + 1];
    __shared__ float rmax[TS][MORPH_TILE + 1], rmin[TS][MORPH_TILE + 1];
    __shared__ double wsum[MORPH_THREADS / 32];
    __shared__ bool is_last;
    const int win = blockIdx.z, slot = first_slot + win, tid = threadIdx.x;
    const int x0 = blockIdx.x * MORPH_TILE, y0 = blockIdx.y * MORPH_TILE;
    const int span = MORPH_TILE + 2 * R;
    const cudaTextureObject_t src = tex0[slot];
    for (int i = tid; i < span * span; i += MORPH_THREADS) {
        const int ly = i / span, lx = i - ly * span;
        tile[ly][lx] = tex2D<float>(src, (float)(x0 + lx - R) + 0.5f, (float)(y0 + ly - R) + 0.5f);
    }
    __syncthreads();
    for (int i = tid; i < span * MORPH_TILE; i += MORPH_THREADS) {
        const int ly = i / MORPH_TILE, lx = i % MORPH_TILE;
        float mx = tile[ly][lx], mn = mx;
        for (int k = 1; k <= 2 * R; ++k) { const float v = tile[ly][lx + k]; mx = fmaxf(mx, v); mn = fminf(mn, v); }
        rmax[ly][lx] = mx;
        rmin[ly][lx] = mn;
    }
    __syncthreads();
    double sq = 0.0;
    for (int i = tid; i < MORPH_TILE * MORPH_TILE; i += MORPH_THREADS) {
        const int ly = i / MORPH_TILE, lx = i % MORPH_TILE;
        const int gx = x0 + lx, gy = y0 + ly;
        if (gx < W && gy < H) {
            float mx = rmax[ly][lx], mn = rmin[ly][lx];
            for (int k = 1; k <= 2 * R; ++k) { mx = fmaxf(mx, rmax[ly + k][lx]); mn = fminf(mn, rmin[ly + k][lx]); }
            const float v = mx + mn;
            surf2Dwrite(v, surfs[slot], gx * (int)sizeof(float), gy);
            sq += (double)v * (double)v;
        }
    }
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if ((tid & 31) == 0) wsum[tid >> 5] = sq;
    __syncthreads();
    const int nctas = gridDim.x * gridDim.y, cta = blockIdx.y * gridDim.x + blockIdx.x;
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < MORPH_THREADS / 32; ++w) s += wsum[w];
        partials[(size_t)win * nctas + cta] = s;
        __threadfence();
        const unsigned t = atomicAdd(&tickets[win], 1u);
        is_last = (t == (unsigned)nctas - 1);
    }
    __syncthreads();
    if (is_last && tid < 32) {
        __threadfence();
        const volatile double* p = partials + (size_t)win * nctas;
        double s = 0.0;
        for (int i = tid; i < nctas; i += 32) s += p[i];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (tid == 0) {
            const double nrm = sqrt(s);
            norms[2 * slot] = nrm;
            norms[2 * slot + 1] = 1.0 / nrm;
            tickets[win] = 0;
        }
    }
}

// Unused dynamic shared memory requested by the blur launch (EDSGPU_BUILD_PAD_KB tunes it, 0 = none).  Measured on the
// benchmark step, where the build of the next window runs beside the batched solve: with the 40-register blur kernel and no
// dynamic shared memory the step takes 0.699 ms, with 1 / 4 / 8-25 / 50 / 100 KB per CTA 0.679 / 0.668 / 0.660 / 0.669 /
// 0.688 ms, although the kernel alone runs equally fast either way; padding the clear and scatter launches as well loses
// 0.010 ms again.  The request changes the shared-memory / L1 split the driver configures on
// the SMs that alternate between build CTAs and the solve's CTAs; asking for the maximum carveout everywhere instead costs
// the solve its L1 (0.72-0.75 ms).
size_t build_pad_bytes() {
    static long kb = -1;
    if (kb < 0) {
        const char* e = getenv("EDSGPU_BUILD_PAD_KB");
        kb = e ? std::max(0, atoi(e)) : 16;
        if (kb > 40) kb = 40;
    }
    return (size_t)kb * 1024;
}

edsgpu_status launch_frames(edsgpu_ctx* ctx, edsgpu_frames* fr, int first_slot, int count, const edsgpu_lut* lut,
                            const uint16_t* x_dev, const uint16_t* y_dev, const uint8_t* pol_dev, int E, int mode, int use_exp, float sigma) {
    EDS_RANGE("edsgpu_event_frame_create (launch)");
    const int H = fr->H, W = fr->W;
    const size_t npix = (size_t)H * W;
    cudaStream_t bs = fr->build_stream;
    {   // wait for the last readers of the slots about to be overwritten
        cudaEvent_t prev = nullptr;
        for (int s = first_slot; s < first_slot + count; ++s) {
            cudaEvent_t e = fr->slot_read[s];
            if (e && e != prev) EDS_CUDA(ctx, cudaStreamWaitEvent(bs, e, 0));
            prev = e;
        }
    }
    double k0 = 0.0, k1 = 1.0;
    if (sigma > 0.f) {  // cv::getGaussianKernel(3, sigma), Utils.cpp:113-119
        double s = (double)sigma;
        k0 = exp(-1.0 / (2.0 * s * s));
        double sum = k0 + 1.0 + k0;
        k0 /= sum;
        k1 = 1.0 / sum;
    }
    fr->k0 = k0;
    fr->k1 = k1;
    if (use_exp && fr->exp_E != E) {  // the time weights of a window of E events
        if (fr->exp_table) { EDS_CUDA(ctx, cudaStreamSynchronize(bs)); cudaFree(fr->exp_table); fr->exp_table = nullptr; fr->exp_E = 0; }
        EDS_CUDA(ctx, cudaMalloc(&fr->exp_table, sizeof(double) * (size_t)E));
        exp_table_kernel<<<(E + 255) / 256, 256, 0, bs>>>(E, fr->exp_table);
        ctx->launches++;
        EDS_CUDA(ctx, cudaGetLastError());
        fr->exp_E = E;
    }
    const bool ring = fr->acc_slots < fr->capacity;
    for (int c0 = 0; c0 < count; c0 += fr->acc_slots) {
        // one chunk of windows: clear -> scatter -> blur on accumulators that stay in L2 (see frames.cuh)
        const int n = std::min(fr->acc_slots, count - c0);
        const int acc0 = ring ? 0 : first_slot + c0;
        if (ring)
            for (int s = 0; s < fr->capacity; ++s)
                if (fr->slot_acc[s] >= 0 && fr->slot_acc[s] < n) fr->slot_acc[s] = -1;
        for (int w = 0; w < n; ++w) fr->slot_acc[first_slot + c0 + w] = acc0 + w;
        if (getenv("EDSGPU_CLEAR_MEMSET") || (((uintptr_t)(fr->acc + (size_t)acc0 * npix)) & 15u)) {
            EDS_CUDA(ctx, cudaMemsetAsync(fr->acc + (size_t)acc0 * npix, 0, sizeof(long long) * npix * n, bs));
        } else {
            const size_t n16 = npix * n / 2, tail = (npix * n) & 1;  // 16-byte stores (the block is 256-byte aligned)
            const int blocks = (int)std::min<size_t>((n16 + 255) / 256, (size_t)8 * ctx->num_sms);
            clear_acc_kernel<<<std::max(1, blocks), 256, 0, bs>>>(reinterpret_cast<ulonglong2*>(fr->acc + (size_t)acc0 * npix), n16,
                                                                   tail ? fr->acc + (size_t)acc0 * npix + 2 * n16 : nullptr);
            ctx->launches++;
            EDS_CUDA(ctx, cudaGetLastError());
        }
        {
            int threads = 256;
            int bx = (E + threads - 1) / threads;
            bx = max(1, min(bx, 4 * ctx->num_sms));
            dim3 grid(bx, n);
            const size_t eoff = (size_t)c0 * E;
            scatter_events_kernel<<<grid, threads, 0, bs>>>(x_dev + eoff, y_dev + eoff, pol_dev + eoff, E, H, W, lut ? lut->mapx : nullptr,
                                                            lut ? lut->mapy : nullptr, mode, use_exp ? fr->exp_table : nullptr,
                                                            fr->acc + (size_t)acc0 * npix);
            ctx->launches++;
            EDS_CUDA(ctx, cudaGetLastError());
        }
        {
            dim3 grid(blur_ctas(H, W), 1, n);
            const size_t pad = build_pad_bytes();
            blur_norm_kernel<false><<<grid, BLUR_THREADS, pad, bs>>>(fr->acc, H, W, k0, k1, fr->surf_dev, fr->partials, fr->tickets, fr->norms,
                                                                    first_slot + c0, acc0, nullptr, nullptr);
            ctx->launches++;
            EDS_CUDA(ctx, cudaGetLastError());
        }
    }
    for (int lvl = 1; lvl < fr->levels; ++lvl) {
        dim3 grid(morph_ctas_x(W), morph_ctas_y(H), count);
        const size_t base = (size_t)lvl * fr->capacity;
        morph_level_kernel<<<grid, MORPH_THREADS, 0, bs>>>(fr->tex_dev, fr->surf_dev + base, H, W, lvl, first_slot, fr->partials, fr->tickets,
                                                           fr->norms + 2 * base);
        ctx->launches++;
        EDS_CUDA(ctx, cudaGetLastError());
    }
    {   // readers of these slots wait for this build
        cudaEvent_t e = fr->built_pool[fr->built_next];
        fr->built_next = (fr->built_next + 1) % edsgpu_frames::kEventPool;
        EDS_CUDA(ctx, cudaEventRecord(e, bs));
        for (int s = first_slot; s < first_slot + count; ++s) fr->slot_built[s] = e;
    }
    return EDSGPU_OK;
}

// blurred image of one slot as doubles (scaled by 1/norm if normalised) into ctx scratch, then to host
edsgpu_status read_image64(edsgpu_ctx* ctx, const edsgpu_frames* fr, int slot, bool normalised, double* host_out) {
    const int H = fr->H, W = fr->W;
    const size_t npix = (size_t)H * W;
    EDS_REQUIRE(ctx, fr->slot_acc[slot] >= 0, "frames_read: the accumulator of this slot was never filled or has been recycled (large sets of slots share a ring of accumulators)");
    edsgpu_status st = edsgpu_ensure_scratch(ctx, npix * sizeof(double));
    if (st == EDSGPU_OK) st = edsgpu_frames_wait_built(fr, slot, 1, ctx->stream);
    if (st != EDSGPU_OK) return st;
    dim3 grid(blur_ctas(H, W), 1, 1);
    blur_norm_kernel<true><<<grid, BLUR_THREADS, 0, ctx->stream>>>(fr->acc, H, W, fr->k0, fr->k1, nullptr, nullptr, nullptr, nullptr, slot,
                                                                    fr->slot_acc[slot], (double*)ctx->scratch, normalised ? fr->norms : nullptr);
    ctx->launches++;
    EDS_CUDA(ctx, cudaGetLastError());
    EDS_CUDA(ctx, cudaMemcpyAsync(host_out, ctx->scratch, npix * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    st = edsgpu_frames_mark_read(fr, slot, 1, ctx->stream);
    if (st != EDSGPU_OK) return st;
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EDSGPU_OK;
}

}  // namespace

edsgpu_status edsgpu_frames_wait_built(const edsgpu_frames* fr, int first, int count, cudaStream_t stream) {
    cudaEvent_t prev = nullptr;
    for (int s = first; s < first + count; ++s) {
        cudaEvent_t e = fr->slot_built[s];
        if (e && e != prev) EDS_CUDA(fr->ctx, cudaStreamWaitEvent(stream, e, 0));
        prev = e;
    }
    return EDSGPU_OK;
}

edsgpu_status edsgpu_frames_mark_read(const edsgpu_frames* cfr, int first, int count, cudaStream_t stream) {
    edsgpu_frames* fr = const_cast<edsgpu_frames*>(cfr);  // bookkeeping only: the images are not touched
    cudaEvent_t e = fr->read_pool[fr->read_next];
    fr->read_next = (fr->read_next + 1) % edsgpu_frames::kEventPool;
    EDS_CUDA(fr->ctx, cudaEventRecord(e, stream));
    for (int s = first; s < first + count; ++s) fr->slot_read[s] = e;
    return EDSGPU_OK;
}

extern "C" {

edsgpu_status edsgpu_lut_create(edsgpu_ctx* ctx, int height, int width, const float* fwd_mapx, const float* fwd_mapy, edsgpu_lut** out) {
    if (!ctx || !out) return EDSGPU_INVALID_ARGUMENT;
    EDS_REQUIRE(ctx, height > 0 && width > 0, "lut_create: bad size");
    EDS_REQUIRE(ctx, (fwd_mapx == nullptr) == (fwd_mapy == nullptr), "lut_create: give both maps or neither");
    DeviceGuard g(ctx->device);
    edsgpu_lut* lut = new edsgpu_lut();
    lut->ctx = ctx; lut->H = height; lut->W = width;
    if (fwd_mapx) {
        size_t bytes = sizeof(float) * (size_t)height * width;
        cudaError_t e = cudaMalloc(&lut->mapx, bytes);
        if (e == cudaSuccess) e = cudaMalloc(&lut->mapy, bytes);
        if (e == cudaSuccess) e = cudaMemcpyAsync(lut->mapx, fwd_mapx, bytes, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(lut->mapy, fwd_mapy, bytes, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { edsgpu_lut_destroy(lut); return edsgpu_fail(ctx, EDSGPU_CUDA_ERROR, cudaGetErrorString(e)); }
    }
    *out = lut;
    return EDSGPU_OK;
}

void edsgpu_lut_destroy(edsgpu_lut* lut) {
    if (!lut) return;
    DeviceGuard g(lut->ctx->device);
    if (lut->mapx) cudaFree(lut->mapx);
    if (lut->mapy) cudaFree(lut->mapy);
    delete lut;
}

edsgpu_status edsgpu_frames_create(edsgpu_ctx* ctx, int height, int width, int capacity, edsgpu_frames** out) {
    return edsgpu_frames_create_pyramid(ctx, height, width, capacity, 1, out);
}

edsgpu_status edsgpu_frames_create_pyramid(edsgpu_ctx* ctx, int height, int width, int capacity, int num_levels, edsgpu_frames** out) {
    if (!ctx || !out) return EDSGPU_INVALID_ARGUMENT;
    EDS_REQUIRE(ctx, height > 0 && width > 0 && capacity > 0, "frames_create: bad size");
    EDS_REQUIRE(ctx, (size_t)height * width < (1ull << 31), "frames_create: image too large");
    EDS_REQUIRE(ctx, num_levels >= 1 && num_levels <= MORPH_MAX_R + 1, "frames_create: num_levels must be in [1,5]");
    DeviceGuard g(ctx->device);
    edsgpu_frames* fr = new edsgpu_frames();
    fr->ctx = ctx; fr->H = height; fr->W = width; fr->capacity = capacity; fr->levels = num_levels;
    const size_t npix = (size_t)height * width;
    const int ntiles = std::max(blur_ctas(height, width), morph_ctas_x(width) * morph_ctas_y(height));
    const int nimg = capacity * num_levels;  // one fp32 image per (level, slot)
    {
        // one accumulator per slot while they all fit the budget, otherwise a ring that stays resident in L2
        // EDSGPU_ACC_RING_MB: size of the ring; 0 (the default) = always one accumulator per slot.  Measured on the 64-window
        // 640x480 benchmark step: the ring keeps the accumulators out of HBM but its chunked launches cost more than they save
        // (0.750 ms per step with a 32 MB ring against 0.733 ms without).
        size_t ring_mb = 0;
        if (const char* e = getenv("EDSGPU_ACC_RING_MB")) ring_mb = (size_t)std::max(0, atoi(e));
        const size_t per = sizeof(long long) * npix;
        fr->acc_slots = (ring_mb == 0 || (size_t)capacity * per <= (ring_mb << 20)) ? capacity
                                                                                    : (int)std::max<size_t>(1, std::min<size_t>(capacity, (ring_mb << 20) / per));
        fr->slot_acc.assign(capacity, -1);
    }
    cudaError_t e = cudaMalloc(&fr->acc, sizeof(long long) * npix * fr->acc_slots);
    fr->arrays = new cudaArray_t[nimg]();
    fr->tex = new cudaTextureObject_t[nimg]();
    fr->surf = new cudaSurfaceObject_t[nimg]();
    float* zeros = nullptr;
    if (e == cudaSuccess) e = cudaMalloc(&zeros, sizeof(float) * npix);
    if (e == cudaSuccess) e = cudaMemsetAsync(zeros, 0, sizeof(float) * npix, ctx->stream);
    for (int i = 0; i < nimg && e == cudaSuccess; ++i) {
        const cudaChannelFormatDesc fmt = cudaCreateChannelDesc<float>();
        e = cudaMallocArray(&fr->arrays[i], &fmt, width, height, cudaArraySurfaceLoadStore | cudaArrayTextureGather);
        cudaResourceDesc res{};
        res.resType = cudaResourceTypeArray;
        res.res.array.array = fr->arrays[i];
        cudaTextureDesc td{};
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModePoint;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 0;
        if (e == cudaSuccess) e = cudaCreateTextureObject(&fr->tex[i], &res, &td, nullptr);
        if (e == cudaSuccess) e = cudaCreateSurfaceObject(&fr->surf[i], &res);
        if (e == cudaSuccess)
            e = cudaMemcpy2DToArrayAsync(fr->arrays[i], 0, 0, zeros, sizeof(float) * width, sizeof(float) * width, height,
                                         cudaMemcpyDeviceToDevice, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaMalloc(&fr->surf_dev, sizeof(cudaSurfaceObject_t) * nimg);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(fr->surf_dev, fr->surf, sizeof(cudaSurfaceObject_t) * nimg, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMalloc(&fr->tex_dev, sizeof(cudaTextureObject_t) * nimg);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(fr->tex_dev, fr->tex, sizeof(cudaTextureObject_t) * nimg, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMalloc(&fr->partials, sizeof(double) * (size_t)ntiles * capacity);
    if (e == cudaSuccess) e = cudaMalloc(&fr->tickets, sizeof(unsigned) * capacity);
    if (e == cudaSuccess) e = cudaMalloc(&fr->norms, sizeof(double) * 2 * nimg);
    if (e == cudaSuccess) e = cudaMemsetAsync(fr->tickets, 0, sizeof(unsigned) * capacity, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(fr->norms, 0, sizeof(double) * 2 * nimg, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(fr->acc, 0, sizeof(long long) * npix * fr->acc_slots, ctx->stream);
    fr->slot_built.assign(capacity, nullptr);
    fr->slot_read.assign(capacity, nullptr);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&fr->build_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&fr->order_evt, cudaEventDisableTiming);
    for (int i = 0; i < edsgpu_frames::kEventPool && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&fr->built_pool[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&fr->read_pool[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&fr->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&fr->copied, cudaEventDisableTiming);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&fr->stage_free[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (zeros) cudaFree(zeros);
    if (e != cudaSuccess) {
        edsgpu_frames_destroy(fr);
        return edsgpu_fail(ctx, e == cudaErrorMemoryAllocation ? EDSGPU_OUT_OF_MEMORY : EDSGPU_CUDA_ERROR, cudaGetErrorString(e));
    }
    ctx->frames_list.push_back(fr);
    *out = fr;
    return EDSGPU_OK;
}

void edsgpu_frames_destroy(edsgpu_frames* fr) {
    if (!fr) return;
    DeviceGuard g(fr->ctx->device);
    cudaStreamSynchronize(fr->ctx->stream);
    {
        auto& list = fr->ctx->frames_list;
        for (size_t i = 0; i < list.size(); ++i)
            if (list[i] == fr) { list.erase(list.begin() + i); break; }
    }
    if (fr->build_stream) { cudaStreamSynchronize(fr->build_stream); cudaStreamDestroy(fr->build_stream); }
    if (fr->order_evt) cudaEventDestroy(fr->order_evt);
    for (int i = 0; i < edsgpu_frames::kEventPool; ++i) {
        if (fr->built_pool[i]) cudaEventDestroy(fr->built_pool[i]);
        if (fr->read_pool[i]) cudaEventDestroy(fr->read_pool[i]);
    }
    if (fr->acc) cudaFree(fr->acc);
    for (int i = 0; i < fr->capacity * fr->levels && fr->arrays; ++i) {
        if (fr->tex[i]) cudaDestroyTextureObject(fr->tex[i]);
        if (fr->surf[i]) cudaDestroySurfaceObject(fr->surf[i]);
        if (fr->arrays[i]) cudaFreeArray(fr->arrays[i]);
    }
    delete[] fr->arrays;
    delete[] fr->tex;
    delete[] fr->surf;
    if (fr->exp_table) cudaFree(fr->exp_table);
    if (fr->surf_dev) cudaFree(fr->surf_dev);
    if (fr->tex_dev) cudaFree(fr->tex_dev);
    if (fr->partials) cudaFree(fr->partials);
    if (fr->tickets) cudaFree(fr->tickets);
    if (fr->norms) cudaFree(fr->norms);
    if (fr->copy_stream) { cudaStreamSynchronize(fr->copy_stream); cudaStreamDestroy(fr->copy_stream); }
    if (fr->copied) cudaEventDestroy(fr->copied);
    for (int i = 0; i < 2; ++i) {
        if (fr->stage_free[i]) cudaEventDestroy(fr->stage_free[i]);
        if (fr->events_dev[i]) cudaFree(fr->events_dev[i]);
    }
    delete fr;
}

static edsgpu_status check_batch_args(edsgpu_ctx* ctx, edsgpu_frames* frames, int first_slot, int count, const edsgpu_lut* lut,
                                      int num_events, int mode) {
    EDS_REQUIRE(ctx, frames != nullptr && frames->ctx == ctx, "event_frame: frames belong to another context");
    EDS_REQUIRE(ctx, count > 0 && first_slot >= 0 && first_slot + count <= frames->capacity, "event_frame: slot range out of bounds");
    EDS_REQUIRE(ctx, num_events > 0, "event_frame: num_events must be positive");
    EDS_REQUIRE(ctx, mode == EDSGPU_DRAW_NN || mode == EDSGPU_DRAW_BILINEAR, "event_frame: unknown mode");
    EDS_REQUIRE(ctx, !lut || (lut->H == frames->H && lut->W == frames->W), "event_frame: LUT size mismatch");
    return EDSGPU_OK;
}

edsgpu_status edsgpu_event_frame_create_batch_dev(edsgpu_ctx* ctx, edsgpu_frames* frames, int first_slot, int count, const edsgpu_lut* lut,
                                                  const uint16_t* x_dev, const uint16_t* y_dev, const uint8_t* polarity_dev, int num_events,
                                                  int mode, int use_exp_weights, float sigma) {
    if (!ctx) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_status st = check_batch_args(ctx, frames, first_slot, count, lut, num_events, mode);
    if (st != EDSGPU_OK) return st;
    EDS_REQUIRE(ctx, x_dev && y_dev && polarity_dev, "event_frame: null event arrays");
    DeviceGuard g(ctx->device);
    // the device arrays were produced by work the caller queued on the context's stream
    EDS_CUDA(ctx, cudaEventRecord(frames->order_evt, ctx->stream));
    EDS_CUDA(ctx, cudaStreamWaitEvent(frames->build_stream, frames->order_evt, 0));
    return launch_frames(ctx, frames, first_slot, count, lut, x_dev, y_dev, polarity_dev, num_events, mode, use_exp_weights, sigma);
}

edsgpu_status edsgpu_event_frame_create_batch(edsgpu_ctx* ctx, edsgpu_frames* frames, int first_slot, int count, const edsgpu_lut* lut,
                                              const uint16_t* x, const uint16_t* y, const uint8_t* polarity, int num_events, int mode,
                                              int use_exp_weights, float sigma, double* norms_out) {
    if (!ctx) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_status st = check_batch_args(ctx, frames, first_slot, count, lut, num_events, mode);
    if (st != EDSGPU_OK) return st;
    EDS_REQUIRE(ctx, x && y && polarity, "event_frame: null event arrays");
    DeviceGuard g(ctx->device);
    // device staging for the events: [x u16 | y u16 | pol u8] per batch, double-buffered.  The copies run
    // on the frames' copy stream and only wait for the kernels that last READ the same staging buffer,
    // so the H2D transfer of this batch overlaps whatever the other streams are still doing (e.g. the LM
    // solve of the previous batch); the build stream then waits for the copy.
    const size_t n = (size_t)count * num_events;
    const size_t off_y = align_up(n * 2, 256), off_p = off_y + align_up(n * 2, 256), total = off_p + align_up(n, 256);
    const int buf = (frames->stage_idx ^= 1);
    if (frames->events_bytes[buf] < total) {
        EDS_CUDA(ctx, cudaStreamSynchronize(frames->build_stream));
        EDS_CUDA(ctx, cudaStreamSynchronize(frames->copy_stream));
        if (frames->events_dev[buf]) cudaFree(frames->events_dev[buf]);
        frames->events_dev[buf] = nullptr;
        frames->events_bytes[buf] = 0;
        EDS_CUDA(ctx, cudaMalloc(&frames->events_dev[buf], total));
        frames->events_bytes[buf] = total;
    }
    char* d = (char*)frames->events_dev[buf];
    cudaStream_t cs = frames->copy_stream;
    EDS_CUDA(ctx, cudaStreamWaitEvent(cs, frames->stage_free[buf], 0));
    // the caller's buffers are the copy source (pinned for a truly asynchronous copy) and must stay
    // unchanged until the next synchronising call
    EDS_CUDA(ctx, cudaMemcpyAsync(d, x, n * 2, cudaMemcpyHostToDevice, cs));
    EDS_CUDA(ctx, cudaMemcpyAsync(d + off_y, y, n * 2, cudaMemcpyHostToDevice, cs));
    EDS_CUDA(ctx, cudaMemcpyAsync(d + off_p, polarity, n, cudaMemcpyHostToDevice, cs));
    EDS_CUDA(ctx, cudaEventRecord(frames->copied, cs));
    EDS_CUDA(ctx, cudaStreamWaitEvent(frames->build_stream, frames->copied, 0));
    st = launch_frames(ctx, frames, first_slot, count, lut, (const uint16_t*)d, (const uint16_t*)(d + off_y), (const uint8_t*)(d + off_p),
                       num_events, mode, use_exp_weights, sigma);
    if (st != EDSGPU_OK) return st;
    EDS_CUDA(ctx, cudaEventRecord(frames->stage_free[buf], frames->build_stream));
    if (norms_out) {
        st = edsgpu_ensure_pinned(ctx, sizeof(double) * 2 * count);
        if (st == EDSGPU_OK) st = edsgpu_frames_wait_built(frames, first_slot, count, ctx->stream);
        if (st != EDSGPU_OK) return st;
        EDS_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, frames->norms + 2 * first_slot, sizeof(double) * 2 * count, cudaMemcpyDeviceToHost, ctx->stream));
        EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < count; ++i) norms_out[i] = ((double*)ctx->pinned)[2 * i];
    }
    return EDSGPU_OK;
}

edsgpu_status edsgpu_event_frame_create(edsgpu_ctx* ctx, edsgpu_frames* frames, int slot, const edsgpu_lut* lut, const uint16_t* x,
                                        const uint16_t* y, const uint8_t* polarity, const int64_t* ts_us, int num_events, int mode,
                                        int use_exp_weights, float sigma, double* norm_out, int64_t* time_us_out, int64_t* delta_time_us_out,
                                        double* host_frame_out) {
    if (!ctx) return EDSGPU_INVALID_ARGUMENT;
    EDS_REQUIRE(ctx, num_events > 0, "event_frame: num_events must be positive");
    if (ts_us) {
        // EventFrame.cpp:319-336: first/last/middle timestamps, throw if first > last
        int64_t first = ts_us[0], last = ts_us[num_events - 1];
        if (first > last) return edsgpu_fail(ctx, EDSGPU_NON_MONOTONIC_TIME, "[EVENT_FRAME] event time[0] > event time[N-1]");
        if (time_us_out) *time_us_out = ts_us[num_events / 2];
        if (delta_time_us_out) *delta_time_us_out = last - first;
    }
    double nrm = 0.0;
    edsgpu_status st = edsgpu_event_frame_create_batch(ctx, frames, slot, 1, lut, x, y, polarity, num_events, mode, use_exp_weights, sigma,
                                                       (norm_out || host_frame_out) ? &nrm : nullptr);
    if (st != EDSGPU_OK) return st;
    if (norm_out) *norm_out = nrm;
    if (host_frame_out) {
        DeviceGuard g(ctx->device);
        return read_image64(ctx, frames, slot, true, host_frame_out);
    }
    return EDSGPU_OK;
}

void* edsgpu_frames_build_stream(const edsgpu_frames* frames) { return frames ? (void*)frames->build_stream : nullptr; }

edsgpu_status edsgpu_frames_read(edsgpu_ctx* ctx, const edsgpu_frames* frames, int slot, double* image_out, double* norm_out) {
    if (!ctx) return EDSGPU_INVALID_ARGUMENT;
    EDS_REQUIRE(ctx, frames && slot >= 0 && slot < frames->capacity, "frames_read: bad slot");
    DeviceGuard g(ctx->device);
    if (image_out) {
        edsgpu_status st = read_image64(ctx, frames, slot, false, image_out);
        if (st != EDSGPU_OK) return st;
    }
    if (norm_out) {
        edsgpu_status st = edsgpu_frames_wait_built(frames, slot, 1, ctx->stream);
        if (st != EDSGPU_OK) return st;
        EDS_CUDA(ctx, cudaMemcpyAsync(norm_out, frames->norms + 2 * slot, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return EDSGPU_OK;
}

namespace {
__global__ void read_level_kernel(cudaTextureObject_t tex, int H, int W, double* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x < W && y < H) out[(size_t)y * W + x] = (double)tex2D<float>(tex, (float)x + 0.5f, (float)y + 0.5f);
}
}  // namespace

edsgpu_status edsgpu_frames_read_level(edsgpu_ctx* ctx, const edsgpu_frames* frames, int slot, int level, double* image_out, double* norm_out) {
    if (!ctx) return EDSGPU_INVALID_ARGUMENT;
    EDS_REQUIRE(ctx, frames && slot >= 0 && slot < frames->capacity && level >= 0 && level < frames->levels, "frames_read_level: bad slot or level");
    DeviceGuard g(ctx->device);
    const int H = frames->H, W = frames->W;
    const size_t npix = (size_t)H * W, idx = (size_t)level * frames->capacity + slot;
    edsgpu_status st = edsgpu_frames_wait_built(frames, slot, 1, ctx->stream);
    if (st != EDSGPU_OK) return st;
    if (image_out) {
        st = edsgpu_ensure_scratch(ctx, npix * sizeof(double));
        if (st != EDSGPU_OK) return st;
        read_level_kernel<<<dim3((W + 127) / 128, H), 128, 0, ctx->stream>>>(frames->tex[idx], H, W, (double*)ctx->scratch);
        ctx->launches++;
        EDS_CUDA(ctx, cudaGetLastError());
        EDS_CUDA(ctx, cudaMemcpyAsync(image_out, ctx->scratch, npix * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (norm_out) EDS_CUDA(ctx, cudaMemcpyAsync(norm_out, frames->norms + 2 * idx, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    st = edsgpu_frames_mark_read(frames, slot, 1, ctx->stream);
    if (st != EDSGPU_OK) return st;
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EDSGPU_OK;
}

edsgpu_status edsgpu_frames_read_accumulator(edsgpu_ctx* ctx, const edsgpu_frames* frames, int slot, int64_t* acc_out) {
    if (!ctx) return EDSGPU_INVALID_ARGUMENT;
    EDS_REQUIRE(ctx, frames && acc_out && slot >= 0 && slot < frames->capacity, "frames_read_accumulator: bad arguments");
    DeviceGuard g(ctx->device);
    const size_t npix = (size_t)frames->H * frames->W;
    EDS_REQUIRE(ctx, frames->slot_acc[slot] >= 0, "frames_read_accumulator: the accumulator of this slot was never filled or has been recycled (large sets of slots share a ring of accumulators)");
    edsgpu_status st = edsgpu_frames_wait_built(frames, slot, 1, ctx->stream);
    if (st != EDSGPU_OK) return st;
    EDS_CUDA(ctx, cudaMemcpyAsync(acc_out, frames->acc + (size_t)frames->slot_acc[slot] * npix, sizeof(int64_t) * npix, cudaMemcpyDeviceToHost, ctx->stream));
    st = edsgpu_frames_mark_read(frames, slot, 1, ctx->stream);
    if (st != EDSGPU_OK) return st;
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EDSGPU_OK;
}

}  // extern "C"
