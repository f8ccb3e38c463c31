// Header-only C++ host layer over the C ABI (include/edsgpu.h).
//
// Mirrors the class surface the reference exposes for the hot path -- same class names, method
// names, argument meaning and error behaviour -- without the reference's third-party types (Eigen,
// OpenCV, Rock base-types are not available in this image): plain arrays stand in for
// Eigen::Vector3d / Quaterniond / base::Transform3d.  INTEGRATION.md shows the same bodies written
// against the reference's own types.
//
//   eds::tracking::EventFrame   src/tracking/EventFrame.hpp:31-107
//   eds::tracking::Tracker      src/tracking/Tracker.hpp:36-114
//   dso::AccumulatedTopHessianSSE / AccumulatedSCHessianSSE   src/bundles/Accumulated{Top,SC}Hessian.h
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/edsgpu.h"

namespace edsgpu_host {

// RAII context; throws when there is no CUDA device (there is no CPU fallback).
class Context {
  public:
    explicit Context(int device = 0, void* stream = nullptr) {
        if (edsgpu_create(device, stream, &ctx_) != EDSGPU_OK) throw std::runtime_error("edsgpu_create failed: no CUDA device");
    }
    ~Context() { edsgpu_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    edsgpu_ctx* get() const { return ctx_; }
    void check(edsgpu_status st) const {
        if (st != EDSGPU_OK) throw std::runtime_error(std::string("edsgpu: ") + edsgpu_last_error_string(ctx_));
    }

  private:
    edsgpu_ctx* ctx_ = nullptr;
};

}  // namespace edsgpu_host

namespace eds { namespace tracking {

struct Event {  // stand-in for base::samples::Event (src/utils/Utils.hpp:40)
    uint16_t x, y;
    int64_t ts_us;
    uint8_t polarity;
};

enum LOSS_FUNCTION { NONE = 0, HUBER = 1, CAUCHY = 2 };
enum LOSS_PARAM_METHOD { CONSTANT = 0, MAD = 1, STD = 2 };  // Tracker.hpp:34

struct SolverOptions {  // tracking/Config.hpp:40-47
    int num_threads = 8;
    std::vector<int> max_num_iterations{30};
    double function_tolerance = 1e-6;
};
struct Config {  // tracking/Config.hpp:49-57
    LOSS_FUNCTION loss_type = HUBER;
    std::vector<double> loss_params{0.05};
    SolverOptions options;
};
struct TrackerInfo {  // tracking/Config.hpp:60-68
    double meas_time_us = 0;
    uint32_t num_points = 0;
    int num_iterations = 0;
    double time_seconds = 0;
    uint8_t success = 0;
};

// EventFrame (EventFrame.hpp:31-107) with its pyramid (EventFrame.cpp:342-364).
class EventFrame {
  public:
    // EventFrame(cam, newcam, ...): the forward undistortion LUT is built once (EventFrame.cpp:53-81).  num_levels is the
    // argument the reference passes to create(); the device images are allocated once, here.
    EventFrame(const edsgpu_host::Context& ctx, uint16_t height, uint16_t width, const float* fwd_mapx = nullptr, const float* fwd_mapy = nullptr,
               int num_levels = 1)
        : ctx_(ctx), height(height), width(width), num_levels(num_levels) {
        if (fwd_mapx) ctx_.check(edsgpu_lut_create(ctx_.get(), height, width, fwd_mapx, fwd_mapy, &lut_));
        ctx_.check(edsgpu_frames_create_pyramid(ctx_.get(), height, width, 1, num_levels, &frames_));
    }
    ~EventFrame() { edsgpu_frames_destroy(frames_); edsgpu_lut_destroy(lut_); }
    EventFrame(const EventFrame&) = delete;
    EventFrame& operator=(const EventFrame&) = delete;

    // void create(idx, events, height, width, num_levels, T, out_size)  (EventFrame.cpp:302-389).
    // Throws std::runtime_error on first_ts > last_ts like the reference (:325-329).
    void create(const uint64_t& idx_, const std::vector<Event>& events, bool keep_host_frame = false) {
        idx = idx_;
        const size_t n = events.size();
        x_.resize(n); y_.resize(n); p_.resize(n); t_.resize(n);
        for (size_t i = 0; i < n; ++i) { x_[i] = events[i].x; y_[i] = events[i].y; p_[i] = events[i].polarity; t_[i] = events[i].ts_us; }
        event_frame.assign(keep_host_frame ? (size_t)height * width : 0, 0.0);
        double nrm = 0.0;
        edsgpu_status st = edsgpu_event_frame_create(ctx_.get(), frames_, 0, lut_, x_.data(), y_.data(), p_.data(), t_.data(), (int)n,
                                                     EDSGPU_DRAW_BILINEAR, 1, 0.5f, &nrm, &time_us, &delta_time_us,
                                                     keep_host_frame ? event_frame.data() : nullptr);
        if (st == EDSGPU_NON_MONOTONIC_TIME) throw std::runtime_error("[EVENT_FRAME] FATAL ERROR Event time[0] > event time [N-1] ");
        ctx_.check(st);
        norm = nrm;
        norms.assign(num_levels, nrm);
        for (int l = 1; l < num_levels; ++l) ctx_.check(edsgpu_frames_read_level(ctx_.get(), frames_, 0, l, nullptr, &norms[l]));
    }
    const edsgpu_frames* frames() const { return frames_; }

    uint64_t idx = 0;
    uint16_t height, width;
    int num_levels;
    double norm = 0;                   // EventFrame::norm[0]
    std::vector<double> norms;         // EventFrame::norm
    int64_t time_us = 0, delta_time_us = 0;
    std::vector<double> event_frame;   // EventFrame::event_frame[0], filled on request

  private:
    const edsgpu_host::Context& ctx_;
    edsgpu_lut* lut_ = nullptr;
    edsgpu_frames* frames_ = nullptr;
    std::vector<uint16_t> x_, y_;
    std::vector<uint8_t> p_;
    std::vector<int64_t> t_;
};

// The KeyFrame arrays the tracker gathers (KeyFrame.hpp:59-96).
class KeyFrame {
  public:
    KeyFrame(const edsgpu_host::Context& ctx, const std::vector<double>& grad_xy, const std::vector<double>& norm_coord_xy,
             const std::vector<double>& idp, const std::vector<double>& weights, int height, int width, double fx, double fy, double cx,
             double cy, int num_blocks)
        : ctx_(ctx) {
        ctx_.check(edsgpu_keyframe_create(ctx_.get(), (int)idp.size(), grad_xy.data(), norm_coord_xy.data(), idp.data(), weights.data(), height,
                                          width, fx, fy, cx, cy, num_blocks, &kf_));
        residuals.resize(idp.size());
    }
    ~KeyFrame() { edsgpu_keyframe_destroy(kf_); }
    KeyFrame(const KeyFrame&) = delete;
    KeyFrame& operator=(const KeyFrame&) = delete;
    const edsgpu_keyframe* get() const { return kf_; }
    std::vector<double> residuals;  // KeyFrame::residuals, written by Tracker::optimize (Tracker.cpp:223-230)

  private:
    const edsgpu_host::Context& ctx_;
    edsgpu_keyframe* kf_ = nullptr;
};

// Tracker (Tracker.hpp:36-114): state px, qx, vx + config.loss_params carried across windows.
class Tracker {
  public:
    Config config;

    Tracker(const edsgpu_host::Context& ctx, std::shared_ptr<KeyFrame> kf, const Config& config_, int level = 0) : config(config_), ctx_(ctx), kf_(kf) {
        edsgpu_tracker_config c{};
        c.num_blocks = config.options.num_threads;
        c.loss_type = (int)config.loss_type;
        c.max_iterations = config.options.max_num_iterations.at(level);
        level_iterations_ = config.options.max_num_iterations;
        c.loss_param_method = MAD;
        c.function_tolerance = config.options.function_tolerance;
        c.gradient_tolerance = 1e-08;   // Tracker.cpp:142
        c.parameter_tolerance = 1e-06;  // Tracker.cpp:143
        ctx_.check(edsgpu_tracker_create(ctx_.get(), &c, config.loss_params.at(0), &tr_));
        ctx_.check(edsgpu_tracker_set_level_iterations(tr_, level_iterations_.data(), (int)level_iterations_.size()));  // Tracker.cpp:139
    }
    ~Tracker() { edsgpu_tracker_destroy(tr_); }
    Tracker(const Tracker&) = delete;
    Tracker& operator=(const Tracker&) = delete;

    // reset(kf, px, qx, velo) (Tracker.cpp:50-73)
    void reset(std::shared_ptr<KeyFrame> kf, const std::array<double, 3>& px, const std::array<double, 4>& qx_xyzw, const std::array<double, 6>* velo = nullptr) {
        kf_ = kf;
        ctx_.check(edsgpu_tracker_set_state(tr_, px.data(), qx_xyzw.data(), velo ? velo->data() : nullptr, nullptr));
    }
    // set(T_kf_ef) (Tracker.cpp:75-82): the tracker works with the inverse transform
    void set(const std::array<double, 3>& t_kf_ef, const std::array<double, 4>& q_kf_ef_xyzw) {
        std::array<double, 3> p;
        std::array<double, 4> q;
        inverse(t_kf_ef, q_kf_ef_xyzw, p, q);
        ctx_.check(edsgpu_tracker_set_state(tr_, p.data(), q.data(), nullptr, nullptr));
    }
    // bool optimize(id, event_frame, T_kf_ef, loss_param_method) (Tracker.cpp:104-241)
    bool optimize(const int& id, const EventFrame& ef, std::array<double, 3>& t_kf_ef, std::array<double, 4>& q_kf_ef_xyzw,
                  const LOSS_PARAM_METHOD loss_param_method = MAD) {
        (void)loss_param_method;  // fixed at construction in this adapter
        edsgpu_tracker_info inf{};
        double tau = 0.0;
        // id = pyramid level: event_frame[id], max_num_iterations[id]
        edsgpu_status st = edsgpu_tracker_optimize_level(tr_, kf_->get(), ef.frames(), 0, id, px.data(), qx.data(), vx.data(), kf_->residuals.data(), &tau, &inf);
        info.num_points = (uint32_t)inf.num_points;
        info.num_iterations = inf.iterations;
        info.success = (uint8_t)inf.usable;
        if (st == EDSGPU_NOT_USABLE) return false;  // !summary.IsSolutionUsable(), Tracker.cpp:237-240
        ctx_.check(st);
        inverse(px, qx, t_kf_ef, q_kf_ef_xyzw);     // T_kf_ef = getTransform().inverse(), Tracker.cpp:220
        config.loss_params = {tau};                 // Tracker.cpp:233
        return true;
    }
    std::array<double, 6>& getVelocity() { return vx; }
    std::vector<double> getLossParams() const { return config.loss_params; }
    TrackerInfo getInfo() const { return info; }

    std::array<double, 3> px{};
    std::array<double, 4> qx{{0, 0, 0, 1}};
    std::array<double, 6> vx{};

  private:
    // inverse of a rigid transform given as (t, unit quaternion xyzw)
    static void inverse(const std::array<double, 3>& t, const std::array<double, 4>& q, std::array<double, 3>& ti, std::array<double, 4>& qi) {
        qi = {-q[0], -q[1], -q[2], q[3]};
        const double x = qi[0], y = qi[1], z = qi[2], w = qi[3];
        const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y), 2 * (x * y + w * z), 1 - 2 * (x * x + z * z),
                             2 * (y * z - w * x),     2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)};
        for (int r = 0; r < 3; ++r) ti[r] = -(R[3 * r] * t[0] + R[3 * r + 1] * t[1] + R[3 * r + 2] * t[2]);
    }
    const edsgpu_host::Context& ctx_;
    std::shared_ptr<KeyFrame> kf_;
    edsgpu_tracker* tr_ = nullptr;
    std::vector<int> level_iterations_;
    TrackerInfo info;
};

} }  // namespace eds::tracking

namespace eds { namespace mapping {
// eds::mapping::DepthPoints (src/mapping/DepthPoints.{hpp,cpp}): filter state on the device
class DepthPoints {
  public:
    DepthPoints(const edsgpu_host::Context& ctx, double fx, double fy, double cx, double cy, const std::vector<double>& inv_depth, double min_depth,
                double max_depth, double init_a = 10.0, double init_b = 10.0)
        : ctx_(ctx), n_((int)inv_depth.size()) {
        ctx_.check(edsgpu_depth_points_create(ctx_.get(), n_, fx, fy, cx, cy, min_depth, max_depth, inv_depth.data(), init_a, init_b, &dp_));
    }
    ~DepthPoints() { edsgpu_depth_points_destroy(dp_); }
    DepthPoints(const DepthPoints&) = delete;
    DepthPoints& operator=(const DepthPoints&) = delete;
    // T_kf_ef: row-major 4x4; coordinates: N x 2 doubles (std::vector<cv::Point2d>)
    void update(const double T_kf_ef[16], const double* kf_coord, const double* ef_coord, bool coords_are_tracks = false) {
        ctx_.check(edsgpu_depth_points_update(dp_, T_kf_ef, kf_coord, ef_coord, coords_are_tracks ? 1 : 0, nullptr));
    }
    void getIDepth(std::vector<double>& x) {
        std::vector<double> st(4 * (size_t)n_);
        ctx_.check(edsgpu_depth_points_get(dp_, st.data()));
        x.resize(n_);
        for (int i = 0; i < n_; ++i) x[i] = st[4 * (size_t)i];
    }

  private:
    const edsgpu_host::Context& ctx_;
    int n_;
    edsgpu_depth_points* dp_ = nullptr;
};
}}  // namespace eds::mapping

namespace dso {

// The accumulator side of dso::EnergyFunctional (EnergyFunctional.cpp:197-261) for one residual graph:
// accumulateAF_MT / accumulateLF_MT / accumulateSCF_MT with their stitches.
class HessianAccumulators {
  public:
    HessianAccumulators(const edsgpu_host::Context& ctx, int nFrames, const std::vector<int32_t>& hostIDX, const std::vector<int32_t>& targetIDX,
                        const std::vector<int32_t>& res_begin)
        : ctx_(ctx), nFrames(nFrames), n(4 + 8 * nFrames) {
        ctx_.check(edsgpu_ba_create(ctx_.get(), nFrames, (int)res_begin.size() - 1, (int)hostIDX.size(), hostIDX.data(), targetIDX.data(),
                                    res_begin.data(), &ba_));
    }
    ~HessianAccumulators() { edsgpu_ba_destroy(ba_); }
    HessianAccumulators(const HessianAccumulators&) = delete;
    HessianAccumulators& operator=(const HessianAccumulators&) = delete;

    // records: R x 76 floats (RawResidualJacobian), flags: EDSGPU_RES_ACTIVE | EDSGPU_RES_LINEARIZED
    void setResiduals(const float* records, const uint8_t* flags, const float* res_toZeroF) { ctx_.check(edsgpu_ba_set_residuals(ba_, records, flags, res_toZeroF)); }
    void setPoints(const float* deltaF, const float* priorF) { ctx_.check(edsgpu_ba_set_points(ba_, deltaF, priorF)); }
    // The feeder on the device instead of setResiduals: FrameHessian::dI per frame, then the state
    // PointFrameResidual::linearize reads (Residuals.cpp:69-265), then linearizeAll() for every residual.
    void setImage(int frame, int height, int width, const float* dI_vec3f) { ctx_.check(edsgpu_ba_set_image(ba_, frame, height, width, dI_vec3f)); }
    void setLinearizeInputs(const float* precalc, const float calib[4], const float* frameEnergyTH, const float* u, const float* v,
                            const float* idepth_zero_scaled, const float* idepth_scaled, const float* color, const float* weights) {
        ctx_.check(edsgpu_ba_set_linearize_inputs(ba_, precalc, calib, frameEnergyTH, u, v, idepth_zero_scaled, idepth_scaled, color, weights));
    }
    // state_NewState / state_NewEnergy of every residual (optional); records, JpJdF and flags stay on the device
    void linearizeAll(const uint8_t* state_state, const uint8_t* isLinearized, const float* res_toZeroF, int32_t* state_NewState,
                      float* state_NewEnergy) {
        ctx_.check(edsgpu_ba_linearize(ba_, state_state, isLinearized, res_toZeroF, state_NewState, state_NewEnergy));
    }
    // linearizeAll() + accumulateAF_MT() as ONE kernel (EnergyFunctional.cpp:838-860): the record of a residual nothing reads later
    // is never written; keepRecords = true before fixLinearizationF / marginalisation, which read them all
    void linearizeAllAndAccumulateAF_MT(const uint8_t* state_state, const uint8_t* isLinearized, const float* res_toZeroF, bool keepRecords,
                                        int32_t* state_NewState, float* state_NewEnergy, std::vector<double>& H, std::vector<double>& b) {
        ctx_.check(edsgpu_ba_linearize_accumulate(ba_, state_state, isLinearized, res_toZeroF, keepRecords ? 1 : 0, state_NewState, state_NewEnergy));
        H.assign((size_t)n * n, 0.0); b.assign(n, 0.0);
        ctx_.check(edsgpu_ba_top_stitch(ba_, 0, 0, nullptr, nullptr, nullptr, H.data(), b.data()));
    }
    // setAdjointsF / setDeltaF results (EnergyFunctional.cpp:46-106,171-194)
    void setFrames(const float* adHTdeltaF, const float* cDeltaF, const double* adHost, const double* adTarget) {
        ctx_.check(edsgpu_ba_set_frames(ba_, adHTdeltaF, cDeltaF, adHost, adTarget));
    }
    // accumulateAF_MT(H, b): addPoint<0> over all points + stitchDoubleMT(usePrior = false)
    void accumulateAF_MT(std::vector<double>& H, std::vector<double>& b) {
        ctx_.check(edsgpu_ba_top_accumulate(ba_, 0, nullptr, nullptr, nullptr, nullptr, nullptr));
        H.assign((size_t)n * n, 0.0); b.assign(n, 0.0);
        ctx_.check(edsgpu_ba_top_stitch(ba_, 0, 0, nullptr, nullptr, nullptr, H.data(), b.data()));
    }
    // accumulateLF_MT(H, b): addPoint<1> + stitchDoubleMT(usePrior = true)
    void accumulateLF_MT(std::vector<double>& H, std::vector<double>& b, const double cPrior[4], const double* framePrior, const double* frameDeltaPrior) {
        ctx_.check(edsgpu_ba_top_accumulate(ba_, 1, nullptr, nullptr, nullptr, nullptr, nullptr));
        H.assign((size_t)n * n, 0.0); b.assign(n, 0.0);
        ctx_.check(edsgpu_ba_top_stitch(ba_, 1, 1, cPrior, framePrior, frameDeltaPrior, H.data(), b.data()));
    }
    // solveSystemF(iteration, lambda, HCalib) (EnergyFunctional.cpp:775-912, default solver mode) after the three accumulations:
    // stitches, damped Schur-reduced system, scaled LDL^T, orthogonalize(&x, 0) when a projector is given, resubstituteF_MT
    void solveSystemF(double lambda, const double* HM, const double* bM, const double* stitchedDeltaF, const double cPrior[4],
                      const double* framePrior, const double* frameDeltaPrior, const double* nullspaceProjector, std::vector<double>& x,
                      std::vector<float>* pointStep = nullptr, int P = 0) {
        x.assign(n, 0.0);
        if (pointStep) pointStep->assign(P, 0.f);
        ctx_.check(edsgpu_ba_solve_system(ba_, lambda, HM, bM, stitchedDeltaF, cPrior, framePrior, frameDeltaPrior, nullspaceProjector, x.data(),
                                          pointStep ? pointStep->data() : nullptr));
    }
    // accumulateSCF_MT(H, b): addPoint(p, shiftPriorToZero = true) + stitchDoubleMT; also EFPoint::HdiF, bdSumF
    void accumulateSCF_MT(std::vector<double>& H, std::vector<double>& b, std::vector<float>* HdiF = nullptr, std::vector<float>* bdSumF = nullptr, int P = 0) {
        if (HdiF) HdiF->assign(P, 0.f);
        if (bdSumF) bdSumF->assign(P, 0.f);
        ctx_.check(edsgpu_ba_sc_accumulate(ba_, 1, nullptr, nullptr, nullptr, nullptr, nullptr, HdiF ? HdiF->data() : nullptr,
                                           bdSumF ? bdSumF->data() : nullptr));
        H.assign((size_t)n * n, 0.0); b.assign(n, 0.0);
        ctx_.check(edsgpu_ba_sc_stitch(ba_, H.data(), b.data()));
    }

    // resubstituteF_MT(x): PointHessian::step of every point (EnergyFunctional.cpp:263-317)
    void resubstituteF_MT(const double* x, float* pointStep) { ctx_.check(edsgpu_ba_resubstitute(ba_, x, pointStep)); }
    // calcLEnergyF_MT() (EnergyFunctional.cpp:396-415)
    double calcLEnergyF_MT(const double cPrior[4], const double* framePrior, const double* frameDeltaPrior) {
        double e = 0.0;
        ctx_.check(edsgpu_ba_calc_l_energy(ba_, cPrior, framePrior, frameDeltaPrior, &e));
        return e;
    }
    // EFResidual::fixLinearizationF for the selected residuals (EnergyFunctionalStructs.cpp:87-113)
    void fixLinearizationF(const uint8_t* select, float* res_toZeroF = nullptr) { ctx_.check(edsgpu_ba_fix_linearization(ba_, select, res_toZeroF)); }

  private:
    const edsgpu_host::Context& ctx_;
    edsgpu_ba* ba_ = nullptr;

  public:
    const int nFrames, n;
};

// dso::CoarseTracker (src/tracking/CoarseTracker.{h,cpp}): the evaluation calcRes + calcGSSSE per pyramid level;
// the Gauss-Newton loop of trackNewestCoarse stays with the caller.
class CoarseTrackerEval {
  public:
    CoarseTrackerEval(const edsgpu_host::Context& ctx, int levels) : ctx_(ctx) { ctx_.check(edsgpu_coarse_create(ctx_.get(), levels, &ct_)); }
    ~CoarseTrackerEval() { edsgpu_coarse_destroy(ct_); }
    CoarseTrackerEval(const CoarseTrackerEval&) = delete;
    CoarseTrackerEval& operator=(const CoarseTrackerEval&) = delete;
    void makeK(int lvl, int w, int h, float fx, float fy, float cx, float cy, const float Ki[9]) { ctx_.check(edsgpu_coarse_set_level(ct_, lvl, w, h, fx, fy, cx, cy, Ki)); }
    void setCoarseTrackingRef(int lvl, int n, const float* pc_u, const float* pc_v, const float* pc_idepth, const float* pc_color) {
        ctx_.check(edsgpu_coarse_set_reference(ct_, lvl, n, pc_u, pc_v, pc_idepth, pc_color));
    }
    void setNewFrame(int lvl, const float* dIp) { ctx_.check(edsgpu_coarse_set_new_frame(ct_, lvl, dIp)); }
    // rs = calcRes(...), H / b = calcGSSSE(...) (either both or none)
    void calcResAndGS(int lvl, const double R[9], const double t[3], const float affLL[2], float b0, float cutoffTH, double rs[6], double* H, double* b) {
        ctx_.check(edsgpu_coarse_calc_res_gs(ct_, lvl, R, t, affLL, b0, cutoffTH, rs, H, b));
    }
    // trackNewestCoarse: false where the reference returns false
    bool trackNewestCoarse(int coarsestLvl, double R[9], double t[3], double aff_g2l[2], const double ref_aff_g2l[2], float refExposure,
                           float newExposure, const double minResForAbort[5], double lastResiduals[5], double lastFlowIndicators[3]) {
        const edsgpu_status st = edsgpu_coarse_track(ct_, coarsestLvl, R, t, aff_g2l, ref_aff_g2l, refExposure, newExposure, minResForAbort,
                                                     lastResiduals, lastFlowIndicators, nullptr);
        if (st == EDSGPU_NOT_USABLE) return false;
        ctx_.check(st);
        return true;
    }

  private:
    const edsgpu_host::Context& ctx_;
    edsgpu_coarse* ct_ = nullptr;
};

}  // namespace dso
