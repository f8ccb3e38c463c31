/*
 * edsgpu.h -- C ABI of the B200-native EDS hot path (libedsgpu.so).
 *
 * Drop-in boundary for the two data-parallel paths of uzh-rpg/slam-eds:
 *   A. event-to-model alignment  (src/tracking: EventFrame::create, Tracker::optimize)
 *   B. windowed-BA Hessian accumulation (src/bundles: Accumulated{Top,SC}HessianSSE)
 * The reference has no plugin/FFI layer; its seam is the C++ class surface that the
 * external Rock task calls.  Each entry point below names the reference
 * method (file:line, relative to the reference root) whose body it replaces; the
 * header-only adapters in slam-eds_b200/host/ keep those method names and forward here.
 *
 * Conventions
 *   - plain C types only; every function returns edsgpu_status; no exception crosses.
 *   - pointers without a _dev suffix are HOST pointers; *_dev entry points take DEVICE
 *     pointers on the context's device (for callers whose data already lives in HBM).
 *   - one edsgpu_ctx = one device + one stream; a context is single-threaded.
 *   - quaternions are (x,y,z,w) like Eigen::Quaterniond::coeffs() (Tracker.hpp:48).
 *   - there is NO CPU fallback: without a CUDA device edsgpu_create fails.
 */
#ifndef EDSGPU_H_
#define EDSGPU_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    EDSGPU_OK = 0,
    EDSGPU_INVALID_ARGUMENT = 1,
    EDSGPU_CUDA_ERROR = 2,
    EDSGPU_NOT_USABLE = 3,          /* !summary.IsSolutionUsable(), Tracker.cpp:217,237-240 */
    EDSGPU_NON_MONOTONIC_TIME = 4,  /* EventFrame.cpp:325-329 throws std::runtime_error   */
    EDSGPU_OUT_OF_MEMORY = 5
} edsgpu_status;

typedef struct edsgpu_ctx edsgpu_ctx;
typedef struct edsgpu_lut edsgpu_lut;           /* device copy of EventFrame::fwd_mapx/fwd_mapy */
typedef struct edsgpu_frames edsgpu_frames;     /* a bank of device event frames (EventFrame::event_frame[0]) */
typedef struct edsgpu_keyframe edsgpu_keyframe; /* device copy of the KeyFrame arrays the tracker gathers */
typedef struct edsgpu_tracker edsgpu_tracker;   /* Tracker state: px, qx, vx, loss_params (Tracker.hpp:47-49) */

/* ---- context ------------------------------------------------------------------------- */
const char* edsgpu_version(void);
/* stream: a cudaStream_t to run on (e.g. torch's current stream), or NULL to create one. */
edsgpu_status edsgpu_create(int device, void* stream, edsgpu_ctx** out);
void edsgpu_destroy(edsgpu_ctx* ctx);
const char* edsgpu_last_error_string(const edsgpu_ctx* ctx);
edsgpu_status edsgpu_synchronize(edsgpu_ctx* ctx);
/* number of kernels this context has launched so far (for gpu_launches accounting) */
int64_t edsgpu_launch_count(const edsgpu_ctx* ctx);

/* ---- A1. event frame ----------------------------------------------------------------- */
enum { EDSGPU_DRAW_NN = 0, EDSGPU_DRAW_BILINEAR = 1 };  /* Utils.cpp:73,83 method strings */

/* EventFrame::EventFrame(cam, newcam, ...) builds fwd_mapx/fwd_mapy once (EventFrame.cpp:53-81).
 * mapx/mapy: H*W float each (CV_32F), row-major.  NULL,NULL = identity map. */
edsgpu_status edsgpu_lut_create(edsgpu_ctx* ctx, int height, int width, const float* fwd_mapx, const float* fwd_mapy,
                                edsgpu_lut** out);
void edsgpu_lut_destroy(edsgpu_lut* lut);

/* capacity event frames of height x width, resident on the device for the tracker. */
edsgpu_status edsgpu_frames_create(edsgpu_ctx* ctx, int height, int width, int capacity, edsgpu_frames** out);
/* The same with the pyramid of EventFrame::create (EventFrame.cpp:342-357, num_levels of the reference's signature):
 * level 0 is the event frame, level i >= 1 is cv::dilate + cv::erode of level 0 with a (2i+1) x (2i+1) rectangle (same
 * resolution), every level with its own L2 norm (:360-364).  Every edsgpu_event_frame_create* call then builds all levels
 * of the slots it fills.  num_levels in [1,5]. */
edsgpu_status edsgpu_frames_create_pyramid(edsgpu_ctx* ctx, int height, int width, int capacity, int num_levels, edsgpu_frames** out);
void edsgpu_frames_destroy(edsgpu_frames* frames);

/* EventFrame::create (EventFrame.cpp:302-389), pyramid level 0, out_scale 1:
 *   undistort via LUT (:313-323), drawValuesPoints (Utils.cpp:50-122) with `mode`,
 *   exp time weights if use_exp_weights (Utils.hpp:542-546), 3x3 Gaussian of `sigma`
 *   if sigma > 0 (Utils.cpp:113-119), L2 norm (:360-364), normalise (:367-383).
 * The reference call is (BILINEAR, use_exp_weights=1, sigma=0.5) (EventFrame.cpp:339).
 * The frame stays on the device in `slot`; norm_out / time / delta / host_frame_out
 * (H*W doubles, normalised) are optional.  ts_us may be NULL (no time outputs, no check). */
edsgpu_status edsgpu_event_frame_create(edsgpu_ctx* ctx, edsgpu_frames* frames, int slot, const edsgpu_lut* lut,
                                        const uint16_t* x, const uint16_t* y, const uint8_t* polarity, const int64_t* ts_us,
                                        int num_events, int mode, int use_exp_weights, float sigma, double* norm_out,
                                        int64_t* time_us_out, int64_t* delta_time_us_out, double* host_frame_out);

/* The CUDA stream (cudaStream_t) the library builds event frames on (created with the frames object).  Builds are ordered
 * against the context's stream per slot by the library itself; the handle is for measurement (CUDA events around a build). */
void* edsgpu_frames_build_stream(const edsgpu_frames* frames);

/* `count` windows of num_events events each, into slots first_slot.. (x,y,polarity are
 * count*num_events long).  norms_out: count doubles or NULL.  Asynchronous if norms_out is NULL.
 * Frames are BUILT on a stream of their own, ordered per slot against the library's readers
 * (a build waits for the last optimize/evaluate/read of the slots it overwrites, those wait for
 * the last build of the slots they use): with two banks of slots the frames of the next windows
 * are built while the context's stream still solves the current ones.  edsgpu_synchronize()
 * covers the build streams too. */
edsgpu_status edsgpu_event_frame_create_batch(edsgpu_ctx* ctx, edsgpu_frames* frames, int first_slot, int count,
                                              const edsgpu_lut* lut, const uint16_t* x, const uint16_t* y,
                                              const uint8_t* polarity, int num_events, int mode, int use_exp_weights,
                                              float sigma, double* norms_out);
/* same, events already in device memory; always asynchronous.  The build is ordered after
 * everything queued on the context stream at the time of the call (the producer of the arrays). */
edsgpu_status edsgpu_event_frame_create_batch_dev(edsgpu_ctx* ctx, edsgpu_frames* frames, int first_slot, int count,
                                                  const edsgpu_lut* lut, const uint16_t* x_dev, const uint16_t* y_dev,
                                                  const uint8_t* polarity_dev, int num_events, int mode,
                                                  int use_exp_weights, float sigma);
/* debug/parity: the un-normalised (blurred) image of a slot as doubles, and the raw
 * fixed-point accumulator (value * 2^40) before the blur. */
edsgpu_status edsgpu_frames_read(edsgpu_ctx* ctx, const edsgpu_frames* frames, int slot, double* image_out /*H*W*/,
                                 double* norm_out);
/* frame[level] as stored on the device (fp32, un-normalised, widened to double) and norm[level]; either may be NULL. */
edsgpu_status edsgpu_frames_read_level(edsgpu_ctx* ctx, const edsgpu_frames* frames, int slot, int level, double* image_out /*H*W*/,
                                       double* norm_out);
edsgpu_status edsgpu_frames_read_accumulator(edsgpu_ctx* ctx, const edsgpu_frames* frames, int slot, int64_t* acc_out /*H*W*/);
#define EDSGPU_ACC_FRACTION_BITS 40

/* ---- A2. tracker ----------------------------------------------------------------------- */
enum { EDSGPU_LOSS_NONE = 0, EDSGPU_LOSS_HUBER = 1, EDSGPU_LOSS_CAUCHY = 2 };     /* Tracker.cpp:146-161 */
enum { EDSGPU_LOSS_PARAM_CONSTANT = 0, EDSGPU_LOSS_PARAM_MAD = 1, EDSGPU_LOSS_PARAM_STD = 2 }; /* Tracker.hpp:34 */
enum { EDSGPU_TERM_CONVERGENCE = 0, EDSGPU_TERM_NO_CONVERGENCE = 1, EDSGPU_TERM_FAILURE = 2 };

typedef struct {
    int num_blocks;              /* config.options.num_threads: residual blocks (Tracker.cpp:138,178-195).
                                    NUMERICAL parameter: model norm and loss are per block.            */
    int loss_type;               /* config.loss_type                                                   */
    int max_iterations;          /* config.options.max_num_iterations[id] (Tracker.cpp:139)            */
    int loss_param_method;       /* LOSS_PARAM_METHOD argument of optimize (Tracker.hpp:80-81)         */
    double function_tolerance;   /* config.options.function_tolerance (Tracker.cpp:140)                */
    double gradient_tolerance;   /* 1e-8  (Tracker.cpp:142)                                            */
    double parameter_tolerance;  /* 1e-6  (Tracker.cpp:143)                                            */
} edsgpu_tracker_config;

typedef struct {                 /* eds::tracking::TrackerInfo (tracking/Config.hpp:60-68) + solver summary */
    int iterations;              /* successful + unsuccessful steps (Tracker.cpp:211) */
    int successful_steps;
    int unsuccessful_steps;
    int termination;             /* EDSGPU_TERM_* */
    int usable;                  /* summary.IsSolutionUsable() (Tracker.cpp:213) */
    int num_points;
    int evaluations;             /* sweeps over the points (residual + Jacobian + reduction) */
    int reserved;
    double initial_cost;
    double final_cost;
    double final_radius;
} edsgpu_tracker_info;

/* KeyFrame arrays gathered by PhotometricError (PhotometricError.hpp:58-112, KeyFrame.hpp:59-96):
 * grad_xy, norm_xy: N x 2 interleaved (std::vector<cv::Point2d>), idp (DepthPoints::getIDepth,
 * Tracker.cpp:167), weights: N.  Intrinsics from kf->K_ref (Tracker.cpp:164-166), size from kf->img.
 * num_blocks fixes the residual-block partition (Tracker.cpp:178-190). */
edsgpu_status edsgpu_keyframe_create(edsgpu_ctx* ctx, int num_points, const double* grad_xy, const double* norm_xy,
                                     const double* idp, const double* weights, int height, int width, double fx, double fy,
                                     double cx, double cy, int num_blocks, edsgpu_keyframe** out);
void edsgpu_keyframe_destroy(edsgpu_keyframe* kf);

/* Tracker::Tracker(config) (Tracker.cpp:41-48): px=0, qx=identity, vx=normalize(1e-3*ones),
 * loss parameter = loss_param (config.loss_params[0]). */
edsgpu_status edsgpu_tracker_create(edsgpu_ctx* ctx, const edsgpu_tracker_config* config, double loss_param,
                                    edsgpu_tracker** out);
void edsgpu_tracker_destroy(edsgpu_tracker* tracker);
/* Tracker::reset(kf, px, qx, velo) / set(T) (Tracker.cpp:50-82): any of px,qx,vx may be NULL (kept). */
edsgpu_status edsgpu_tracker_set_state(edsgpu_tracker* tracker, const double px[3], const double qx_xyzw[4],
                                       const double vx[6], const double* loss_param);
/* Tracker::getTransform / getVelocity / getLossParams: synchronises the stream. */
edsgpu_status edsgpu_tracker_get_state(edsgpu_tracker* tracker, double px[3], double qx_xyzw[4], double vx[6],
                                       double* loss_param, edsgpu_tracker_info* info);

/* bool Tracker::optimize(id, event_frame, T_kf_ef, loss_param_method) (Tracker.cpp:104-241):
 * LM on 0.5*sum_b rho(||r_b||^2) from the tracker's current state, write-back of px,qx,vx
 * when usable, residuals (no loss) and the next loss parameter (MAD, Tracker.cpp:223-233).
 * residuals_out: N doubles or NULL.  Returns EDSGPU_NOT_USABLE when optimize() would return false. */
edsgpu_status edsgpu_tracker_optimize(edsgpu_tracker* tracker, const edsgpu_keyframe* kf, const edsgpu_frames* frames,
                                      int slot, double px[3], double qx_xyzw[4], double vx[6], double* residuals_out,
                                      double* next_loss_param_out, edsgpu_tracker_info* info);
/* Tracker::optimize(id, &event_frame[id], ...) (Tracker.cpp:85-241) at pyramid level id of frames made with
 * edsgpu_frames_create_pyramid: the solve samples event_frame[level] (with norm[level]) and runs at most
 * config.options.max_num_iterations[level] iterations (Tracker.cpp:139), set per tracker with
 * edsgpu_tracker_set_level_iterations (without it every level uses edsgpu_tracker_config::max_iterations).
 * The reference's caller goes coarse to fine: level num_levels-1 down to 0, each solve warm-started by the last. */
edsgpu_status edsgpu_tracker_set_level_iterations(edsgpu_tracker* tracker, const int* max_num_iterations, int num_levels);
edsgpu_status edsgpu_tracker_optimize_level(edsgpu_tracker* tracker, const edsgpu_keyframe* kf, const edsgpu_frames* frames, int slot,
                                            int level, double px[3], double qx_xyzw[4], double vx[6], double* residuals_out,
                                            double* next_loss_param_out, edsgpu_tracker_info* info);

/* `count` independent trackers (sequences), tracker i against keyframe i and frame slot
 * first_slot+i, in one launch.  Asynchronous; read results with edsgpu_tracker_get_state
 * or edsgpu_trackers_gather. */
edsgpu_status edsgpu_trackers_optimize_batch(edsgpu_ctx* ctx, edsgpu_tracker* const* trackers,
                                             const edsgpu_keyframe* const* keyframes, int count, const edsgpu_frames* frames,
                                             int first_slot);
/* A fixed set of (tracker, keyframe, frame slot) problems whose device descriptors are built
 * once: edsgpu_batch_optimize is then two kernel launches and nothing else (asynchronous). */
typedef struct edsgpu_batch edsgpu_batch;
edsgpu_status edsgpu_batch_create(edsgpu_ctx* ctx, edsgpu_tracker* const* trackers, const edsgpu_keyframe* const* keyframes,
                                  int count, const edsgpu_frames* frames, int first_slot, edsgpu_batch** out);
/* the same against pyramid level `level` of the frames (see edsgpu_tracker_optimize_level) */
edsgpu_status edsgpu_batch_create_level(edsgpu_ctx* ctx, edsgpu_tracker* const* trackers, const edsgpu_keyframe* const* keyframes,
                                        int count, const edsgpu_frames* frames, int first_slot, int level, edsgpu_batch** out);
void edsgpu_batch_destroy(edsgpu_batch* batch);
edsgpu_status edsgpu_batch_optimize(edsgpu_batch* batch);
edsgpu_status edsgpu_batch_count(const edsgpu_batch* batch, int* count); /* problems of the batch */
/* How edsgpu_batch_optimize will launch: evaluator CTAs (they sweep residual blocks of any problem, taken from a
 * global task queue), leader CTAs (one warp per problem runs its Levenberg-Marquardt loop) and the number of problems
 * kept in flight at once (chosen when the batch is created; informational -- results do not depend on it). */
edsgpu_status edsgpu_batch_launch_shape(const edsgpu_batch* batch, int* evaluator_ctas, int* leader_ctas, int* problems_in_flight);
/* copies the count x 14 state records into one contiguous DEVICE buffer (e.g. the send buffer
 * of the caller's final NCCL all-gather); asynchronous on the context stream. */
edsgpu_status edsgpu_batch_pack_states_dev(edsgpu_batch* batch, double* states_dev);
/* states_out: count x 14 doubles [px(3) qx(4) vx(6) loss_param]; infos_out optional. Synchronises. */
edsgpu_status edsgpu_trackers_gather(edsgpu_ctx* ctx, edsgpu_tracker* const* trackers, int count, double* states_out,
                                     edsgpu_tracker_info* infos_out);
/* device address of tracker i's 14-double state record (for NCCL gathers by the caller) */
void* edsgpu_tracker_state_dev(edsgpu_tracker* tracker);

/* ceres::CostFunction::Evaluate as used at Tracker.cpp:228-229, plus the tangent-space
 * Jacobian: residuals (no loss) N, jacobian N x 12 row-major [t(3) theta(3) v(6)] (may be NULL),
 * robustified cost 0.5*sum_b rho(||r_b||^2), and the reduced normal equations (may be NULL):
 * H 12x12 row-major = sum_b rho'_b J_b^T J_b, g 12 = sum_b rho'_b J_b^T r_b. */
edsgpu_status edsgpu_tracker_evaluate(edsgpu_ctx* ctx, const edsgpu_keyframe* kf, const edsgpu_frames* frames, int slot,
                                      int loss_type, double loss_param, const double px[3], const double qx_xyzw[4],
                                      const double vx[6], double* residuals_out, double* jacobian_out, double* cost_out,
                                      double* H_out, double* g_out);

/* ---- B. windowed-BA Hessian accumulation -------------------------------------------------- */
/* One BA window = the residual graph the EnergyFunctional owns (EnergyFunctional.h:63-69,137-148):
 * F frames (<= 8; setting_maxFrames = 7), P points, R residuals stored POINT-MAJOR like
 * EFPoint::residualsAll: residuals of point p are [res_begin[p], res_begin[p+1]); host_idx/target_idx
 * are EFResidual::hostIDX/targetIDX.  The (host,target) tile plan is built once per graph. */
typedef struct edsgpu_ba edsgpu_ba;
#define EDSGPU_RAWJAC_FLOATS 76  /* dso::RawResidualJacobian (RawResidualJacobian.h:32-61) as Eigen lays it out
                                    with 16-byte alignment (304 B): resF[8] @0, Jpdxi[2][6] @8, Jpdc[2][4] @20, Jpdd[2] @28,
                                    pad[2], JIdx[2][8] @32, JabF[2][8] @48, JIdx2 @64, JabJIdx @68, Jab2 @72 (Mat22f column-major) */
enum { EDSGPU_RES_ACTIVE = 1, EDSGPU_RES_LINEARIZED = 2 };  /* EFResidual::isActive(), isLinearized */
edsgpu_status edsgpu_ba_create(edsgpu_ctx* ctx, int num_frames, int num_points, int num_residuals, const int32_t* host_idx,
                               const int32_t* target_idx, const int32_t* res_begin, edsgpu_ba** out);
void edsgpu_ba_destroy(edsgpu_ba* ba);
/* per linearisation: records (R x 76 floats), flags (R), EFResidual::res_toZeroF (R x 8, NULL if unused).
 * Also evaluates EFResidual::takeDataF's JpJdF (EnergyFunctionalStructs.cpp:38-48) on the device. */
edsgpu_status edsgpu_ba_set_residuals(edsgpu_ba* ba, const float* records, const uint8_t* flags, const float* res_toZero);
/* EFPoint::deltaF, priorF (EnergyFunctionalStructs.cpp:77-83); NULL = zeros. */
edsgpu_status edsgpu_ba_set_points(edsgpu_ba* ba, const float* deltaF, const float* priorF);
/* EnergyFunctional::adHTdeltaF (F*F x 8), cDeltaF (4) (EnergyFunctional.cpp:171-184), adHost/adTarget
 * (F*F Mat88, column-major = reinterpret_cast<const double*>(EF->adHost), EnergyFunctional.cpp:46-106); any may be NULL (kept). */
edsgpu_status edsgpu_ba_set_frames(edsgpu_ba* ba, const float* adHTdeltaF, const float* cDeltaF, const double* adHost,
                                   const double* adTarget);
/* AccumulatedTopHessianSSE::addPoint<mode> over all points (EnergyFunctional.cpp:197-238; mode 0 active,
 * 1 linearized, 2 marginalize; AccumulatedTopHessian.cpp:39-159).  acc_out: F*F x 13x13 doubles
 * (AccumulatorApprox::finish()'s H, index order [C(4) xi(6) a b r], accumulator h + t*F);
 * Hdd/bd (P) and Hcd (P x 4) are EFPoint::{Hdd,bd,Hcd}_acc{A,L}F.  All outputs optional; with none the
 * call is asynchronous and the results stay on the device for the stitch / SC calls. */
edsgpu_status edsgpu_ba_top_accumulate(edsgpu_ba* ba, int mode, double* acc_out, float* Hdd_out, float* bd_out, float* Hcd_out,
                                       int64_t* nres_out);
/* AccumulatedTopHessianSSE::stitchDoubleMT (AccumulatedTopHessian.h:91-139): which = 0 the mode-0
 * accumulators, 1 the mode-1/2 ones.  H: (4+8F)^2 column-major (Eigen MatXX), b: 4+8F.
 * use_prior adds cPrior / EFFrame::prior / delta_prior (.cpp:292-302). */
edsgpu_status edsgpu_ba_top_stitch(edsgpu_ba* ba, int which, int use_prior, const double* cPrior, const double* frame_prior,
                                   const double* frame_delta_prior, double* H, double* b);
/* AccumulatedSCHessianSSE::addPoint over all points (EnergyFunctional.cpp:244-261, AccumulatedSCHessian.cpp:34-77)
 * from the device-resident results of the two top accumulations.  accD: F^3 x 8x8 (index h + t1*F + t2*F*F),
 * accE: F*F x 8x4, accEB: F*F x 8, accHcc 4x4, accbc 4 (row-major doubles); HdiF/bdSum: EFPoint::HdiF, bdSumF. */
edsgpu_status edsgpu_ba_sc_accumulate(edsgpu_ba* ba, int shift_prior_to_zero, double* accD, double* accE, double* accEB,
                                      double* accHcc, double* accbc, float* HdiF_out, float* bdSum_out);
/* AccumulatedSCHessianSSE::stitchDoubleMT (AccumulatedSCHessian.h:93-133). */
edsgpu_status edsgpu_ba_sc_stitch(edsgpu_ba* ba, double* H, double* b);
/* EFResidual::JpJdF of every residual (R x 8). */
edsgpu_status edsgpu_ba_get_jpjd(edsgpu_ba* ba, float* JpJdF_out);

/* ---- the feeder on the device --------------------------------------
 * PointFrameResidual::linearize (src/tracking/Residuals.cpp:69-265) for every residual of the
 * window, computed on the device: the 304-byte records are produced in HBM where the
 * accumulators read them and never cross PCIe (edsgpu_ba_set_residuals is then not needed).
 * Replaces the loop FullSystem::linearizeAll_Reductor runs over activeResiduals. */
#define EDSGPU_PRECALC_FLOATS 28
/* FrameHessian::dI of frame `frame`: height*width Vec3f {I, dx, dy} (HessianBlocks.h:118). */
edsgpu_status edsgpu_ba_set_image(edsgpu_ba* ba, int frame, int height, int width, const float* dI);
/* precalc: F*F records of EDSGPU_PRECALC_FLOATS floats at index host + F*target, the members of
 *   FrameFramePrecalc (HessianBlocks.h:77-104) the feeder reads, 3x3 blocks column-major as Eigen
 *   stores them: PRE_RTll_0 (9), PRE_tTll_0 (3), PRE_KRKiTll (9), PRE_KtTll (3), PRE_aff_mode (2),
 *   PRE_b0_mode (1), pad (1).
 * calib: CalibHessian fxl, fyl, cxl, cyl.  frame_energy_th: FrameHessian::frameEnergyTH, F floats.
 * per point (P): PointHessian u, v, idepth_zero_scaled, idepth_scaled, color[8], weights[8]. */
edsgpu_status edsgpu_ba_set_linearize_inputs(edsgpu_ba* ba, const float* precalc, const float calib[4], const float* frame_energy_th,
                                             const float* u, const float* v, const float* idepth_zero_scaled,
                                             const float* idepth_scaled, const float* color, const float* weights);
/* state_in: ResState per residual (0 IN, 1 OOB, 2 OUTLIER) or NULL = none is OOB yet (OOB ones
 *   are skipped, :73-74).  linearized: EFResidual::isLinearized per residual or NULL = none;
 *   res_toZero as in edsgpu_ba_set_residuals (needed for linearized residuals only).
 * Afterwards the window holds the records, JpJdF and the flags (ACTIVE <=> new state IN).
 * state_out / energy_out: state_NewState / state_NewEnergy (R each) or NULL; the call is
 * asynchronous when every host pointer is NULL.  Records of residuals that leave as OOB are zero
 * (the reference leaves J stale there and never reads it); projectedTo / centerProjectedTo
 * (visualisation only) are not produced. */
edsgpu_status edsgpu_ba_linearize(edsgpu_ba* ba, const uint8_t* state_in, const uint8_t* linearized, const float* res_toZero,
                                  int32_t* state_out, float* energy_out);
/* edsgpu_ba_linearize FUSED with edsgpu_ba_top_accumulate(mode 0): linearizeAll_Reductor followed by accumulateAF_MT
 * (EnergyFunctional.cpp:838-860; Residuals.cpp:69-265 feeding AccumulatedTopHessian.cpp:102-135 and
 * EnergyFunctionalStructs.cpp:38-48) in ONE kernel: the thread that linearises a residual accumulates it from its
 * registers, so the 304-byte record is neither written nor re-read.  Results are bit-identical to the two-call sequence.
 * write_records == 0 keeps only the records later stages read (the LINEARIZED residuals: mode-1 pass, linearised
 * energy); pass 1 before edsgpu_ba_fix_linearization or edsgpu_ba_get_residuals, which need them all.  Asynchronous
 * when every host pointer is NULL; the active-side accumulation (acc / Hdd / bd / Hcd) stays on the device for
 * edsgpu_ba_top_stitch / edsgpu_ba_sc_accumulate (edsgpu_ba_top_read copies it out). */
edsgpu_status edsgpu_ba_linearize_accumulate(edsgpu_ba* ba, const uint8_t* state_in, const uint8_t* linearized, const float* res_toZero,
                                             int write_records, int32_t* state_out, float* energy_out);
/* Read-back of an accumulation that is already on the device (by edsgpu_ba_top_accumulate or
 * edsgpu_ba_linearize_accumulate): which = 0 (active side) or 1 (linearized side); same outputs as top_accumulate. */
edsgpu_status edsgpu_ba_top_read(edsgpu_ba* ba, int which, double* acc_out, float* Hdd_out, float* bd_out, float* Hcd_out, int64_t* nres_out);
/* ---- after the solve -------------------------------------------------
 * EnergyFunctional::resubstituteF_MT (EnergyFunctional.cpp:263-317): per-point inverse-depth step
 * from the solved update x (4 + 8F doubles).  Needs the adjoints (edsgpu_ba_set_frames) and the
 * results of top_accumulate(0), top_accumulate(1) and sc_accumulate of this linearisation, which
 * are still on the device.  point_step_out: PointHessian::step, P floats. */
edsgpu_status edsgpu_ba_resubstitute(edsgpu_ba* ba, const double* x, float* point_step_out);
/* EnergyFunctional::solveSystemF (EnergyFunctional.cpp:775-912) on the device, default solver mode (setting_solverMode =
 * SOLVER_FIX_LAMBDA | SOLVER_ORTHOGONALIZE_X_LATER, settings.cpp:60 -- the caller passes the lambda that mode fixes, 1e-5):
 * the three stitches, HFinal = HL + HM + HA with diag * (1 + lambda) - H_sc / (1 + lambda), bFinal = bL + (bM + HM delta) +
 * bA - b_sc, the diagonally scaled (4+8F)^2 LDL^T solve, orthogonalize(&x, 0) if a projector is given, then
 * resubstituteF_MT with the x that never left the device.  Closes a Gauss-Newton iteration of the window optimiser: after
 * edsgpu_ba_linearize_accumulate, edsgpu_ba_top_accumulate(1) and edsgpu_ba_sc_accumulate this is the only call, and
 * only x (4+8F doubles) and, on request, the P point steps come back.
 * HM (n x n, column-major), bM: the marginalisation prior, or both NULL; delta: getStitchedDeltaF() (n) or NULL = 0;
 * cPrior (4), frame_prior / frame_delta_prior (8F each): the priors of accumulateLF_MT's stitch, or all NULL;
 * nullspace_projector: N (N^T N)^-1 N^T (n x n) of EnergyFunctional::orthogonalize (:718-772), which depends on the frames'
 * null spaces only and is made by the host, or NULL (iterations 0 and 1).  Needs edsgpu_ba_set_frames with the adjoints. */
edsgpu_status edsgpu_ba_solve_system(edsgpu_ba* ba, double lambda, const double* HM, const double* bM, const double* delta,
                                     const double* cPrior, const double* frame_prior, const double* frame_delta_prior,
                                     const double* nullspace_projector, double* x_out, float* point_step_out);
/* EFResidual::fixLinearizationF (EnergyFunctionalStructs.cpp:87-113) for the residuals with
 * select[r] != 0 (NULL = every active residual): res_toZero = resF - J delta with the current
 * deltas, isLinearized = true.  res_toZero_out: R x 8 floats or NULL. */
edsgpu_status edsgpu_ba_fix_linearization(edsgpu_ba* ba, const uint8_t* select, float* res_toZero_out);
/* EnergyFunctional::calcLEnergyF_MT (EnergyFunctional.cpp:332-415): energy of the linearised part
 * at the current deltas; cPrior (4), frame_prior / frame_delta_prior (F x 8) may be NULL. */
edsgpu_status edsgpu_ba_calc_l_energy(edsgpu_ba* ba, const double* cPrior, const double* frame_prior, const double* frame_delta_prior,
                                      double* energy_out);
/* debug/parity: the records and flags as they sit on the device (R x 76 floats, R bytes). */
edsgpu_status edsgpu_ba_get_residuals(edsgpu_ba* ba, float* recs_out, uint8_t* flags_out);

/* ---- DSO coarse tracker evaluation ----------------------------------
 * CoarseTracker::calcRes fused with CoarseTracker::calcGSSSE (src/tracking/CoarseTracker.cpp:
 * 287-498): direct alignment of a new frame against the reference point cloud at one pyramid
 * level.  The Gauss-Newton loop of trackNewestCoarse (:520-701) stays on the host (8x8 solves);
 * each of its evaluations is one call. */
typedef struct edsgpu_coarse edsgpu_coarse;
edsgpu_status edsgpu_coarse_create(edsgpu_ctx* ctx, int num_levels, edsgpu_coarse** out);
void edsgpu_coarse_destroy(edsgpu_coarse* coarse);
/* makeK (:67-100): size and intrinsics of pyramid level lvl, Ki = K[lvl].inverse() as Eigen stores
 * it (9 floats, column-major). */
edsgpu_status edsgpu_coarse_set_level(edsgpu_coarse* coarse, int lvl, int width, int height, float fx, float fy, float cx, float cy,
                                      const float Ki[9]);
/* setCoarseTrackingRef / makeCoarseDepthL0 output (:103-283): pc_u, pc_v, pc_idepth, pc_color of level lvl. */
edsgpu_status edsgpu_coarse_set_reference(edsgpu_coarse* coarse, int lvl, int n, const float* pc_u, const float* pc_v,
                                          const float* pc_idepth, const float* pc_color);
/* CoarseTracker::makeCoarseDepthL0 (:127-283) on the device: the reference frame's inverse-depth map from the points projected
 * into it (proj_u, proj_v, proj_idepth = PointFrameResidual::centerProjectedTo of the n points whose last residual is IN, HdiF =
 * EFPoint::HdiF), summed down the pyramid, dilated by one pixel, normalised and compacted in scan-line order into pc_u / pc_v /
 * pc_idepth / pc_color of levels 0 .. levels_used-1 -- the level's reference point cloud is then in place for
 * edsgpu_coarse_calc_res_gs / edsgpu_coarse_track, as after edsgpu_coarse_set_reference.  Needs edsgpu_coarse_set_level (levels
 * must halve, makeK) and edsgpu_coarse_set_reference_frame (lastRef->dIp[lvl], height*width Vec3f) for every level used.
 * pc_n_out: pc_n of every level or NULL; edsgpu_coarse_get_reference reads a level's point cloud back (any pointer may be NULL). */
edsgpu_status edsgpu_coarse_set_reference_frame(edsgpu_coarse* coarse, int lvl, const float* dI);
edsgpu_status edsgpu_coarse_make_depth_l0(edsgpu_coarse* coarse, int levels_used, int n, const float* proj_u, const float* proj_v,
                                          const float* proj_idepth, const float* HdiF, int* pc_n_out);
edsgpu_status edsgpu_coarse_get_reference(edsgpu_coarse* coarse, int lvl, int* n_out, float* pc_u, float* pc_v, float* pc_idepth,
                                          float* pc_color);
/* newFrame->dIp[lvl]: height*width Vec3f {I, dx, dy}. */
edsgpu_status edsgpu_coarse_set_new_frame(edsgpu_coarse* coarse, int lvl, const float* dI);
/* R (row-major), t: refToNew; affLL = AffLight::fromToVecExposure(...) as floats, b0 = lastRef_aff_g2l.b.
 * rs: the Vec6 calcRes returns {E, numTermsInE, flow_t, 0, flow_rt, saturated ratio}; H (8x8), b (8):
 * calcGSSSE's outputs including the SCALE_* factors; H and b may be NULL (residual only). */
edsgpu_status edsgpu_coarse_calc_res_gs(edsgpu_coarse* coarse, int lvl, const double R[9], const double t[3], const float affLL[2],
                                        float b0, float cutoffTH, double rs[6], double H[64], double b[8]);

/* CoarseTracker::trackNewestCoarse (:520-701): the coarse-to-fine Gauss-Newton loop (levels coarsest_lvl..0,
 * Levenberg damping, step extrapolation, accept / reject, cutoff repeat) run on the host around the device
 * evaluation; Sophus' SE3::exp and Eigen's 8x8 LDLT are restated.  R (row-major), t: lastToNew_out, in-out;
 * aff_g2l: {a, b} of the new frame, in-out; ref_aff_g2l: lastRef_aff_g2l; exposures: ab_exposure of both frames;
 * min_res_for_abort: 5 values or NULL (never abort); last_residuals (5, NaN where a level was not reached) and
 * last_flow (3): lastResiduals / lastFlowIndicators.  Returns EDSGPU_NOT_USABLE where the reference returns
 * false (pose and affine parameters are then left untouched, except for the affine range check, as in the
 * reference). */
edsgpu_status edsgpu_coarse_track(edsgpu_coarse* coarse, int coarsest_lvl, double R[9], double t[3], double aff_g2l[2],
                                  const double ref_aff_g2l[2], float ref_exposure, float new_exposure, const double min_res_for_abort[5],
                                  double last_residuals[5], double last_flow[3], int* evaluations_out);

/* ---- per-point depth filter -----------------------------------------
 * eds::mapping::DepthPoints (src/mapping/DepthPoints.{hpp,cpp}): the filter state {mu = inverse
 * depth, sigma2, a, b} of every key-frame point stays on the device; update() is one launch. */
typedef struct edsgpu_depth_points edsgpu_depth_points;
/* DepthPoints::init (:52-91): inv_depth NULL = every point at the mean depth with sigma2 = range^2,
 * else the given inverse depths with sigma2 = range^2 / 36.  px_noise is the reference's 3 px. */
edsgpu_status edsgpu_depth_points_create(edsgpu_ctx* ctx, int num_points, double fx, double fy, double cx, double cy, double min_depth,
                                         double max_depth, const double* inv_depth, double init_a, double init_b,
                                         edsgpu_depth_points** out);
void edsgpu_depth_points_destroy(edsgpu_depth_points* dp);
/* DepthPoints::update (:93-176): T_kf_ef row-major 4x4; kf_coord / ef_coord: N x 2 pixel coordinates
 * (std::vector<cv::Point2d> is layout-compatible); coords_are_tracks != 0: ef_coord holds the offsets
 * (KeyFrame::tracks) instead.  ok_out: filterVogiatzis' return value per point, N bytes or NULL. */
edsgpu_status edsgpu_depth_points_update(edsgpu_depth_points* dp, const double T_kf_ef[16], const double* kf_coord, const double* ef_coord,
                                         int coords_are_tracks, uint8_t* ok_out);
/* N x 4 doubles {mu, sigma2, a, b} (getIDepth is column 0). */
edsgpu_status edsgpu_depth_points_get(edsgpu_depth_points* dp, double* state_out);
/* Tracker::getCoord (Tracker.cpp:319-376) with the filter's current inverse depths and the tracker's
 * pose, all on the device: coord_out N x 2 pixel coordinates in the event frame, outlier_out[i] != 0
 * where the reference would erase the point (outside the image); either may be NULL (asynchronous). */
edsgpu_status edsgpu_tracker_get_coord(edsgpu_tracker* tracker, const edsgpu_keyframe* kf, const edsgpu_depth_points* dp, double* coord_out,
                                       uint8_t* outlier_out);
/* The key frame's inverse depths <- the filter's means, on the device (what re-uploading the key frame
 * after KeyFrame::inv_depth changed would do; Tracker.cpp:167). */
edsgpu_status edsgpu_keyframe_refresh_idepth(edsgpu_keyframe* kf, const edsgpu_depth_points* dp);
/* One call per tracked window: getCoord -> DepthPoints::update(T_kf_ef = getTransform().inverse(), kf_coord,
 * coord) -> optional key-frame refresh, without leaving the device.  kf_coord: KeyFrame::coord (N x 2),
 * needed on the first call for a key frame, NULL afterwards.  Asynchronous when kf_coord is NULL. */
edsgpu_status edsgpu_depth_points_update_from_tracker(edsgpu_depth_points* dp, edsgpu_tracker* tracker, edsgpu_keyframe* kf,
                                                      const double* kf_coord, int refresh_keyframe);

#ifdef __cplusplus
}
#endif
#endif /* EDSGPU_H_ */
