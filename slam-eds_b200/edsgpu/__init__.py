"""ctypes binding of libedsgpu.so plus thin Python mirrors of the reference classes.

`EventFrame.create`, `Tracker.optimize` keep the reference method names and argument
meaning (src/tracking/EventFrame.hpp:83-85, src/tracking/Tracker.hpp:73-81); everything
below them is the C ABI of include/edsgpu.h.  There is no CPU fallback: if the shared
library or a CUDA device is missing, `load()` / `Context()` raise.
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
# EDSGPU_LIBRARY: developer override to A/B an experimental build of the same ABI
LIB_PATH = os.environ.get("EDSGPU_LIBRARY") or os.path.join(os.path.dirname(_PKG), "libedsgpu.so")
_lib = None

OK, INVALID_ARGUMENT, CUDA_ERROR, NOT_USABLE, NON_MONOTONIC_TIME, OUT_OF_MEMORY = range(6)
DRAW_NN, DRAW_BILINEAR = 0, 1
LOSS_NONE, LOSS_HUBER, LOSS_CAUCHY = 0, 1, 2
LOSS_PARAM_CONSTANT, LOSS_PARAM_MAD, LOSS_PARAM_STD = 0, 1, 2
TERM_CONVERGENCE, TERM_NO_CONVERGENCE, TERM_FAILURE = 0, 1, 2
ACC_FRACTION_BITS = 40

# every symbol include/edsgpu.h declares (tests check that the library exports them all)
SYMBOLS = [
    "edsgpu_version", "edsgpu_create", "edsgpu_destroy", "edsgpu_last_error_string", "edsgpu_synchronize",
    "edsgpu_launch_count", "edsgpu_lut_create", "edsgpu_lut_destroy", "edsgpu_frames_create", "edsgpu_frames_destroy",
    "edsgpu_event_frame_create", "edsgpu_event_frame_create_batch", "edsgpu_event_frame_create_batch_dev",
    "edsgpu_frames_read", "edsgpu_frames_read_accumulator", "edsgpu_frames_create_pyramid", "edsgpu_frames_read_level",
    "edsgpu_tracker_set_level_iterations", "edsgpu_tracker_optimize_level", "edsgpu_batch_create_level", "edsgpu_frames_build_stream", "edsgpu_keyframe_create", "edsgpu_keyframe_destroy",
    "edsgpu_tracker_create", "edsgpu_tracker_destroy", "edsgpu_tracker_set_state", "edsgpu_tracker_get_state",
    "edsgpu_tracker_optimize", "edsgpu_trackers_optimize_batch", "edsgpu_trackers_gather", "edsgpu_tracker_state_dev",
    "edsgpu_batch_create", "edsgpu_batch_destroy", "edsgpu_batch_optimize", "edsgpu_batch_count", "edsgpu_batch_pack_states_dev", "edsgpu_batch_launch_shape",
    "edsgpu_tracker_evaluate",
    "edsgpu_ba_create", "edsgpu_ba_destroy", "edsgpu_ba_set_residuals", "edsgpu_ba_set_points", "edsgpu_ba_set_frames",
    "edsgpu_ba_top_accumulate", "edsgpu_ba_top_stitch", "edsgpu_ba_sc_accumulate", "edsgpu_ba_sc_stitch", "edsgpu_ba_get_jpjd",
    "edsgpu_ba_set_image", "edsgpu_ba_set_linearize_inputs", "edsgpu_ba_linearize", "edsgpu_ba_linearize_accumulate", "edsgpu_ba_top_read", "edsgpu_ba_solve_system", "edsgpu_coarse_set_reference_frame", "edsgpu_coarse_make_depth_l0", "edsgpu_coarse_get_reference", "edsgpu_ba_get_residuals",
    "edsgpu_ba_resubstitute", "edsgpu_ba_fix_linearization", "edsgpu_ba_calc_l_energy",
    "edsgpu_coarse_create", "edsgpu_coarse_destroy", "edsgpu_coarse_set_level", "edsgpu_coarse_set_reference",
    "edsgpu_coarse_set_new_frame", "edsgpu_coarse_calc_res_gs", "edsgpu_coarse_track",
    "edsgpu_depth_points_create", "edsgpu_depth_points_destroy", "edsgpu_depth_points_update", "edsgpu_depth_points_get",
    "edsgpu_tracker_get_coord", "edsgpu_keyframe_refresh_idepth", "edsgpu_depth_points_update_from_tracker",
]


NCCL_LIB_PATH = os.path.join(os.path.dirname(_PKG), "libedsgpu_nccl.so")
NCCL_SYMBOLS = ["edsgpu_comm_unique_id", "edsgpu_comm_create", "edsgpu_comm_adopt", "edsgpu_comm_destroy",
                "edsgpu_gather_states_nccl", "edsgpu_batch_gather_states_nccl"]
_nccl_lib = None


def load_nccl():
    """libedsgpu_nccl.so: the final gather of a sharded batch (include/edsgpu_nccl.h).  Loaded on first use only."""
    global _nccl_lib
    if _nccl_lib is None:
        load()
        if not os.path.exists(NCCL_LIB_PATH):
            raise ImportError("libedsgpu_nccl.so not built: run `make -C slam-eds_b200`")
        # PyTorch bundles its own libnccl.so.2 (same SONAME, newer version).  Whichever copy is loaded first serves the
        # whole process, and torch cannot start on an older one: let it load its copy before ours asks for the SONAME.
        try:
            import torch  # noqa: F401
        except ImportError:
            pass
        _nccl_lib = C.CDLL(NCCL_LIB_PATH)
    return _nccl_lib


class EdsGpuError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("edsgpu status %d: %s" % (status, msg))
        self.status = status


class TrackerConfig(C.Structure):
    _fields_ = [("num_blocks", C.c_int), ("loss_type", C.c_int), ("max_iterations", C.c_int),
                ("loss_param_method", C.c_int), ("function_tolerance", C.c_double),
                ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double)]


class TrackerInfo(C.Structure):
    _fields_ = [("iterations", C.c_int), ("successful_steps", C.c_int), ("unsuccessful_steps", C.c_int),
                ("termination", C.c_int), ("usable", C.c_int), ("num_points", C.c_int),
                ("evaluations", C.c_int), ("reserved", C.c_int), ("initial_cost", C.c_double), ("final_cost", C.c_double), ("final_radius", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def load():
    """Load libedsgpu.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libedsgpu.so not built: run `make -C slam-eds_b200` (or __graft_entry__.build())")
        lib = C.CDLL(LIB_PATH)
        lib.edsgpu_version.restype = C.c_char_p
        lib.edsgpu_last_error_string.restype = C.c_char_p
        lib.edsgpu_last_error_string.argtypes = [C.c_void_p]
        lib.edsgpu_launch_count.restype = C.c_int64
        lib.edsgpu_launch_count.argtypes = [C.c_void_p]
        lib.edsgpu_tracker_state_dev.restype = C.c_void_p
        lib.edsgpu_tracker_state_dev.argtypes = [C.c_void_p]
        lib.edsgpu_frames_build_stream.restype = C.c_void_p
        lib.edsgpu_frames_build_stream.argtypes = [C.c_void_p]
        for name in ("edsgpu_destroy", "edsgpu_lut_destroy", "edsgpu_frames_destroy", "edsgpu_keyframe_destroy",
                     "edsgpu_tracker_destroy", "edsgpu_batch_destroy", "edsgpu_ba_destroy"):
            getattr(lib, name).restype = None
            getattr(lib, name).argtypes = [C.c_void_p]
        _lib = lib
    return _lib


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class Context:
    """edsgpu_ctx: one device + one stream."""

    def __init__(self, device=0, stream=None):
        self.lib = load()
        h = C.c_void_p()
        st = self.lib.edsgpu_create(C.c_int(device), C.c_void_p(stream) if stream else None, C.byref(h))
        if st != OK:
            raise EdsGpuError(st, "edsgpu_create failed (no CUDA device? there is no CPU fallback)")
        self.h = h
        self.device = device

    def check(self, st):
        if st != OK:
            raise EdsGpuError(st, self.lib.edsgpu_last_error_string(self.h).decode())

    def synchronize(self):
        self.check(self.lib.edsgpu_synchronize(self.h))

    @property
    def launches(self):
        return int(self.lib.edsgpu_launch_count(self.h))

    def close(self):
        if self.h:
            self.lib.edsgpu_destroy(self.h)
            self.h = None


class Lut:
    def __init__(self, ctx, H, W, mapx=None, mapy=None):
        self.ctx = ctx
        self.h = C.c_void_p()
        mx = np.ascontiguousarray(mapx, np.float32) if mapx is not None else None
        my = np.ascontiguousarray(mapy, np.float32) if mapy is not None else None
        ctx.check(ctx.lib.edsgpu_lut_create(ctx.h, C.c_int(H), C.c_int(W), _ptr(mx, C.c_float), _ptr(my, C.c_float),
                                            C.byref(self.h)))

    def close(self):
        if self.h:
            self.ctx.lib.edsgpu_lut_destroy(self.h)
            self.h = None


class Frames:
    """A bank of device event frames (EventFrame::event_frame[0 .. levels-1] of several windows)."""

    def __init__(self, ctx, H, W, capacity=1, levels=1):
        self.ctx, self.H, self.W, self.capacity, self.levels = ctx, H, W, capacity, levels
        self.h = C.c_void_p()
        ctx.check(ctx.lib.edsgpu_frames_create_pyramid(ctx.h, C.c_int(H), C.c_int(W), C.c_int(capacity), C.c_int(levels), C.byref(self.h)))

    def read_level(self, slot=0, level=0):
        """frame[level] of the slot as it is stored (fp32 widened to double, un-normalised) and norm[level]"""
        img = np.zeros((self.H, self.W), np.float64)
        norm = C.c_double(0)
        self.ctx.check(self.ctx.lib.edsgpu_frames_read_level(self.ctx.h, self.h, C.c_int(slot), C.c_int(level), _ptr(img, C.c_double),
                                                             C.byref(norm)))
        return img, norm.value

    def read(self, slot=0):
        img = np.zeros((self.H, self.W), np.float64)
        norm = C.c_double(0)
        self.ctx.check(self.ctx.lib.edsgpu_frames_read(self.ctx.h, self.h, C.c_int(slot), _ptr(img, C.c_double),
                                                       C.byref(norm)))
        return img, norm.value

    def build_stream(self):
        """cudaStream_t (int) the frames of this bank are built on."""
        return int(self.ctx.lib.edsgpu_frames_build_stream(self.h))

    def read_accumulator(self, slot=0):
        acc = np.zeros((self.H, self.W), np.int64)
        self.ctx.check(self.ctx.lib.edsgpu_frames_read_accumulator(self.ctx.h, self.h, C.c_int(slot), _ptr(acc, C.c_int64)))
        return acc

    def close(self):
        if self.h:
            self.ctx.lib.edsgpu_frames_destroy(self.h)
            self.h = None


class EventFrame:
    """Mirror of eds::tracking::EventFrame (EventFrame.hpp:31-107) for pyramid level 0.

    create() fills `norm`, `time`, `delta_time` and (on request) `event_frame`, like the public
    members of the reference class (EventFrame.hpp:41,57-61); the frame itself stays in `frames[slot]`.
    """

    def __init__(self, ctx, H, W, mapx=None, mapy=None, frames=None, slot=0):
        self.ctx, self.H, self.W = ctx, H, W
        self.lut = Lut(ctx, H, W, mapx, mapy) if mapx is not None else None
        self.frames = frames or Frames(ctx, H, W, 1)
        self.slot = slot
        self.norm = self.time = self.delta_time = None
        self.event_frame = None

    def create(self, x, y, polarity, ts_us=None, mode=DRAW_BILINEAR, use_exp_weights=True, sigma=0.5, want_host_frame=False):
        x = np.ascontiguousarray(x, np.uint16)
        y = np.ascontiguousarray(y, np.uint16)
        p = np.ascontiguousarray(polarity, np.uint8)
        ts = np.ascontiguousarray(ts_us, np.int64) if ts_us is not None else None
        norm, t, d = C.c_double(0), C.c_int64(0), C.c_int64(0)
        host = np.zeros((self.H, self.W), np.float64) if want_host_frame else None
        self.ctx.check(self.ctx.lib.edsgpu_event_frame_create(
            self.ctx.h, self.frames.h, C.c_int(self.slot), self.lut.h if self.lut else None, _ptr(x, C.c_uint16),
            _ptr(y, C.c_uint16), _ptr(p, C.c_uint8), _ptr(ts, C.c_int64), C.c_int(len(x)), C.c_int(mode),
            C.c_int(int(use_exp_weights)), C.c_float(sigma), C.byref(norm), C.byref(t), C.byref(d), _ptr(host, C.c_double)))
        self.norm, self.time, self.delta_time, self.event_frame = norm.value, t.value, d.value, host
        return self


class KeyFrame:
    """Device copy of the KeyFrame arrays the tracker gathers (KeyFrame.hpp:59-96)."""

    def __init__(self, ctx, kf, num_blocks):
        self.ctx = ctx
        self.N = len(kf["idp"])
        self.num_blocks = num_blocks
        g = np.ascontiguousarray(kf["grad"], np.float64)
        nc = np.ascontiguousarray(kf["norm_coord"], np.float64)
        idp = np.ascontiguousarray(kf["idp"], np.float64)
        w = np.ascontiguousarray(kf["weights"], np.float64)
        self.h = C.c_void_p()
        ctx.check(ctx.lib.edsgpu_keyframe_create(ctx.h, C.c_int(self.N), _ptr(g, C.c_double), _ptr(nc, C.c_double),
                                                 _ptr(idp, C.c_double), _ptr(w, C.c_double), C.c_int(kf["H"]), C.c_int(kf["W"]),
                                                 C.c_double(kf["fx"]), C.c_double(kf["fy"]), C.c_double(kf["cx"]),
                                                 C.c_double(kf["cy"]), C.c_int(num_blocks), C.byref(self.h)))

    def close(self):
        if self.h:
            self.ctx.lib.edsgpu_keyframe_destroy(self.h)
            self.h = None


class Tracker:
    """Mirror of eds::tracking::Tracker (Tracker.hpp:36-114): optimize() + state accessors."""

    def __init__(self, ctx, num_blocks=8, loss_type=LOSS_HUBER, loss_param=0.05, max_iterations=30,
                 function_tolerance=1e-6, gradient_tolerance=1e-8, parameter_tolerance=1e-6,
                 loss_param_method=LOSS_PARAM_MAD):
        self.ctx = ctx
        self.cfg = TrackerConfig(num_blocks, loss_type, max_iterations, loss_param_method, function_tolerance,
                                 gradient_tolerance, parameter_tolerance)
        self.h = C.c_void_p()
        ctx.check(ctx.lib.edsgpu_tracker_create(ctx.h, C.byref(self.cfg), C.c_double(loss_param), C.byref(self.h)))
        self.info = None

    def set_state(self, px=None, qx=None, vx=None, loss_param=None):
        arrs = [np.ascontiguousarray(a, np.float64) if a is not None else None for a in (px, qx, vx)]
        lp = C.c_double(loss_param) if loss_param is not None else None
        self.ctx.check(self.ctx.lib.edsgpu_tracker_set_state(self.h, *[_ptr(a, C.c_double) for a in arrs],
                                                             C.byref(lp) if lp is not None else None))

    def get_state(self):
        px, qx, vx = np.zeros(3), np.zeros(4), np.zeros(6)
        lp = C.c_double(0)
        info = TrackerInfo()
        self.ctx.check(self.ctx.lib.edsgpu_tracker_get_state(self.h, _ptr(px, C.c_double), _ptr(qx, C.c_double),
                                                             _ptr(vx, C.c_double), C.byref(lp), C.byref(info)))
        return px, qx, vx, lp.value, info.as_dict()

    def set_level_iterations(self, max_num_iterations):
        """config.options.max_num_iterations: one cap per pyramid level (Tracker.cpp:139)"""
        a = (C.c_int * len(max_num_iterations))(*[int(v) for v in max_num_iterations])
        self.ctx.check(self.ctx.lib.edsgpu_tracker_set_level_iterations(self.h, a, C.c_int(len(max_num_iterations))))

    def optimize(self, kf, frames, slot=0, want_residuals=False, level=0):
        """bool Tracker::optimize(id, event_frame, T_kf_ef, MAD) (Tracker.cpp:104-241), id = level.
        Returns dict(usable, px, qx, vx, residuals, next_loss_param, info)."""
        px, qx, vx = np.zeros(3), np.zeros(4), np.zeros(6)
        res = np.zeros(kf.N) if want_residuals else None
        tau = C.c_double(0)
        info = TrackerInfo()
        st = self.ctx.lib.edsgpu_tracker_optimize_level(self.h, kf.h, frames.h, C.c_int(slot), C.c_int(level), _ptr(px, C.c_double),
                                                        _ptr(qx, C.c_double), _ptr(vx, C.c_double), _ptr(res, C.c_double),
                                                        C.byref(tau), C.byref(info))
        if st not in (OK, NOT_USABLE):
            self.ctx.check(st)
        self.info = info.as_dict()
        return dict(usable=(st == OK), px=px, qx=qx, vx=vx, x=np.concatenate([px, qx, vx]), residuals=res,
                    next_loss_param=tau.value, info=self.info)

    def state_dev_ptr(self):
        return int(self.ctx.lib.edsgpu_tracker_state_dev(self.h))

    def close(self):
        if self.h:
            self.ctx.lib.edsgpu_tracker_destroy(self.h)
            self.h = None


def event_frames_batch(ctx, frames, first_slot, count, x, y, pol, num_events, lut=None, mode=DRAW_BILINEAR,
                       use_exp_weights=True, sigma=0.5, want_norms=False):
    """Host arrays (count*num_events) -> `count` device frames (edsgpu_event_frame_create_batch)."""
    # the C ABI takes uint16 / uint16 / uint8 arrays: convert here rather than reinterpret whatever dtype came in
    x, y = np.ascontiguousarray(x, np.uint16), np.ascontiguousarray(y, np.uint16)
    pol = np.ascontiguousarray(pol, np.uint8)
    if not (len(x) == len(y) == len(pol) == count * num_events):
        raise ValueError("event_frames_batch: need count * num_events = %d events per array" % (count * num_events))
    norms = np.zeros(count) if want_norms else None
    ctx.check(ctx.lib.edsgpu_event_frame_create_batch(
        ctx.h, frames.h, C.c_int(first_slot), C.c_int(count), lut.h if lut else None, _ptr(x, C.c_uint16),
        _ptr(y, C.c_uint16), _ptr(pol, C.c_uint8), C.c_int(num_events), C.c_int(mode), C.c_int(int(use_exp_weights)),
        C.c_float(sigma), _ptr(norms, C.c_double)))
    return norms


def event_frames_batch_dev(ctx, frames, first_slot, count, x_ptr, y_ptr, pol_ptr, num_events, lut=None,
                           mode=DRAW_BILINEAR, use_exp_weights=True, sigma=0.5):
    """Device pointers (ints) -> `count` device frames, asynchronous."""
    ctx.check(ctx.lib.edsgpu_event_frame_create_batch_dev(
        ctx.h, frames.h, C.c_int(first_slot), C.c_int(count), lut.h if lut else None, C.c_void_p(x_ptr),
        C.c_void_p(y_ptr), C.c_void_p(pol_ptr), C.c_int(num_events), C.c_int(mode), C.c_int(int(use_exp_weights)),
        C.c_float(sigma)))


class TrackerBatch:
    """`count` independent trackers advanced by one launch (edsgpu_batch_*): tracker i runs against
    keyframe i and frame slot first_slot+i; descriptors are built once."""

    def __init__(self, ctx, trackers, keyframes, frames, first_slot=0, level=0):
        self.ctx, self.trackers, self.keyframes, self.frames = ctx, trackers, keyframes, frames
        n = len(trackers)
        self._tr = (C.c_void_p * n)(*[t.h for t in trackers])
        self._kf = (C.c_void_p * n)(*[k.h for k in keyframes])
        self.count = n
        self.h = C.c_void_p()
        ctx.check(ctx.lib.edsgpu_batch_create_level(ctx.h, self._tr, self._kf, C.c_int(n), frames.h, C.c_int(first_slot), C.c_int(level),
                                                    C.byref(self.h)))

    def optimize(self):
        self.ctx.check(self.ctx.lib.edsgpu_batch_optimize(self.h))

    def launch_shape(self):
        """(evaluator CTAs, leader CTAs, problems in flight) of the batched launch."""
        a, b, c = C.c_int(0), C.c_int(0), C.c_int(0)
        self.ctx.check(self.ctx.lib.edsgpu_batch_launch_shape(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def pack_states_dev(self, dev_ptr):
        self.ctx.check(self.ctx.lib.edsgpu_batch_pack_states_dev(self.h, C.c_void_p(dev_ptr)))

    def gather(self, want_infos=True):
        states = np.zeros((self.count, 14))
        infos = (TrackerInfo * self.count)() if want_infos else None
        self.ctx.check(self.ctx.lib.edsgpu_trackers_gather(self.ctx.h, self._tr, C.c_int(self.count),
                                                           _ptr(states, C.c_double), infos))
        return states, ([i.as_dict() for i in infos] if want_infos else None)

    def close(self):
        if self.h:
            self.ctx.lib.edsgpu_batch_destroy(self.h)
            self.h = None


def tracker_evaluate(ctx, kf, frames, slot, x, loss_type=LOSS_HUBER, loss_param=0.05, want_jacobian=True):
    """ceres::CostFunction::Evaluate-style probe (edsgpu_tracker_evaluate)."""
    x = np.ascontiguousarray(x, np.float64)
    px, qx, vx = x[:3].copy(), x[3:7].copy(), x[7:].copy()
    res = np.zeros(kf.N)
    jac = np.zeros((kf.N, 12)) if want_jacobian else None
    cost = C.c_double(0)
    H, g = np.zeros((12, 12)), np.zeros(12)
    ctx.check(ctx.lib.edsgpu_tracker_evaluate(ctx.h, kf.h, frames.h, C.c_int(slot), C.c_int(loss_type), C.c_double(loss_param),
                                              _ptr(px, C.c_double), _ptr(qx, C.c_double), _ptr(vx, C.c_double),
                                              _ptr(res, C.c_double), _ptr(jac, C.c_double), C.byref(cost),
                                              _ptr(H, C.c_double), _ptr(g, C.c_double)))
    return dict(residuals=res, jacobian=jac, cost=cost.value, H=H, g=g)


class BaWindow:
    """Mirror of the accumulator side of dso::EnergyFunctional (accumulateAF_MT / LF_MT / SCF_MT,
    EnergyFunctional.cpp:197-261) on top of edsgpu_ba_*: the residual graph is given once, the
    per-linearisation records per call."""

    def __init__(self, ctx, F, host_idx, target_idx, res_begin):
        self.ctx, self.F = ctx, F
        h = np.ascontiguousarray(host_idx, np.int32)
        t = np.ascontiguousarray(target_idx, np.int32)
        rb = np.ascontiguousarray(res_begin, np.int32)
        self.R, self.P = len(h), len(rb) - 1
        self.n = 4 + 8 * F
        self.h = C.c_void_p()
        ctx.check(ctx.lib.edsgpu_ba_create(ctx.h, C.c_int(F), C.c_int(self.P), C.c_int(self.R), _ptr(h, C.c_int32),
                                           _ptr(t, C.c_int32), _ptr(rb, C.c_int32), C.byref(self.h)))

    def set_residuals(self, recs, flags, res_toZero=None):
        recs = np.ascontiguousarray(recs, np.float32)
        fl = np.ascontiguousarray(flags, np.uint8)
        rtz = np.ascontiguousarray(res_toZero, np.float32) if res_toZero is not None else None
        assert recs.shape == (self.R, 76) and fl.shape == (self.R,)
        self.ctx.check(self.ctx.lib.edsgpu_ba_set_residuals(self.h, _ptr(recs, C.c_float), _ptr(fl, C.c_uint8), _ptr(rtz, C.c_float)))

    # ---- the feeder on the device: PointFrameResidual::linearize (Residuals.cpp:69-265) ----
    def set_images(self, dI):
        """dI: (F, H, W, 3) float32, FrameHessian::dI of every frame of the window."""
        dI = np.ascontiguousarray(dI, np.float32)
        assert dI.ndim == 4 and dI.shape[0] == self.F and dI.shape[3] == 3
        for f in range(self.F):
            self.ctx.check(self.ctx.lib.edsgpu_ba_set_image(self.h, C.c_int(f), C.c_int(dI.shape[1]), C.c_int(dI.shape[2]), _ptr(dI[f], C.c_float)))

    def set_linearize_inputs(self, precalc, calib, frame_energy_th, u, v, idepth_zero_scaled, idepth_scaled, color, weights):
        f32 = lambda a: np.ascontiguousarray(a, np.float32)  # noqa: E731
        a = [f32(x) for x in (precalc, calib, frame_energy_th, u, v, idepth_zero_scaled, idepth_scaled, color, weights)]
        assert a[0].shape == (self.F * self.F, 28) and a[1].shape == (4,) and a[7].shape == (self.P, 8) and a[8].shape == (self.P, 8)
        self.ctx.check(self.ctx.lib.edsgpu_ba_set_linearize_inputs(self.h, *[_ptr(x, C.c_float) for x in a]))

    def linearize(self, state_in=None, linearized=None, res_toZero=None, want_outputs=True):
        """-> (state_new, energy_new) or None; the records, JpJdF and flags stay on the device."""
        u8 = lambda a: np.ascontiguousarray(a, np.uint8) if a is not None else None  # noqa: E731
        si, li = u8(state_in), u8(linearized)
        rtz = np.ascontiguousarray(res_toZero, np.float32) if res_toZero is not None else None
        st = np.zeros(self.R, np.int32) if want_outputs else None
        en = np.zeros(self.R, np.float32) if want_outputs else None
        self.ctx.check(self.ctx.lib.edsgpu_ba_linearize(self.h, _ptr(si, C.c_uint8), _ptr(li, C.c_uint8), _ptr(rtz, C.c_float),
                                                        _ptr(st, C.c_int32), _ptr(en, C.c_float)))
        return (st, en) if want_outputs else None

    def linearize_accumulate(self, state_in=None, linearized=None, res_toZero=None, write_records=False, want_outputs=True):
        """linearize fused with top_accumulate(0): one kernel, the 304-byte records of the residuals nothing reads later are
        not written.  -> (state_new, energy_new) or None; read the accumulation with top_result(0)."""
        u8 = lambda a: np.ascontiguousarray(a, np.uint8) if a is not None else None  # noqa: E731
        si, li = u8(state_in), u8(linearized)
        rtz = np.ascontiguousarray(res_toZero, np.float32) if res_toZero is not None else None
        st = np.zeros(self.R, np.int32) if want_outputs else None
        en = np.zeros(self.R, np.float32) if want_outputs else None
        self.ctx.check(self.ctx.lib.edsgpu_ba_linearize_accumulate(self.h, _ptr(si, C.c_uint8), _ptr(li, C.c_uint8), _ptr(rtz, C.c_float),
                                                                   C.c_int(1 if write_records else 0), _ptr(st, C.c_int32), _ptr(en, C.c_float)))
        return (st, en) if want_outputs else None

    # ---- after the solve (EnergyFunctional.cpp:263-415, EnergyFunctionalStructs.cpp:87-113) ----
    def resubstitute(self, x):
        """resubstituteF_MT: per-point step from the solved update x (4 + 8F)."""
        x = np.ascontiguousarray(x, np.float64)
        assert x.shape == (self.n,)
        step = np.zeros(self.P, np.float32)
        self.ctx.check(self.ctx.lib.edsgpu_ba_resubstitute(self.h, _ptr(x, C.c_double), _ptr(step, C.c_float)))
        return step

    def solve_system(self, lam=1e-5, HM=None, bM=None, delta=None, cPrior=None, frame_prior=None, frame_delta_prior=None, projector=None,
                     want_step=True):
        """solveSystemF (default solver mode) + resubstituteF_MT on the device -> (x, point steps)"""
        f64 = lambda a, order="C": np.ascontiguousarray(np.asarray(a, np.float64).ravel(order=order)) if a is not None else None  # noqa: E731
        hm, pr = f64(HM, "F"), f64(projector, "F")
        bm, dl, cp, fp, fd = f64(bM), f64(delta), f64(cPrior), f64(frame_prior), f64(frame_delta_prior)
        x = np.zeros(self.n)
        step = np.zeros(self.P, np.float32) if want_step else None
        self.ctx.check(self.ctx.lib.edsgpu_ba_solve_system(self.h, C.c_double(lam), _ptr(hm, C.c_double), _ptr(bm, C.c_double), _ptr(dl, C.c_double),
                                                           _ptr(cp, C.c_double), _ptr(fp, C.c_double), _ptr(fd, C.c_double), _ptr(pr, C.c_double),
                                                           _ptr(x, C.c_double), _ptr(step, C.c_float)))
        return x, step

    def fix_linearization(self, select=None, want_output=True):
        sel = np.ascontiguousarray(select, np.uint8) if select is not None else None
        out = np.zeros((self.R, 8), np.float32) if want_output else None
        self.ctx.check(self.ctx.lib.edsgpu_ba_fix_linearization(self.h, _ptr(sel, C.c_uint8), _ptr(out, C.c_float)))
        return out

    def calc_l_energy(self, cPrior=None, frame_prior=None, frame_delta_prior=None):
        f64 = lambda a: np.ascontiguousarray(a, np.float64) if a is not None else None  # noqa: E731
        cp, fp, fd = f64(cPrior), f64(frame_prior), f64(frame_delta_prior)
        e = C.c_double(0.0)
        self.ctx.check(self.ctx.lib.edsgpu_ba_calc_l_energy(self.h, _ptr(cp, C.c_double), _ptr(fp, C.c_double), _ptr(fd, C.c_double), C.byref(e)))
        return e.value

    def get_residuals(self):
        recs, flags = np.zeros((self.R, 76), np.float32), np.zeros(self.R, np.uint8)
        self.ctx.check(self.ctx.lib.edsgpu_ba_get_residuals(self.h, _ptr(recs, C.c_float), _ptr(flags, C.c_uint8)))
        return recs, flags

    def set_points(self, deltaF=None, priorF=None):
        d = np.ascontiguousarray(deltaF, np.float32) if deltaF is not None else None
        p = np.ascontiguousarray(priorF, np.float32) if priorF is not None else None
        self.ctx.check(self.ctx.lib.edsgpu_ba_set_points(self.h, _ptr(d, C.c_float), _ptr(p, C.c_float)))

    def set_frames(self, adHTdeltaF=None, cDeltaF=None, adHost=None, adTarget=None):
        a = np.ascontiguousarray(adHTdeltaF, np.float32) if adHTdeltaF is not None else None
        c = np.ascontiguousarray(cDeltaF, np.float32) if cDeltaF is not None else None
        ah = np.ascontiguousarray(adHost, np.float64) if adHost is not None else None
        at = np.ascontiguousarray(adTarget, np.float64) if adTarget is not None else None
        self.ctx.check(self.ctx.lib.edsgpu_ba_set_frames(self.h, _ptr(a, C.c_float), _ptr(c, C.c_float), _ptr(ah, C.c_double),
                                                         _ptr(at, C.c_double)))

    def top_accumulate(self, mode, want_outputs=True):
        if not want_outputs:
            self.ctx.check(self.ctx.lib.edsgpu_ba_top_accumulate(self.h, C.c_int(mode), None, None, None, None, None))
            return None
        acc = np.zeros((self.F * self.F, 13, 13))
        Hdd, bd, Hcd = np.zeros(self.P, np.float32), np.zeros(self.P, np.float32), np.zeros((self.P, 4), np.float32)
        nres = C.c_int64(0)
        self.ctx.check(self.ctx.lib.edsgpu_ba_top_accumulate(self.h, C.c_int(mode), _ptr(acc, C.c_double), _ptr(Hdd, C.c_float),
                                                             _ptr(bd, C.c_float), _ptr(Hcd, C.c_float), C.byref(nres)))
        return dict(acc=acc, Hdd=Hdd, bd=bd, Hcd=Hcd, nres=nres.value)

    def top_result(self, which):
        """the accumulation of side `which` (0 active, 1 linearized) that is already on the device"""
        acc = np.zeros((self.F * self.F, 13, 13))
        Hdd, bd, Hcd = np.zeros(self.P, np.float32), np.zeros(self.P, np.float32), np.zeros((self.P, 4), np.float32)
        nres = C.c_int64(0)
        self.ctx.check(self.ctx.lib.edsgpu_ba_top_read(self.h, C.c_int(which), _ptr(acc, C.c_double), _ptr(Hdd, C.c_float), _ptr(bd, C.c_float),
                                                       _ptr(Hcd, C.c_float), C.byref(nres)))
        return dict(acc=acc, Hdd=Hdd, bd=bd, Hcd=Hcd, nres=nres.value)

    def top_stitch(self, which, use_prior=False, cPrior=None, frame_prior=None, frame_delta_prior=None):
        H = np.zeros((self.n, self.n), order="F")
        b = np.zeros(self.n)
        cp = np.ascontiguousarray(cPrior, np.float64) if cPrior is not None else None
        fp = np.ascontiguousarray(frame_prior, np.float64) if frame_prior is not None else None
        fd = np.ascontiguousarray(frame_delta_prior, np.float64) if frame_delta_prior is not None else None
        self.ctx.check(self.ctx.lib.edsgpu_ba_top_stitch(self.h, C.c_int(which), C.c_int(int(use_prior)), _ptr(cp, C.c_double),
                                                         _ptr(fp, C.c_double), _ptr(fd, C.c_double), _ptr(H, C.c_double), _ptr(b, C.c_double)))
        return H, b

    def sc_accumulate(self, shift_prior_to_zero=True, want_outputs=True):
        if not want_outputs:
            self.ctx.check(self.ctx.lib.edsgpu_ba_sc_accumulate(self.h, C.c_int(int(shift_prior_to_zero)), None, None, None, None, None, None, None))
            return None
        F = self.F
        accD, accE, accEB = np.zeros((F ** 3, 8, 8)), np.zeros((F * F, 8, 4)), np.zeros((F * F, 8))
        accHcc, accbc = np.zeros((4, 4)), np.zeros(4)
        HdiF, bdSum = np.zeros(self.P, np.float32), np.zeros(self.P, np.float32)
        self.ctx.check(self.ctx.lib.edsgpu_ba_sc_accumulate(self.h, C.c_int(int(shift_prior_to_zero)), _ptr(accD, C.c_double),
                                                            _ptr(accE, C.c_double), _ptr(accEB, C.c_double), _ptr(accHcc, C.c_double),
                                                            _ptr(accbc, C.c_double), _ptr(HdiF, C.c_float), _ptr(bdSum, C.c_float)))
        return dict(accD=accD, accE=accE, accEB=accEB, accHcc=accHcc, accbc=accbc, HdiF=HdiF, bdSum=bdSum)

    def sc_stitch(self):
        H = np.zeros((self.n, self.n), order="F")
        b = np.zeros(self.n)
        self.ctx.check(self.ctx.lib.edsgpu_ba_sc_stitch(self.h, _ptr(H, C.c_double), _ptr(b, C.c_double)))
        return H, b

    def jpjd(self):
        out = np.zeros((self.R, 8), np.float32)
        self.ctx.check(self.ctx.lib.edsgpu_ba_get_jpjd(self.h, _ptr(out, C.c_float)))
        return out

    def close(self):
        if self.h:
            self.ctx.lib.edsgpu_ba_destroy(self.h)
            self.h = None


class CoarseTracker:
    """Evaluation side of dso::CoarseTracker (src/tracking/CoarseTracker.cpp): calcRes fused with
    calcGSSSE per pyramid level on the device; the Gauss-Newton loop of trackNewestCoarse stays with
    the caller."""

    def __init__(self, ctx, num_levels):
        self.ctx, self.num_levels = ctx, num_levels
        self.h = C.c_void_p()
        ctx.check(ctx.lib.edsgpu_coarse_create(ctx.h, C.c_int(num_levels), C.byref(self.h)))

    def set_level(self, lvl, width, height, fx, fy, cx, cy, Ki):
        ki = np.ascontiguousarray(Ki, np.float32).reshape(-1)
        assert ki.shape == (9,)
        self.ctx.check(self.ctx.lib.edsgpu_coarse_set_level(self.h, C.c_int(lvl), C.c_int(width), C.c_int(height), C.c_float(fx), C.c_float(fy),
                                                            C.c_float(cx), C.c_float(cy), _ptr(ki, C.c_float)))

    def set_reference(self, lvl, pc_u, pc_v, pc_idepth, pc_color):
        a = [np.ascontiguousarray(x, np.float32) for x in (pc_u, pc_v, pc_idepth, pc_color)]
        self.ctx.check(self.ctx.lib.edsgpu_coarse_set_reference(self.h, C.c_int(lvl), C.c_int(len(a[0])), *[_ptr(x, C.c_float) for x in a]))

    def set_reference_frame(self, lvl, dI):
        """lastRef->dIp[lvl] (h, w, 3): makeCoarseDepthL0 takes the point colours from it"""
        d = np.ascontiguousarray(dI, np.float32)
        self.ctx.check(self.ctx.lib.edsgpu_coarse_set_reference_frame(self.h, C.c_int(lvl), _ptr(d, C.c_float)))

    def make_depth_l0(self, levels_used, proj_u, proj_v, proj_idepth, HdiF):
        """CoarseTracker::makeCoarseDepthL0 on the device -> pc_n per level; the levels' reference point clouds are in place"""
        a = [np.ascontiguousarray(x, np.float32) for x in (proj_u, proj_v, proj_idepth, HdiF)]
        n = (C.c_int * levels_used)()
        self.ctx.check(self.ctx.lib.edsgpu_coarse_make_depth_l0(self.h, C.c_int(levels_used), C.c_int(len(a[0])), *[_ptr(x, C.c_float) for x in a], n))
        return list(n)

    def get_reference(self, lvl):
        n = C.c_int(0)
        self.ctx.check(self.ctx.lib.edsgpu_coarse_get_reference(self.h, C.c_int(lvl), C.byref(n), None, None, None, None))
        out = [np.zeros(n.value, np.float32) for _ in range(4)]
        self.ctx.check(self.ctx.lib.edsgpu_coarse_get_reference(self.h, C.c_int(lvl), C.byref(n), *[_ptr(x, C.c_float) for x in out]))
        return dict(n=n.value, pc_u=out[0], pc_v=out[1], pc_idepth=out[2], pc_color=out[3])

    def set_new_frame(self, lvl, dI):
        d = np.ascontiguousarray(dI, np.float32)
        self.ctx.check(self.ctx.lib.edsgpu_coarse_set_new_frame(self.h, C.c_int(lvl), _ptr(d, C.c_float)))

    def calc_res_gs(self, lvl, R, t, affLL, b0, cutoffTH, want_system=True):
        R = np.ascontiguousarray(R, np.float64).reshape(-1)
        t = np.ascontiguousarray(t, np.float64)
        aff = np.ascontiguousarray(affLL, np.float32)
        rs = np.zeros(6)
        H = np.zeros((8, 8)) if want_system else None
        b = np.zeros(8) if want_system else None
        self.ctx.check(self.ctx.lib.edsgpu_coarse_calc_res_gs(self.h, C.c_int(lvl), _ptr(R, C.c_double), _ptr(t, C.c_double), _ptr(aff, C.c_float),
                                                              C.c_float(b0), C.c_float(cutoffTH), _ptr(rs, C.c_double), _ptr(H, C.c_double),
                                                              _ptr(b, C.c_double)))
        return dict(rs=rs, H=H, b=b)

    def track(self, coarsest_lvl, R, t, aff=(0.0, 0.0), ref_aff=(0.0, 0.0), ref_exposure=1.0, new_exposure=1.0, min_res_for_abort=None):
        """trackNewestCoarse -> dict(ok, R, t, aff, last_residuals, last_flow, evaluations)."""
        Rm = np.array(R, np.float64).reshape(-1).copy()
        tv, af = np.array(t, np.float64).copy(), np.array(aff, np.float64).copy()
        raf = np.ascontiguousarray(ref_aff, np.float64)
        mra = np.ascontiguousarray(min_res_for_abort, np.float64) if min_res_for_abort is not None else None
        lr, lf, ev = np.zeros(5), np.zeros(3), C.c_int(0)
        st = self.ctx.lib.edsgpu_coarse_track(self.h, C.c_int(coarsest_lvl), _ptr(Rm, C.c_double), _ptr(tv, C.c_double), _ptr(af, C.c_double),
                                              _ptr(raf, C.c_double), C.c_float(ref_exposure), C.c_float(new_exposure), _ptr(mra, C.c_double),
                                              _ptr(lr, C.c_double), _ptr(lf, C.c_double), C.byref(ev))
        if st not in (OK, NOT_USABLE):
            self.ctx.check(st)
        return dict(ok=(st == OK), R=Rm.reshape(3, 3), t=tv, aff=af, last_residuals=lr, last_flow=lf, evaluations=ev.value)

    def close(self):
        if self.h:
            self.ctx.lib.edsgpu_coarse_destroy(self.h)
            self.h = None


class DepthPoints:
    """eds::mapping::DepthPoints (src/mapping/DepthPoints.{hpp,cpp}): per-point Vogiatzis depth filter with
    the state {mu, sigma2, a, b} resident on the device."""

    def __init__(self, ctx, num_points, fx, fy, cx, cy, min_depth, max_depth, inv_depth=None, init_a=10.0, init_b=10.0):
        self.ctx, self.N = ctx, num_points
        idp = np.ascontiguousarray(inv_depth, np.float64) if inv_depth is not None else None
        assert idp is None or idp.shape == (num_points,)
        self.h = C.c_void_p()
        ctx.check(ctx.lib.edsgpu_depth_points_create(ctx.h, C.c_int(num_points), C.c_double(fx), C.c_double(fy), C.c_double(cx), C.c_double(cy),
                                                     C.c_double(min_depth), C.c_double(max_depth), _ptr(idp, C.c_double), C.c_double(init_a),
                                                     C.c_double(init_b), C.byref(self.h)))

    def update(self, T_kf_ef, kf_coord, ef_coord, coords_are_tracks=False):
        T = np.ascontiguousarray(T_kf_ef, np.float64).reshape(-1)
        kf, ef = np.ascontiguousarray(kf_coord, np.float64), np.ascontiguousarray(ef_coord, np.float64)
        assert T.shape == (16,) and kf.shape == (self.N, 2) and ef.shape == (self.N, 2)
        ok = np.zeros(self.N, np.uint8)
        self.ctx.check(self.ctx.lib.edsgpu_depth_points_update(self.h, _ptr(T, C.c_double), _ptr(kf, C.c_double), _ptr(ef, C.c_double),
                                                               C.c_int(int(coords_are_tracks)), _ptr(ok, C.c_uint8)))
        return ok

    def get(self):
        st = np.zeros((self.N, 4))
        self.ctx.check(self.ctx.lib.edsgpu_depth_points_get(self.h, _ptr(st, C.c_double)))
        return st

    def get_coord(self, tracker, keyframe):
        """Tracker::getCoord with this filter's inverse depths -> (coord N x 2, outlier flags)."""
        coord, out = np.zeros((self.N, 2)), np.zeros(self.N, np.uint8)
        self.ctx.check(self.ctx.lib.edsgpu_tracker_get_coord(tracker.h, keyframe.h, self.h, _ptr(coord, C.c_double), _ptr(out, C.c_uint8)))
        return coord, out

    def update_from_tracker(self, tracker, keyframe, kf_coord=None, refresh_keyframe=True):
        """getCoord -> update -> key-frame refresh, on the device (one call per tracked window)."""
        kc = np.ascontiguousarray(kf_coord, np.float64) if kf_coord is not None else None
        assert kc is None or kc.shape == (self.N, 2)
        self.ctx.check(self.ctx.lib.edsgpu_depth_points_update_from_tracker(self.h, tracker.h, keyframe.h, _ptr(kc, C.c_double),
                                                                            C.c_int(int(refresh_keyframe))))

    def close(self):
        if self.h:
            self.ctx.lib.edsgpu_depth_points_destroy(self.h)
            self.h = None


class Comm:
    """One rank's end of the NCCL communicator used for the final gather (edsgpu_comm_*)."""

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(128)
        st = load_nccl().edsgpu_comm_unique_id(buf)
        if st != OK:
            raise EdsGpuError(st, "ncclGetUniqueId failed")
        return buf.raw

    def __init__(self, ctx, world_size, rank, unique_id):
        self.ctx, self.world, self.rank = ctx, world_size, rank
        self.lib = load_nccl()
        self.h = C.c_void_p()
        ctx.check(self.lib.edsgpu_comm_create(ctx.h, C.c_int(world_size), C.c_int(rank), C.c_char_p(unique_id), C.byref(self.h)))

    def gather_states_dev(self, local_ptr, n_local, num_sequences, out_ptr):
        """Device pointers (ints): [n_local x 14] -> [num_sequences x 14] in global order; asynchronous."""
        self.ctx.check(self.lib.edsgpu_gather_states_nccl(self.h, C.c_void_p(local_ptr), C.c_int(n_local), C.c_int(num_sequences), C.c_void_p(out_ptr)))

    def gather_batch(self, batch, num_sequences):
        """States of a TrackerBatch's trackers from every rank -> [num_sequences x 14] host array (synchronises)."""
        out = np.zeros((num_sequences, 14))
        self.ctx.check(self.lib.edsgpu_batch_gather_states_nccl(batch.h, self.h, C.c_int(num_sequences), _ptr(out, C.c_double)))
        return out

    def close(self):
        if self.h:
            self.lib.edsgpu_comm_destroy(self.h)
            self.h = None
