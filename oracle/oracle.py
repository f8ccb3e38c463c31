"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product (slam-eds_b200/) never does.
PARITY UNPINNED: see the header of eds_oracle_tracking.cpp.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libeds_oracle.so")
BUILD_FLAGS = "-O3 -march=x86-64-v3 -ffp-contract=off"
_lib = None

REC = 76  # floats per RawResidualJacobian record (304 B)


def build(force=False):
    """Compile the oracle with the committed Makefile (g++ only)."""
    srcs = [os.path.join(_HERE, f) for f in ("eds_oracle_tracking.cpp", "eds_oracle_ba.cpp", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return _LIB_PATH


class SolverConfig(C.Structure):
    _fields_ = [("num_blocks", C.c_int), ("loss_type", C.c_int), ("loss_param", C.c_double),
                ("max_iterations", C.c_int), ("function_tolerance", C.c_double),
                ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double),
                ("jacobian_mode", C.c_int), ("threads", C.c_int)]


class SolverInfo(C.Structure):
    _fields_ = [("iterations", C.c_int), ("successful_steps", C.c_int), ("unsuccessful_steps", C.c_int),
                ("initial_cost", C.c_double), ("final_cost", C.c_double), ("usable", C.c_int),
                ("termination", C.c_int), ("solve_time_us", C.c_double), ("final_radius", C.c_double)]


def use_native_build():
    """bench.py's CPU arms: switch to a build tuned for THIS machine (-O3 -march=native, contraction allowed), compiled on
    the spot into a per-host directory.  Must be called before the first use; returns the flags in effect."""
    global _LIB_PATH, BUILD_FLAGS
    assert _lib is None, "oracle already loaded"
    import platform
    d = os.path.join("_build", "native_" + platform.node().replace("/", "_"))
    try:
        subprocess.run(["make", "-C", _HERE, "native", "NATIVE_DIR=" + d], check=True, capture_output=True)
        C.CDLL(os.path.join(_HERE, d, "libeds_oracle.so"))
        _LIB_PATH = os.path.join(_HERE, d, "libeds_oracle.so")
        BUILD_FLAGS = "-O3 -march=native -ffp-contract=fast"
    except (subprocess.CalledProcessError, OSError):
        pass  # keep the default build
    return BUILD_FLAGS


def lib():
    global _lib
    if _lib is None:
        build()
        try:
            _lib = C.CDLL(_LIB_PATH)
        except OSError:
            build(force=True)
            _lib = C.CDLL(os.path.join(_HERE, "_build", "libeds_oracle.so"))
        _lib.eds_oracle_mad_tau.restype = C.c_double
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def draw_values(px, py, val, H, W, method="bilinear", use_exp=True, sigma=0.5):
    px, py = _f64(px), _f64(py)
    val = np.ascontiguousarray(val, dtype=np.int8)
    img = np.zeros((H, W), np.float64)
    rc = lib().eds_oracle_draw_values(_p(px, C.c_double), _p(py, C.c_double), _p(val, C.c_int8), C.c_int(len(px)),
                                      C.c_int(H), C.c_int(W), C.c_int(0 if method == "nn" else 1), C.c_int(int(use_exp)),
                                      C.c_double(float(np.float32(sigma))), _p(img, C.c_double))  # `const float s`, Utils.cpp:50
    assert rc == 0
    return img


def event_frame(x, y, pol, ts_us, H, W, mapx=None, mapy=None, method="bilinear", use_exp=True, sigma=0.5):
    """Returns dict(img, frame, norm, time, delta, status)."""
    x = np.ascontiguousarray(x, dtype=np.uint16)
    y = np.ascontiguousarray(y, dtype=np.uint16)
    pol = np.ascontiguousarray(pol, dtype=np.uint8)
    ts = np.ascontiguousarray(ts_us, dtype=np.int64) if ts_us is not None else None
    mx = _f32(mapx) if mapx is not None else None
    my = _f32(mapy) if mapy is not None else None
    img = np.zeros((H, W), np.float64)
    frame = np.zeros((H, W), np.float64)
    norm = C.c_double(0)
    t = C.c_int64(0)
    d = C.c_int64(0)
    rc = lib().eds_oracle_event_frame(_p(x, C.c_uint16), _p(y, C.c_uint16), _p(pol, C.c_uint8), _p(ts, C.c_int64),
                                      C.c_int(len(x)), C.c_int(H), C.c_int(W), _p(mx, C.c_float), _p(my, C.c_float),
                                      C.c_int(0 if method == "nn" else 1), C.c_int(int(use_exp)), C.c_double(float(np.float32(sigma))),
                                      _p(img, C.c_double), _p(frame, C.c_double), C.byref(norm), C.byref(t), C.byref(d))
    return dict(img=img, frame=frame, norm=norm.value, time=t.value, delta=d.value, status=rc)


def event_frame_levels(img, num_levels):
    """Pyramid of EventFrame::create (reference src/tracking/EventFrame.cpp:342-364) from the level-0 image `img`
    (un-normalised): frame[0] = img, frame[i] = cv::dilate(img, rect(2i+1)) + cv::erode(img, rect(2i+1)) at the same
    resolution, norm[i] = cv::norm(frame[i]) (L2).  The default border of cv::dilate / cv::erode keeps pixels outside
    the image out of the maximum / minimum: windows are clipped.  Arithmetic in the dtype of `img` (the reference: double).
    Pinned against cv2.dilate / cv2.erode in tests/test_oracle_tracking.py.  -> (frames, norms)"""
    img = np.asarray(img)
    H, W = img.shape
    frames, norms = [img.copy()], []
    for i in range(1, num_levels):
        hi = np.pad(img, i, constant_values=-np.inf)
        lo = np.pad(img, i, constant_values=np.inf)
        mx = np.full_like(img, -np.inf)
        mn = np.full_like(img, np.inf)
        for dy in range(2 * i + 1):
            for dx in range(2 * i + 1):
                mx = np.maximum(mx, hi[dy:dy + H, dx:dx + W])
                mn = np.minimum(mn, lo[dy:dy + H, dx:dx + W])
        frames.append(mx + mn)
    for f in frames:
        norms.append(float(np.sqrt(np.sum(np.asarray(f, np.float64) ** 2))))
    return frames, norms


def bicubic(grid, rows, cols):
    grid = _f64(grid)
    rows, cols = _f64(rows), _f64(cols)
    n = len(rows)
    f, dr, dc = np.zeros(n), np.zeros(n), np.zeros(n)
    lib().eds_oracle_bicubic(_p(grid, C.c_double), C.c_int(grid.shape[0]), C.c_int(grid.shape[1]), C.c_int(n),
                             _p(rows, C.c_double), _p(cols, C.c_double), _p(f, C.c_double), _p(dr, C.c_double), _p(dc, C.c_double))
    return f, dr, dc


def _kf_args(kf, frame):
    g, nc, idp, w = _f64(kf["grad"]), _f64(kf["norm_coord"]), _f64(kf["idp"]), _f64(kf["weights"])
    fr = _f64(frame)
    keep = (g, nc, idp, w, fr)
    args = [C.c_int(len(idp)), _p(g, C.c_double), _p(nc, C.c_double), _p(idp, C.c_double), _p(w, C.c_double),
            _p(fr, C.c_double), C.c_int(kf["H"]), C.c_int(kf["W"]), C.c_double(kf["fx"]), C.c_double(kf["fy"]),
            C.c_double(kf["cx"]), C.c_double(kf["cy"])]
    return args, keep


def tracker_evaluate(kf, frame, x, num_blocks, loss_type=1, loss_param=0.05, jacobian_mode=0, want_jac=True):
    """kf: dict(grad Nx2, norm_coord Nx2, idp N, weights N, H, W, fx, fy, cx, cy). x: 13 = p, q(xyzw), v."""
    args, keep = _kf_args(kf, frame)
    N = len(keep[2])
    x = _f64(x)
    res = np.zeros(N)
    jac = np.zeros((N, 12)) if want_jac else None
    cost = C.c_double(0)
    Hm, g, bs = np.zeros((12, 12)), np.zeros(12), np.zeros(num_blocks)
    rc = lib().eds_oracle_tracker_evaluate(*args, C.c_int(num_blocks), C.c_int(loss_type), C.c_double(loss_param),
                                           C.c_int(jacobian_mode), _p(x, C.c_double), _p(res, C.c_double), _p(jac, C.c_double),
                                           C.byref(cost), _p(Hm, C.c_double), _p(g, C.c_double), _p(bs, C.c_double))
    assert rc == 0, rc
    return dict(residuals=res, jacobian=jac, cost=cost.value, H=Hm, g=g, block_sqnorm=bs)


def tracker_solve(kf, frame, x, num_blocks=8, loss_type=1, loss_param=0.05, max_iterations=30, function_tolerance=1e-6,
                  gradient_tolerance=1e-8, parameter_tolerance=1e-6, jacobian_mode=0, threads=1, want_trace=False):
    args, keep = _kf_args(kf, frame)
    N = len(keep[2])
    x = _f64(x).copy()
    cfg = SolverConfig(num_blocks, loss_type, loss_param, max_iterations, function_tolerance, gradient_tolerance,
                       parameter_tolerance, jacobian_mode, threads)
    info = SolverInfo()
    res = np.zeros(N)
    tau = C.c_double(0)
    trace = np.zeros((max_iterations + 1, 16)) if want_trace else None
    rc = lib().eds_oracle_tracker_solve(*args, C.byref(cfg), _p(x, C.c_double), _p(res, C.c_double), C.byref(tau),
                                        C.byref(info), _p(trace, C.c_double))
    out = dict(status=rc, x=x, residuals=res, next_loss_param=tau.value,
               info={k: getattr(info, k) for k, _ in SolverInfo._fields_})
    if want_trace:
        out["trace"] = trace
    return out


def mad_tau(residuals):
    r = _f64(residuals).copy()
    return lib().eds_oracle_mad_tau(_p(r, C.c_double), C.c_int(len(r)))


# ----------------------------------------------------------------------------- BA
def ba_jpjd(recs):
    recs = _f32(recs)
    R = recs.shape[0]
    out = np.zeros((R, 8), np.float32)
    lib().eds_oracle_ba_jpjd(C.c_int(R), _p(recs, C.c_float), _p(out, C.c_float))
    return out


def ba_linearize(F, H, W, dI, precalc, calib, pu, pv, idepth_zero_scaled, idepth_scaled, color, weights, host_idx, target_idx,
                 res_begin, frame_energy_th, state_in=None):
    """PointFrameResidual::linearize for every residual -> (recs [R,76], state_new [R], energy_new [R])."""
    dI, precalc, calib = _f32(dI), _f32(precalc), _f32(calib)
    pu, pv, iz, isc = _f32(pu), _f32(pv), _f32(idepth_zero_scaled), _f32(idepth_scaled)
    color, weights, th = _f32(color), _f32(weights), _f32(frame_energy_th)
    h, t, rb = _i32(host_idx), _i32(target_idx), _i32(res_begin)
    R, P = len(h), len(rb) - 1
    st = None if state_in is None else np.ascontiguousarray(state_in, np.uint8)
    recs = np.zeros((R, 76), np.float32)
    state = np.zeros(R, np.int32)
    energy = np.zeros(R, np.float32)
    lib().eds_oracle_ba_linearize(C.c_int(F), C.c_int(P), C.c_int(R), C.c_int(H), C.c_int(W), _p(dI, C.c_float), _p(precalc, C.c_float),
                                  _p(calib, C.c_float), _p(pu, C.c_float), _p(pv, C.c_float), _p(iz, C.c_float), _p(isc, C.c_float),
                                  _p(color, C.c_float), _p(weights, C.c_float), _p(h, C.c_int32), _p(t, C.c_int32), _p(rb, C.c_int32),
                                  _p(st, C.c_uint8) if st is not None else None, _p(th, C.c_float), _p(recs, C.c_float),
                                  _p(state, C.c_int32), _p(energy, C.c_float))
    return recs, state, energy


def ba_fix_linearization(F, recs, host_idx, target_idx, res_begin, deltaF, adHTdeltaF, cDeltaF):
    recs = _f32(recs)
    R, P = recs.shape[0], len(res_begin) - 1
    out = np.zeros((R, 8), np.float32)
    h, t, rb = _i32(host_idx), _i32(target_idx), _i32(res_begin)
    d, a, c = _f32(deltaF), _f32(adHTdeltaF), _f32(cDeltaF)
    lib().eds_oracle_ba_fix_linearization(C.c_int(F), C.c_int(P), C.c_int(R), _p(recs, C.c_float), _p(h, C.c_int32),
                                          _p(t, C.c_int32), _p(rb, C.c_int32), _p(d, C.c_float), _p(a, C.c_float),
                                          _p(c, C.c_float), _p(out, C.c_float))
    return out


def ba_resubstitute(F, x, adHost, adTarget, host_idx, target_idx, res_begin, flags, JpJdF, bdSumF, Hcd_A, Hcd_L, HdiF):
    """resubstituteF_MT: per-point step (P floats). adHost/adTarget: F*F column-major 8x8 doubles."""
    h, t, rb = _i32(host_idx), _i32(target_idx), _i32(res_begin)
    R, P = len(h), len(rb) - 1
    x, ah, at = _f64(x), _f64(adHost), _f64(adTarget)
    fl = np.ascontiguousarray(flags, np.uint8)
    j, b, ca, cl, hd = _f32(JpJdF), _f32(bdSumF), _f32(Hcd_A), _f32(Hcd_L), _f32(HdiF)
    out = np.zeros(P, np.float32)
    lib().eds_oracle_ba_resubstitute(C.c_int(F), C.c_int(P), C.c_int(R), _p(x, C.c_double), _p(ah, C.c_double), _p(at, C.c_double),
                                     _p(h, C.c_int32), _p(t, C.c_int32), _p(rb, C.c_int32), _p(fl, C.c_uint8), _p(j, C.c_float),
                                     _p(b, C.c_float), _p(ca, C.c_float), _p(cl, C.c_float), _p(hd, C.c_float), _p(out, C.c_float))
    return out


def ba_nullspace_projector(nullspaces, delta=1e-5):
    """N (N^T N)^-1 N^T as EnergyFunctional::orthogonalize builds it (reference src/bundles/EnergyFunctional.cpp:738-760):
    columns normalised, pseudo-inverse by SVD with singular values below setting_solverModeDelta * max cut, symmetrised."""
    N = np.stack([np.asarray(v, np.float64) / np.linalg.norm(v) for v in nullspaces], axis=1)
    U, S, Vt = np.linalg.svd(N, full_matrices=False)
    Si = np.where(S > delta * S.max(), 1.0 / S, 0.0)
    Npi = U @ np.diag(Si) @ Vt
    NNpiT = N @ Npi.T
    return 0.5 * (NNpiT + NNpiT.T)


def ba_solve_system(HA, bA, HL, bL, Hsc, bsc, lam=1e-5, HM=None, bM=None, delta=None, projector=None):
    """EnergyFunctional::solveSystemF (reference src/bundles/EnergyFunctional.cpp:775-905) in the default solver mode
    (setting_solverMode = SOLVER_FIX_LAMBDA | SOLVER_ORTHOGONALIZE_X_LATER, settings.cpp:60): the non-orthogonalised, non-SVD
    branch (:838-893) and orthogonalize(&x, 0) (:898-902) when a projector is given.  numpy float64; np.linalg.solve (LU with
    partial pivoting) on the lower-triangle-symmetrised matrix stands for Eigen's pivoted LDLT (which reads the lower triangle):
    both are backward stable on the scaled SPD system.  -> x"""
    n = len(bA)
    HM = np.zeros((n, n)) if HM is None else np.asarray(HM, np.float64)
    bM = np.zeros(n) if bM is None else np.asarray(bM, np.float64)
    bM_top = bM + (HM @ np.asarray(delta, np.float64) if delta is not None else 0.0)
    H = np.asarray(HL, np.float64) + HM + np.asarray(HA, np.float64)
    b = np.asarray(bL, np.float64) + bM_top + np.asarray(bA, np.float64) - np.asarray(bsc, np.float64)
    H[np.diag_indices(n)] *= (1.0 + lam)
    H = H - np.asarray(Hsc, np.float64) * (1.0 / (1.0 + lam))
    sv = 1.0 / np.sqrt(np.diag(H) + 10.0)
    Hs = sv[:, None] * H * sv[None, :]
    # Eigen's LDLT reads the LOWER triangle only; H_sc is symmetric only up to float32 rounding (its float accumulators reach
    # D_ijk and D_ikj^T through different summation orders), so the triangle that is read matters at the 1e-6 level
    Hs = np.tril(Hs) + np.tril(Hs, -1).T
    x = sv * np.linalg.solve(Hs, sv * b)
    if projector is not None:
        x = x - np.asarray(projector, np.float64) @ x
    return x


def ba_calc_l_energy(F, recs, host_idx, target_idx, res_begin, flags, res_toZero, deltaF, priorF, adHTdeltaF, cDeltaF, cPrior=None,
                     frame_prior=None, frame_delta_prior=None):
    """calcLEnergyF_MT (double)."""
    recs = _f32(recs)
    h, t, rb = _i32(host_idx), _i32(target_idx), _i32(res_begin)
    R, P = len(h), len(rb) - 1
    fl = np.ascontiguousarray(flags, np.uint8)
    rtz, d, pr, a, c = _f32(res_toZero), _f32(deltaF), _f32(priorF), _f32(adHTdeltaF), _f32(cDeltaF)
    cp = _f64(cPrior) if cPrior is not None else None
    fp = _f64(frame_prior) if frame_prior is not None else None
    fd = _f64(frame_delta_prior) if frame_delta_prior is not None else None
    f = lib().eds_oracle_ba_calc_l_energy
    f.restype = C.c_double
    return f(C.c_int(F), C.c_int(P), C.c_int(R), _p(recs, C.c_float), _p(h, C.c_int32), _p(t, C.c_int32), _p(rb, C.c_int32),
             _p(fl, C.c_uint8), _p(rtz, C.c_float), _p(d, C.c_float), _p(pr, C.c_float), _p(a, C.c_float), _p(c, C.c_float),
             _p(cp, C.c_double) if cp is not None else None, _p(fp, C.c_double) if fp is not None else None,
             _p(fd, C.c_double) if fd is not None else None)


def ba_top_accumulate(mode, F, recs, host_idx, target_idx, res_begin, flags, res_toZero=None, deltaF=None,
                      adHTdeltaF=None, cDeltaF=None, threads=1):
    recs = _f32(recs)
    R, P = recs.shape[0], len(res_begin) - 1
    h, t, rb = _i32(host_idx), _i32(target_idx), _i32(res_begin)
    fl = np.ascontiguousarray(flags, dtype=np.uint8)
    rtz = _f32(res_toZero) if res_toZero is not None else np.zeros((R, 8), np.float32)
    d = _f32(deltaF) if deltaF is not None else np.zeros(P, np.float32)
    a = _f32(adHTdeltaF) if adHTdeltaF is not None else np.zeros((F * F, 8), np.float32)
    c = _f32(cDeltaF) if cDeltaF is not None else np.zeros(4, np.float32)
    acc = np.zeros((F * F, 13, 13))
    num = np.zeros(F * F, np.int64)
    Hdd, bd, Hcd = np.zeros(P, np.float32), np.zeros(P, np.float32), np.zeros((P, 4), np.float32)
    nres = C.c_int64(0)
    rc = lib().eds_oracle_ba_top_accumulate(C.c_int(mode), C.c_int(F), C.c_int(P), C.c_int(R), _p(recs, C.c_float),
                                            _p(h, C.c_int32), _p(t, C.c_int32), _p(rb, C.c_int32), _p(fl, C.c_uint8),
                                            _p(rtz, C.c_float), _p(d, C.c_float), _p(a, C.c_float), _p(c, C.c_float),
                                            C.c_int(threads), _p(acc, C.c_double), _p(num, C.c_int64), _p(Hdd, C.c_float),
                                            _p(bd, C.c_float), _p(Hcd, C.c_float), C.byref(nres))
    assert rc == 0
    return dict(acc=acc, num=num, Hdd=Hdd, bd=bd, Hcd=Hcd, nres=nres.value)


def ba_top_stitch(F, acc, adHost, adTarget, use_prior=False, cPrior=None, cDeltaF=None, frame_prior=None,
                  frame_delta_prior=None):
    n = 4 + 8 * F
    acc, ah, at = _f64(acc), _f64(adHost), _f64(adTarget)
    cp = _f64(cPrior) if cPrior is not None else np.zeros(4)
    cd = _f32(cDeltaF) if cDeltaF is not None else np.zeros(4, np.float32)
    fp = _f64(frame_prior) if frame_prior is not None else np.zeros((F, 8))
    fd = _f64(frame_delta_prior) if frame_delta_prior is not None else np.zeros((F, 8))
    H = np.zeros((n, n), order="F")
    b = np.zeros(n)
    lib().eds_oracle_ba_top_stitch(C.c_int(F), _p(acc, C.c_double), _p(ah, C.c_double), _p(at, C.c_double),
                                   C.c_int(int(use_prior)), _p(cp, C.c_double), _p(cd, C.c_float), _p(fp, C.c_double),
                                   _p(fd, C.c_double), _p(H, C.c_double), _p(b, C.c_double))
    return H, b


def ba_sc_accumulate(F, host_idx, target_idx, res_begin, flags, JpJdF, Hdd_A, Hdd_L, bd_A, bd_L, Hcd_A, Hcd_L, priorF,
                     deltaF, shift_prior_to_zero=True, threads=1):
    P = len(res_begin) - 1
    h, t, rb = _i32(host_idx), _i32(target_idx), _i32(res_begin)
    R = len(h)
    fl = np.ascontiguousarray(flags, dtype=np.uint8)
    arrs = [_f32(a) for a in (JpJdF, Hdd_A, Hdd_L, bd_A, bd_L, Hcd_A, Hcd_L, priorF, deltaF)]
    accD, accE, accEB = np.zeros((F ** 3, 8, 8)), np.zeros((F * F, 8, 4)), np.zeros((F * F, 8))
    accHcc, accbc = np.zeros((4, 4)), np.zeros(4)
    HdiF, bdSum = np.zeros(P, np.float32), np.zeros(P, np.float32)
    rc = lib().eds_oracle_ba_sc_accumulate(C.c_int(F), C.c_int(P), C.c_int(R), _p(h, C.c_int32), _p(t, C.c_int32),
                                           _p(rb, C.c_int32), _p(fl, C.c_uint8), *[_p(a, C.c_float) for a in arrs],
                                           C.c_int(int(shift_prior_to_zero)), C.c_int(threads), _p(accD, C.c_double),
                                           _p(accE, C.c_double), _p(accEB, C.c_double), _p(accHcc, C.c_double),
                                           _p(accbc, C.c_double), _p(HdiF, C.c_float), _p(bdSum, C.c_float))
    assert rc == 0
    return dict(accD=accD, accE=accE, accEB=accEB, accHcc=accHcc, accbc=accbc, HdiF=HdiF, bdSum=bdSum)


def ba_sc_stitch(F, sc, adHost, adTarget):
    n = 4 + 8 * F
    ah, at = _f64(adHost), _f64(adTarget)
    H = np.zeros((n, n), order="F")
    b = np.zeros(n)
    lib().eds_oracle_ba_sc_stitch(C.c_int(F), _p(_f64(sc["accD"]), C.c_double), _p(_f64(sc["accE"]), C.c_double),
                                  _p(_f64(sc["accEB"]), C.c_double), _p(_f64(sc["accHcc"]), C.c_double),
                                  _p(_f64(sc["accbc"]), C.c_double), _p(ah, C.c_double), _p(at, C.c_double),
                                  _p(H, C.c_double), _p(b, C.c_double))
    return H, b


# ----------------------------------------------------------------------------- coarse tracker
def coarse_calc_res_gs(lvl, dI, fx, fy, cx, cy, Ki, R, t, affLL, b0, cutoffTH, pc_u, pc_v, pc_idepth, pc_color):
    """CoarseTracker::calcRes + calcGSSSE at one level -> dict(rs[6], H[8,8], b[8], counts[3])."""
    dI = _f32(dI)
    hl, wl = dI.shape[0], dI.shape[1]
    Ki, aff = _f32(Ki).reshape(-1), _f32(affLL)
    R, t = _f64(R).reshape(-1), _f64(t)
    u, v, idp, col = _f32(pc_u), _f32(pc_v), _f32(pc_idepth), _f32(pc_color)
    rs, H, b = np.zeros(6), np.zeros((8, 8)), np.zeros(8)
    counts = np.zeros(3, np.int64)
    lib().eds_oracle_coarse_calc_res_gs(C.c_int(lvl), C.c_int(wl), C.c_int(hl), _p(dI, C.c_float), C.c_float(fx), C.c_float(fy), C.c_float(cx),
                                        C.c_float(cy), _p(Ki, C.c_float), _p(R, C.c_double), _p(t, C.c_double), _p(aff, C.c_float),
                                        C.c_float(b0), C.c_float(cutoffTH), C.c_int(len(u)), _p(u, C.c_float), _p(v, C.c_float),
                                        _p(idp, C.c_float), _p(col, C.c_float), _p(rs, C.c_double), _p(H, C.c_double), _p(b, C.c_double),
                                        _p(counts, C.c_int64))
    return dict(rs=rs, H=H, b=b, counts=counts)


# ----------------------------------------------------------------------------- depth filter
def depth_update(fx, fy, cx, cy, mu_range, px_error_angle, T_kf_ef, kf_coord, ef_coord, state, coords_are_tracks=False):
    """DepthPoints::update on a copy of state (N x 4: mu, sigma2, a, b) -> (new state, ok flags)."""
    T = _f64(T_kf_ef).reshape(-1)
    kf, ef = _f64(kf_coord), _f64(ef_coord)
    st = np.array(state, np.float64, copy=True, order="C")
    N = len(st)
    ok = np.zeros(N, np.uint8)
    lib().eds_oracle_depth_update(C.c_int(N), C.c_double(fx), C.c_double(fy), C.c_double(cx), C.c_double(cy), C.c_double(mu_range),
                                  C.c_double(px_error_angle), _p(T, C.c_double), _p(kf, C.c_double), _p(ef, C.c_double),
                                  C.c_int(int(coords_are_tracks)), _p(st, C.c_double), _p(ok, C.c_uint8))
    return st, ok


def tracker_get_coord(kf, inv_depth, px, qx):
    """Tracker::getCoord (Tracker.cpp:319-376), numpy double: key-frame points at inverse depths `inv_depth` warped with
    (px, qx = xyzw) into the event frame -> (coord N x 2, outlier flags)."""
    x, y, z, w = [float(v) for v in qx]
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    nc = np.asarray(kf["norm_coord"], np.float64)
    pz = 1.0 / np.asarray(inv_depth, np.float64)
    p = np.stack([nc[:, 0] * pz, nc[:, 1] * pz, pz], 1) @ R.T + np.asarray(px, np.float64)
    xp = kf["fx"] * (p[:, 0] / p[:, 2]) + kf["cx"]
    yp = kf["fy"] * (p[:, 1] / p[:, 2]) + kf["cy"]
    outlier = (xp < 0.0) | (xp > kf["W"]) | (yp < 0.0) | (yp > kf["H"])
    return np.stack([xp, yp], 1), outlier.astype(np.uint8)


def coarse_track(pb, coarsest_lvl, R, t, aff=(0.0, 0.0), ref_aff=(0.0, 0.0), ref_exposure=1.0, new_exposure=1.0, min_res_for_abort=None):
    """CoarseTracker::trackNewestCoarse on a synth_coarse problem -> dict(ok, R, t, aff, last_residuals, last_flow, evaluations)."""
    L = pb["levels"]
    nl = len(L)
    keep = []

    def arr_of_ptrs(key):
        a = [_f32(l[key]) for l in L]
        keep.append(a)
        return (C.POINTER(C.c_float) * nl)(*[_p(x, C.c_float) for x in a])

    wl = _i32([l["w"] for l in L]); hl = _i32([l["h"] for l in L]); n = _i32([len(l["pc_u"]) for l in L])
    fx, fy = _f32([l["fx"] for l in L]), _f32([l["fy"] for l in L])
    cx, cy = _f32([l["cx"] for l in L]), _f32([l["cy"] for l in L])
    Rm, tv, af, raf = np.array(R, np.float64).reshape(-1).copy(), np.array(t, np.float64).copy(), np.array(aff, np.float64), _f64(ref_aff)
    mra = _f64(min_res_for_abort if min_res_for_abort is not None else [np.inf] * 5)
    lr, lf = np.zeros(5), np.zeros(3)
    ev = C.c_int(0)
    f = lib().eds_oracle_coarse_track
    f.restype = C.c_int
    ok = f(C.c_int(coarsest_lvl), _p(wl, C.c_int32), _p(hl, C.c_int32), arr_of_ptrs("dI_new"), _p(fx, C.c_float), _p(fy, C.c_float),
           _p(cx, C.c_float), _p(cy, C.c_float), arr_of_ptrs("Ki"), _p(n, C.c_int32), arr_of_ptrs("pc_u"), arr_of_ptrs("pc_v"),
           arr_of_ptrs("pc_idepth"), arr_of_ptrs("pc_color"), _p(Rm, C.c_double), _p(tv, C.c_double), _p(af, C.c_double), _p(raf, C.c_double),
           C.c_float(ref_exposure), C.c_float(new_exposure), _p(mra, C.c_double), _p(lr, C.c_double), _p(lf, C.c_double), C.byref(ev))
    return dict(ok=bool(ok), R=Rm.reshape(3, 3), t=tv, aff=af, last_residuals=lr, last_flow=lf, evaluations=ev.value)


def make_coarse_depth_l0(levels, cu, cv, cid, HdiF):
    """CoarseTracker::makeCoarseDepthL0.  levels: list of dict(w, h, dI_ref (h, w, 3) float32).  -> list of dict(n, pc_u, pc_v,
    pc_idepth, pc_color, idepth, weightSums) per level."""
    L = len(levels)
    w = (C.c_int * L)(*[int(l["w"]) for l in levels])
    h = (C.c_int * L)(*[int(l["h"]) for l in levels])
    cu, cv, cid, HdiF = _f32(cu), _f32(cv), _f32(cid), _f32(HdiF)
    dI = [_f32(l["dI_ref"]) for l in levels]
    bufs = {k: [np.zeros(int(l["w"]) * int(l["h"]), np.float32) for l in levels] for k in ("idepth", "weightSums", "pc_u", "pc_v", "pc_idepth", "pc_color")}
    PP = C.POINTER(C.c_float) * L
    ptrs = {k: PP(*[_p(a, C.c_float) for a in v]) for k, v in bufs.items()}
    dIp = PP(*[_p(a, C.c_float) for a in dI])
    pc_n = (C.c_int * L)()
    lib().eds_oracle_make_coarse_depth_l0(C.c_int(L), w, h, C.c_int(len(cu)), _p(cu, C.c_float), _p(cv, C.c_float), _p(cid, C.c_float),
                                          _p(HdiF, C.c_float), dIp, ptrs["idepth"], ptrs["weightSums"], ptrs["pc_u"], ptrs["pc_v"],
                                          ptrs["pc_idepth"], ptrs["pc_color"], pc_n)
    out = []
    for i, l in enumerate(levels):
        n = pc_n[i]
        out.append(dict(n=n, pc_u=bufs["pc_u"][i][:n].copy(), pc_v=bufs["pc_v"][i][:n].copy(), pc_idepth=bufs["pc_idepth"][i][:n].copy(),
                        pc_color=bufs["pc_color"][i][:n].copy(), idepth=bufs["idepth"][i].reshape(int(l["h"]), int(l["w"])),
                        weightSums=bufs["weightSums"][i].reshape(int(l["h"]), int(l["w"]))))
    return out
