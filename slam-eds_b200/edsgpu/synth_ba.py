"""Deterministic synthetic inputs for the windowed-BA Hessian accumulators.

F keyframes look at a textured plane from poses on a smooth random walk; every hosted point gets
a residual to every other frame.  The per-residual RawResidualJacobian records (76-float layout
of include/edsgpu.h) are produced by a float32 numpy restatement of the reference's feeder,
PointFrameResidual::linearize (src/tracking/Residuals.cpp:69-265, projectPoint of
ResidualProjections.h:46-86, getInterpolatedElement33 of globalFuncs.h:78-92), so the accumulators
see inputs with the right structure and magnitudes.  numpy only; shared by tests/ and bench.py.
"""
import numpy as np

PATTERN = np.array([[0, -2], [-1, -1], [1, -1], [-2, 0], [0, 0], [2, 0], [-1, 1], [0, 2]], np.float32)  # settings.cpp:276
SETTING_HUBER_TH = np.float32(9.0)            # settings.cpp:127
SETTING_OUTLIER_TH_SUMC = np.float32(50 * 50)  # settings.cpp:91
SCALE_A, SCALE_B = 10.0, 1000.0                # HessianBlocks.h:64-65
REC = 76
O_RES, O_JPDXI0, O_JPDXI1, O_JPDC0, O_JPDC1, O_JPDD = 0, 8, 14, 20, 24, 28
O_JIDX0, O_JIDX1, O_JAB0, O_JAB1, O_JIDX2, O_JABJIDX, O_JAB2 = 32, 40, 48, 56, 64, 68, 72


def _so3_exp(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K


def _hat(t):
    return np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])


def _interp33(dI, x, y):
    """getInterpolatedElement33 (globalFuncs.h:78-92), float32, vectorised. dI: (H,W,3)."""
    W = dI.shape[1]
    ix = x.astype(np.int32)
    iy = y.astype(np.int32)
    dx = (x - ix).astype(np.float32)
    dy = (y - iy).astype(np.float32)
    dxdy = dx * dy
    flat = dI.reshape(-1, 3)
    bp = ix + iy * W
    return (dxdy[:, None] * flat[bp + 1 + W] + (dy - dxdy)[:, None] * flat[bp + W]
            + (dx - dxdy)[:, None] * flat[bp + 1] + (1 - dx - dy + dxdy)[:, None] * flat[bp])


def make_ba_problem(F=7, points_per_frame=2048, H=480, W=640, seed=4234, linearized_frac=0.25, perturb=1e-3):
    """Returns a dict with every array the Top / SC accumulators and their stitches consume."""
    rng = np.random.default_rng(seed)
    f32 = np.float32
    fx = fy = f32(520.0 * W / 640.0)
    cx, cy = f32(W / 2), f32(H / 2)
    Z0 = 2.0
    # ---- texture on the world plane z = Z0 (intensities 0..255 like DSO) -----------------
    comps = [(rng.uniform(1.0, 12.0), rng.uniform(0, 2 * np.pi), rng.uniform(0, 2 * np.pi)) for _ in range(48)]

    def texture(X, Y):
        v = np.zeros_like(X)
        for fr, ang, ph in comps:
            v += (1.0 / fr) * np.cos(2 * np.pi * fr * (np.cos(ang) * X + np.sin(ang) * Y) / 3.0 + ph)
        return v

    # ---- poses: world->cam, smooth random walk around identity ---------------------------
    poses = []
    t, w = np.zeros(3), np.zeros(3)
    for f in range(F):
        poses.append((_so3_exp(w), t.copy()))
        t = t + rng.normal(scale=0.02, size=3)
        w = w + rng.normal(scale=0.008, size=3)
    v, u = np.mgrid[0:H, 0:W].astype(np.float64)
    dIs, depths = [], []
    lo, hi = None, None
    raw = []
    for R, tt in poses:
        # ray-plane intersection: X_w = R^T (lambda d - t), z_w = Z0
        d = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u)], -1)
        dw = d @ R  # R^T d
        ow = -R.T @ tt
        lam = (Z0 - ow[2]) / dw[..., 2]
        Xw = ow[0] + lam * dw[..., 0]
        Yw = ow[1] + lam * dw[..., 1]
        raw.append(texture(Xw, Yw))
        depths.append(lam)  # depth along the optical axis (d_z = 1)
    lo, hi = min(r.min() for r in raw), max(r.max() for r in raw)
    for r in raw:
        I = ((r - lo) / (hi - lo) * 255.0).astype(f32)
        dI = np.zeros((H, W, 3), f32)
        dI[..., 0] = I
        dI[1:-1, 1:-1, 1] = 0.5 * (I[1:-1, 2:] - I[1:-1, :-2])  # makeImages, HessianBlocks.cpp:180-186
        dI[1:-1, 1:-1, 2] = 0.5 * (I[2:, 1:-1] - I[:-2, 1:-1])
        dIs.append(dI)

    # ---- affine brightness states (a, b) per frame, exposures 1 -----------------------------
    aff = [(0.0, 0.0)] + [(rng.normal(scale=0.02), rng.normal(scale=2.0)) for _ in range(F - 1)]

    def from_to_vec_exposure(g2F, g2T):  # NumType.h:178-190 with exposures 1
        a = np.exp(g2T[0] - g2F[0])
        return a, g2T[1] - a * g2F[1]

    # ---- points: strongest-gradient pixels of a random candidate set ------------------------
    P = F * points_per_frame
    pu, pv, pid, phost = np.zeros(P, f32), np.zeros(P, f32), np.zeros(P, f32), np.zeros(P, np.int32)
    colors, pweights = np.zeros((P, 8), f32), np.zeros((P, 8), f32)
    for h in range(F):
        g2 = dIs[h][..., 1] ** 2 + dIs[h][..., 2] ** 2
        cu = rng.integers(8, W - 8, 6 * points_per_frame)
        cv = rng.integers(8, H - 8, 6 * points_per_frame)
        order = np.argsort(-g2[cv, cu], kind="stable")[:points_per_frame]
        sel = np.sort(order)
        s = slice(h * points_per_frame, (h + 1) * points_per_frame)
        pu[s], pv[s], phost[s] = cu[sel], cv[sel], h
        pid[s] = (1.0 / depths[h][cv[sel], cu[sel]]) * (1.0 + rng.normal(scale=perturb * 5, size=points_per_frame))
        for k in range(8):  # ImmaturePoint.cpp:39-59: colour + gradient weight per pattern pixel
            px = dIs[h][(cv[sel] + int(PATTERN[k, 1])), (cu[sel] + int(PATTERN[k, 0]))]
            colors[s, k] = px[:, 0]
            pweights[s, k] = np.sqrt(SETTING_OUTLIER_TH_SUMC / (SETTING_OUTLIER_TH_SUMC + px[:, 1] ** 2 + px[:, 2] ** 2))

    # ---- residual graph: every point x every other frame, point-major ------------------------
    R = P * (F - 1)
    host_idx = np.repeat(phost, F - 1).astype(np.int32)
    target_idx = np.zeros(R, np.int32)
    point_of_res = np.repeat(np.arange(P, dtype=np.int32), F - 1)
    res_begin = (np.arange(P + 1) * (F - 1)).astype(np.int32)
    for h in range(F):
        others = np.array([t for t in range(F) if t != h], np.int32)
        m = host_idx == h
        target_idx[m] = np.tile(others, points_per_frame)

    # ---- linearize (Residuals.cpp:69-265), float32, vectorised over residuals ---------------
    recs = np.zeros((R, REC), f32)
    state = np.zeros(R, np.int32)  # 0 IN, 1 OOB, 2 OUTLIER
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64)
    Ki = np.linalg.inv(K)
    fxi, fyi = f32(1.0 / fx), f32(1.0 / fy)
    frame_energy_th = f32(12 * 12 * 8)  # setting_outlierTH * patternNum (FullSystem default)
    # FrameFramePrecalc of (host,target) at host + F*target, layout of include/edsgpu.h (3x3 blocks column-major like Eigen)
    precalc = np.zeros((F * F, 28), f32)
    for h in range(F):
        Rh, th = poses[h]
        for t in range(F):
            if t == h:
                continue
            m = np.where((host_idx == h) & (target_idx == t))[0]
            Rt, tt = poses[t]
            Rll = (Rt @ Rh.T)
            tll = (tt - Rll @ th)
            PRE_R = Rll.astype(f32)
            PRE_t = tll.astype(f32)
            KRKi = (K @ Rll @ Ki).astype(f32)
            Kt = (K @ tll).astype(f32)
            a_ll, b_ll = from_to_vec_exposure(aff[h], aff[t])
            a_ll, b_ll, b0 = f32(a_ll), f32(b_ll), f32(aff[h][1])
            pc = precalc[h + F * t]
            pc[0:9] = PRE_R.T.reshape(-1); pc[9:12] = PRE_t; pc[12:21] = KRKi.T.reshape(-1); pc[21:24] = Kt
            pc[24], pc[25], pc[26] = a_ll, b_ll, b0
            p = point_of_res[m]
            u0, v0, idp = pu[p], pv[p], pid[p]
            # centre projection (ResidualProjections.h:60-86)
            kx = (u0 - cx) * fxi
            ky = (v0 - cy) * fyi
            ptp = np.stack([PRE_R[r, 0] * kx + PRE_R[r, 1] * ky + PRE_R[r, 2] + PRE_t[r] * idp for r in range(3)], 0).astype(f32)
            drescale = (f32(1.0) / ptp[2]).astype(f32)
            new_id = idp * drescale
            uu = ptp[0] * drescale
            vv = ptp[1] * drescale
            Ku = uu * fx + cx
            Kv = vv * fy + cy
            oob = ~((drescale > 0) & (Ku > 1.1) & (Kv > 1.1) & (Ku < W - 3) & (Kv < H - 3))
            rec = recs[m]
            rec[:, O_JPDD] = drescale * (PRE_t[0] - PRE_t[2] * uu) * fx
            rec[:, O_JPDD + 1] = drescale * (PRE_t[1] - PRE_t[2] * vv) * fy
            cx2 = drescale * (PRE_R[2, 0] * uu - PRE_R[0, 0])
            cx3 = fx * drescale * (PRE_R[2, 1] * uu - PRE_R[0, 1]) * fyi
            cy2 = fy * drescale * (PRE_R[2, 0] * vv - PRE_R[1, 0]) * fxi
            cy3 = drescale * (PRE_R[2, 1] * vv - PRE_R[1, 1])
            rec[:, O_JPDC0 + 0] = kx * cx2 + uu
            rec[:, O_JPDC0 + 1] = ky * cx3
            rec[:, O_JPDC0 + 2] = cx2 + 1
            rec[:, O_JPDC0 + 3] = cx3
            rec[:, O_JPDC1 + 0] = kx * cy2
            rec[:, O_JPDC1 + 1] = ky * cy3 + vv
            rec[:, O_JPDC1 + 2] = cy2
            rec[:, O_JPDC1 + 3] = cy3 + 1
            rec[:, O_JPDXI0 + 0] = new_id * fx
            rec[:, O_JPDXI0 + 1] = 0
            rec[:, O_JPDXI0 + 2] = -new_id * uu * fx
            rec[:, O_JPDXI0 + 3] = -uu * vv * fx
            rec[:, O_JPDXI0 + 4] = (1 + uu * uu) * fx
            rec[:, O_JPDXI0 + 5] = -vv * fx
            rec[:, O_JPDXI1 + 0] = 0
            rec[:, O_JPDXI1 + 1] = new_id * fy
            rec[:, O_JPDXI1 + 2] = -new_id * vv * fy
            rec[:, O_JPDXI1 + 3] = -(1 + vv * vv) * fy
            rec[:, O_JPDXI1 + 4] = uu * vv * fy
            rec[:, O_JPDXI1 + 5] = uu * fy
            J2 = np.zeros((len(m), 3), f32)   # JIdx2 00, 10, 11
            JabJ = np.zeros((len(m), 4), f32)  # 00 01 10 11
            Jab2 = np.zeros((len(m), 3), f32)  # 00 01 11
            energy = np.zeros(len(m), f32)
            wJI2 = np.zeros(len(m), f32)
            for k in range(8):
                upt = u0 + PATTERN[k, 0]
                vpt = v0 + PATTERN[k, 1]
                q = np.stack([KRKi[r, 0] * upt + KRKi[r, 1] * vpt + KRKi[r, 2] + Kt[r] * idp for r in range(3)], 0).astype(f32)
                Kuk = q[0] / q[2]
                Kvk = q[1] / q[2]
                inb = (Kuk > 1.1) & (Kvk > 1.1) & (Kuk < W - 3) & (Kvk < H - 3)
                oob |= ~inb
                Kus = np.where(inb, Kuk, f32(4.0))
                Kvs = np.where(inb, Kvk, f32(4.0))
                hit = _interp33(dIs[t], Kus, Kvs)
                residual = hit[:, 0] - (a_ll * colors[p, k] + b_ll)
                drdA = colors[p, k] - b0
                wgt = np.sqrt(SETTING_OUTLIER_TH_SUMC / (SETTING_OUTLIER_TH_SUMC + (hit[:, 1] ** 2 + hit[:, 2] ** 2))).astype(f32)  # tail<2>().squaredNorm() first
                wgt = f32(0.5) * (wgt + pweights[p, k])
                ar = np.abs(residual)
                hw = np.where(ar < SETTING_HUBER_TH, f32(1.0), SETTING_HUBER_TH / np.maximum(ar, f32(1e-20))).astype(f32)
                energy += wgt * wgt * hw * residual * residual * (2 - hw)
                hw = np.where(hw < 1, np.sqrt(hw), hw).astype(f32) * wgt
                gx = hit[:, 1] * hw
                gy = hit[:, 2] * hw
                rec[:, O_RES + k] = residual * hw
                rec[:, O_JIDX0 + k] = gx
                rec[:, O_JIDX1 + k] = gy
                rec[:, O_JAB0 + k] = drdA * hw
                rec[:, O_JAB1 + k] = hw
                J2[:, 0] += gx * gx; J2[:, 2] += gy * gy; J2[:, 1] += gx * gy
                JabJ[:, 0] += drdA * hw * gx; JabJ[:, 1] += drdA * hw * gy; JabJ[:, 2] += hw * gx; JabJ[:, 3] += hw * gy
                Jab2[:, 0] += drdA * drdA * hw * hw; Jab2[:, 1] += drdA * hw * hw; Jab2[:, 2] += hw * hw
                wJI2 += hw * hw * (gx * gx + gy * gy)
            # Mat22f column-major: (0,0) (1,0) (0,1) (1,1)
            rec[:, O_JIDX2 + 0] = J2[:, 0]; rec[:, O_JIDX2 + 1] = J2[:, 1]; rec[:, O_JIDX2 + 2] = J2[:, 1]; rec[:, O_JIDX2 + 3] = J2[:, 2]
            rec[:, O_JABJIDX + 0] = JabJ[:, 0]; rec[:, O_JABJIDX + 1] = JabJ[:, 2]; rec[:, O_JABJIDX + 2] = JabJ[:, 1]; rec[:, O_JABJIDX + 3] = JabJ[:, 3]
            rec[:, O_JAB2 + 0] = Jab2[:, 0]; rec[:, O_JAB2 + 1] = Jab2[:, 1]; rec[:, O_JAB2 + 2] = Jab2[:, 1]; rec[:, O_JAB2 + 3] = Jab2[:, 2]
            recs[m] = rec
            st = np.where(oob, 1, np.where((energy > frame_energy_th) | (wJI2 < 2), 2, 0))
            state[m] = st
    recs[state == 1] = 0  # OOB residuals never get a Jacobian (Residuals.cpp:73-74,103)

    # ---- EnergyFunctional-level inputs ----------------------------------------------------
    active = state == 0
    linearized = active & (rng.random(R) < linearized_frac)
    flags = (active.astype(np.uint8) | (linearized.astype(np.uint8) << 1)).astype(np.uint8)
    # adjoints (EnergyFunctional.cpp:46-106)
    adHost = np.zeros((F * F, 8, 8))
    adTarget = np.zeros((F * F, 8, 8))
    for h in range(F):
        for t in range(F):
            Rh, th = poses[h]
            Rt, tt = poses[t]
            Rll = Rt @ Rh.T
            tll = tt - Rll @ th
            Adj = np.zeros((6, 6))
            Adj[:3, :3] = Rll; Adj[:3, 3:] = _hat(tll) @ Rll; Adj[3:, 3:] = Rll
            AH, AT = np.eye(8), np.eye(8)
            AH[:6, :6] = -Adj.T
            a_ll, _ = from_to_vec_exposure(aff[h], aff[t])
            a_ll = float(np.float32(a_ll))
            AT[6, 6] = -a_ll; AH[6, 6] = a_ll; AT[7, 7] = -1; AH[7, 7] = a_ll
            AH[6, :] *= SCALE_A; AH[7, :] *= SCALE_B; AT[6, :] *= SCALE_A; AT[7, :] *= SCALE_B
            adHost[h + t * F] = AH
            adTarget[h + t * F] = AT
    frame_delta = rng.normal(scale=perturb, size=(F, 8))
    frame_delta[0] = 0
    adHTdeltaF = np.zeros((F * F, 8), f32)
    for h in range(F):
        for t in range(F):
            k = h + t * F
            adHTdeltaF[k] = (frame_delta[h].astype(f32) @ adHost[k].astype(f32) + frame_delta[t].astype(f32) @ adTarget[k].astype(f32))
    cDeltaF = rng.normal(scale=perturb, size=4).astype(f32)
    deltaF = rng.normal(scale=perturb, size=P).astype(f32)
    priorF = np.where(rng.random(P) < 0.1, f32(50 * 50), f32(0)).astype(f32)  # setting_idepthFixPrior on a few points
    return dict(F=F, P=P, R=R, H=H, W=W, recs=recs, host_idx=host_idx, target_idx=target_idx, point_of_res=point_of_res,
                res_begin=res_begin, flags=flags, state=state, deltaF=deltaF, priorF=priorF, adHTdeltaF=adHTdeltaF,
                cDeltaF=cDeltaF, adHost=adHost, adTarget=adTarget, cPrior=np.full(4, 5e9),
                frame_prior=np.abs(rng.normal(scale=1e3, size=(F, 8))), frame_delta_prior=rng.normal(scale=perturb, size=(F, 8)),
                # inputs of the feeder itself (PointFrameResidual::linearize), for the device-side linearisation
                dI=np.stack(dIs), precalc=precalc, calib=np.array([fx, fy, cx, cy], f32), pu=pu, pv=pv, idepth=pid,
                color=colors, weights=pweights, frame_energy_th=np.full(F, frame_energy_th, f32))


def col_major(mats):
    """(n,8,8) row-major numpy -> flat column-major blocks as Eigen stores Mat88 arrays."""
    return np.ascontiguousarray(np.transpose(mats, (0, 2, 1))).reshape(len(mats), 64)
