"""Deterministic synthetic inputs for the coarse tracker evaluation (DSO CoarseTracker::calcRes / calcGSSSE):
a reference and a new view of a textured plane at every pyramid level, the reference's point cloud with
inverse depths, and a slightly wrong relative pose to evaluate at.  numpy only; shared by tests/ and bench.py."""
import numpy as np


def _so3_exp(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K


def make_coarse_problem(W=640, H=480, levels=4, points=20000, seed=7, pose_error=2e-3):
    rng = np.random.default_rng(seed)
    f32 = np.float32
    Z0 = 2.0
    comps = [(rng.uniform(0.5, 6.0), rng.uniform(0, 2 * np.pi), rng.uniform(0, 2 * np.pi)) for _ in range(32)]

    def texture(X, Y):
        v = np.zeros_like(X)
        for fr, ang, ph in comps:
            v += (1.0 / fr) * np.cos(2 * np.pi * fr * (np.cos(ang) * X + np.sin(ang) * Y) / 3.0 + ph)
        return v

    R_true = _so3_exp(rng.normal(scale=0.01, size=3))      # refToNew
    t_true = rng.normal(scale=0.03, size=3)
    fx0 = fy0 = 520.0 * W / 640.0
    cx0, cy0 = W / 2 - 0.5, H / 2 - 0.5
    out = dict(levels=[], R_true=R_true, t_true=t_true, R=(_so3_exp(rng.normal(scale=pose_error, size=3)) @ R_true), t=t_true + rng.normal(scale=pose_error, size=3),
               affLL=np.array([1.02, -1.5], f32), b0=f32(0.7), cutoffTH=f32(20.0))
    lo, hi = -3.0, 3.0
    for lvl in range(levels):
        wl, hl = W >> lvl, H >> lvl
        # makeK (CoarseTracker.cpp:67-100): fx/2^l, (cx + 0.5)/2^l - 0.5
        fx, fy = fx0 / 2 ** lvl, fy0 / 2 ** lvl
        cx, cy = (cx0 + 0.5) / 2 ** lvl - 0.5, (cy0 + 0.5) / 2 ** lvl - 0.5
        K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
        Ki = np.linalg.inv(K).astype(f32)
        v, u = np.mgrid[0:hl, 0:wl].astype(np.float64)

        def render(R, t):  # camera x_c = R x_ref + t sees the plane z_ref = Z0
            d = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones_like(u)], -1)
            dw = d @ R          # R^T d
            ow = -R.T @ t
            lam = (Z0 - ow[2]) / dw[..., 2]
            return texture(ow[0] + lam * dw[..., 0], ow[1] + lam * dw[..., 1]), lam

        I_ref, depth_ref = render(np.eye(3), np.zeros(3))
        I_new, _ = render(R_true, t_true)

        def to_dI(I):
            I = ((I - lo) / (hi - lo) * 255.0).astype(f32)
            dI = np.zeros((hl, wl, 3), f32)
            dI[..., 0] = I
            dI[1:-1, 1:-1, 1] = 0.5 * (I[1:-1, 2:] - I[1:-1, :-2])
            dI[1:-1, 1:-1, 2] = 0.5 * (I[2:, 1:-1] - I[:-2, 1:-1])
            return dI

        dI_ref, dI_new = to_dI(I_ref), to_dI(I_new)
        n = max(64, points >> (2 * lvl))
        pu = rng.integers(2, wl - 2, n)
        pv = rng.integers(2, hl - 2, n)
        idp = (1.0 / depth_ref[pv, pu]) * (1.0 + rng.normal(scale=0.01, size=n))
        # a few points that must be rejected: negative depth after the warp, far outside, saturated residual
        idp[::97] = -0.2
        color = dI_ref[pv, pu, 0].copy()
        color[::53] += 60.0
        out["levels"].append(dict(w=wl, h=hl, fx=f32(fx), fy=f32(fy), cx=f32(cx), cy=f32(cy), Ki=Ki.T.reshape(-1).copy(),  # column-major
                                  dI_new=dI_new, pc_u=pu.astype(f32), pc_v=pv.astype(f32), pc_idepth=idp.astype(f32), pc_color=color.astype(f32)))
    return out
