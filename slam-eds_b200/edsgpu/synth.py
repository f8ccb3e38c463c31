"""Deterministic ESIM-style synthetic inputs for the tracking hot path.

numpy only; shared by tests/ and bench.py.  Shapes follow BASELINE.json configs:
a textured scene (band-limited log-intensity noise) with smooth depth relief, a
keyframe rendered at identity (Sobel-3 gradients of log(img+0.2) as in the
reference's KeyFrame.cpp:373-385, points = strongest gradients per 20x20 cell,
KeyFrame.cpp:409), and event windows generated from the brightness increment
-grad(L).flow at a random 6-DoF twist.
"""
import numpy as np

CONFIGS = {
    # name: (H, W, fx, fy, cx, cy, N points, E events)
    "davis240c": dict(H=180, W=240, fx=200.0, fy=200.0, cx=120.0, cy=90.0, N=2048, E=20000),
    "gen3_vga": dict(H=480, W=640, fx=520.0, fy=520.0, cx=320.0, cy=240.0, N=10240, E=50000),
    "gen4_hd": dict(H=720, W=1280, fx=1040.0, fy=1040.0, cx=640.0, cy=360.0, N=51200, E=200000),
    "tiny": dict(H=48, W=64, fx=60.0, fy=60.0, cx=32.0, cy=24.0, N=256, E=3000),
}
CONFIG_ID = {"davis240c": 1, "gen3_vga": 2, "gen4_hd": 3, "tiny": 9}
Z0 = 2.0


def seed_for(config, sequence_id=0):
    return 1234 + CONFIG_ID[config] * 1000 + sequence_id


def rotvec_to_quat_xyzw(rv):
    th = np.linalg.norm(rv)
    if th < 1e-12:
        return np.array([0.5 * rv[0], 0.5 * rv[1], 0.5 * rv[2], 1.0])
    ax = rv / th
    return np.concatenate([ax * np.sin(th / 2), [np.cos(th / 2)]])


def quat_to_rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def quat_angle(q1, q2):
    """Rotation angle (rad) between two unit quaternions (xyzw)."""
    d = abs(float(np.dot(q1 / np.linalg.norm(q1), q2 / np.linalg.norm(q2))))
    return 2.0 * np.arccos(min(1.0, d))


def sobel3(img):
    """cv::Sobel(ksize=3, BORDER_REFLECT_101) in x and y."""
    p = np.pad(img, 1, mode="reflect")
    gx = (p[:-2, 2:] - p[:-2, :-2]) + 2 * (p[1:-1, 2:] - p[1:-1, :-2]) + (p[2:, 2:] - p[2:, :-2])
    gy = (p[2:, :-2] - p[:-2, :-2]) + 2 * (p[2:, 1:-1] - p[:-2, 1:-1]) + (p[2:, 2:] - p[:-2, 2:])
    return gx, gy


def make_scene(config, sequence_id=0):
    c = CONFIGS[config]
    H, W = c["H"], c["W"]
    rng = np.random.default_rng(seed_for(config, sequence_id))
    v, u = np.mgrid[0:H, 0:W].astype(np.float64)
    # texture: 64 random Fourier components, amplitude ~ 1/f
    img = np.zeros((H, W))
    for _ in range(64):
        f = rng.uniform(2.0, 40.0)
        ang = rng.uniform(0, 2 * np.pi)
        ph = rng.uniform(0, 2 * np.pi)
        img += (1.0 / f) * np.cos(2 * np.pi * f * (np.cos(ang) * u / W + np.sin(ang) * v / W) + ph)
    img = (img - img.min()) / (img.max() - img.min())
    # depth: plane at Z0 with +-0.5 m relief from 8 low-frequency sinusoids
    rel = np.zeros((H, W))
    for _ in range(8):
        f = rng.uniform(0.3, 1.5)
        ang = rng.uniform(0, 2 * np.pi)
        ph = rng.uniform(0, 2 * np.pi)
        rel += np.cos(2 * np.pi * f * (np.cos(ang) * u / W + np.sin(ang) * v / W) + ph)
    rel = 0.5 * rel / np.abs(rel).max()
    depth = Z0 + rel
    log_img = np.log(img + 0.2)
    gx, gy = sobel3(log_img)
    return dict(config=config, H=H, W=W, fx=c["fx"], fy=c["fy"], cx=c["cx"], cy=c["cy"], img=img, log_img=log_img,
                gx=gx, gy=gy, depth=depth, rng=rng)


def make_keyframe(scene, N=None):
    """Point arrays the tracker gathers (KeyFrame.hpp:59-96): grad, norm_coord, idp, weights."""
    c = CONFIGS[scene["config"]]
    N = N or c["N"]
    H, W = scene["H"], scene["W"]
    mag = np.hypot(scene["gx"], scene["gy"])
    margin = 4
    cell = 20
    cand = []
    ncell = ((H - 2 * margin + cell - 1) // cell) * ((W - 2 * margin + cell - 1) // cell)
    k = int(np.ceil(1.3 * N / ncell))
    for r0 in range(margin, H - margin, cell):
        for c0 in range(margin, W - margin, cell):
            blk = mag[r0:min(r0 + cell, H - margin), c0:min(c0 + cell, W - margin)]
            idx = np.argsort(-blk, axis=None, kind="stable")[:k]
            rr, cc = np.unravel_index(idx, blk.shape)
            cand.append(np.stack([rr + r0, cc + c0], 1))
    cand = np.concatenate(cand, 0)
    if len(cand) < N:
        raise ValueError("not enough candidate points")
    m = mag[cand[:, 0], cand[:, 1]]
    keep = np.sort(np.argsort(-m, kind="stable")[:N])  # keep grid order
    pts = cand[keep]
    rows, cols = pts[:, 0], pts[:, 1]
    rng = np.random.default_rng(seed_for(scene["config"], 777))
    kf = dict(H=H, W=W, fx=scene["fx"], fy=scene["fy"], cx=scene["cx"], cy=scene["cy"],
              coord=np.stack([cols, rows], 1).astype(np.float64),
              grad=np.stack([scene["gx"][rows, cols], scene["gy"][rows, cols]], 1),
              norm_coord=np.stack([(cols - scene["cx"]) / scene["fx"], (rows - scene["cy"]) / scene["fy"]], 1),
              idp=1.0 / scene["depth"][rows, cols],
              weights=rng.uniform(0.7, 1.0, N))
    return kf


def flow_field(X, Y, idp, tw):
    """Feature flow of PhotometricError.hpp:114-122 for twist tw = [v(3), w(3)]."""
    fx = -idp * tw[0] + X * idp * tw[2] + X * Y * tw[3] - (1 + X * X) * tw[4] + Y * tw[5]
    fy = -idp * tw[1] + Y * idp * tw[2] + (1 + Y * Y) * tw[3] - X * Y * tw[4] - X * tw[5]
    return fx, fy


def random_twist(rng):
    """Twist [v, w] with |t| in U[0.01,0.05]*Z0 and angle in U[0.5,3] deg."""
    vd = rng.normal(size=3)
    vd /= np.linalg.norm(vd)
    wd = rng.normal(size=3)
    wd /= np.linalg.norm(wd)
    return np.concatenate([vd * rng.uniform(0.01, 0.05) * Z0, wd * np.deg2rad(rng.uniform(0.5, 3.0))])


def make_window(scene, twist, E=None, rng=None, noise_frac=0.05, duration_us=10000):
    """Events for one window.  Returns dict(x,y,pol,ts,truth state, init state)."""
    c = CONFIGS[scene["config"]]
    E = E or c["E"]
    rng = rng or scene["rng"]
    H, W = scene["H"], scene["W"]
    v, u = np.mgrid[0:H, 0:W].astype(np.float64)
    X = (u - scene["cx"]) / scene["fx"]
    Y = (v - scene["cy"]) / scene["fy"]
    idp = 1.0 / scene["depth"]
    flx, fly = flow_field(X, Y, idp, twist)
    dL = -(scene["gx"] * flx + scene["gy"] * fly)
    # true state: camera moved by the twist for unit time: P_ef = R P_kf + t, R = exp(-w), t = -v
    t_true = -twist[:3]
    q_true = rotvec_to_quat_xyzw(-twist[3:])
    R = quat_to_rot(q_true)
    Z = scene["depth"]
    Pk = np.stack([X * Z, Y * Z, Z], 0).reshape(3, -1)
    Pe = R @ Pk + t_true[:, None]
    ue = np.rint(scene["fx"] * Pe[0] / Pe[2] + scene["cx"]).astype(np.int64)
    ve = np.rint(scene["fy"] * Pe[1] / Pe[2] + scene["cy"]).astype(np.int64)
    ok = (ue >= 0) & (ue < W) & (ve >= 0) & (ve < H)
    a = np.abs(dL).reshape(-1) * ok
    n_sig = int(round(E * (1.0 - noise_frac)))
    # contrast threshold chosen so that sum round(|dL|/C) ~ n_sig (bisection)
    lo, hi = 1e-9, a.max() + 1e-9
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if np.rint(a / mid).sum() > n_sig:
            lo = mid
        else:
            hi = mid
    cnt = np.rint(a / hi).astype(np.int64)
    xs = np.repeat(ue, cnt)
    ys = np.repeat(ve, cnt)
    ps = np.repeat((dL.reshape(-1) > 0).astype(np.uint8), cnt)
    n_noise = max(0, E - len(xs))
    xs = np.concatenate([xs, rng.integers(0, W, n_noise)])
    ys = np.concatenate([ys, rng.integers(0, H, n_noise)])
    ps = np.concatenate([ps, rng.integers(0, 2, n_noise).astype(np.uint8)])
    perm = rng.permutation(len(xs))[:E]
    ts = np.sort(rng.integers(0, duration_us, E)).astype(np.int64)
    tw_unit = twist / np.linalg.norm(twist)
    x_true = np.concatenate([t_true, q_true, tw_unit])
    x_init = np.concatenate([0.8 * t_true, rotvec_to_quat_xyzw(-0.8 * twist[3:]),
                             np.full(6, 1e-3) / np.linalg.norm(np.full(6, 1e-3))])  # Tracker.cpp:45-46
    return dict(x=xs[perm].astype(np.uint16), y=ys[perm].astype(np.uint16), pol=ps[perm].astype(np.uint8), ts=ts,
                x_true=x_true, x_init=x_init, twist=twist)


def radtan_lut(H, W, fx, fy, cx, cy, k1=-0.08, k2=0.01):
    """A smooth non-identity forward undistortion LUT (float32 HxW, like EventFrame.cpp:72-81)."""
    v, u = np.mgrid[0:H, 0:W].astype(np.float64)
    x = (u - cx) / fx
    y = (v - cy) / fy
    r2 = x * x + y * y
    s = 1 + k1 * r2 + k2 * r2 * r2
    return (fx * x * s + cx).astype(np.float32), (fy * y * s + cy).astype(np.float32)


def make_problem(config, sequence_id=0, n_windows=1):
    """Scene + keyframe + n_windows event windows along a smooth random walk of twists."""
    scene = make_scene(config, sequence_id)
    kf = make_keyframe(scene)
    rng = scene["rng"]
    tw = random_twist(rng)
    wins = []
    for _ in range(n_windows):
        wins.append(make_window(scene, tw, rng=rng))
        tw = 0.8 * tw + 0.2 * random_twist(rng)
    return scene, kf, wins
