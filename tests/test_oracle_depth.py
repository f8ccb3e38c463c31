"""CPU tests of the depth-filter oracle (DepthPoints::update).  The reference has no
golden vectors for this path (parity unpinned): the checks are what the algorithm must do on exact geometry."""
import numpy as np

from oracle import oracle as O

FX, FY, CX, CY = 520.0, 515.0, 320.0, 240.0


def scene(n=400, seed=0, noise=0.0):
    """Points with known depth in the key frame, observed from a second pose T_kf_ef (point_kf = T_kf_ef point_ef)."""
    rng = np.random.default_rng(seed)
    kf = np.stack([rng.uniform(20, 620, n), rng.uniform(20, 460, n)], 1)
    depth = rng.uniform(1.0, 6.0, n)
    P = np.stack([(kf[:, 0] - CX) / FX * depth, (kf[:, 1] - CY) / FY * depth, depth], 1)
    w = np.array([0.01, -0.02, 0.015])
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    R = np.eye(3) + K + 0.5 * K @ K
    u, _, vt = np.linalg.svd(R)
    R = u @ vt
    t = np.array([0.25, -0.08, 0.05])
    T = np.eye(4); T[:3, :3] = R; T[:3, 3] = t           # T_kf_ef
    Pe = (P - t) @ R                                      # R^T (P - t)
    ef = np.stack([FX * Pe[:, 0] / Pe[:, 2] + CX, FY * Pe[:, 1] / Pe[:, 2] + CY], 1) + rng.normal(scale=noise, size=(n, 2))
    return T, kf, ef, depth


def px_angle():
    return np.arctan(3.0 / (2 * FX)) + np.arctan(3.0 / (2 * FY))


def test_triangulation_recovers_the_true_inverse_depth():
    T, kf, ef, depth = scene()
    n = len(depth)
    st0 = np.tile([0.4, 1e6, 1e9, 1e-9], (n, 1))        # flat prior, certain inlier: the posterior mean is the measurement
    st, ok = O.depth_update(FX, FY, CX, CY, 5.0, px_angle(), T, kf, ef, st0)
    assert ok.all()
    assert np.allclose(st[:, 0], 1.0 / depth, rtol=1e-6)
    assert np.all(st[:, 1] > 0) and np.all(st[:, 1] < 1e6)
    # tracks (offsets) are the same measurement
    st2, _ = O.depth_update(FX, FY, CX, CY, 5.0, px_angle(), T, kf, ef - kf, st0, coords_are_tracks=True)
    assert np.allclose(st2, st, rtol=1e-12)


def test_repeated_updates_converge_and_outliers_lower_the_inlier_ratio():
    T, kf, ef, depth = scene(noise=0.3, seed=4)
    n = len(depth)
    st = np.tile([1.0 / 3.0, 25.0 / 36.0, 10.0, 10.0], (n, 1))
    for k in range(8):
        st, ok = O.depth_update(FX, FY, CX, CY, 5.0, px_angle(), T, kf, ef, st)
    assert np.median(np.abs(st[:, 0] - 1.0 / depth) * depth) < 0.02       # relative inverse-depth error
    assert np.all(st[:, 1] < 25.0 / 36.0)                                 # uncertainty shrank
    inlier = st[:, 2] / (st[:, 2] + st[:, 3])
    assert np.median(inlier) > 0.5
    # gross outliers: measurements far from the converged mean pull a/(a+b) down and barely move mu
    bad = ef + np.array([40.0, -25.0])
    st_bad, _ = O.depth_update(FX, FY, CX, CY, 5.0, px_angle(), T, kf, bad, st)
    assert np.median(st_bad[:, 2] / (st_bad[:, 2] + st_bad[:, 3])) < np.median(inlier)
    assert np.median(np.abs(st_bad[:, 0] - st[:, 0]) / st[:, 0]) < 0.05


def test_negative_mean_is_reset_like_the_reference():
    T, kf, ef, depth = scene(n=8)
    st0 = np.tile([-0.5, 1e-12, 10.0, 10.0], (8, 1))    # a (broken) negative prior the measurement cannot move
    st, ok = O.depth_update(FX, FY, CX, CY, 5.0, px_angle(), T, kf, ef, st0)
    assert not ok.any() and np.all(st[:, 0] == 1.0)
