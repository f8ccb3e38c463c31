"""CPU tests that pin the oracle (the reference has no tests of its own):
analytic known-answer tests, real OpenCV (cv2) for blur/norm, the in-repo second statements of
the same maths, finite differences, and the committed golden fixtures."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from edsgpu import synth
import drift

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def tiny():
    scene, kf, wins = synth.make_problem("tiny", 0, 1)
    w = wins[0]
    ef = O.event_frame(w["x"], w["y"], w["pol"], w["ts"], kf["H"], kf["W"])
    return kf, w, ef


# ------------------------------------------------------------------ event frame
def test_exp_weight_matches_reference_formula():
    # Utils.hpp:542-546 with idx=i/E, window=1: exp(-0.5*((i/E-0.5)*6)^2)
    E = 7
    img = O.draw_values(np.zeros(E), np.zeros(E), np.ones(E, np.int8), 1, 1, "nn", True, 0.0)
    expect = sum(np.exp(-0.5 * ((i / E - 0.5) / (1 / 6.0)) ** 2) for i in range(E))
    assert abs(img[0, 0] - expect) < 1e-14


def test_single_event_integer_location_deposits_one_corner():
    img = O.draw_values([5.0], [3.0], [1], 8, 10, "bilinear", False, 0.0)
    assert img[3, 5] == 1.0 and img.sum() == 1.0


def test_bilinear_weights_and_oob_corners():
    img = O.draw_values([2.25], [1.5], [-1], 4, 4, "bilinear", False, 0.0)
    np.testing.assert_allclose(img[1, 2], -0.75 * 0.5)
    np.testing.assert_allclose(img[2, 2], -0.75 * 0.5)
    np.testing.assert_allclose(img[1, 3], -0.25 * 0.5)
    np.testing.assert_allclose(img.sum(), -1.0)
    # right/bottom corners outside: weight zero (Utils.cpp:92-95), nothing wraps onto the clipped index
    img = O.draw_values([3.5], [3.5], [1], 4, 4, "bilinear", False, 0.0)
    np.testing.assert_allclose(img[3, 3], 0.25)
    np.testing.assert_allclose(img.sum(), 0.25)
    img = O.draw_values([-0.5], [0.0], [1], 4, 4, "bilinear", False, 0.0)
    np.testing.assert_allclose(img[0, 0], 0.5)
    np.testing.assert_allclose(img.sum(), 0.5)


def test_nn_rounds_half_to_even_and_clips():
    # cv::Point2i(Point2d) == cvRound: 0.5 -> 0, 1.5 -> 2, 2.5 -> 2 ; then clip (Utils.cpp:75-78)
    img = O.draw_values([0.5, 1.5, 2.5, 99.0, -7.0], [0, 0, 0, 0, 0], [1, 1, 1, 1, 1], 1, 4, "nn", False, 0.0)
    np.testing.assert_array_equal(img[0], [2, 0, 2, 1])


def test_blur_and_norm_match_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    H, W, E = 37, 53, 4000
    px, py = rng.uniform(-1, W, E), rng.uniform(-1, H, E)
    val = rng.choice([-1, 1], E).astype(np.int8)
    raw = O.draw_values(px, py, val, H, W, "bilinear", True, 0.0)
    blur = O.draw_values(px, py, val, H, W, "bilinear", True, 0.5)
    ref = cv2.GaussianBlur(raw, (3, 3), 0.5, sigmaY=0.5)  # Utils.cpp:118, default BORDER_REFLECT_101
    np.testing.assert_allclose(blur, ref, rtol=0, atol=1e-14 * np.abs(ref).max())
    const = O.draw_values([], [], [], 5, 6, "nn", False, 0.5) + 3.0
    np.testing.assert_allclose(cv2.GaussianBlur(const, (3, 3), 0.5), 3.0, atol=1e-15)
    k = cv2.getGaussianKernel(3, 0.5).ravel()
    np.testing.assert_allclose(k, [0.10650698, 0.78698604, 0.10650698], atol=1e-8)


def test_event_frame_norm_time_and_error(tiny):
    cv2 = pytest.importorskip("cv2")
    kf, w, ef = tiny
    assert ef["status"] == 0
    assert abs(ef["norm"] - cv2.norm(ef["img"])) < 1e-12 * ef["norm"]  # EventFrame.cpp:360-364
    np.testing.assert_allclose(np.linalg.norm(ef["frame"]), 1.0, rtol=1e-14)
    assert ef["time"] == w["ts"][len(w["ts"]) // 2] and ef["delta"] == w["ts"][-1] - w["ts"][0]
    bad = w["ts"][::-1].copy()
    assert O.event_frame(w["x"], w["y"], w["pol"], bad, kf["H"], kf["W"])["status"] == 4  # EventFrame.cpp:325-329


def test_identity_lut_equals_no_lut(tiny):
    kf, w, ef = tiny
    v, u = np.mgrid[0:kf["H"], 0:kf["W"]].astype(np.float32)
    ef2 = O.event_frame(w["x"], w["y"], w["pol"], w["ts"], kf["H"], kf["W"], u, v)
    np.testing.assert_array_equal(ef2["img"], ef["img"])


# ------------------------------------------------------------------ bicubic (Ceres semantics)
def test_bicubic_reproduces_quadratics_value_and_partials():
    H, W = 12, 14
    r, c = np.mgrid[0:H, 0:W].astype(np.float64)
    grid = 0.3 + 0.2 * r - 0.1 * c + 0.05 * r * c + 0.02 * r * r - 0.03 * c * c
    rng = np.random.default_rng(1)
    rows, cols = rng.uniform(2, H - 3, 50), rng.uniform(2, W - 3, 50)
    f, dr, dc = O.bicubic(grid, rows, cols)
    np.testing.assert_allclose(f, 0.3 + 0.2 * rows - 0.1 * cols + 0.05 * rows * cols + 0.02 * rows ** 2 - 0.03 * cols ** 2, atol=1e-13)
    np.testing.assert_allclose(dr, 0.2 + 0.05 * cols + 0.04 * rows, atol=1e-13)
    np.testing.assert_allclose(dc, -0.1 + 0.05 * rows - 0.06 * cols, atol=1e-13)


def test_bicubic_interpolates_nodes_and_clamps():
    rng = np.random.default_rng(2)
    grid = rng.normal(size=(6, 7))
    rr, cc = np.mgrid[0:6, 0:7]
    f, _, _ = O.bicubic(grid, rr.ravel().astype(float), cc.ravel().astype(float))
    np.testing.assert_allclose(f, grid.ravel(), atol=1e-15)
    # far outside: clamped Grid2D => edge value, zero derivative
    f, dr, dc = O.bicubic(grid, [-50.0, 60.0, 2.0], [3.0, 3.0, 1e9])
    np.testing.assert_allclose(f[:2], [grid[0, 3], grid[5, 3]])
    np.testing.assert_allclose(dr[:2], 0)
    np.testing.assert_allclose(dc[2], 0)


# ------------------------------------------------------------------ residual / Jacobian
def test_flow_formula_against_second_in_repo_statement():
    # src/utils/Utils.hpp:165-173 states the same feature flow as PhotometricError.hpp:114-122
    rng = np.random.default_rng(3)
    X, Y, d = rng.normal(size=20), rng.normal(size=20), rng.uniform(0.2, 2, 20)
    tw = rng.normal(size=6)
    fx, fy = synth.flow_field(X, Y, d, tw)
    B = lambda x, y, z: np.array([[-z, 0, x * z, x * y, -(1 + x * x), y], [0, -z, y * z, 1 + y * y, -x * y, -x]])
    for i in range(20):
        np.testing.assert_allclose(B(X[i], Y[i], d[i]) @ tw, [fx[i], fy[i]], rtol=1e-13)


def test_analytic_jacobian_equals_dual_number_jacobian(tiny):
    kf, w, ef = tiny
    a = O.tracker_evaluate(kf, ef["frame"], w["x_init"], 4, jacobian_mode=0)
    d = O.tracker_evaluate(kf, ef["frame"], w["x_init"], 4, jacobian_mode=1)
    np.testing.assert_allclose(a["residuals"], d["residuals"], atol=1e-14)
    np.testing.assert_allclose(a["jacobian"], d["jacobian"], atol=1e-11 * np.abs(d["jacobian"]).max())
    np.testing.assert_allclose(a["H"], d["H"], rtol=1e-10, atol=1e-12)
    assert abs(a["cost"] - d["cost"]) < 1e-14


def _plus(x, delta):
    """Ceres Plus: p+d, q <- [sin|d| d/|d|, cos|d|]*q (xyzw), v <- normalize(v+d)."""
    out = x.copy()
    out[:3] += delta[:3]
    d = delta[3:6]
    n = np.linalg.norm(d)
    if n > 0:
        dq = np.concatenate([np.sin(n) * d / n, [np.cos(n)]])
        ax, ay, az, aw = dq
        bx, by, bz, bw = x[3:7]
        out[3:7] = [aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                    aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz]
    v = x[7:] + delta[6:]
    out[7:] = v / np.linalg.norm(v)
    return out


def test_jacobian_against_finite_differences(tiny):
    kf, w, ef = tiny
    x = w["x_init"]
    base = O.tracker_evaluate(kf, ef["frame"], x, 4, loss_type=0)
    J = base["jacobian"]
    h = 1e-6
    for k in range(12):
        dp = np.zeros(12); dp[k] = h
        dm = np.zeros(12); dm[k] = -h
        rp = O.tracker_evaluate(kf, ef["frame"], _plus(x, dp), 4, loss_type=0, want_jac=False)["residuals"]
        rm = O.tracker_evaluate(kf, ef["frame"], _plus(x, dm), 4, loss_type=0, want_jac=False)["residuals"]
        fd = (rp - rm) / (2 * h)
        scale = np.abs(J[:, k]).max()
        assert np.abs(fd - J[:, k]).max() < 2e-5 * scale + 1e-9, k
    # the unit-norm retraction makes v itself a null direction of the v-block
    np.testing.assert_allclose(J[:, 6:] @ x[7:], 0, atol=1e-12 * np.abs(J[:, 6:]).max())
    ev = np.linalg.eigvalsh(base["H"])
    assert ev[0] < 1e-10 * ev[-1] and ev[1] > 1e-8 * ev[-1]  # rank 11


def test_block_loss_semantics(tiny):
    kf, w, ef = tiny
    x = w["x_init"]
    e0 = O.tracker_evaluate(kf, ef["frame"], x, 4, loss_type=0)
    a = 0.05
    e1 = O.tracker_evaluate(kf, ef["frame"], x, 4, loss_type=1, loss_param=a)
    s = e0["block_sqnorm"]
    assert np.all(s > a * a)  # block-level Huber sits in its linear regime
    assert abs(e1["cost"] - 0.5 * np.sum(2 * a * np.sqrt(s) - a * a)) < 1e-14
    assert abs(e0["cost"] - 0.5 * s.sum()) < 1e-14
    e2 = O.tracker_evaluate(kf, ef["frame"], x, 4, loss_type=2, loss_param=a)
    assert abs(e2["cost"] - 0.5 * np.sum(a * a * np.log1p(s / (a * a)))) < 1e-14
    # residuals returned are never robustified (Tracker.cpp:223-230)
    np.testing.assert_array_equal(e0["residuals"], e1["residuals"])
    # block partition: last block takes the remainder (Tracker.cpp:178-190)
    e3 = O.tracker_evaluate(kf, ef["frame"], x, 3, loss_type=0)
    n = len(kf["idp"]) // 3
    r = e3["residuals"]
    np.testing.assert_allclose(e3["block_sqnorm"], [np.sum(r[:n] ** 2), np.sum(r[n:2 * n] ** 2), np.sum(r[2 * n:] ** 2)], rtol=1e-13)


def test_mad_tau_matches_reference_definition():
    rng = np.random.default_rng(5)
    r = rng.normal(size=1001)
    srt = np.sort(r)
    med = srt[len(r) // 2]
    mad = np.sort(np.abs(r - med))[len(r) // 2]
    assert abs(O.mad_tau(r) - 1.345 * 1.4826 * mad) < 1e-15
    r = rng.normal(size=1000)  # even size: upper median (nth_element at size/2)
    med = np.sort(r)[500]
    mad = np.sort(np.abs(r - med))[500]
    assert abs(O.mad_tau(r) - 1.345 * 1.4826 * mad) < 1e-15


def test_lm_decreases_cost_and_is_deterministic(tiny):
    kf, w, ef = tiny
    s1 = O.tracker_solve(kf, ef["frame"], w["x_init"], num_blocks=4, max_iterations=20, want_trace=True)
    s2 = O.tracker_solve(kf, ef["frame"], w["x_init"], num_blocks=4, max_iterations=20, threads=4)
    assert s1["status"] == 0 and s1["info"]["usable"] == 1
    assert s1["info"]["final_cost"] < s1["info"]["initial_cost"]
    np.testing.assert_array_equal(s1["x"], s2["x"])  # threading does not change the arithmetic
    assert abs(np.linalg.norm(s1["x"][3:7]) - 1) < 1e-12 and abs(np.linalg.norm(s1["x"][7:]) - 1) < 1e-12
    acc = s1["trace"][:, 3] == 1
    costs = s1["trace"][acc, 0]
    assert np.all(np.diff(costs[costs > 0]) <= 0)  # accepted costs are monotone
    d = O.tracker_solve(kf, ef["frame"], w["x_init"], num_blocks=4, max_iterations=20, jacobian_mode=1)
    assert synth.quat_angle(d["x"][3:7], s1["x"][3:7]) < 1e-9 and np.linalg.norm(d["x"][:3] - s1["x"][:3]) < 1e-9
    z = O.tracker_solve(kf, ef["frame"], w["x_init"], num_blocks=4, max_iterations=0)
    np.testing.assert_array_equal(z["x"], w["x_init"])
    assert z["info"]["termination"] == 1 and z["info"]["iterations"] == 0


# ------------------------------------------------------------------ golden fixtures
@pytest.mark.parametrize("name", ["tracking_tiny.npz", "tracking_davis240c.npz"])
def test_oracle_reproduces_golden(name):
    g = np.load(os.path.join(GOLD, name))
    H, W = (int(v) for v in g["kf_size"])
    nn = O.event_frame(g["ev_x"], g["ev_y"], g["ev_pol"], g["ev_ts"], H, W, method="nn", use_exp=False, sigma=0.0)
    np.testing.assert_array_equal(np.rint(nn["img"]).astype(np.int32), g["nn_counts"])
    bl = O.event_frame(g["ev_x"], g["ev_y"], g["ev_pol"], g["ev_ts"], H, W)
    np.testing.assert_allclose(bl["img"], g["bl_img"], rtol=0, atol=1e-13)
    assert abs(bl["norm"] - float(g["bl_norm"])) < 1e-12 * bl["norm"]
    fx, fy, cx, cy = g["kf_intr"]
    kf = dict(grad=g["kf_grad"], norm_coord=g["kf_norm_coord"], idp=g["kf_idp"], weights=g["kf_weights"], H=H, W=W,
              fx=fx, fy=fy, cx=cx, cy=cy)
    B = int(g["num_blocks"])
    ev = O.tracker_evaluate(kf, bl["frame"], g["x0"], B)
    np.testing.assert_allclose(ev["residuals"], g["ev_residuals"], atol=1e-12)
    np.testing.assert_allclose(ev["jacobian"], g["ev_jacobian"], atol=1e-10 * np.abs(g["ev_jacobian"]).max())
    so = O.tracker_solve(kf, bl["frame"], g["x0"], num_blocks=B, max_iterations=int(g["max_iterations"]))
    assert so["info"]["iterations"] == int(g["so_info"][0])
    assert synth.quat_angle(so["x"][3:7], g["so_x"][3:7]) < 1e-7
    np.testing.assert_allclose(so["x"][:3], g["so_x"][:3], atol=1e-7)
    assert abs(so["next_loss_param"] - float(g["so_tau"])) < 1e-7


# ------------------------------------------------------------------ where the pose gate is meaningful
@pytest.mark.parametrize("config", ["gen3_vga", "gen4_hd"])
def test_oracle_self_drift_bounds_the_pose_gate(config):
    """BASELINE.md section 3 gates the converged pose at 1e-4 rad / 1e-4 x depth.  This test pins, on the oracle ALONE, where
    that gate can be met by any implementation: perturbing the inputs by one ulp of double (or rounding them to the fp32 the
    device stores), or switching between analytic and dual-number Jacobians, moves the oracle's own pose by
      * <= 1e-6 rad / 1e-6 m at max_num_iterations = 20 (two orders below the gate), but
      * more than the gate at max_num_iterations = 30 (measured 1.8e-4 rad / 3..4e-4 m on configs 2 and 3),
    because the trust-region radius has passed 1e15 by then (every step accepted with gain ratio > 1, radius x3 per
    iteration, Tracker.cpp:138-143 leaves Ceres' defaults) and the damping of the null velocity direction is below double
    rounding.  tests/test_gpu_fullsize.py therefore gates 1e-4 strictly at 20 iterations and bounds the difference at 30 by
    drift.DRIFT_FACTOR x this self-drift."""
    scene, kf, wins = synth.make_problem(config, 3, 2)
    w = wins[0]
    frame = O.event_frame(w["x"], w["y"], w["pol"], w["ts"], kf["H"], kf["W"])["frame"]
    a20, t20, _, ref20 = drift.oracle_self_drift(kf, frame, w["x_init"], 20)
    assert a20 <= 1e-6 and t20 <= 1e-6, (a20, t20)
    assert ref20["info"]["final_radius"] < 1e15
    a30, t30, detail, ref30 = drift.oracle_self_drift(kf, frame, w["x_init"], 30)
    assert ref30["info"]["final_radius"] >= 1e15 and ref30["info"]["iterations"] == 30
    assert a30 > drift.ANGLE_GATE / 4 or t30 > drift.DEPTH_GATE / 4, detail   # the strict gate has no headroom left here ...
    assert a30 < 5e-3 and t30 < 5e-3 * synth.Z0, detail                         # ... yet the solution stays put at the 1e-3 level
    # the cost moves too, but only at the 1e-4 level (measured 1.1e-4): the valley is flat along the drift direction
    c = [O.tracker_solve(k2, f2, w["x_init"], num_blocks=8, max_iterations=30, threads=8)["info"]["final_cost"]
         for _, k2, f2 in drift.perturbed_inputs(kf, frame)[:3]]
    assert max(abs(v - ref30["info"]["final_cost"]) for v in c) < 5e-4 * ref30["info"]["final_cost"]


def test_pyramid_levels_match_opencv_morphology():
    """EventFrame.cpp:349-357: level i = dilate + erode with a (2i+1)^2 rectangle.  The numpy restatement against real
    OpenCV, bit for bit (selection of existing values plus one addition), on a frame with flat regions and on noise."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for img in (rng.normal(size=(37, 53)), np.round(rng.normal(size=(24, 31)) * 2) / 2):
        frames, norms = O.event_frame_levels(img, 4)
        assert np.array_equal(frames[0], img)
        for i in range(1, 4):
            el = cv2.getStructuringElement(cv2.MORPH_RECT, (2 * i + 1, 2 * i + 1), (i, i))
            ref = cv2.dilate(img, el) + cv2.erode(img, el)
            assert np.array_equal(frames[i], ref)
            assert abs(norms[i] - cv2.norm(ref)) <= 1e-12 * norms[i]


def test_oracle_reproduces_round2_fixture():
    """tests/golden/round2_small.npz: the oracle's outputs for the paths added in round 2 (event-frame pyramid, per-level solve,
    makeCoarseDepthL0, solveSystemF) are frozen; regenerating them here must give the committed numbers."""
    from edsgpu import synth_ba as SB, synth_coarse as SCo
    g = np.load(os.path.join(GOLD, "round2_small.npz"))
    scene, kf, wins = synth.make_problem("tiny", 0, 1)
    w = wins[0]
    o = O.event_frame(w["x"], w["y"], w["pol"], w["ts"], kf["H"], kf["W"])
    frames, norms = O.event_frame_levels(o["img"], 3)
    assert np.array_equal(frames[1], g["pyr_level1"]) and np.array_equal(frames[2], g["pyr_level2"]) and np.array_equal(np.array(norms), g["pyr_norms"])
    x, tau = w["x_init"].copy(), 0.05
    for k, lvl in enumerate((2, 1, 0)):
        so = O.tracker_solve(kf, frames[lvl] / norms[lvl], x, num_blocks=4, loss_param=tau, max_iterations=int(g["pyr_caps"][lvl]))
        x, tau = so["x"], so["next_loss_param"]
        assert np.allclose(np.concatenate([x, [tau, so["info"]["iterations"], so["info"]["final_cost"]]]), g["pyr_solves"][k], rtol=1e-12, atol=1e-14)
    W, H, L, pts, seed = [int(v) for v in g["cd_kw"]]
    pb = SCo.make_coarse_problem(W=W, H=H, levels=L, points=pts, seed=seed)
    ref = O.make_coarse_depth_l0([dict(w=Lv["w"], h=Lv["h"], dI_ref=Lv["dI_new"]) for Lv in pb["levels"]], *g["cd_points"])
    for lvl, r in enumerate(ref):
        assert np.array_equal(np.stack([r["pc_u"], r["pc_v"], r["pc_idepth"], r["pc_color"]]), g["cd_pc%d" % lvl])
