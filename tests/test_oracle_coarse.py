"""CPU tests of the coarse-tracker oracle (CoarseTracker::calcRes + calcGSSSE).
The reference has no golden vectors for this path (parity unpinned): the checks are properties the
algorithm must have on the synthetic plane scene."""
import numpy as np

from edsgpu import synth_coarse as SC
from oracle import oracle as O


def _eval(pb, lvl, R, t, aff=None):
    L = pb["levels"][lvl]
    return O.coarse_calc_res_gs(lvl, L["dI_new"], L["fx"], L["fy"], L["cx"], L["cy"], L["Ki"], R, t, pb["affLL"] if aff is None else aff,
                                pb["b0"], pb["cutoffTH"], L["pc_u"], L["pc_v"], L["pc_idepth"], L["pc_color"])


def test_counts_rejections_and_system_shape():
    pb = SC.make_coarse_problem(W=160, H=120, levels=3, points=3000)
    for lvl in range(3):
        r = _eval(pb, lvl, pb["R"], pb["t"])
        n = len(pb["levels"][lvl]["pc_u"])
        nE, nW, nSat = r["counts"]
        assert 0 < nW <= nE <= n and nE == nW + nSat and nSat > 0       # the +60 colours saturate, the negative depths are dropped
        assert nE <= n - len(range(0, n, 97))
        assert r["rs"][1] == nE and np.isclose(r["rs"][5], np.float32(nSat) / np.float32(nE))
        assert np.allclose(r["H"], r["H"].T) and np.linalg.eigvalsh(r["H"]).min() > -1e-6 * np.abs(r["H"]).max()
        assert np.all(np.isfinite(r["b"]))
        if lvl == 0:
            assert r["rs"][2] > 0 and r["rs"][4] > 0                    # mean flow is sampled on level 0 only (:403)
        else:
            assert r["rs"][2] == 0 and r["rs"][4] == 0


def test_energy_is_lower_at_the_true_pose_and_a_newton_step_reduces_it():
    pb = SC.make_coarse_problem(W=160, H=120, levels=1, points=4000, seed=3, pose_error=0.0)
    L = pb["levels"][0]
    L["pc_color"][::53] -= 60.0  # remove the planted outliers for this test
    aff = np.array([1.0, 0.0], np.float32)
    good = _eval(pb, 0, pb["R"], pb["t"], aff)
    bad_t = pb["t"] + np.array([0.02, -0.015, 0.0])
    bad = _eval(pb, 0, pb["R"], bad_t, aff)
    assert good["rs"][0] / good["rs"][1] < 0.5 * bad["rs"][0] / bad["rs"][1]
    # b is the gradient of the mean energy wrt the left pose increment: moving the translation along -H^-1 b lowers E
    inc = -np.linalg.solve(bad["H"] + 1e-3 * np.eye(8), bad["b"])
    stepped = _eval(pb, 0, pb["R"], bad_t + inc[:3], aff)   # SE3 increment: translation first (NumType: Vec6 = [trans, rot])
    assert stepped["rs"][0] / stepped["rs"][1] < bad["rs"][0] / bad["rs"][1]


def test_oracle_against_committed_coarse_fixture():
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "coarse_small.npz"))
    W, H, levels, points, seed = [int(v) for v in g["kw"]]
    pb = SC.make_coarse_problem(W=W, H=H, levels=levels, points=points, seed=seed)
    for lvl in range(levels):
        r = _eval(pb, lvl, pb["R"], pb["t"])
        assert np.array_equal(r["counts"], g["counts%d" % lvl])
        assert np.array_equal(r["rs"], g["rs%d" % lvl]) and np.array_equal(r["H"], g["H%d" % lvl]) and np.array_equal(r["b"], g["b%d" % lvl])


def _angle(Ra, Rb):
    return float(np.arccos(np.clip((np.trace(Ra.T @ Rb) - 1) / 2, -1, 1)))


def test_track_newest_coarse_recovers_the_pose():
    """trackNewestCoarse (oracle restatement incl. Sophus SE3::exp and the 8x8 LDLT) on the synthetic pyramid: from a
    wrong start the coarse-to-fine loop must land near the true relative pose (the point depths carry 1 % noise)."""
    pb = SC.make_coarse_problem(W=320, H=240, levels=4, points=8000, seed=5, pose_error=6e-3)
    for L in pb["levels"]:
        L["pc_color"][::53] -= 60.0   # no planted outliers here
    e0 = (_angle(pb["R"], pb["R_true"]), np.linalg.norm(pb["t"] - pb["t_true"]))
    r = O.coarse_track(pb, 3, pb["R"], pb["t"])
    e1 = (_angle(r["R"], pb["R_true"]), np.linalg.norm(r["t"] - pb["t_true"]))
    assert r["ok"] and 8 <= r["evaluations"] <= 200
    assert e1[0] < 0.15 * e0[0] and e1[1] < 0.1 * e0[1]
    assert np.all(np.isfinite(r["last_residuals"][:4])) and np.isnan(r["last_residuals"][4])
    assert abs(r["aff"][0]) < 0.01 and abs(r["aff"][1]) < 0.5        # both frames were rendered with the same brightness
    # the abort threshold of the caller is honoured (:668)
    bad = O.coarse_track(pb, 3, pb["R"], pb["t"], min_res_for_abort=[1e-3] * 5)
    assert not bad["ok"]


def test_make_coarse_depth_l0_hand_checked():
    """makeCoarseDepthL0 (CoarseTracker.cpp:127-283) on a case small enough to follow by hand: one point -> its pixel at level 0,
    the 2x2 sum at level 1, the diagonal dilation ring on levels 0-1, weight-independent normalisation, scan-line order, and
    a point with a non-finite reference colour dropped."""
    W, H, L = 32, 24, 3
    levels = [dict(w=W >> l, h=H >> l, dI_ref=np.full((H >> l, W >> l, 3), 7.0 + l, np.float32)) for l in range(L)]
    out = O.make_coarse_depth_l0(levels, [10.2], [8.4], [0.5], [1e-3])
    l0 = out[0]
    # level 0: the point at (10, 8) and its four diagonal neighbours (dilation), in scan-line order
    assert l0["n"] == 5
    assert list(zip(l0["pc_u"], l0["pc_v"])) == [(9, 7), (11, 7), (10, 8), (9, 9), (11, 9)]
    assert np.allclose(l0["pc_idepth"], 0.5, rtol=1e-6) and np.all(l0["pc_color"] == 7.0)
    # level 1: pixel (5, 4) + diagonal ring; level 2: pixel (2, 2) + axis neighbours, of which only those inside [2, w-2) x [2, h-2)
    assert (5.0, 4.0) in list(zip(out[1]["pc_u"], out[1]["pc_v"])) and out[1]["n"] == 5
    assert list(zip(out[2]["pc_u"], out[2]["pc_v"])) == [(2, 2), (3, 2), (2, 3)]
    assert np.allclose(out[2]["pc_idepth"], 0.5, rtol=1e-6)
    # two points in one pixel: weighted mean with weights sqrt(1e-3 / HdiF)
    out = O.make_coarse_depth_l0(levels, [10.0, 10.3], [8.0, 8.2], [0.4, 0.8], [1e-3, 4e-3])
    w1, w2 = np.sqrt(1.0), np.sqrt(0.25)
    k = list(zip(out[0]["pc_u"], out[0]["pc_v"])).index((10, 8))
    assert abs(out[0]["pc_idepth"][k] - (0.4 * w1 + 0.8 * w2) / (w1 + w2)) < 1e-6
    # non-finite reference colour: the pixel is skipped
    levels[0]["dI_ref"][8, 10, 0] = np.nan
    out = O.make_coarse_depth_l0(levels, [10.2], [8.4], [0.5], [1e-3])
    assert out[0]["n"] == 4 and (10, 8) not in list(zip(out[0]["pc_u"], out[0]["pc_v"]))
