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
