"""GPU parity tests for the coarse-tracker evaluation (run with -m gpu): CoarseTracker::calcRes fused with
calcGSSSE through the C ABI against the CPU oracle.  Inclusion decisions (in bounds, cutoff) are float32
comparisons evaluated in the same operation order on both sides, so the counts must be exact; the sums
(E, H, b) are gated at 1e-5 relative (the reference accumulates them in float)."""
import numpy as np
import pytest

import edsgpu
from edsgpu import synth_coarse as SC
from oracle import oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rel(a, b):
    return np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(np.asarray(b, np.float64)).max(), 1e-300)


@pytest.mark.parametrize("size", ["small", "vga"])
def test_calc_res_and_gs_match_oracle(gpu_ctx, size):
    pb = SC.make_coarse_problem(W=160, H=120, levels=3, points=3000) if size == "small" else SC.make_coarse_problem()
    ct = edsgpu.CoarseTracker(gpu_ctx, len(pb["levels"]))
    for lvl, L in enumerate(pb["levels"]):
        ct.set_level(lvl, L["w"], L["h"], L["fx"], L["fy"], L["cx"], L["cy"], L["Ki"])
        ct.set_reference(lvl, L["pc_u"], L["pc_v"], L["pc_idepth"], L["pc_color"])
        ct.set_new_frame(lvl, L["dI_new"])
    for lvl, L in enumerate(pb["levels"]):
        ref = O.coarse_calc_res_gs(lvl, L["dI_new"], L["fx"], L["fy"], L["cx"], L["cy"], L["Ki"], pb["R"], pb["t"], pb["affLL"], pb["b0"],
                                   pb["cutoffTH"], L["pc_u"], L["pc_v"], L["pc_idepth"], L["pc_color"])
        got = ct.calc_res_gs(lvl, pb["R"], pb["t"], pb["affLL"], pb["b0"], pb["cutoffTH"])
        assert got["rs"][1] == ref["rs"][1] and got["rs"][5] == ref["rs"][5] and got["rs"][3] == 0   # counts are exact
        assert abs(got["rs"][0] - ref["rs"][0]) <= TOL * ref["rs"][0]
        for k in (2, 4):
            assert abs(got["rs"][k] - ref["rs"][k]) <= TOL * max(ref["rs"][k], 1e-30)
        assert rel(got["H"], ref["H"]) < TOL and rel(got["b"], ref["b"]) < TOL
        assert np.allclose(got["H"], got["H"].T, rtol=0, atol=0)
        # residual-only evaluation (the accept / reject test of trackNewestCoarse)
        only = ct.calc_res_gs(lvl, pb["R"], pb["t"], pb["affLL"], pb["b0"], pb["cutoffTH"], want_system=False)
        assert np.array_equal(only["rs"], got["rs"])
        # deterministic
        again = ct.calc_res_gs(lvl, pb["R"], pb["t"], pb["affLL"], pb["b0"], pb["cutoffTH"])
        assert np.array_equal(again["H"], got["H"]) and np.array_equal(again["b"], got["b"]) and np.array_equal(again["rs"], got["rs"])
    ct.close()


def test_coarse_edge_cases(gpu_ctx):
    pb = SC.make_coarse_problem(W=96, H=64, levels=1, points=500)
    L = pb["levels"][0]
    ct = edsgpu.CoarseTracker(gpu_ctx, 1)
    with pytest.raises(edsgpu.EdsGpuError):
        ct.set_new_frame(0, L["dI_new"])           # level not described yet
    ct.set_level(0, L["w"], L["h"], L["fx"], L["fy"], L["cx"], L["cy"], L["Ki"])
    with pytest.raises(edsgpu.EdsGpuError):
        ct.calc_res_gs(0, pb["R"], pb["t"], pb["affLL"], pb["b0"], pb["cutoffTH"])   # no reference, no frame
    ct.set_new_frame(0, L["dI_new"])
    # every point behind the camera / out of the image: nothing survives, like the reference (0 terms)
    ct.set_reference(0, L["pc_u"], L["pc_v"], -np.abs(L["pc_idepth"]), L["pc_color"])
    r = ct.calc_res_gs(0, pb["R"], pb["t"], pb["affLL"], pb["b0"], pb["cutoffTH"], want_system=False)
    assert r["rs"][0] == 0 and r["rs"][1] == 0
    # a ragged point count (not a multiple of the warp or of 4)
    n = 333
    ct.set_reference(0, L["pc_u"][:n], L["pc_v"][:n], L["pc_idepth"][:n], L["pc_color"][:n])
    ref = O.coarse_calc_res_gs(0, L["dI_new"], L["fx"], L["fy"], L["cx"], L["cy"], L["Ki"], pb["R"], pb["t"], pb["affLL"], pb["b0"], pb["cutoffTH"],
                               L["pc_u"][:n], L["pc_v"][:n], L["pc_idepth"][:n], L["pc_color"][:n])
    got = ct.calc_res_gs(0, pb["R"], pb["t"], pb["affLL"], pb["b0"], pb["cutoffTH"])
    assert got["rs"][1] == ref["rs"][1] and rel(got["H"], ref["H"]) < TOL and rel(got["b"], ref["b"]) < TOL
    ct.close()


def test_gpu_against_committed_coarse_fixture(gpu_ctx):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "coarse_small.npz"))
    W, H, levels, points, seed = [int(v) for v in g["kw"]]
    pb = SC.make_coarse_problem(W=W, H=H, levels=levels, points=points, seed=seed)
    ct = edsgpu.CoarseTracker(gpu_ctx, levels)
    for lvl, L in enumerate(pb["levels"]):
        ct.set_level(lvl, L["w"], L["h"], L["fx"], L["fy"], L["cx"], L["cy"], L["Ki"])
        ct.set_reference(lvl, L["pc_u"], L["pc_v"], L["pc_idepth"], L["pc_color"])
        ct.set_new_frame(lvl, L["dI_new"])
        got = ct.calc_res_gs(lvl, pb["R"], pb["t"], pb["affLL"], pb["b0"], pb["cutoffTH"])
        ref = g["rs%d" % lvl]
        assert got["rs"][1] == ref[1] and got["rs"][5] == ref[5] and abs(got["rs"][0] - ref[0]) <= TOL * ref[0]
        assert rel(got["H"], g["H%d" % lvl]) < TOL and rel(got["b"], g["b%d" % lvl]) < TOL
    ct.close()


@pytest.mark.parametrize("seed,pose_error", [(5, 6e-3), (11, 2e-3), (23, 1e-2)])
def test_track_newest_coarse_matches_the_oracle_loop(gpu_ctx, seed, pose_error):
    """edsgpu_coarse_track (the device-resident loop: one cooperative launch) against the oracle's loop: same number of
    evaluations, same accept / reject history, final pose and brightness parameters to 1e-6."""
    pb = SC.make_coarse_problem(W=320, H=240, levels=4, points=8000, seed=seed, pose_error=pose_error)
    ct = edsgpu.CoarseTracker(gpu_ctx, 4)
    for lvl, L in enumerate(pb["levels"]):
        ct.set_level(lvl, L["w"], L["h"], L["fx"], L["fy"], L["cx"], L["cy"], L["Ki"])
        ct.set_reference(lvl, L["pc_u"], L["pc_v"], L["pc_idepth"], L["pc_color"])
        ct.set_new_frame(lvl, L["dI_new"])
    ref = O.coarse_track(pb, 3, pb["R"], pb["t"])
    got = ct.track(3, pb["R"], pb["t"])
    assert got["ok"] and ref["ok"] and got["evaluations"] == ref["evaluations"]
    assert np.abs(got["R"] - ref["R"]).max() < 1e-6 and np.abs(got["t"] - ref["t"]).max() < 1e-6
    assert np.abs(got["aff"] - ref["aff"]).max() < 1e-4
    assert np.allclose(got["last_residuals"][:4], ref["last_residuals"][:4], rtol=1e-5) and np.isnan(got["last_residuals"][4])
    assert np.allclose(got["last_flow"], ref["last_flow"], rtol=1e-5)
    # failure is reported like the reference's `return false`, the caller's pose is left alone
    bad = ct.track(3, pb["R"], pb["t"], min_res_for_abort=[1e-3] * 5)
    assert not bad["ok"] and np.array_equal(bad["R"], np.asarray(pb["R"], np.float64)) and np.array_equal(bad["t"], pb["t"])
    ct.close()


@pytest.mark.parametrize("size", ["small", "vga"])
def test_make_coarse_depth_l0_matches_oracle(gpu_ctx, size):
    """CoarseTracker::makeCoarseDepthL0 on the device against the CPU restatement: the per-level point clouds have the same
    points in the same (scan-line) order, colours bit for bit, inverse depths to float rounding (two points in one pixel are
    summed in arrival order on the device), and the levels are ready for calcRes afterwards."""
    pb = SC.make_coarse_problem(W=160, H=120, levels=3, points=3000) if size == "small" else SC.make_coarse_problem()
    L = len(pb["levels"])
    rng = np.random.default_rng(17)
    W0, H0 = pb["levels"][0]["w"], pb["levels"][0]["h"]
    n = 2500 if size == "small" else 14000
    cu = rng.uniform(3.0, W0 - 4.0, n).astype(np.float32)
    cv = rng.uniform(3.0, H0 - 4.0, n).astype(np.float32)
    cid = rng.uniform(0.2, 1.5, n).astype(np.float32)
    cid[::211] = -0.3  # rejected: not a positive inverse depth
    hdi = rng.uniform(1e-4, 1e-2, n).astype(np.float32)
    levels = [dict(w=Lv["w"], h=Lv["h"], dI_ref=Lv["dI_new"]) for Lv in pb["levels"]]  # any image serves as the reference frame
    ref = O.make_coarse_depth_l0(levels, cu, cv, cid, hdi)
    ct = edsgpu.CoarseTracker(gpu_ctx, L)
    for lvl, Lv in enumerate(pb["levels"]):
        ct.set_level(lvl, Lv["w"], Lv["h"], Lv["fx"], Lv["fy"], Lv["cx"], Lv["cy"], Lv["Ki"])
        ct.set_reference_frame(lvl, Lv["dI_new"])
        ct.set_new_frame(lvl, Lv["dI_new"])
    pc_n = ct.make_depth_l0(L, cu, cv, cid, hdi)
    assert pc_n == [r["n"] for r in ref] and pc_n[0] > n
    for lvl in range(L):
        g = ct.get_reference(lvl)
        r = ref[lvl]
        assert np.array_equal(g["pc_u"], r["pc_u"]) and np.array_equal(g["pc_v"], r["pc_v"]) and np.array_equal(g["pc_color"], r["pc_color"])
        assert np.abs(g["pc_idepth"] - r["pc_idepth"]).max() <= 2e-6 * np.abs(r["pc_idepth"]).max()
        # the level can be evaluated right away, and gives what an upload of the oracle's point cloud gives
        a = ct.calc_res_gs(lvl, pb["R"], pb["t"], pb["affLL"], pb["b0"], pb["cutoffTH"])
        o = O.coarse_calc_res_gs(lvl, pb["levels"][lvl]["dI_new"], pb["levels"][lvl]["fx"], pb["levels"][lvl]["fy"], pb["levels"][lvl]["cx"],
                                 pb["levels"][lvl]["cy"], pb["levels"][lvl]["Ki"], pb["R"], pb["t"], pb["affLL"], pb["b0"], pb["cutoffTH"],
                                 r["pc_u"], r["pc_v"], r["pc_idepth"], r["pc_color"])
        assert abs(a["rs"][1] - o["rs"][1]) <= 2 and abs(a["rs"][0] - o["rs"][0]) <= 1e-4 * o["rs"][0]
    with pytest.raises(edsgpu.EdsGpuError):
        edsgpu.CoarseTracker(gpu_ctx, 2).make_depth_l0(1, cu, cv, cid, hdi)  # no level set


def test_make_coarse_depth_l0_against_round2_fixture(gpu_ctx):
    """makeCoarseDepthL0 on the device against tests/golden/round2_small.npz, without calling the oracle."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "round2_small.npz"))
    W, H, L, pts, seed = [int(v) for v in g["cd_kw"]]
    pb = SC.make_coarse_problem(W=W, H=H, levels=L, points=pts, seed=seed)
    ct = edsgpu.CoarseTracker(gpu_ctx, L)
    for lvl, Lv in enumerate(pb["levels"]):
        ct.set_level(lvl, Lv["w"], Lv["h"], Lv["fx"], Lv["fy"], Lv["cx"], Lv["cy"], Lv["Ki"])
        ct.set_reference_frame(lvl, Lv["dI_new"])
    pc_n = ct.make_depth_l0(L, *g["cd_points"])
    for lvl in range(L):
        ref = g["cd_pc%d" % lvl]
        got = ct.get_reference(lvl)
        assert pc_n[lvl] == ref.shape[1] == got["n"]
        assert np.array_equal(got["pc_u"], ref[0]) and np.array_equal(got["pc_v"], ref[1]) and np.array_equal(got["pc_color"], ref[3])
        assert np.abs(got["pc_idepth"] - ref[2]).max() <= 2e-6 * np.abs(ref[2]).max()
