"""GPU parity tests for the windowed-BA Hessian accumulators (run with -m gpu): the CUDA path
through the C ABI against the double-precision CPU oracle.  The reference accumulates in float
with a run-to-run varying order, so the gate is norm-wise 1e-5 relative on every accumulator and
on the stitched H, b, not bit-exactness."""
import numpy as np
import pytest

import edsgpu
from edsgpu import synth_ba as SB
from oracle import oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rel(a, b):
    return np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(np.asarray(b, np.float64)).max(), 1e-300)


@pytest.fixture(scope="module", params=["small", "config4", "ragged"])
def problem(request):
    if request.param == "small":
        return SB.make_ba_problem(F=4, points_per_frame=250, H=120, W=160)
    if request.param == "config4":
        return SB.make_ba_problem()  # F=7, 2048 points per keyframe, R = 86 016
    # ragged: drop residuals so that points have 0..F-1 of them, some points none, some keys empty
    pb = SB.make_ba_problem(F=5, points_per_frame=300, H=120, W=160, seed=99)
    rng = np.random.default_rng(5)
    keep = rng.random(pb["R"]) < 0.6
    keep[pb["target_idx"] == 3] &= pb["host_idx"][pb["target_idx"] == 3] != 1  # empty (host 1, target 3) accumulator
    pts = pb["point_of_res"][keep]
    out = dict(pb)
    out["recs"], out["host_idx"], out["target_idx"], out["flags"] = pb["recs"][keep], pb["host_idx"][keep], pb["target_idx"][keep], pb["flags"][keep]
    out["res_begin"] = np.concatenate([[0], np.cumsum(np.bincount(pts, minlength=pb["P"]))]).astype(np.int32)
    out["R"] = int(keep.sum())
    return out


@pytest.fixture(scope="module")
def solved(gpu_ctx, problem):
    pb = problem
    F = pb["F"]
    rtz = O.ba_fix_linearization(F, pb["recs"], pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["deltaF"], pb["adHTdeltaF"], pb["cDeltaF"])
    w = edsgpu.BaWindow(gpu_ctx, F, pb["host_idx"], pb["target_idx"], pb["res_begin"])
    w.set_residuals(pb["recs"], pb["flags"], rtz)
    w.set_points(pb["deltaF"], pb["priorF"])
    w.set_frames(pb["adHTdeltaF"], pb["cDeltaF"], SB.col_major(pb["adHost"]), SB.col_major(pb["adTarget"]))
    yield pb, rtz, w
    w.close()


def _oracle_top(pb, rtz, mode):
    return O.ba_top_accumulate(mode, pb["F"], pb["recs"], pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["flags"], rtz, pb["deltaF"],
                               pb["adHTdeltaF"], pb["cDeltaF"], threads=4)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_top_accumulators(solved, mode):
    pb, rtz, w = solved
    g = w.top_accumulate(mode)
    o = _oracle_top(pb, rtz, mode)
    assert g["nres"] == o["nres"]
    for k in range(pb["F"] ** 2):
        if o["num"][k] == 0:
            assert not g["acc"][k].any()
        else:
            assert rel(g["acc"][k], o["acc"][k]) < TOL, k
            assert np.array_equal(g["acc"][k], g["acc"][k].T)
    assert rel(g["Hdd"], o["Hdd"]) < TOL and rel(g["bd"], o["bd"]) < TOL and rel(g["Hcd"], o["Hcd"]) < TOL
    # deterministic: fixed tile plan and summation order
    g2 = w.top_accumulate(mode)
    assert np.array_equal(g["acc"], g2["acc"]) and np.array_equal(g["Hdd"], g2["Hdd"])


def test_jpjd(solved):
    pb, rtz, w = solved
    assert rel(w.jpjd(), O.ba_jpjd(pb["recs"])) < 1e-6


@pytest.mark.parametrize("use_prior", [False, True])
def test_top_stitch(solved, use_prior):
    pb, rtz, w = solved
    F = pb["F"]
    for which, mode in ((0, 0), (1, 1)):
        w.top_accumulate(mode, want_outputs=False)
        o = _oracle_top(pb, rtz, mode)
        args = (True, pb["cPrior"], pb["cDeltaF"], pb["frame_prior"], pb["frame_delta_prior"]) if use_prior else ()
        Ho, bo = O.ba_top_stitch(F, o["acc"], SB.col_major(pb["adHost"]), SB.col_major(pb["adTarget"]), *args)
        Hg, bg = w.top_stitch(which, use_prior, pb["cPrior"], pb["frame_prior"], pb["frame_delta_prior"])
        assert rel(Hg, Ho) < TOL and rel(bg, bo) < TOL
        assert rel(Hg, Hg.T) < 1e-12


@pytest.mark.parametrize("shift", [True, False])
def test_schur_complement(solved, shift):
    pb, rtz, w = solved
    F = pb["F"]
    w.top_accumulate(0, want_outputs=False)
    w.top_accumulate(1, want_outputs=False)
    A, L = _oracle_top(pb, rtz, 0), _oracle_top(pb, rtz, 1)
    g = w.sc_accumulate(shift)
    o = O.ba_sc_accumulate(F, pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["flags"], O.ba_jpjd(pb["recs"]), A["Hdd"], L["Hdd"],
                           A["bd"], L["bd"], A["Hcd"], L["Hcd"], pb["priorF"], pb["deltaF"], shift, threads=4)
    assert rel(g["HdiF"], o["HdiF"]) < TOL and rel(g["bdSum"], o["bdSum"]) < TOL
    for name in ("accD", "accE", "accEB", "accHcc", "accbc"):
        assert rel(g[name], o[name]) < TOL, name
    Hg, bg = w.sc_stitch()
    Ho, bo = O.ba_sc_stitch(F, o, SB.col_major(pb["adHost"]), SB.col_major(pb["adTarget"]))
    assert rel(Hg, Ho) < TOL and rel(bg, bo) < TOL
    # a point without active residuals contributes nothing and gets HdiF = 0 (AccumulatedSCHessian.cpp:38-45)
    none = [p for p in range(pb["P"]) if not (pb["flags"][pb["res_begin"][p]:pb["res_begin"][p + 1]] & 1).any()]
    assert all(g["HdiF"][p] == 0 and g["bdSum"][p] == 0 for p in none[:50])


def test_marginalisation_sequence_mode2_then_schur(solved):
    """marginalizePointsF (EnergyFunctional.cpp:538-560): addPoint<2> followed by the Schur addPoint(p, false) with no
    separate active pass -- mode 2 defines both point-term sides (the active one as zeros, AccumulatedTopHessian.cpp:152-157)."""
    pb, rtz, w = solved
    F = pb["F"]
    w.set_residuals(pb["recs"], pb["flags"], rtz)  # resets every freshness flag
    with pytest.raises(edsgpu.EdsGpuError):
        w.sc_accumulate(False)
    w.top_accumulate(2, want_outputs=False)
    g = w.sc_accumulate(False)
    M = _oracle_top(pb, rtz, 2)
    z1, z4 = np.zeros(pb["P"], np.float32), np.zeros((pb["P"], 4), np.float32)
    o = O.ba_sc_accumulate(F, pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["flags"], O.ba_jpjd(pb["recs"]), z1, M["Hdd"],
                           z1, M["bd"], z4, M["Hcd"], pb["priorF"], pb["deltaF"], False, threads=4)
    assert rel(g["HdiF"], o["HdiF"]) < TOL and rel(g["bdSum"], o["bdSum"]) < TOL
    for name in ("accD", "accE", "accEB", "accHcc", "accbc"):
        assert rel(g[name], o[name]) < TOL, name
    Hg, bg = w.sc_stitch()
    Ho, bo = O.ba_sc_stitch(F, o, SB.col_major(pb["adHost"]), SB.col_major(pb["adTarget"]))
    assert rel(Hg, Ho) < TOL and rel(bg, bo) < TOL


def test_stale_accumulations_are_refused(solved):
    """Whatever changes an input of an accumulation invalidates it: new deltas (mode 1 and the Schur terms), new frame
    deltas, fixLinearizationF (flags and res_toZero), a newer top accumulation (the Schur terms were built from the old one)."""
    pb, rtz, w = solved
    x = np.zeros(4 + 8 * pb["F"])

    def all_three():
        w.top_accumulate(0, want_outputs=False); w.top_accumulate(1, want_outputs=False); w.sc_accumulate(True, want_outputs=False)

    all_three(); w.resubstitute(x)
    w.set_points(pb["deltaF"], pb["priorF"])
    with pytest.raises(edsgpu.EdsGpuError):
        w.sc_accumulate(True)  # the linearized side is stale
    all_three()
    w.set_frames(pb["adHTdeltaF"], pb["cDeltaF"], SB.col_major(pb["adHost"]), SB.col_major(pb["adTarget"]))
    with pytest.raises(edsgpu.EdsGpuError):
        w.resubstitute(x)
    all_three()
    w.top_accumulate(0, want_outputs=False)
    with pytest.raises(edsgpu.EdsGpuError):
        w.resubstitute(x)  # Schur terms older than the active accumulation
    all_three()
    w.fix_linearization(np.zeros(pb["R"], np.uint8))
    with pytest.raises(edsgpu.EdsGpuError):
        w.sc_accumulate(True)
    w.set_residuals(pb["recs"], pb["flags"], rtz)


@pytest.mark.parametrize("cfg", ["small", "config4"])
def test_device_linearize_is_bit_exact_and_feeds_the_accumulators(gpu_ctx, cfg):
    """PointFrameResidual::linearize on the device against the float32 oracle: same
    bits for the records, same ResState, same energies; the accumulators fed by the device-made records
    give the same result as when the records are uploaded."""
    pb = SB.make_ba_problem(F=4, points_per_frame=250, H=120, W=160) if cfg == "small" else SB.make_ba_problem()
    ref_recs, ref_state, ref_energy = O.ba_linearize(pb["F"], pb["H"], pb["W"], pb["dI"], pb["precalc"], pb["calib"], pb["pu"], pb["pv"],
                                                     pb["idepth"], pb["idepth"], pb["color"], pb["weights"], pb["host_idx"],
                                                     pb["target_idx"], pb["res_begin"], pb["frame_energy_th"])
    w = edsgpu.BaWindow(gpu_ctx, pb["F"], pb["host_idx"], pb["target_idx"], pb["res_begin"])
    w.set_images(pb["dI"])
    w.set_linearize_inputs(pb["precalc"], pb["calib"], pb["frame_energy_th"], pb["pu"], pb["pv"], pb["idepth"], pb["idepth"],
                           pb["color"], pb["weights"])
    linearized = (pb["flags"] >> 1) & 1
    rtz = O.ba_fix_linearization(pb["F"], pb["recs"], pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["deltaF"], pb["adHTdeltaF"],
                                 pb["cDeltaF"])
    state, energy = w.linearize(linearized=linearized, res_toZero=rtz)
    recs, flags = w.get_residuals()
    assert np.array_equal(state, ref_state) and {0, 1} <= set(np.unique(state))
    assert np.array_equal(recs, ref_recs)
    assert np.array_equal(energy, ref_energy)
    assert np.array_equal(flags, pb["flags"])
    assert np.array_equal(w.jpjd(), O.ba_jpjd(ref_recs)) or rel(w.jpjd(), O.ba_jpjd(ref_recs)) < 1e-6
    # downstream: identical to the upload path
    w.set_points(pb["deltaF"], pb["priorF"])
    w.set_frames(pb["adHTdeltaF"], pb["cDeltaF"], SB.col_major(pb["adHost"]), SB.col_major(pb["adTarget"]))
    a = w.top_accumulate(0)
    w2 = edsgpu.BaWindow(gpu_ctx, pb["F"], pb["host_idx"], pb["target_idx"], pb["res_begin"])
    w2.set_residuals(pb["recs"], pb["flags"], rtz)
    w2.set_points(pb["deltaF"], pb["priorF"])
    w2.set_frames(pb["adHTdeltaF"], pb["cDeltaF"], SB.col_major(pb["adHost"]), SB.col_major(pb["adTarget"]))
    b = w2.top_accumulate(0)
    assert np.array_equal(a["acc"], b["acc"]) and np.array_equal(a["Hdd"], b["Hdd"]) and a["nres"] == b["nres"]
    # a sticky OOB input state is honoured
    st_in = np.zeros(pb["R"], np.uint8); st_in[::7] = 1
    state2, _ = w.linearize(state_in=st_in)
    assert np.all(state2[::7] == 1) and np.array_equal(np.delete(state2, np.s_[::7]), np.delete(ref_state, np.s_[::7]))
    w.close(); w2.close()


@pytest.mark.parametrize("cfg", ["small", "config4"])
def test_fused_linearize_accumulate_is_bit_identical_to_the_two_calls(gpu_ctx, cfg):
    """edsgpu_ba_linearize_accumulate (linearize + addPoint<0> in one kernel): every output equals the
    output of linearize followed by top_accumulate(0) bit for bit; without write_records only the linearized residuals'
    records reach memory, and the rest of the chain (mode 1, Schur complement) gives the same result."""
    pb = SB.make_ba_problem(F=4, points_per_frame=250, H=120, W=160) if cfg == "small" else SB.make_ba_problem()
    linearized = (pb["flags"] >> 1) & 1
    rtz = O.ba_fix_linearization(pb["F"], pb["recs"], pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["deltaF"], pb["adHTdeltaF"],
                                 pb["cDeltaF"])

    def window():
        w = edsgpu.BaWindow(gpu_ctx, pb["F"], pb["host_idx"], pb["target_idx"], pb["res_begin"])
        w.set_images(pb["dI"])
        w.set_linearize_inputs(pb["precalc"], pb["calib"], pb["frame_energy_th"], pb["pu"], pb["pv"], pb["idepth"], pb["idepth"],
                               pb["color"], pb["weights"])
        w.set_points(pb["deltaF"], pb["priorF"])
        w.set_frames(pb["adHTdeltaF"], pb["cDeltaF"], SB.col_major(pb["adHost"]), SB.col_major(pb["adTarget"]))
        return w

    a, b = window(), window()
    sa, ea = a.linearize(linearized=linearized, res_toZero=rtz)
    ta = a.top_accumulate(0)
    for write_all in (True, False):
        sb, eb = b.linearize_accumulate(linearized=linearized, res_toZero=rtz, write_records=write_all)
        tb = b.top_result(0)
        assert np.array_equal(sa, sb) and np.array_equal(ea, eb)
        for k in ("acc", "Hdd", "bd", "Hcd"):
            assert np.array_equal(ta[k], tb[k]), k
        assert ta["nres"] == tb["nres"] and ta["nres"] > 0
        assert np.array_equal(a.jpjd(), b.jpjd())
        ra, fa = a.get_residuals()
        rb, fb = b.get_residuals()
        assert np.array_equal(fa, fb)
        if write_all:
            assert np.array_equal(ra, rb)
        else:
            assert np.array_equal(ra[linearized == 1], rb[linearized == 1])
        # the rest of the chain
        la, lb = a.top_accumulate(1), b.top_accumulate(1)
        for k in ("acc", "Hdd", "bd", "Hcd"):
            assert np.array_equal(la[k], lb[k]), k
        ca, cb = a.sc_accumulate(True), b.sc_accumulate(True)
        for k in ca:
            assert np.array_equal(ca[k], cb[k]), k
        if write_all:  # a fresh window for the sparse-record pass: no record left over from this one
            b.close()
            b = window()
    a.close(); b.close()


def test_back_substitution_energy_and_fix_linearization(solved):
    """The steps right after the solve on data that is still on the device:
    resubstituteF_MT, calcLEnergyF_MT, fixLinearizationF against the CPU oracle."""
    pb, rtz, w = solved
    F = pb["F"]
    rng = np.random.default_rng(11)
    x = rng.normal(scale=1e-3, size=4 + 8 * F)
    top0, top1, sc = w.top_accumulate(0), w.top_accumulate(1), w.sc_accumulate(True)
    step = w.resubstitute(x)
    ref = O.ba_resubstitute(F, x, SB.col_major(pb["adHost"]), SB.col_major(pb["adTarget"]), pb["host_idx"], pb["target_idx"],
                            pb["res_begin"], pb["flags"], w.jpjd(), sc["bdSum"], top0["Hcd"], top1["Hcd"], sc["HdiF"])
    assert rel(step, ref) < TOL
    nores = np.diff(pb["res_begin"]) == 0
    assert np.all(step[nores] == 0)
    # linearised energy
    e = w.calc_l_energy(pb["cPrior"], pb["frame_prior"], pb["frame_delta_prior"])
    e_ref = O.ba_calc_l_energy(F, pb["recs"], pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["flags"], rtz, pb["deltaF"], pb["priorF"],
                               pb["adHTdeltaF"], pb["cDeltaF"], pb["cPrior"], pb["frame_prior"], pb["frame_delta_prior"])
    assert abs(e - e_ref) <= TOL * abs(e_ref)
    e0 = w.calc_l_energy()
    e0_ref = O.ba_calc_l_energy(F, pb["recs"], pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["flags"], rtz, pb["deltaF"], pb["priorF"],
                                pb["adHTdeltaF"], pb["cDeltaF"])
    assert abs(e0 - e0_ref) <= TOL * max(abs(e0_ref), 1e-30)
    # fixLinearizationF for a subset: same arithmetic as the oracle, so the same bits; flags gain LINEARIZED
    sel = (rng.random(pb["R"]) < 0.3).astype(np.uint8)
    out = w.fix_linearization(sel)
    ref_rtz = O.ba_fix_linearization(F, pb["recs"], pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["deltaF"], pb["adHTdeltaF"], pb["cDeltaF"])
    m = sel.astype(bool)
    assert rel(out[m], ref_rtz[m]) < 1e-6
    assert np.array_equal(out[~m], rtz[~m])
    _, flags = w.get_residuals()
    assert np.all(flags[m] & 2) and np.array_equal(flags[~m], pb["flags"][~m])
    w.set_residuals(pb["recs"], pb["flags"], rtz)  # restore for the other tests of the module


@pytest.mark.parametrize("with_prior", [False, True])
def test_device_solve_closes_the_gauss_newton_iteration(solved, with_prior):
    """edsgpu_ba_solve_system = EnergyFunctional::solveSystemF (default solver mode) + resubstituteF_MT without x leaving the
    device: against the restatement fed with the GPU's own stitched systems, and the point steps against resubstitute(x)."""
    pb, rtz, w = solved
    F, n = pb["F"], 4 + 8 * pb["F"]
    rng = np.random.default_rng(23)
    w.top_accumulate(0); w.top_accumulate(1); w.sc_accumulate(True)
    pri = (pb["cPrior"], pb["frame_prior"], pb["frame_delta_prior"]) if with_prior else (None, None, None)
    HA, bA = w.top_stitch(0)
    HL, bL = w.top_stitch(1, with_prior, *pri)
    Hsc, bsc = w.sc_stitch()
    HM = bM = delta = P = None
    if with_prior:
        a = rng.normal(size=(n, n))
        HM = 1e-2 * np.abs(HA).max() * (a @ a.T) / n
        bM = 1e-2 * np.abs(bA).max() * rng.normal(size=n)
        delta = 1e-3 * rng.normal(size=n)
        P = O.ba_nullspace_projector([rng.normal(size=n) for _ in range(7)])
    x, step = w.solve_system(1e-5, HM, bM, delta, *pri, projector=P)
    ref = O.ba_solve_system(HA, bA, HL, bL, Hsc, bsc, 1e-5, HM, bM, delta, P)
    assert np.isfinite(x).all() and np.abs(x).max() > 0
    assert np.abs(x - ref).max() <= 1e-8 * np.abs(ref).max()
    # the point steps are those of resubstituteF_MT for this x (same float arithmetic on the device as on the host path)
    assert np.array_equal(step, w.resubstitute(x))
    with pytest.raises(edsgpu.EdsGpuError):
        w.solve_system(1e-5, HM=np.eye(n))  # HM without bM


def test_gpu_against_committed_ba_fixture(gpu_ctx):
    """The CUDA path against tests/golden/ba_small.npz, without rebuilding or calling the oracle."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ba_small.npz"))
    F, ppf, H, W, seed = [int(v) for v in g["kw"]]
    pb = SB.make_ba_problem(F=F, points_per_frame=ppf, H=H, W=W, seed=seed)
    w = edsgpu.BaWindow(gpu_ctx, F, pb["host_idx"], pb["target_idx"], pb["res_begin"])
    w.set_images(pb["dI"])
    w.set_linearize_inputs(pb["precalc"], pb["calib"], pb["frame_energy_th"], pb["pu"], pb["pv"], pb["idepth"], pb["idepth"], pb["color"], pb["weights"])
    state, energy = w.linearize(linearized=(pb["flags"] >> 1) & 1, res_toZero=g["res_toZero"])
    recs, _ = w.get_residuals()
    assert np.array_equal(recs, g["recs"]) and np.array_equal(state, g["state"]) and np.array_equal(energy, g["energy"])
    w.set_points(pb["deltaF"], pb["priorF"])
    w.set_frames(pb["adHTdeltaF"], pb["cDeltaF"], SB.col_major(pb["adHost"]), SB.col_major(pb["adTarget"]))
    t0, t1, sc = w.top_accumulate(0), w.top_accumulate(1), w.sc_accumulate(True)
    assert rel(t0["acc"], g["acc0"]) < TOL and rel(t1["acc"], g["acc1"]) < TOL and rel(t0["Hdd"], g["Hdd0"]) < TOL
    assert rel(sc["accD"], g["accD"]) < TOL and rel(sc["HdiF"], g["HdiF"]) < TOL and rel(sc["bdSum"], g["bdSum"]) < TOL
    assert rel(w.resubstitute(g["x"]), g["step"]) < TOL
    e = w.calc_l_energy(pb["cPrior"], pb["frame_prior"], pb["frame_delta_prior"])
    assert abs(e - float(g["l_energy"])) <= TOL * abs(float(g["l_energy"]))
    w.close()


def test_device_solve_against_round2_fixture(gpu_ctx):
    """edsgpu_ba_solve_system on the committed small window against tests/golden/round2_small.npz (the oracle's accumulators,
    stitches and solve), without calling the oracle: float accumulators on both sides, so the gate is the accumulators' 1e-5
    scaled by the conditioning the solve adds."""
    import os
    gb = np.load(os.path.join(os.path.dirname(__file__), "golden", "ba_small.npz"))
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "round2_small.npz"))
    F, ppf, H, W, seed = [int(v) for v in gb["kw"]]
    pb = SB.make_ba_problem(F=F, points_per_frame=ppf, H=H, W=W, seed=seed)
    w = edsgpu.BaWindow(gpu_ctx, F, pb["host_idx"], pb["target_idx"], pb["res_begin"])
    w.set_residuals(gb["recs"], pb["flags"], gb["res_toZero"])
    w.set_points(pb["deltaF"], pb["priorF"])
    w.set_frames(pb["adHTdeltaF"], pb["cDeltaF"], SB.col_major(pb["adHost"]), SB.col_major(pb["adTarget"]))
    w.top_accumulate(0); w.top_accumulate(1); w.sc_accumulate(True)
    pri = (pb["cPrior"], pb["frame_prior"], pb["frame_delta_prior"])
    x_plain, _ = w.solve_system(1e-5, None, None, None, *pri)
    x_full, _ = w.solve_system(1e-5, g["solve_HM"], g["solve_bM"], g["solve_delta"], *pri, projector=g["solve_P"])
    assert rel(x_plain, g["solve_x_plain"]) < 1e-3 and rel(x_full, g["solve_x_full"]) < 1e-3
    w.close()


def test_linearize_needs_its_inputs(gpu_ctx):
    pb = SB.make_ba_problem(F=3, points_per_frame=50, H=64, W=80)
    w = edsgpu.BaWindow(gpu_ctx, pb["F"], pb["host_idx"], pb["target_idx"], pb["res_begin"])
    with pytest.raises(edsgpu.EdsGpuError):
        w.linearize()
    # the steps after the solve need the accumulations of the current linearisation
    w.set_residuals(pb["recs"], pb["flags"], np.zeros((pb["R"], 8), np.float32))
    w.set_points(pb["deltaF"], pb["priorF"])
    w.set_frames(pb["adHTdeltaF"], pb["cDeltaF"], SB.col_major(pb["adHost"]), SB.col_major(pb["adTarget"]))
    with pytest.raises(edsgpu.EdsGpuError):
        w.sc_accumulate(True)
    with pytest.raises(edsgpu.EdsGpuError):
        w.resubstitute(np.zeros(4 + 8 * pb["F"]))
    w.set_linearize_inputs(pb["precalc"], pb["calib"], pb["frame_energy_th"], pb["pu"], pb["pv"], pb["idepth"], pb["idepth"],
                           pb["color"], pb["weights"])
    with pytest.raises(edsgpu.EdsGpuError):
        w.linearize()  # images missing
    w.close()


def test_argument_validation(gpu_ctx):
    with pytest.raises(edsgpu.EdsGpuError):
        edsgpu.BaWindow(gpu_ctx, 9, [0], [1], [0, 1])  # F > 8
    with pytest.raises(edsgpu.EdsGpuError):
        edsgpu.BaWindow(gpu_ctx, 3, [0, 1], [1, 2], [0, 2])  # residuals of one point with two hosts
    with pytest.raises(edsgpu.EdsGpuError):
        edsgpu.BaWindow(gpu_ctx, 3, [0, 5], [1, 2], [0, 1, 2])  # frame index out of range
    w = edsgpu.BaWindow(gpu_ctx, 3, [0, 0], [1, 2], [0, 2])
    with pytest.raises(edsgpu.EdsGpuError):
        w.top_accumulate(3)
    w.close()
