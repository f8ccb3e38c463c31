"""CPU tests pinning the BA oracle: the accumulators against the dense J^T J of the stacked 8x13
rows, the two in-repo statements of each stitch, and the re-linearisation identity
(AccumulatedTopHessian.cpp:84-98 adds J*delta, EnergyFunctionalStructs.cpp:101-110 subtracts it)."""
import numpy as np
import pytest

from oracle import oracle as O
from edsgpu import synth_ba as SB


@pytest.fixture(scope="module")
def small():
    return SB.make_ba_problem(F=4, points_per_frame=250, H=120, W=160)


def _dense_top(pb, mode, res_override=None):
    F = pb["F"]
    recs = pb["recs"].astype(np.float64)
    acc = np.zeros((F * F, 13, 13))
    for r in range(pb["R"]):
        fl = pb["flags"][r]
        act, lin = fl & 1, fl & 2
        if (mode == 0 and (lin or not act)) or (mode == 1 and (not lin or not act)) or (mode == 2 and not act):
            continue
        J = recs[r]
        x = np.concatenate([J[20:24], J[8:14]])
        y = np.concatenate([J[24:28], J[14:20]])
        res = J[0:8] if res_override is None else res_override[r].astype(np.float64)
        rows = np.zeros((8, 13))
        rows[:, :10] = np.outer(J[32:40], x) + np.outer(J[40:48], y)
        rows[:, 10], rows[:, 11], rows[:, 12] = J[48:56], J[56:64], res
        acc[pb["host_idx"][r] + F * pb["target_idx"][r]] += rows.T @ rows
    return acc


def test_generator_produces_a_sane_graph(small):
    pb = small
    assert pb["recs"].shape == (pb["R"], 76) and pb["R"] == pb["P"] * (pb["F"] - 1)
    assert (pb["state"] == 0).mean() > 0.8
    assert np.all(np.isfinite(pb["recs"]))
    act = (pb["flags"] & 1).astype(bool)
    # shorthand 2x2 products are consistent with the stored columns (Residuals.cpp:216-247)
    J = pb["recs"][act].astype(np.float64)
    np.testing.assert_allclose(J[:, 64], np.sum(J[:, 32:40] ** 2, 1), rtol=1e-5)
    np.testing.assert_allclose(J[:, 66], np.sum(J[:, 32:40] * J[:, 40:48], 1), rtol=1e-4, atol=1e-2)
    np.testing.assert_allclose(J[:, 75], np.sum(J[:, 56:64] ** 2, 1), rtol=1e-5)


def test_top_accumulator_equals_dense_normal_equations(small):
    pb = small
    top = O.ba_top_accumulate(0, pb["F"], pb["recs"], pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["flags"])
    dense = _dense_top(pb, 0)
    # records store the 2x2 shorthands in float32: agreement to float32 rounding of those sums
    assert np.abs(top["acc"] - dense).max() < 1e-6 * np.abs(dense).max()
    assert top["nres"] == int(np.sum(((pb["flags"] & 1) == 1) & ((pb["flags"] & 2) == 0)))
    np.testing.assert_array_equal(top["num"].sum(), top["nres"])
    for k in range(pb["F"] ** 2):
        np.testing.assert_allclose(top["acc"][k], top["acc"][k].T)


def test_threads_do_not_change_the_double_result(small):
    pb = small
    a = O.ba_top_accumulate(0, pb["F"], pb["recs"], pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["flags"], threads=1)
    b = O.ba_top_accumulate(0, pb["F"], pb["recs"], pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["flags"], threads=6)
    np.testing.assert_allclose(a["acc"], b["acc"], rtol=1e-12, atol=1e-6)
    np.testing.assert_array_equal(a["Hdd"], b["Hdd"])


def test_relinearisation_identity(small):
    """mode 1 re-adds J*delta to res_toZero = resF - J*delta, so it must reproduce resF."""
    pb = small
    F = pb["F"]
    rtz = O.ba_fix_linearization(F, pb["recs"], pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["deltaF"], pb["adHTdeltaF"], pb["cDeltaF"])
    lin = O.ba_top_accumulate(1, F, pb["recs"], pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["flags"], rtz, pb["deltaF"],
                              pb["adHTdeltaF"], pb["cDeltaF"])
    dense = _dense_top(pb, 1)  # with resF
    assert np.abs(lin["acc"] - dense).max() < 2e-5 * np.abs(dense).max()
    marg = O.ba_top_accumulate(2, F, pb["recs"], pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["flags"], rtz)
    dense2 = _dense_top(pb, 2, rtz)
    assert np.abs(marg["acc"] - dense2).max() < 1e-6 * np.abs(dense2).max()


def test_per_point_terms(small):
    pb = small
    top = O.ba_top_accumulate(0, pb["F"], pb["recs"], pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["flags"])
    J = pb["recs"].astype(np.float64)
    p = 17
    Hdd = bd = 0.0
    Hcd = np.zeros(4)
    for r in range(pb["res_begin"][p], pb["res_begin"][p + 1]):
        if (pb["flags"][r] & 1) and not (pb["flags"][r] & 2):
            M = np.array([[J[r, 64], J[r, 66]], [J[r, 65], J[r, 67]]])
            d = J[r, 28:30]
            JI_r = np.array([J[r, 0:8] @ J[r, 32:40], J[r, 0:8] @ J[r, 40:48]])
            Hdd += d @ M @ d
            bd += JI_r @ d
            Hcd += J[r, 20:24] * (M @ d)[0] + J[r, 24:28] * (M @ d)[1]
    assert abs(top["Hdd"][p] - Hdd) <= 1e-5 * abs(Hdd) + 1e-6 and abs(top["bd"][p] - bd) <= 1e-5 * abs(bd) + 1e-3
    np.testing.assert_allclose(top["Hcd"][p], Hcd, rtol=1e-5, atol=1e-3)


def _stitch_top_reference_single(F, acc, adHost, adTarget):
    """AccumulatedTopHessianSSE::stitchDouble (single-thread statement, AccumulatedTopHessian.cpp:171-238)."""
    n = 4 + 8 * F
    H, b = np.zeros((n, n)), np.zeros(n)
    for h in range(F):
        for t in range(F):
            k = h + F * t
            A = acc[k]
            AH, AT = adHost[k], adTarget[k]
            hI, tI = 4 + 8 * h, 4 + 8 * t
            H[hI:hI + 8, hI:hI + 8] += AH @ A[4:12, 4:12] @ AH.T
            H[tI:tI + 8, tI:tI + 8] += AT @ A[4:12, 4:12] @ AT.T
            H[hI:hI + 8, tI:tI + 8] += AH @ A[4:12, 4:12] @ AT.T
            H[hI:hI + 8, 0:4] += AH @ A[4:12, 0:4]
            H[tI:tI + 8, 0:4] += AT @ A[4:12, 0:4]
            H[0:4, 0:4] += A[0:4, 0:4]
            b[hI:hI + 8] += AH @ A[4:12, 12]
            b[tI:tI + 8] += AT @ A[4:12, 12]
            b[0:4] += A[0:4, 12]
    for h in range(F):
        hI = 4 + 8 * h
        H[0:4, hI:hI + 8] = H[hI:hI + 8, 0:4].T
        for t in range(h + 1, F):
            tI = 4 + 8 * t
            H[hI:hI + 8, tI:tI + 8] += H[tI:tI + 8, hI:hI + 8].T
            H[tI:tI + 8, hI:hI + 8] = H[hI:hI + 8, tI:tI + 8].T
    return H, b


def test_top_stitch_matches_single_thread_statement(small):
    pb = small
    F = pb["F"]
    top = O.ba_top_accumulate(0, F, pb["recs"], pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["flags"])
    H, b = O.ba_top_stitch(F, top["acc"], SB.col_major(pb["adHost"]), SB.col_major(pb["adTarget"]))
    Hr, br = _stitch_top_reference_single(F, top["acc"], pb["adHost"], pb["adTarget"])
    np.testing.assert_allclose(H, Hr, rtol=1e-12, atol=1e-9 * np.abs(Hr).max())
    np.testing.assert_allclose(b, br, rtol=1e-12, atol=1e-9 * np.abs(br).max())
    np.testing.assert_allclose(H, H.T, atol=1e-9 * np.abs(H).max())
    Hp, bp = O.ba_top_stitch(F, top["acc"], SB.col_major(pb["adHost"]), SB.col_major(pb["adTarget"]), True, pb["cPrior"], pb["cDeltaF"],
                             pb["frame_prior"], pb["frame_delta_prior"])
    np.testing.assert_allclose(np.diag(Hp)[:4] - np.diag(H)[:4], pb["cPrior"], rtol=1e-6)
    np.testing.assert_allclose(np.diag(Hp)[4:] - np.diag(H)[4:], pb["frame_prior"].ravel(), rtol=1e-6, atol=1e-3)
    np.testing.assert_allclose(bp[4:] - b[4:], (pb["frame_prior"] * pb["frame_delta_prior"]).ravel(), atol=1e-9 * np.abs(b).max() + 1e-9)


def test_sc_accumulator_and_stitch(small):
    pb = small
    F, P = pb["F"], pb["P"]
    A = O.ba_top_accumulate(0, F, pb["recs"], pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["flags"])
    rtz = O.ba_fix_linearization(F, pb["recs"], pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["deltaF"], pb["adHTdeltaF"], pb["cDeltaF"])
    L = O.ba_top_accumulate(1, F, pb["recs"], pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["flags"], rtz, pb["deltaF"], pb["adHTdeltaF"], pb["cDeltaF"])
    jp = O.ba_jpjd(pb["recs"])
    sc = O.ba_sc_accumulate(F, pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["flags"], jp, A["Hdd"], L["Hdd"], A["bd"], L["bd"],
                            A["Hcd"], L["Hcd"], pb["priorF"], pb["deltaF"])
    # direct restatement of AccumulatedSCHessian.cpp:34-77 for a handful of accumulators
    accD = np.zeros_like(sc["accD"]); accE = np.zeros_like(sc["accE"]); Hcc = np.zeros((4, 4))
    for p in range(P):
        rs = [r for r in range(pb["res_begin"][p], pb["res_begin"][p + 1]) if pb["flags"][r] & 1]
        if not rs:
            assert sc["HdiF"][p] == 0
            continue
        Hh = max(np.float32(A["Hdd"][p] + L["Hdd"][p] + pb["priorF"][p]), np.float32(1e-10))
        w = float(np.float32(1.0 / float(Hh)))
        Hcd = (A["Hcd"][p] + L["Hcd"][p]).astype(np.float64)
        Hcc += w * np.outer(Hcd, Hcd)
        for r1 in rs:
            k1 = pb["host_idx"][r1] + F * pb["target_idx"][r1]
            accE[k1] += w * np.outer(jp[r1].astype(np.float64), Hcd)
            for r2 in rs:
                accD[k1 + F * F * pb["target_idx"][r2]] += w * np.outer(jp[r1].astype(np.float64), jp[r2].astype(np.float64))
    np.testing.assert_allclose(sc["accD"], accD, rtol=1e-10, atol=1e-12 * np.abs(accD).max())
    np.testing.assert_allclose(sc["accE"], accE, rtol=1e-10, atol=1e-12 * np.abs(accE).max())
    np.testing.assert_allclose(sc["accHcc"], Hcc, rtol=1e-10)
    # stitchDouble (single) vs stitchDoubleInternal (MT) statements agree: AccumulatedSCHessian.cpp:159-219 vs :78-157
    H, b = O.ba_sc_stitch(F, sc, SB.col_major(pb["adHost"]), SB.col_major(pb["adTarget"]))
    n = 4 + 8 * F
    Hr, br = np.zeros((n, n)), np.zeros(n)
    for i in range(F):
        for j in range(F):
            ij = i + F * j
            iI, jI = 4 + 8 * i, 4 + 8 * j
            Hr[iI:iI + 8, 0:4] += pb["adHost"][ij] @ sc["accE"][ij]
            Hr[jI:jI + 8, 0:4] += pb["adTarget"][ij] @ sc["accE"][ij]
            br[iI:iI + 8] += pb["adHost"][ij] @ sc["accEB"][ij]
            br[jI:jI + 8] += pb["adTarget"][ij] @ sc["accEB"][ij]
            for k in range(F):
                kI, ik = 4 + 8 * k, i + F * k
                Dm = sc["accD"][ij + k * F * F]
                Hr[iI:iI + 8, iI:iI + 8] += pb["adHost"][ij] @ Dm @ pb["adHost"][ik].T
                Hr[jI:jI + 8, kI:kI + 8] += pb["adTarget"][ij] @ Dm @ pb["adTarget"][ik].T
                Hr[jI:jI + 8, iI:iI + 8] += pb["adTarget"][ij] @ Dm @ pb["adHost"][ik].T
                Hr[iI:iI + 8, kI:kI + 8] += pb["adHost"][ij] @ Dm @ pb["adTarget"][ik].T
    Hr[0:4, 0:4] = sc["accHcc"]; br[0:4] = sc["accbc"]
    for h in range(F):
        Hr[0:4, 4 + 8 * h:12 + 8 * h] = Hr[4 + 8 * h:12 + 8 * h, 0:4].T
    np.testing.assert_allclose(H, Hr, rtol=1e-11, atol=1e-10 * np.abs(Hr).max())
    np.testing.assert_allclose(b, br, rtol=1e-11, atol=1e-10 * np.abs(br).max())


def _linearize_inputs(pb):
    return (pb["F"], pb["H"], pb["W"], pb["dI"], pb["precalc"], pb["calib"], pb["pu"], pb["pv"], pb["idepth"], pb["idepth"],
            pb["color"], pb["weights"], pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["frame_energy_th"])


def test_linearize_two_independent_restatements_agree_bit_for_bit():
    """PointFrameResidual::linearize (Residuals.cpp:69-265): the C++ oracle and the numpy restatement that
    generates the synthetic records were written separately from the reference; in float32 without FMA
    contraction they must produce the same bits (records, ResState)."""
    pb = SB.make_ba_problem(F=4, points_per_frame=300, H=120, W=160, seed=3)
    recs, state, energy = O.ba_linearize(*_linearize_inputs(pb))
    assert np.array_equal(state, pb["state"])
    assert set(np.unique(state)) >= {0, 1}  # the problem has residuals that stay in and that leave the image
    assert np.array_equal(recs, pb["recs"])
    th = pb["frame_energy_th"][0]
    assert np.all(energy[state == 2] == th) and np.all(energy[state == 1] == 0) and np.all(energy[state == 0] <= th)


def test_linearize_state_machine_and_known_answers():
    """Hand-checkable cases: identity motion on a linear intensity ramp."""
    F, H, W = 2, 32, 40
    dI = np.zeros((F, H, W, 3), np.float32)
    xs = np.arange(W, dtype=np.float32)[None, :].repeat(H, 0)
    for f in range(F):
        dI[f, ..., 0] = 2.0 * xs + 10.0   # I = 2x + 10
        dI[f, ..., 1] = 2.0               # dI/dx
    fx, fy, cx, cy = 30.0, 30.0, 20.0, 16.0
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    pre = np.zeros((F * F, 28), np.float32)
    for k in range(F * F):
        pre[k, 0:9] = np.eye(3).T.reshape(-1); pre[k, 12:21] = (K @ np.eye(3) @ np.linalg.inv(K)).T.reshape(-1)
        pre[k, 24] = 1.0  # affLL = (1, 0), b0 = 0
    pu = np.array([10.0, 3.0, 20.0], np.float32)   # second point: pattern pixel u-2 = 1 < 1.1 -> OOB
    pv = np.array([12.0, 12.0, 12.0], np.float32)
    idp = np.ones(3, np.float32)
    color = np.zeros((3, 8), np.float32)
    pat = np.array([[0, -2], [-1, -1], [1, -1], [-2, 0], [0, 0], [2, 0], [-1, 1], [0, 2]], np.float32)
    for p in range(3):
        color[p] = 2.0 * (pu[p] + pat[:, 0]) + 10.0
    color[2] += 100.0  # third point: |residual| = 100 on every pixel -> Huber, energy above the threshold -> OUTLIER
    weights = np.ones((3, 8), np.float32)
    host = np.zeros(3, np.int32); target = np.ones(3, np.int32); rb = np.arange(4, dtype=np.int32)
    th = np.full(F, 12 * 12 * 8, np.float32)
    recs, state, energy = O.ba_linearize(F, H, W, dI, pre, [fx, fy, cx, cy], pu, pv, idp, idp, color, weights, host, target, rb, th)
    assert list(state) == [0, 1, 2]
    assert np.all(recs[1] == 0) and energy[1] == 0
    w = 0.5 * (np.sqrt(np.float32(2500.0) / np.float32(2504.0)) + 1.0)
    assert np.allclose(recs[0, 0:8], 0.0, atol=1e-4)                 # perfect match: zero residual
    assert np.allclose(recs[0, 32:40], 2.0 * w, rtol=1e-6)           # JIdx[0] = dx * hw
    assert np.allclose(recs[0, 40:48], 0.0)                          # JIdx[1] = dy * hw
    assert np.allclose(recs[0, 56:64], w, rtol=1e-6)                 # JabF[1] = hw
    assert np.isclose(recs[0, 8], fx) and recs[0, 9] == 0           # Jpdxi[0] = (new_idepth fx, 0, ...)
    assert energy[2] == th[0]
    # an input state of OOB is sticky (:73-74)
    _, state2, _ = O.ba_linearize(F, H, W, dI, pre, [fx, fy, cx, cy], pu, pv, idp, idp, color, weights, host, target, rb, th,
                                  state_in=np.array([1, 0, 0], np.uint8))
    assert list(state2) == [1, 1, 2]


def test_resubstitute_and_l_energy_known_answers():
    """resubstituteFPt / calcLEnergyPt (EnergyFunctional.cpp:291-392) on hand-checkable inputs."""
    F, P = 2, 3
    host = np.array([0, 0, 1], np.int32); target = np.array([1, 1, 0], np.int32)
    rb = np.array([0, 2, 3, 3], np.int32)               # point 2 has no residual
    flags = np.array([1, 0, 1], np.uint8)               # the second residual of point 0 is inactive
    eye = np.stack([np.eye(8)] * (F * F))
    x = np.zeros(4 + 8 * F); x[0] = 2.0; x[4 + 1] = 3.0; x[4 + 8 + 1] = 5.0   # calib[0], frame0[1], frame1[1]
    JpJd = np.zeros((3, 8), np.float32); JpJd[:, 1] = [1.0, 100.0, 2.0]
    bd = np.array([10.0, 20.0, 30.0], np.float32)
    HcdA = np.zeros((P, 4), np.float32); HcdA[:, 0] = 1.0
    HcdL = np.zeros((P, 4), np.float32); HcdL[:, 0] = 0.5
    Hdi = np.array([0.5, 0.25, 1.0], np.float32)
    step = O.ba_resubstitute(F, x, SB.col_major(eye), SB.col_major(eye), host, target, rb, flags, JpJd, bd, HcdA, HcdL, Hdi)
    # xAd[h,t][1] = x_h[1] + x_t[1] = 8 for both pairs; b = bd - 2*(1.5) - 8*JpJd[.,1]
    assert np.allclose(step, [-(10 - 3 - 8 * 1.0) * 0.5, -(20 - 3 - 8 * 2.0) * 0.25, 0.0])
    # energy: zero deltas leave only the depth prior term; a single linearised residual adds (2 rtz + J d) . J d
    recs = np.zeros((3, 76), np.float32)
    recs[0, 56:64] = 1.0                                 # JabF[1] = 1: J delta = delta_b
    rtz = np.zeros((3, 8), np.float32); rtz[0] = 0.5
    deltaF = np.array([0.1, 0.0, 0.2], np.float32); priorF = np.array([0.0, 7.0, 2500.0], np.float32)
    ad = np.zeros((F * F, 8), np.float32)
    lin = np.array([3, 0, 1], np.uint8)                  # residual 0 active + linearised
    e = O.ba_calc_l_energy(F, recs, host, target, rb, lin, rtz, deltaF, priorF, ad, np.zeros(4, np.float32))
    assert np.isclose(e, 0.2 * 0.2 * 2500.0)
    ad[0 + F * 1, 7] = 0.25                              # delta_b of (host 0, target 1)
    e = O.ba_calc_l_energy(F, recs, host, target, rb, lin, rtz, deltaF, priorF, ad, np.zeros(4, np.float32))
    assert np.isclose(e, 100.0 + 8 * (2 * 0.5 + 0.25) * 0.25)


def test_oracle_against_committed_ba_fixture():
    """tests/golden/ba_small.npz (generate_golden.py): freezes the oracle's outputs for the whole BA chain."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ba_small.npz"))
    F, ppf, H, W, seed = [int(v) for v in g["kw"]]
    pb = SB.make_ba_problem(F=F, points_per_frame=ppf, H=H, W=W, seed=seed)
    recs, state, energy = O.ba_linearize(*_linearize_inputs(pb))
    assert np.array_equal(recs, g["recs"]) and np.array_equal(state, g["state"]) and np.array_equal(energy, g["energy"])
    rtz = O.ba_fix_linearization(F, recs, pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["deltaF"], pb["adHTdeltaF"], pb["cDeltaF"])
    assert np.array_equal(rtz, g["res_toZero"])
    top0 = O.ba_top_accumulate(0, F, recs, pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["flags"], rtz, pb["deltaF"], pb["adHTdeltaF"],
                               pb["cDeltaF"], threads=1)
    assert np.array_equal(top0["acc"], g["acc0"]) and np.array_equal(top0["Hdd"], g["Hdd0"])
    e = O.ba_calc_l_energy(F, recs, pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["flags"], rtz, pb["deltaF"], pb["priorF"],
                           pb["adHTdeltaF"], pb["cDeltaF"], pb["cPrior"], pb["frame_prior"], pb["frame_delta_prior"])
    assert e == float(g["l_energy"])


def test_solve_system_restatement_properties():
    """solveSystemF (EnergyFunctional.cpp:775-905, default solver mode): the restatement solves the system it states (residual of
    the damped, Schur-reduced equations at rounding level), the marginalisation prior enters as bM + HM delta, and the
    null-space projector is an orthogonal projector that removes exactly the given directions from x."""
    rng = np.random.default_rng(3)
    n = 4 + 8 * 5

    def spd(scale):
        a = rng.normal(size=(n, n))
        return scale * (a @ a.T + n * np.eye(n))

    HA, HL, HM, Hsc = spd(1.0), spd(0.3), spd(0.1), spd(0.02)
    bA, bL, bM, bsc, delta = (rng.normal(size=n) for _ in range(5))
    lam = 1e-5
    x = O.ba_solve_system(HA, bA, HL, bL, Hsc, bsc, lam, HM, bM, delta)
    Hf = HL + HM + HA
    Hf[np.diag_indices(n)] *= 1 + lam
    Hf -= Hsc / (1 + lam)
    bf = bL + (bM + HM @ delta) + bA - bsc
    assert np.abs(Hf @ x - bf).max() < 1e-9 * np.abs(bf).max()
    x0 = O.ba_solve_system(HA, bA, HL, bL, Hsc, bsc, lam)
    H0 = HL + HA
    H0[np.diag_indices(n)] *= 1 + lam
    H0 -= Hsc / (1 + lam)
    assert np.abs(H0 @ x0 - (bL + bA - bsc)).max() < 1e-9 * np.abs(bL + bA - bsc).max()  # no marginalisation prior
    ns = [rng.normal(size=n) for _ in range(7)]
    P = O.ba_nullspace_projector(ns)
    assert np.allclose(P, P.T, atol=1e-14) and np.allclose(P @ P, P, atol=1e-12) and abs(np.trace(P) - 7) < 1e-10
    xo = O.ba_solve_system(HA, bA, HL, bL, Hsc, bsc, lam, HM, bM, delta, projector=P)
    assert max(abs(np.dot(v, xo)) / np.linalg.norm(v) for v in ns) < 1e-12 * np.linalg.norm(x)
    assert np.allclose(xo, x - P @ x, atol=1e-15)
