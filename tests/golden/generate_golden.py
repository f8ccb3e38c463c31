"""Generates tests/golden/*.npz from the CPU oracle (run here, commit the outputs).

The reference ships no golden vectors and cannot be built or imported
(C++ with absent dependencies), so these fixtures pin the ORACLE's outputs on seeded inputs:
they catch regressions of the oracle itself (-m "not gpu") and give the GPU tests a
committed target that does not depend on rebuilding the oracle (-m gpu).

    python tests/golden/generate_golden.py            # everything
    python tests/golden/generate_golden.py round2     # only round2_small.npz (pyramid, per-level solve, makeCoarseDepthL0, solveSystemF)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "slam-eds_b200"))
from oracle import oracle as O  # noqa: E402
from edsgpu import synth, synth_ba, synth_coarse  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def tracking_fixture(config, name, num_blocks, max_iterations, with_lut=True):
    scene, kf, wins = synth.make_problem(config, 0, 1)
    w = wins[0]
    H, W = kf["H"], kf["W"]
    mx, my = synth.radtan_lut(H, W, kf["fx"], kf["fy"], kf["cx"], kf["cy"])
    nn = O.event_frame(w["x"], w["y"], w["pol"], w["ts"], H, W, method="nn", use_exp=False, sigma=0.0)
    bl = O.event_frame(w["x"], w["y"], w["pol"], w["ts"], H, W)
    lut = O.event_frame(w["x"], w["y"], w["pol"], w["ts"], H, W, mx, my)
    x0 = w["x_init"]
    ev = O.tracker_evaluate(kf, bl["frame"], x0, num_blocks, loss_type=1, loss_param=0.05)
    so = O.tracker_solve(kf, bl["frame"], x0, num_blocks=num_blocks, loss_type=1, loss_param=0.05, max_iterations=max_iterations)
    np.savez_compressed(
        os.path.join(HERE, name), config=config, num_blocks=num_blocks, max_iterations=max_iterations,
        ev_x=w["x"], ev_y=w["y"], ev_pol=w["pol"], ev_ts=w["ts"], lut_x=mx if with_lut else np.zeros(0), lut_y=my if with_lut else np.zeros(0),
        nn_counts=np.rint(nn["img"]).astype(np.int32), bl_img=bl["img"], bl_norm=bl["norm"], bl_time=bl["time"],
        bl_delta=bl["delta"], lut_img=lut["img"] if with_lut else np.zeros(0), lut_norm=lut["norm"],
        kf_grad=kf["grad"], kf_norm_coord=kf["norm_coord"], kf_idp=kf["idp"], kf_weights=kf["weights"],
        kf_intr=np.array([kf["fx"], kf["fy"], kf["cx"], kf["cy"]]), kf_size=np.array([H, W]),
        x0=x0, ev_residuals=ev["residuals"], ev_jacobian=ev["jacobian"], ev_cost=ev["cost"], ev_H=ev["H"], ev_g=ev["g"],
        so_x=so["x"], so_residuals=so["residuals"], so_tau=so["next_loss_param"],
        so_info=np.array([so["info"]["iterations"], so["info"]["successful_steps"], so["info"]["unsuccessful_steps"],
                          so["info"]["termination"]]), so_costs=np.array([so["info"]["initial_cost"], so["info"]["final_cost"]]))
    print(name, "iterations", so["info"]["iterations"], "cost", so["info"]["initial_cost"], "->", so["info"]["final_cost"])


def ba_fixture(name):
    """Windowed BA on a small window: the feeder (linearize), both top accumulators, the Schur complement, the
    stitched systems, back-substitution and linearised energy; inputs are regenerated from the seed by the tests."""
    kw = dict(F=4, points_per_frame=120, H=96, W=128, seed=21)
    pb = synth_ba.make_ba_problem(**kw)
    F = pb["F"]
    recs, state, energy = O.ba_linearize(F, pb["H"], pb["W"], pb["dI"], pb["precalc"], pb["calib"], pb["pu"], pb["pv"], pb["idepth"], pb["idepth"],
                                         pb["color"], pb["weights"], pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["frame_energy_th"])
    rtz = O.ba_fix_linearization(F, recs, pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["deltaF"], pb["adHTdeltaF"], pb["cDeltaF"])
    top = [O.ba_top_accumulate(m, F, recs, pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["flags"], rtz, pb["deltaF"], pb["adHTdeltaF"],
                               pb["cDeltaF"], threads=1) for m in (0, 1)]
    jp = O.ba_jpjd(recs)
    sc = O.ba_sc_accumulate(F, pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["flags"], jp, top[0]["Hdd"], top[1]["Hdd"], top[0]["bd"],
                            top[1]["bd"], top[0]["Hcd"], top[1]["Hcd"], pb["priorF"], pb["deltaF"], True)
    x = np.random.default_rng(5).normal(scale=1e-3, size=4 + 8 * F)
    step = O.ba_resubstitute(F, x, synth_ba.col_major(pb["adHost"]), synth_ba.col_major(pb["adTarget"]), pb["host_idx"], pb["target_idx"],
                             pb["res_begin"], pb["flags"], jp, sc["bdSum"], top[0]["Hcd"], top[1]["Hcd"], sc["HdiF"])
    e = O.ba_calc_l_energy(F, recs, pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["flags"], rtz, pb["deltaF"], pb["priorF"],
                           pb["adHTdeltaF"], pb["cDeltaF"], pb["cPrior"], pb["frame_prior"], pb["frame_delta_prior"])
    np.savez_compressed(os.path.join(HERE, name), kw=np.array([kw["F"], kw["points_per_frame"], kw["H"], kw["W"], kw["seed"]]), recs=recs,
                        state=state, energy=energy, res_toZero=rtz, acc0=top[0]["acc"], acc1=top[1]["acc"], Hdd0=top[0]["Hdd"], bd0=top[0]["bd"],
                        Hcd0=top[0]["Hcd"], JpJdF=jp, accD=sc["accD"], accE=sc["accE"], accEB=sc["accEB"], HdiF=sc["HdiF"], bdSum=sc["bdSum"],
                        x=x, step=step, l_energy=e)
    print(name, "R", pb["R"], "states", np.bincount(state), "l_energy", e)


def coarse_fixture(name):
    kw = dict(W=128, H=96, levels=3, points=2000, seed=13)
    pb = synth_coarse.make_coarse_problem(**kw)
    out = {"kw": np.array([kw["W"], kw["H"], kw["levels"], kw["points"], kw["seed"]])}
    for lvl, L in enumerate(pb["levels"]):
        r = O.coarse_calc_res_gs(lvl, L["dI_new"], L["fx"], L["fy"], L["cx"], L["cy"], L["Ki"], pb["R"], pb["t"], pb["affLL"], pb["b0"],
                                 pb["cutoffTH"], L["pc_u"], L["pc_v"], L["pc_idepth"], L["pc_color"])
        out["rs%d" % lvl], out["H%d" % lvl], out["b%d" % lvl], out["counts%d" % lvl] = r["rs"], r["H"], r["b"], r["counts"]
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, [out["counts%d" % l].tolist() for l in range(kw["levels"])])


def round2_fixture(name):
    """The paths added in round 2, all from the oracle: the event-frame pyramid (levels 1-2 of the tiny window), the
    coarse-to-fine per-level solve, makeCoarseDepthL0 on a small pyramid, and solveSystemF on the small BA window
    (the oracle's own accumulators and stitches feed the oracle's solve)."""
    out = {}
    scene, kf, wins = synth.make_problem("tiny", 0, 1)
    w = wins[0]
    o = O.event_frame(w["x"], w["y"], w["pol"], w["ts"], kf["H"], kf["W"])
    frames, norms = O.event_frame_levels(o["img"], 3)
    out["pyr_norms"] = np.array(norms)
    out["pyr_level1"], out["pyr_level2"] = frames[1], frames[2]
    caps = [6, 4, 3]
    x, tau = w["x_init"].copy(), 0.05
    xs = []
    for lvl in (2, 1, 0):
        so = O.tracker_solve(kf, frames[lvl] / norms[lvl], x, num_blocks=4, loss_param=tau, max_iterations=caps[lvl])
        x, tau = so["x"], so["next_loss_param"]
        xs.append(np.concatenate([x, [tau, so["info"]["iterations"], so["info"]["final_cost"]]]))
    out["pyr_caps"], out["pyr_solves"] = np.array(caps), np.array(xs)
    # makeCoarseDepthL0
    kw = dict(W=128, H=96, levels=3, points=2000, seed=13)
    pb = synth_coarse.make_coarse_problem(**kw)
    rng = np.random.default_rng(29)
    n = 1500
    cu = rng.uniform(3.0, kw["W"] - 4.0, n).astype(np.float32)
    cv = rng.uniform(3.0, kw["H"] - 4.0, n).astype(np.float32)
    cid = rng.uniform(0.2, 1.5, n).astype(np.float32)
    hdi = rng.uniform(1e-4, 1e-2, n).astype(np.float32)
    ref = O.make_coarse_depth_l0([dict(w=L["w"], h=L["h"], dI_ref=L["dI_new"]) for L in pb["levels"]], cu, cv, cid, hdi)
    out["cd_kw"] = np.array([kw["W"], kw["H"], kw["levels"], kw["points"], kw["seed"]])
    out["cd_points"] = np.stack([cu, cv, cid, hdi])
    for lvl, r in enumerate(ref):
        out["cd_pc%d" % lvl] = np.stack([r["pc_u"], r["pc_v"], r["pc_idepth"], r["pc_color"]])
    # solveSystemF on the small BA window
    kb = dict(F=4, points_per_frame=120, H=96, W=128, seed=21)
    pbb = synth_ba.make_ba_problem(**kb)
    F = pbb["F"]
    g = np.load(os.path.join(HERE, "ba_small.npz"))
    top = [O.ba_top_accumulate(m, F, g["recs"], pbb["host_idx"], pbb["target_idx"], pbb["res_begin"], pbb["flags"], g["res_toZero"], pbb["deltaF"],
                               pbb["adHTdeltaF"], pbb["cDeltaF"], threads=1) for m in (0, 1)]
    sc = O.ba_sc_accumulate(F, pbb["host_idx"], pbb["target_idx"], pbb["res_begin"], pbb["flags"], g["JpJdF"], top[0]["Hdd"], top[1]["Hdd"],
                            top[0]["bd"], top[1]["bd"], top[0]["Hcd"], top[1]["Hcd"], pbb["priorF"], pbb["deltaF"], True)
    ah, at = synth_ba.col_major(pbb["adHost"]), synth_ba.col_major(pbb["adTarget"])
    HA, bA = O.ba_top_stitch(F, top[0]["acc"], ah, at)
    HL, bL = O.ba_top_stitch(F, top[1]["acc"], ah, at, True, pbb["cPrior"], pbb["cDeltaF"], pbb["frame_prior"], pbb["frame_delta_prior"])
    Hs, bs = O.ba_sc_stitch(F, sc, ah, at)
    nn = 4 + 8 * F
    r2 = np.random.default_rng(31)
    a = r2.normal(size=(nn, nn))
    HM = 1e-2 * np.abs(HA).max() * (a @ a.T) / nn
    bM = 1e-2 * np.abs(bA).max() * r2.normal(size=nn)
    delta = 1e-3 * r2.normal(size=nn)
    P = O.ba_nullspace_projector([r2.normal(size=nn) for _ in range(7)])
    out["solve_HM"], out["solve_bM"], out["solve_delta"], out["solve_P"] = HM, bM, delta, P
    out["solve_x_plain"] = O.ba_solve_system(HA, bA, HL, bL, Hs, bs, 1e-5)
    out["solve_x_full"] = O.ba_solve_system(HA, bA, HL, bL, Hs, bs, 1e-5, HM, bM, delta, P)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "pyramid norms", norms, "coarse depth points", [r["n"] for r in ref], "|x|", np.abs(out["solve_x_plain"]).max())


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "round2":  # the round-1 fixtures stay byte-identical
        round2_fixture("round2_small.npz")
        sys.exit(0)
    ba_fixture("ba_small.npz")
    coarse_fixture("coarse_small.npz")
    tracking_fixture("tiny", "tracking_tiny.npz", 4, 20)
    tracking_fixture("davis240c", "tracking_davis240c.npz", 8, 30, with_lut=False)
    round2_fixture("round2_small.npz")
