"""Generates tests/golden/*.npz from the CPU oracle (run here, commit the outputs).

The reference ships no golden vectors and cannot be built or imported
(C++ with absent dependencies), so these fixtures pin the ORACLE's outputs on seeded inputs:
they catch regressions of the oracle itself (-m "not gpu") and give the GPU tests a
committed target that does not depend on rebuilding the oracle (-m gpu).

    python tests/golden/generate_golden.py
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


if __name__ == "__main__":
    ba_fixture("ba_small.npz")
    coarse_fixture("coarse_small.npz")
    tracking_fixture("tiny", "tracking_tiny.npz", 4, 20)
    tracking_fixture("davis240c", "tracking_davis240c.npz", 8, 30, with_lut=False)
