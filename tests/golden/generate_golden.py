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
from edsgpu import synth  # noqa: E402

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


if __name__ == "__main__":
    tracking_fixture("tiny", "tracking_tiny.npz", 4, 20)
    tracking_fixture("davis240c", "tracking_davis240c.npz", 8, 30, with_lut=False)
