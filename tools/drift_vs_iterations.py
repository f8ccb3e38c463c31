import sys, numpy as np
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'slam-eds_b200'))
import edsgpu
from edsgpu import synth
from oracle import oracle as O
ctx = edsgpu.Context(0)
for cfg in ("gen4_hd", "gen3_vga"):
    scene, kf, wins = synth.make_problem(cfg, 3, 2)
    w = wins[0]; H, W = kf["H"], kf["W"]
    ef = edsgpu.EventFrame(ctx, H, W).create(w["x"], w["y"], w["pol"], w["ts"])
    o = O.event_frame(w["x"], w["y"], w["pol"], w["ts"], H, W)
    kfd = edsgpu.KeyFrame(ctx, kf, 8)
    x0 = w["x_init"]
    for it in (3, 6, 10, 15, 20, 30):
        tr = edsgpu.Tracker(ctx, num_blocks=8, max_iterations=it)
        tr.set_state(x0[:3], x0[3:7], x0[7:], 0.05)
        r = tr.optimize(kfd, ef.frames, 0)
        s = O.tracker_solve(kf, o["frame"], x0, num_blocks=8, max_iterations=it, threads=8)
        s32 = None
        print(cfg, it, "angle %.2e t %.2e v %.2e | gpu succ %d unsucc %d cost %.10f rad %.3e | cpu succ %d unsucc %d cost %.10f rad %.3e" % (
            synth.quat_angle(r["qx"], s["x"][3:7]), np.linalg.norm(r["px"] - s["x"][:3]), np.linalg.norm(r["vx"] - s["x"][7:]),
            r["info"]["successful_steps"], r["info"]["unsuccessful_steps"], r["info"]["final_cost"], r["info"]["final_radius"],
            s["info"]["successful_steps"], s["info"]["unsuccessful_steps"], s["info"]["final_cost"], s["info"]["final_radius"]))
        tr.close()
