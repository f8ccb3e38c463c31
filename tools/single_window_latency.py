"""Developer tool: latency of one tracking window (event frame + LM solve + MAD) per sensor configuration,
one sequence alone on the GPU.  python tools/single_window_latency.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "slam-eds_b200"))
import edsgpu  # noqa: E402
from edsgpu import synth  # noqa: E402


def main():
    ctx = edsgpu.Context(0)
    for name, B, iters in (("davis240c", 8, 30), ("gen3_vga", 8, 30), ("gen4_hd", 8, 30)):
        scene, kf, wins = synth.make_problem(name, 0, 2)
        kfd = edsgpu.KeyFrame(ctx, kf, B)
        ef = edsgpu.EventFrame(ctx, kf["H"], kf["W"])
        tr = edsgpu.Tracker(ctx, num_blocks=B, max_iterations=iters, function_tolerance=1e-6)
        w = wins[0]
        t_ef, t_lm, evals = [], [], 0
        for rep in range(12):
            x0 = w["x_init"]
            tr.set_state(x0[:3], x0[3:7], x0[7:], 0.05)
            ctx.synchronize()
            t0 = time.perf_counter()
            ef.create(w["x"], w["y"], w["pol"], w["ts"])
            ctx.synchronize()
            t1 = time.perf_counter()
            r = tr.optimize(kfd, ef.frames, 0)
            t2 = time.perf_counter()
            if rep >= 2:
                t_ef.append(t1 - t0)
                t_lm.append(t2 - t1)
            evals = r["info"]["evaluations"]
        print("%-10s N=%6d E=%7d  event frame %.3f ms   optimize (LM + MAD, host-synchronous) %.3f ms  (%d evaluations, %.1f us each)" % (
            name, len(kf["idp"]), len(w["x"]), 1e3 * np.median(t_ef), 1e3 * np.median(t_lm), evals, 1e6 * np.median(t_lm) / max(evals, 1)))
        tr.close(); kfd.close()


if __name__ == "__main__":
    main()
