"""Developer tool: bit-level fingerprint of edsgpu_tracker_evaluate (residuals, Jacobian rows, cost, H, g) for the library
EDSGPU_LIBRARY selects -- two builds that must agree bit for bit can be compared across processes."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "slam-eds_b200"))
import edsgpu  # noqa: E402
from edsgpu import synth  # noqa: E402

ctx = edsgpu.Context(0)
for cfg, B in (("gen3_vga", 8), ("davis240c", 5), ("tiny", 16)):
    scene, kf, wins = synth.make_problem(cfg, 1, 1)
    w = wins[0]
    ef = edsgpu.EventFrame(ctx, kf["H"], kf["W"]).create(w["x"], w["y"], w["pol"], w["ts"])
    kfd = edsgpu.KeyFrame(ctx, kf, B)
    g = edsgpu.tracker_evaluate(ctx, kfd, ef.frames, 0, w["x_init"])
    h = lambda a: hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:10]
    print(os.path.basename(edsgpu.LIB_PATH), cfg, "B", B, "res", h(g["residuals"]), "jac", h(g["jacobian"]), "cost", h(np.float64(g["cost"])), "H", h(g["H"]), "g", h(g["g"]))
    if len(sys.argv) > 1:
        np.savez(sys.argv[1] + "_" + cfg + ".npz", res=g["residuals"], jac=g["jacobian"], H=g["H"], g=g["g"], cost=g["cost"])
