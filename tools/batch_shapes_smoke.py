import sys, time, numpy as np
sys.path.insert(0, "slam-eds_b200")
import edsgpu
from edsgpu import synth
ctx = edsgpu.Context(0)
for name, S, B in (("gen4_hd", 24, 8), ("gen4_hd", 70, 16), ("davis240c", 200, 5), ("gen3_vga", 33, 3)):
    scene, kf, wins = synth.make_problem(name, 1, 1)
    w = wins[0]
    H, W = kf["H"], kf["W"]
    fr = edsgpu.Frames(ctx, H, W, S)
    E = len(w["x"])
    ev = [np.tile(w[k], S) for k in ("x", "y", "pol")]
    edsgpu.event_frames_batch(ctx, fr, 0, S, *ev, E)
    kfd = edsgpu.KeyFrame(ctx, kf, B)
    trs = [edsgpu.Tracker(ctx, num_blocks=B, max_iterations=10 + (i % 4)) for i in range(S)]
    for t in trs:
        x0 = w["x_init"]; t.set_state(x0[:3], x0[3:7], x0[7:], 0.05)
    b = edsgpu.TrackerBatch(ctx, trs, [kfd] * S, fr, 0)
    t0 = time.perf_counter(); b.optimize(); ctx.synchronize(); dt = time.perf_counter() - t0
    st, inf = b.gather()
    same = all(np.array_equal(st[i], st[i % 4]) for i in range(S))
    print(name, "S", S, "B", B, "shape", b.launch_shape(), "ms %.2f" % (1e3 * dt), "usable", sum(i["usable"] for i in inf), "consistent", same)
    b.close(); [t.close() for t in trs]; kfd.close(); fr.close()
