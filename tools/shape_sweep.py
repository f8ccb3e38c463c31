import os, sys, time, numpy as np
sys.path.insert(0, "slam-eds_b200")
import edsgpu
from edsgpu import synth
ctx = edsgpu.Context(0)
scene, kf, wins = synth.make_problem("gen3_vga", 1, 1)
w = wins[0]; H, W = kf["H"], kf["W"]; E = len(w["x"])
kfd = edsgpu.KeyFrame(ctx, kf, 8)
for S in (24, 48, 70, 100):
    fr = edsgpu.Frames(ctx, H, W, S)
    ev = [np.tile(w[k], S) for k in ("x", "y", "pol")]
    edsgpu.event_frames_batch(ctx, fr, 0, S, *ev, E)
    res = []
    for c, k in ((0, 0), (144, 4), (140, 8), (128, 4), (112, 4), (96, 4), (64, 4)):  # (evaluator CTAs, leader CTAs)
        if c:
            os.environ["EDSGPU_EVAL_CTAS"] = str(c); os.environ["EDSGPU_LEADER_CTAS"] = str(k)
        else:
            os.environ.pop("EDSGPU_EVAL_CTAS", None); os.environ.pop("EDSGPU_LEADER_CTAS", None)
        trs = [edsgpu.Tracker(ctx, num_blocks=8, max_iterations=30) for i in range(S)]
        b = edsgpu.TrackerBatch(ctx, trs, [kfd] * S, fr, 0)
        ts = []
        for rep in range(3):
            for t in trs:
                x0 = w["x_init"]; t.set_state(x0[:3], x0[3:7], x0[7:], 0.05)
            ctx.synchronize(); t0 = time.perf_counter(); b.optimize(); ctx.synchronize(); ts.append(time.perf_counter() - t0)
        res.append(((c, k), b.launch_shape(), 1e3 * min(ts)))
        b.close(); [t.close() for t in trs]
    print("S", S, " ".join("%s%s:%.2f" % (("auto" if not r[0][0] else ""), r[1], r[2]) for r in res))
    fr.close()
