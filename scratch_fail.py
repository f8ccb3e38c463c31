import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/slam-eds_b200')
import bench, edsgpu
from edsgpu import synth
from oracle import oracle as O
ctx = edsgpu.Context(0)
data = bench.make_data(0)
c = synth.CONFIGS["gen3_vga"]; H, W, E = c["H"], c["W"], c["E"]
S = 64; n_sc, n_win = len(data), len(data[0][1])
kfs = [edsgpu.KeyFrame(ctx, kf, 8) for kf, _ in data]
frames = edsgpu.Frames(ctx, H, W, S)
trs = [edsgpu.Tracker(ctx, num_blocks=8, max_iterations=30) for _ in range(S)]
for s, t in enumerate(trs):
    x0 = data[s % n_sc][1][(s // n_sc) % n_win]["x_init"]; t.set_state(x0[:3], x0[3:7], x0[7:], 0.05)
batch = edsgpu.TrackerBatch(ctx, trs, [kfs[s % n_sc] for s in range(S)], frames, 0)
prev, _ = batch.gather(False)
for k in range(30):
    xs = np.concatenate([data[s % n_sc][1][(s // n_sc + k) % n_win]["x"] for s in range(S)])
    ys = np.concatenate([data[s % n_sc][1][(s // n_sc + k) % n_win]["y"] for s in range(S)])
    ps = np.concatenate([data[s % n_sc][1][(s // n_sc + k) % n_win]["pol"] for s in range(S)])
    edsgpu.event_frames_batch(ctx, frames, 0, S, xs, ys, ps, E)
    batch.optimize()
    st, infos = batch.gather()
    bad = [s for s in range(S) if not infos[s]["usable"]]
    print("step", k, "unusable", bad, "mean iters %.1f" % np.mean([i["iterations"] for i in infos]), "mean evals %.1f" % np.mean([i["evaluations"] for i in infos]))
    for s in bad[:2]:
        print("  seq", s, infos[s], "state before", prev[s])
        kf, wins = data[s % n_sc]; w = wins[(s // n_sc + k) % n_win]
        ef = O.event_frame(w["x"], w["y"], w["pol"], w["ts"], H, W)
        r = O.tracker_solve(kf, ef["frame"], prev[s][:13], num_blocks=8, loss_param=prev[s][13], max_iterations=30, want_trace=True)
        print("  oracle:", r["status"], r["info"]); print(r["trace"][:8, :5])
    if bad: break
    prev = st
