import sys, numpy as np
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'slam-eds_b200'))
import edsgpu
from edsgpu import synth, synth_ba
ctx = edsgpu.Context(0)
scene, kf, wins = synth.make_problem("tiny", 0, 2)
H, W = kf["H"], kf["W"]
mx, my = synth.radtan_lut(H, W, kf["fx"], kf["fy"], kf["cx"], kf["cy"], k1=-0.3)
ef = edsgpu.EventFrame(ctx, H, W, mx, my)
for w in wins:
    ef.create(w["x"], w["y"], w["pol"], w["ts"], want_host_frame=True)
    ef.create(w["x"], w["y"], w["pol"], w["ts"], mode=edsgpu.DRAW_NN, use_exp_weights=False, sigma=0.0)
ef = edsgpu.EventFrame(ctx, H, W)
ef.create(wins[0]["x"], wins[0]["y"], wins[0]["pol"], wins[0]["ts"])
for B in (4, 3):
    kfd = edsgpu.KeyFrame(ctx, kf, B)
    g = edsgpu.tracker_evaluate(ctx, kfd, ef.frames, 0, wins[0]["x_init"])
    tr = edsgpu.Tracker(ctx, num_blocks=B, max_iterations=8)
    x0 = wins[0]["x_init"]; tr.set_state(x0[:3], x0[3:7], x0[7:], 0.05)
    r = tr.optimize(kfd, ef.frames, 0, want_residuals=True)
    print("B", B, r["info"]["iterations"], r["usable"])
n = 5
fr = edsgpu.Frames(ctx, H, W, n)
E = len(wins[0]["x"])
edsgpu.event_frames_batch(ctx, fr, 0, n, np.tile(wins[0]["x"], n), np.tile(wins[0]["y"], n), np.tile(wins[0]["pol"], n), E)
kfd = edsgpu.KeyFrame(ctx, kf, 4)
trs = [edsgpu.Tracker(ctx, num_blocks=4, max_iterations=6) for _ in range(n)]
for t in trs: t.set_state(x0[:3], x0[3:7], x0[7:], 0.05)
b = edsgpu.TrackerBatch(ctx, trs, [kfd] * n, fr, 0); b.optimize(); st, inf = b.gather(); print("batch ok", [i["iterations"] for i in inf])
pb = synth_ba.make_ba_problem(F=4, points_per_frame=120, H=120, W=160)
w = edsgpu.BaWindow(ctx, pb["F"], pb["host_idx"], pb["target_idx"], pb["res_begin"])
w.set_residuals(pb["recs"], pb["flags"], np.zeros((pb["R"], 8), np.float32)); w.set_points(pb["deltaF"], pb["priorF"])
w.set_frames(pb["adHTdeltaF"], pb["cDeltaF"], synth_ba.col_major(pb["adHost"]), synth_ba.col_major(pb["adTarget"]))
for m in (0, 1, 2): w.top_accumulate(m)
w.top_stitch(0); w.sc_accumulate(True); w.sc_stitch(); print("ba ok")
# the feeder and the after-solve steps on the device
w.set_images(pb["dI"])
w.set_linearize_inputs(pb["precalc"], pb["calib"], pb["frame_energy_th"], pb["pu"], pb["pv"], pb["idepth"], pb["idepth"], pb["color"], pb["weights"])
st, en = w.linearize(linearized=(pb["flags"] >> 1) & 1, res_toZero=np.zeros((pb["R"], 8), np.float32))
for m in (0, 1): w.top_accumulate(m)
w.sc_accumulate(True)
w.resubstitute(np.full(4 + 8 * pb["F"], 1e-3)); w.calc_l_energy(); w.fix_linearization(None); print("ba linearize / after-solve ok", np.bincount(st))
# frames built ahead in two banks, K problems in flight per cluster
fr2 = edsgpu.Frames(ctx, H, W, 2 * n)
banks = [edsgpu.TrackerBatch(ctx, trs, [kfd] * n, fr2, k * n) for k in (0, 1)]
ev = (np.tile(wins[0]["x"], n), np.tile(wins[0]["y"], n), np.tile(wins[0]["pol"], n))
edsgpu.event_frames_batch(ctx, fr2, 0, n, *ev, E)
for k in range(3):
    edsgpu.event_frames_batch(ctx, fr2, ((k + 1) & 1) * n, n, *ev, E)
    banks[k & 1].optimize()
ctx.synchronize(); print("pipelined ok")
# coarse tracker evaluation and depth filter
from edsgpu import synth_coarse
cp = synth_coarse.make_coarse_problem(W=96, H=64, levels=2, points=700)
ct = edsgpu.CoarseTracker(ctx, 2)
for lvl, L in enumerate(cp["levels"]):
    ct.set_level(lvl, L["w"], L["h"], L["fx"], L["fy"], L["cx"], L["cy"], L["Ki"]); ct.set_reference(lvl, L["pc_u"], L["pc_v"], L["pc_idepth"], L["pc_color"])
    ct.set_new_frame(lvl, L["dI_new"]); r = ct.calc_res_gs(lvl, cp["R"], cp["t"], cp["affLL"], cp["b0"], cp["cutoffTH"])
print("coarse ok", r["rs"][1])
rng = np.random.default_rng(0); nd = 777
dp = edsgpu.DepthPoints(ctx, nd, 500.0, 500.0, 80.0, 60.0, 0.5, 5.5, inv_depth=np.full(nd, 0.4))
Td = np.eye(4); Td[:3, 3] = [0.2, 0.0, 0.05]
kfc = np.stack([rng.uniform(5, 150, nd), rng.uniform(5, 110, nd)], 1)
dp.update(Td, kfc, kfc + rng.normal(scale=2.0, size=(nd, 2))); dp.update(Td, kfc, rng.normal(scale=2.0, size=(nd, 2)), coords_are_tracks=True); print("depth ok", dp.get()[:2, 0])
