# scratch: first GPU bring-up (not a test)
import sys, time, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/slam-eds_b200')
import edsgpu
from edsgpu import synth
from oracle import oracle as O
cfgname = sys.argv[1] if len(sys.argv) > 1 else "davis240c"
ctx = edsgpu.Context(0)
print(edsgpu.load().edsgpu_version())
scene, kf, wins = synth.make_problem(cfgname, 0, 2)
w = wins[0]
H, W = kf["H"], kf["W"]
# event frame nn-int
ef = edsgpu.EventFrame(ctx, H, W)
ef.create(w["x"], w["y"], w["pol"], w["ts"], mode=edsgpu.DRAW_NN, use_exp_weights=False, sigma=0.0)
acc = ef.frames.read_accumulator(0)
o = O.event_frame(w["x"], w["y"], w["pol"], w["ts"], H, W, method="nn", use_exp=False, sigma=0.0)
print("nn-int bitexact:", np.array_equal(acc, (o["img"] * 2**40).astype(np.int64)), "norm", ef.norm, o["norm"])
ef.create(w["x"], w["y"], w["pol"], w["ts"], want_host_frame=True)
o = O.event_frame(w["x"], w["y"], w["pol"], w["ts"], H, W)
img, nrm = ef.frames.read(0)
print("bilinear img maxabs diff", np.abs(img - o["img"]).max(), "norm rel", abs(nrm - o["norm"]) / o["norm"], "frame diff", np.abs(ef.event_frame - o["frame"]).max(), ef.time, ef.delta_time)
mx, my = synth.radtan_lut(H, W, kf["fx"], kf["fy"], kf["cx"], kf["cy"])
ef2 = edsgpu.EventFrame(ctx, H, W, mx, my)
ef2.create(w["x"], w["y"], w["pol"], w["ts"])
o2 = O.event_frame(w["x"], w["y"], w["pol"], w["ts"], H, W, mx, my)
img2, nrm2 = ef2.frames.read(0)
print("LUT img maxabs diff", np.abs(img2 - o2["img"]).max(), "norm rel", abs(nrm2 - o2["norm"]) / o2["norm"])
# tracker evaluate
B = 8
kfd = edsgpu.KeyFrame(ctx, kf, B)
x0 = w["x_init"]
g = edsgpu.tracker_evaluate(ctx, kfd, ef.frames, 0, x0)
oe = O.tracker_evaluate(kf, o["frame"], x0, B)
print("res rel", np.abs(g["residuals"] - oe["residuals"]).max() / np.abs(oe["residuals"]).max())
cm = np.abs(oe["jacobian"]).max(0)
print("jac col rel", (np.abs(g["jacobian"] - oe["jacobian"]).max(0) / cm))
print("cost", g["cost"], oe["cost"], "H rel", np.abs(g["H"] - oe["H"]).max() / np.abs(oe["H"]).max(), "g rel", np.abs(g["g"] - oe["g"]).max() / np.abs(oe["g"]).max())
# solve
tr = edsgpu.Tracker(ctx, num_blocks=B)
tr.set_state(x0[:3], x0[3:7], x0[7:], 0.05)
t = time.time(); r = tr.optimize(kfd, ef.frames, 0, want_residuals=True); dt = time.time() - t
so = O.tracker_solve(kf, o["frame"], x0, num_blocks=B, threads=8)
print("gpu", r["info"], "time", dt)
print("cpu", so["info"])
print("x gpu", r["x"]); print("x cpu", so["x"])
print("angle diff", synth.quat_angle(r["x"][3:7], so["x"][3:7]), "t diff", np.linalg.norm(r["x"][:3] - so["x"][:3]), "v diff", np.linalg.norm(r["x"][7:] - so["x"][7:]))
print("tau", r["next_loss_param"], so["next_loss_param"], "res diff", np.abs(r["residuals"] - so["residuals"]).max())
for i in range(3):
    tr.set_state(x0[:3], x0[3:7], x0[7:], 0.05)
    ctx.synchronize(); t = time.time(); r = tr.optimize(kfd, ef.frames, 0); dt = time.time() - t
    print("optimize wall ms", dt * 1e3)
print("launches", ctx.launches)
