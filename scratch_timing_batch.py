import sys, time, ctypes as C, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/slam-eds_b200')
import edsgpu
edsgpu.LIB_PATH = edsgpu.LIB_PATH.replace("libedsgpu.so", "libedsgpu_timing.so")
import bench
from edsgpu import synth
ctx = edsgpu.Context(0)
data = bench.make_data(0, 4, 2)
c = synth.CONFIGS["gen3_vga"]; H, W, E = c["H"], c["W"], c["E"]
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64; n_sc, n_win = len(data), len(data[0][1])
kfs = [edsgpu.KeyFrame(ctx, kf, 8) for kf, _ in data]
frames = edsgpu.Frames(ctx, H, W, S)
trs = [edsgpu.Tracker(ctx, num_blocks=8, max_iterations=30) for _ in range(S)]
batch = edsgpu.TrackerBatch(ctx, trs, [kfs[s % n_sc] for s in range(S)], frames, 0)
buf = (C.c_ulonglong * 16)()
for k in range(3):
    for s, t in enumerate(trs):
        x0 = data[s % n_sc][1][(s // n_sc) % n_win]["x_init"]; t.set_state(x0[:3], x0[3:7], x0[7:], 0.05)
    xs = np.concatenate([data[s % n_sc][1][(s // n_sc + k) % n_win]["x"] for s in range(S)])
    ys = np.concatenate([data[s % n_sc][1][(s // n_sc + k) % n_win]["y"] for s in range(S)])
    ps = np.concatenate([data[s % n_sc][1][(s // n_sc + k) % n_win]["pol"] for s in range(S)])
    edsgpu.event_frames_batch(ctx, frames, 0, S, xs, ys, ps, E)
    ctx.synchronize(); ctx.lib.edsgpu_debug_timing(buf, 1)
    t = time.time(); batch.optimize(); ctx.synchronize(); dt = time.time() - t
    ctx.lib.edsgpu_debug_timing(buf, 1)
    n = buf[4]; m = max(1, buf[10])
    print("S=%d wall ms %.3f evals %d | per eval ns: waitA %.0f evaluate %.0f syncB %.0f advance %.0f adv+publish %.0f | sum %.0f take %.0f steploop %.0f plus %.0f" % (
        S, dt*1e3, n, buf[0]/n, buf[1]/n, buf[2]/n, buf[3]/n, buf[5]/n, buf[6]/m, buf[7]/m, buf[8]/m, buf[9]/m))
