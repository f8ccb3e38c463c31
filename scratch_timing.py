import sys, time, ctypes as C, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/slam-eds_b200')
import edsgpu
edsgpu.LIB_PATH = edsgpu.LIB_PATH.replace("libedsgpu.so", "libedsgpu_timing.so")
from edsgpu import synth
cfgname = sys.argv[1] if len(sys.argv) > 1 else "davis240c"
ctx = edsgpu.Context(0)
scene, kf, wins = synth.make_problem(cfgname, 0, 1)
w = wins[0]
ef = edsgpu.EventFrame(ctx, kf["H"], kf["W"]); ef.create(w["x"], w["y"], w["pol"], w["ts"])
kfd = edsgpu.KeyFrame(ctx, kf, 8)
tr = edsgpu.Tracker(ctx, num_blocks=8)
x0 = w["x_init"]
buf = (C.c_ulonglong * 16)()
for i in range(3):
    tr.set_state(x0[:3], x0[3:7], x0[7:], 0.05)
    ctx.lib.edsgpu_debug_timing(buf, 1)
    t = time.time(); r = tr.optimize(kfd, ef.frames, 0); dt = time.time() - t
    ctx.lib.edsgpu_debug_timing(buf, 1)
    n = buf[4]
    m = max(1, buf[10]); print("  advance split ns: sum %.0f take %.0f steploop %.0f plus %.0f (n=%d)" % (buf[6]/m, buf[7]/m, buf[8]/m, buf[9]/m, m))
    print("wall ms %.3f evals %d | per eval ns: waitA %.0f evaluate %.0f syncB %.0f advance %.0f adv+publish %.0f" % (dt*1e3, n, buf[0]/n, buf[1]/n, buf[2]/n, buf[3]/n, buf[5]/n))
