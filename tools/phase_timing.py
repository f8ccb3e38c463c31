"""Developer tool: per-phase timing of track_lm_kernel (needs a library built with -DEDS_TIMING).

EDSGPU_LIBRARY=$PWD/libedsgpu_timing.so python tools/phase_timing.py [sequences]
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload constants and synthetic data of the benchmark)
import edsgpu  # noqa: E402
from edsgpu import synth  # noqa: E402


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    c = synth.CONFIGS[bench.CONFIG]
    ctx = edsgpu.Context(0)
    data = bench.make_data(0)
    n_sc, n_win = len(data), len(data[0][1])
    kfs = [edsgpu.KeyFrame(ctx, kf, bench.NUM_BLOCKS) for kf, _ in data]
    frames = edsgpu.Frames(ctx, c["H"], c["W"], S)
    trackers = [edsgpu.Tracker(ctx, num_blocks=bench.NUM_BLOCKS, loss_type=edsgpu.LOSS_HUBER, loss_param=bench.TAU0,
                               max_iterations=bench.MAX_ITER, function_tolerance=1e-6, loss_param_method=edsgpu.LOSS_PARAM_MAD)
                for _ in range(S)]
    batch = edsgpu.TrackerBatch(ctx, trackers, [kfs[s % n_sc] for s in range(S)], frames, 0)
    xs = np.concatenate([data[s % n_sc][1][(s // n_sc) % n_win]["x"] for s in range(S)])
    ys = np.concatenate([data[s % n_sc][1][(s // n_sc) % n_win]["y"] for s in range(S)])
    ps = np.concatenate([data[s % n_sc][1][(s // n_sc) % n_win]["pol"] for s in range(S)])
    ctx.check(ctx.lib.edsgpu_event_frame_create_batch(ctx.h, frames.h, 0, S, None, xs.ctypes.data_as(C.c_void_p), ys.ctypes.data_as(C.c_void_p),
                                                      ps.ctypes.data_as(C.c_void_p), c["E"], edsgpu.DRAW_BILINEAR, 1, C.c_float(0.5), None))
    ctx.synchronize()
    out = (C.c_ulonglong * 32)()
    for rep in range(3):
        for s, t in enumerate(trackers):
            x0 = data[s % n_sc][1][(s // n_sc) % n_win]["x_init"]
            t.set_state(x0[:3], x0[3:7], x0[7:], bench.TAU0)
        ctx.lib.edsgpu_debug_timing.restype = None
        ctx.lib.edsgpu_debug_timing(out, C.c_int(1))
        batch.optimize()
        ctx.synchronize()
        ctx.lib.edsgpu_debug_timing(out, C.c_int(0))
    t = np.array(list(out), dtype=np.float64)
    n = max(t[2], 1)
    print("leader warps: %d steps, per step [us]: work %.2f  waiting for results %.2f" % (n, t[0] / n / 1e3, t[1] / n / 1e3))
    m = max(t[5], 1)
    print("last CTA of each cluster, per visit [us]: consumer work %.2f wait %.2f | producer 0 work %.2f wait %.2f  (%d visits)" % (
        t[3] / m / 1e3, t[4] / m / 1e3, t[13] / m / 1e3, t[14] / m / 1e3, m))
    print("  producer 6 work %.2f, last producer work %.2f, consumer 1 work %.2f | consumer 0 wait-full %.0f cycles/batch, producers wait-empty %.0f cycles/batch" % (
        t[21] / m / 1e3, t[20] / m / 1e3, t[22] / m / 1e3, t[16] / max(t[17], 1), t[18] / max(t[19], 1)))
    k = max(t[10], 1)
    print("leader step parts [us]: tree %.2f  decide %.2f  solve %.2f  plus %.2f  publish %.2f (steps %d)" % (
        t[6] / k / 1e3, t[7] / k / 1e3, t[8] / k / 1e3, t[9] / k / 1e3, t[15] / k / 1e3, k))


if __name__ == "__main__":
    main()
