"""Developer tool: CUDA-event time of the batched LM solve (track_lm_kernel + MAD) on the bench workload, for whichever
library build EDSGPU_LIBRARY selects, plus a checksum of the final states (two builds that must agree bit for bit can be
compared across processes).  With a -DEDS_TIMING build it also prints the clocks taken inside the kernel.

    EDSGPU_LIBRARY=$PWD/build/libedsgpu_x.so python tools/lm_time.py [sequences] [reps] [config] [max_iter]
"""
import ctypes as C
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload constants and synthetic data of the benchmark)
import edsgpu  # noqa: E402
from edsgpu import synth  # noqa: E402


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    config = sys.argv[3] if len(sys.argv) > 3 else bench.CONFIG
    max_iter = int(sys.argv[4]) if len(sys.argv) > 4 else bench.MAX_ITER
    c = synth.CONFIGS[config]
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = edsgpu.Context(0, stream.cuda_stream)
    n_sc = min(8, S)
    data = []
    for s in range(n_sc):
        scene, kf, wins = synth.make_problem(config, s, 2)
        data.append((kf, wins))
    kfs = [edsgpu.KeyFrame(ctx, kf, bench.NUM_BLOCKS) for kf, _ in data]
    frames = edsgpu.Frames(ctx, c["H"], c["W"], S)
    trackers = [edsgpu.Tracker(ctx, num_blocks=bench.NUM_BLOCKS, loss_type=edsgpu.LOSS_HUBER, loss_param=bench.TAU0,
                               max_iterations=max_iter, function_tolerance=1e-6, loss_param_method=edsgpu.LOSS_PARAM_MAD)
                for _ in range(S)]
    batch = edsgpu.TrackerBatch(ctx, trackers, [kfs[s % n_sc] for s in range(S)], frames, 0)
    win = lambda s: data[s % n_sc][1][(s // n_sc) % 2]
    ev = [np.concatenate([win(s)[k] for s in range(S)]) for k in ("x", "y", "pol")]
    edsgpu.event_frames_batch(ctx, frames, 0, S, *ev, c["E"])
    ctx.synchronize()
    has_timing = hasattr(ctx.lib, "edsgpu_debug_timing")
    out = (C.c_ulonglong * 32)()
    ms = []
    for rep in range(reps + 2):
        for s, t in enumerate(trackers):
            x0 = win(s)["x_init"]
            t.set_state(x0[:3], x0[3:7], x0[7:], bench.TAU0)
        if has_timing:
            ctx.lib.edsgpu_debug_timing.restype = None
            ctx.lib.edsgpu_debug_timing(out, C.c_int(1))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        batch.optimize()
        b.record(stream)
        stream.synchronize()
        if rep >= 2:
            ms.append(a.elapsed_time(b))
    states, infos = batch.gather()
    evals = sum(i["evaluations"] for i in infos)
    digest = hashlib.sha1(np.ascontiguousarray(states).tobytes()).hexdigest()[:12]
    shape = batch.launch_shape()
    print("lm_time %s S=%d %s max_iter=%d: %.4f ms (min %.4f) per launch, %d evaluations, usable %d/%d, shape %s, states sha1 %s" % (
        os.path.basename(edsgpu.LIB_PATH), S, config, max_iter, float(np.mean(ms)), float(np.min(ms)), evals, sum(i["usable"] for i in infos), S,
        shape, digest))
    if "successful_steps" in infos[0]:
        print("  steps: %d successful, %d unsuccessful" % (sum(i["successful_steps"] for i in infos), sum(i["unsuccessful_steps"] for i in infos)))
    if has_timing:
        ctx.lib.edsgpu_debug_timing(out, C.c_int(0))
        t = np.array(list(out), dtype=np.float64)
        n = max(t[2], 1)
        print("  leader warps: %d steps, per step [us]: work %.2f  waiting for results %.2f" % (n, t[0] / n / 1e3, t[1] / n / 1e3))
        m = max(t[5], 1)
        print("  last evaluator CTA, per task [us]: consumer 0 work %.2f wait-for-task %.2f | producer 0 work %.2f wait-for-task %.2f  (%d tasks)" % (
            t[3] / m / 1e3, t[4] / m / 1e3, t[13] / m / 1e3, t[14] / m / 1e3, m))
        print("  consumer 0 wait-full %.0f cycles/batch, producers wait-empty %.0f cycles/batch, control warp fetch %.2f us/task" % (
            t[16] / max(t[17], 1), t[18] / max(t[19], 1), t[20] / max(t[21], 1) / 1e3))
        q = max(t[31], 1)
        print("  producer 0, cycles per batch: loads+geometry %.0f  gathers %.0f  finish %.0f  ring %.0f  (%d batches)" % (t[27] / q, t[28] / q, t[29] / q, t[30] / q, q))
        k = max(t[10], 1)
        print("  leader step parts [us]: tree %.2f  decide %.2f  solve %.2f  plus %.2f  publish %.2f (steps %d)" % (
            t[6] / k / 1e3, t[7] / k / 1e3, t[8] / k / 1e3, t[9] / k / 1e3, t[15] / k / 1e3, k))
        print("  decide parts [us]: accept test %.2f  load rows %.2f  projection %.2f  scaling+stores %.2f  gradient norm %.2f" % tuple(t[22:27] / k / 1e3))


if __name__ == "__main__":
    main()
