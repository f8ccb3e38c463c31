#!/usr/bin/env python
"""bench.py -- tracking windows/s at 640x480 (BASELINE.json metric) on N B200s of one node.

Workload (config.workload): BASELINE.json configs[4] = 64 independent sequences of configs[1] -- Prophesee Gen3 640x480,
10 240 keyframe points, 50 000 events/window, B = 8 residual blocks, Huber (tau0 = 0.05, MAD-updated), max_num_iterations = 30,
function_tolerance = 1e-6.  One step = one event window of every sequence: event-frame construction +
device-side LM solve + MAD, warm-started from the previous window (px, qx, vx, tau stay in HBM).  Every sequence has its own
scene and key frame and runs through WINDOWS consecutive windows of its own trajectory (then round again).

  --scaling strong (default)  the 64 sequences are dealt round-robin to the N ranks; `weak` (64 per rank) is
                              reported next to it as the `weak` object when N > 1
  value  device-resident inputs (events already in HBM), CUDA-event timed, max over ranks
  e2e    the same step through the host-facing C ABI: events from pinned host memory every step (H2D inside the timed
         region) and the state records read back (D2H)
  config1 / config3 objects   the other two sensor shapes (240x180, 1280x720), same step, own roofline (N = 1 only)
  --impl reference   the CPU restatement of the reference path (oracle, dual-number Jacobians = the cost structure of the
         Ceres autodiff functor) on the box's host cores, warm-started and tau-carried like the GPU arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "slam-eds_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

CONFIG = "gen3_vga"
NUM_BLOCKS = 8
MAX_ITER = 30
TAU0 = 0.05
WINDOWS = 16             # consecutive windows generated per sequence of the headline workload
SIDE_SCENES, SIDE_WINDOWS = 8, 4   # the config1 / config3 side measurements cycle through fewer distinct inputs
RESERVE_SMS = 16         # SMs the batched solve leaves to the event-frame builds of the next window (other stream)
PARITY_NOTE = ("pose gate 1e-4 rad / 1e-4 x depth met at <= 20 LM iterations; at this cap (30) GPU-vs-oracle differences are bounded by "
               "4 x the oracle's own drift under 1-ulp input perturbations (tests/test_oracle_tracking.py, tests/test_gpu_fullsize.py)")


def algorithmic_bytes_lm(N, H, W, B, evaluations):
    """per LM Jacobian evaluation 24 N + min(64 N, 4 H W) + 364 B (+ 4 N write-back per window, added by the caller)."""
    return evaluations * (24 * N + min(64 * N, 4 * H * W) + 364 * B)


def algorithmic_bytes_ef(E, H, W):
    """per window 16 E + 12 H W."""
    return 16 * E + 12 * H * W


def _make_sequence(args):
    from edsgpu import synth
    config, seq, n_windows = args
    scene, kf, wins = synth.make_problem(config, seq, n_windows)
    return kf, wins


def make_sequences(config, seq_ids, n_windows, procs=None):
    """[(keyframe, windows)] for the given sequence ids (seed = 1234 + 1000 config_id + id), generated in parallel."""
    seq_ids = list(seq_ids)
    procs = procs or max(1, min(len(seq_ids), (os.cpu_count() or 1)))
    jobs = [(config, int(s), n_windows) for s in seq_ids]
    if procs == 1 or len(jobs) == 1:
        return [_make_sequence(j) for j in jobs]
    with ProcessPoolExecutor(max_workers=min(procs, len(jobs))) as ex:
        return list(ex.map(_make_sequence, jobs))


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------- CPU arm
def cpu_windows_per_s(data, windows_per_sequence, concurrent, threads_per_window, jacobian_mode=1):
    """Times the oracle on `windows_per_sequence` consecutive windows of every sequence in `data`: event frame + Ceres-style
    solve + MAD, warm-started from the previous window's state and tau exactly like the GPU arm, `concurrent` sequences at a
    time, `threads_per_window` workers each (the reference parallelises over its B residual blocks, Tracker.cpp:138,178-195).
    -> (windows/s, seconds, windows, mean LM iterations)"""
    from oracle import oracle as O
    O.lib()
    iters = []

    def run_sequence(kf, wins):
        x, tau = wins[0]["x_init"].copy(), TAU0
        for k in range(windows_per_sequence):
            w = wins[k % len(wins)]
            ef = O.event_frame(w["x"], w["y"], w["pol"], w["ts"], kf["H"], kf["W"])
            s = O.tracker_solve(kf, ef["frame"], x, num_blocks=NUM_BLOCKS, loss_type=1, loss_param=tau, max_iterations=MAX_ITER,
                                function_tolerance=1e-6, jacobian_mode=jacobian_mode, threads=threads_per_window)
            iters.append(s["info"]["iterations"])
            if s["info"]["usable"]:  # Tracker.cpp:217-233
                x, tau = s["x"], s["next_loss_param"]

    kf0, w0 = data[0]
    O.tracker_solve(kf0, O.event_frame(w0[0]["x"], w0[0]["y"], w0[0]["pol"], w0[0]["ts"], kf0["H"], kf0["W"])["frame"], w0[0]["x_init"],
                    num_blocks=NUM_BLOCKS, max_iterations=2, jacobian_mode=jacobian_mode, threads=threads_per_window)  # warm (page in)
    iters.clear()
    t0 = time.perf_counter()
    if concurrent <= 1:
        for kf, wins in data:
            run_sequence(kf, wins)
    else:
        nxt = [0]
        lock = threading.Lock()

        def worker():
            while True:
                with lock:
                    k = nxt[0]
                    nxt[0] += 1
                if k >= len(data):
                    return
                run_sequence(*data[k])  # ctypes releases the GIL inside the oracle

        ths = [threading.Thread(target=worker) for _ in range(concurrent)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
    dt = time.perf_counter() - t0
    n = len(data) * windows_per_sequence
    return n / dt, dt, n, float(np.mean(iters))


def workload_string(c):
    return "config2: Gen3 640x480, %d points, %d events/window, B=%d, Huber+MAD, max_iter %d" % (c["N"], c["E"], NUM_BLOCKS, MAX_ITER)


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; the reference itself cannot be built here) on all host
    cores; rank 0 only.  A step = `concurrent` sequences advancing by 2 consecutive windows each."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import oracle as O
    flags = O.use_native_build()
    from edsgpu import synth
    cores = os.cpu_count() or 1
    concurrent = max(1, cores // NUM_BLOCKS)
    per_seq = 2
    data = make_sequences(CONFIG, range(concurrent), per_seq)
    c = synth.CONFIGS[CONFIG]
    for _ in range(args.warmup):
        cpu_windows_per_s(data, per_seq, concurrent, NUM_BLOCKS)
    t_total, n_total, it = 0.0, 0, []
    for _ in range(args.steps):
        wps, dt, n, iters = cpu_windows_per_s(data, per_seq, concurrent, NUM_BLOCKS)
        t_total += dt
        n_total += n
        it.append(iters)
    value = n_total / t_total
    _, dt_a, n_a, _ = cpu_windows_per_s(data, per_seq, concurrent, NUM_BLOCKS, jacobian_mode=0)
    sample = "%d windows/step x %d steps of config2 (event frame + dual-number LM solve + MAD, warm-started, tau carried), %d concurrent sequences x %d threads" % (
        concurrent * per_seq, args.steps, concurrent, NUM_BLOCKS)
    print(json.dumps({
        "impl": "reference", "metric": "tracking_windows_per_s_640x480", "value": value, "unit": "windows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(1, args.steps), "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "mevents_per_s": value * c["E"] / 1e6,
        "config": {"workload": workload_string(c),
                   "note": "oracle restatement of the reference CPU/Ceres path (not Ceres itself), dual-number Jacobians; built " + flags,
                   "mean_lm_iterations": float(np.mean(it))},
        "cpu_baseline": {"value": value, "unit": "windows/s", "cores": min(cores, concurrent * NUM_BLOCKS), "kind": "port", "sample": sample,
                         "build": flags, "analytic_jacobian_value": n_a / dt_a},
        "e2e": {"value": value, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------- BA side line
def bench_ba(ctx, stream, reps=50):
    """BASELINE.json configs[3]: F=7, 2048 points/keyframe, R=86016: top accumulate (modes 0 and 1),
    SC accumulate and both stitches with the records resident in HBM; CUDA-event timed."""
    import torch
    import edsgpu
    from edsgpu import synth_ba
    from oracle import oracle as O  # CPU baseline leg of this side line only: the timed CPU port, never on the GPU path
    pb = synth_ba.make_ba_problem()
    F, P, R = pb["F"], pb["P"], pb["R"]
    rtz = np.zeros((R, 8), np.float32)
    w = edsgpu.BaWindow(ctx, F, pb["host_idx"], pb["target_idx"], pb["res_begin"])
    w.set_residuals(pb["recs"], pb["flags"], rtz)
    w.set_points(pb["deltaF"], pb["priorF"])
    w.set_frames(pb["adHTdeltaF"], pb["cDeltaF"], synth_ba.col_major(pb["adHost"]), synth_ba.col_major(pb["adTarget"]))

    def once():
        w.top_accumulate(0, want_outputs=False)
        w.top_accumulate(1, want_outputs=False)
        w.sc_accumulate(True, want_outputs=False)

    for _ in range(5):
        once()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        once()
    b.record(stream)
    stream.synchronize()
    ms = a.elapsed_time(b) / reps
    t0 = time.perf_counter()
    n_cpu = 20
    for _ in range(n_cpu):
        O.ba_top_accumulate(0, F, pb["recs"], pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["flags"], threads=6)
        O.ba_top_accumulate(1, F, pb["recs"], pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["flags"], rtz, pb["deltaF"],
                            pb["adHTdeltaF"], pb["cDeltaF"], threads=6)
    cpu_ms = 1e3 * (time.perf_counter() - t0) / n_cpu
    alg = 2 * (296 * R + 12 * R + 364 * F * F + 24 * P) + (44 * R + 48 * P + 4 * (64 * F ** 3 + 40 * F * F + 20))
    # the feeder on the device: PointFrameResidual::linearize + takeDataF, inputs resident
    w.set_images(pb["dI"])
    w.set_linearize_inputs(pb["precalc"], pb["calib"], pb["frame_energy_th"], pb["pu"], pb["pv"], pb["idepth"], pb["idepth"],
                           pb["color"], pb["weights"])
    for _ in range(5):
        w.linearize(want_outputs=False)
    a.record(stream)
    for _ in range(reps):
        w.linearize(want_outputs=False)
    b.record(stream)
    stream.synchronize()
    lin_ms = a.elapsed_time(b) / reps
    t0 = time.perf_counter()
    for _ in range(5):
        O.ba_linearize(F, pb["H"], pb["W"], pb["dI"], pb["precalc"], pb["calib"], pb["pu"], pb["pv"], pb["idepth"], pb["idepth"],
                       pb["color"], pb["weights"], pb["host_idx"], pb["target_idx"], pb["res_begin"], pb["frame_energy_th"])
    lin_cpu_ms = 1e3 * (time.perf_counter() - t0) / 5
    # one Gauss-Newton iteration's accumulation as the library now runs it: linearize FUSED with addPoint<0> (records of the
    # residuals nothing reads later are never written), then the linearized side and the Schur complement
    linearized = (pb["flags"] >> 1) & 1
    w.linearize_accumulate(linearized=linearized, write_records=True, want_outputs=False)  # the linearized records exist once

    def fused():
        w.ctx.check(w.ctx.lib.edsgpu_ba_linearize_accumulate(w.h, None, None, None, 0, None, None))
        w.top_accumulate(1, want_outputs=False)
        w.sc_accumulate(True, want_outputs=False)

    for _ in range(5):
        fused()
    a.record(stream)
    for _ in range(reps):
        fused()
    b.record(stream)
    stream.synchronize()
    fused_ms = a.elapsed_time(b) / reps
    # per residual: 8 pattern pixels x 4 taps x 16 B gathered, 80 B of point state, 112 B of precalc, 304 B record + 32 B JpJdF
    # written (and the record read once more by takeDataF), 9 B of state / energy / flag
    lin_alg = R * (8 * 4 * 16 + 80 + 112 + 304 + 304 + 32 + 9)
    # after the solve: back-substitution + linearised energy, host-synchronous calls
    once()
    xs = np.random.default_rng(0).normal(scale=1e-3, size=4 + 8 * F)
    for _ in range(3):
        w.resubstitute(xs); w.calc_l_energy()
    t0 = time.perf_counter()
    for _ in range(20):
        w.resubstitute(xs)
    resub_ms = 1e3 * (time.perf_counter() - t0) / 20
    t0 = time.perf_counter()
    for _ in range(20):
        w.calc_l_energy()
    energy_ms = 1e3 * (time.perf_counter() - t0) / 20
    for _ in range(3):
        w.solve_system(1e-5, want_step=False)
    t0 = time.perf_counter()
    for _ in range(20):
        w.solve_system(1e-5, want_step=False)
    solve_ms = 1e3 * (time.perf_counter() - t0) / 20
    w.close()
    return {"workload": "config4: F=7, P=%d, R=%d; top<0> + top<1> + SC accumulate, device-resident" % (P, R),
            "accumulations_per_s": 1e3 / ms, "ms": ms, "algorithmic_bytes": alg, "achieved_gbs": alg / (ms * 1e-3) / 1e9,
            "cpu_port_top_ms_6_threads": cpu_ms,
            "linearize": {"what": "PointFrameResidual::linearize + takeDataF for all R residuals on the device (records stay in HBM)",
                          "ms": lin_ms, "residuals_per_s": R / (lin_ms * 1e-3), "algorithmic_bytes": lin_alg,
                          "achieved_gbs": lin_alg / (lin_ms * 1e-3) / 1e9, "cpu_port_ms_1_thread": lin_cpu_ms,
                          "upload_avoided_bytes": 304 * R},
            "gn_iteration": {"what": "linearize fused with addPoint<0> (edsgpu_ba_linearize_accumulate, records not materialised) + top<1> + SC "
                                     "accumulate: everything one Gauss-Newton iteration accumulates, 5 launches",
                             "ms": fused_ms, "unfused_ms": lin_ms + ms},
            "after_solve": {"what": "resubstituteF_MT (P steps back to the host) and calcLEnergyF_MT, wall clock per host-synchronous call",
                            "resubstitute_ms": resub_ms, "calc_l_energy_ms": energy_ms,
                            "solve_system_ms": solve_ms,
                            "solve_system": "solveSystemF on the device (three stitches, %d x %d LDL^T, resubstituteF_MT), only x comes back" % (4 + 8 * F, 4 + 8 * F)}}


def bench_coarse(ctx, stream, reps=200):
    """one evaluation of the DSO coarse tracker (calcRes fused with calcGSSSE) per pyramid
    level on a 640x480 frame with 20 000 reference points at level 0; wall clock of the host-synchronous call
    (that is how trackNewestCoarse uses it: every evaluation is followed by an 8x8 solve on the host)."""
    import edsgpu
    from edsgpu import synth_coarse
    from oracle import oracle as O  # CPU baseline leg of this side line only
    pb = synth_coarse.make_coarse_problem()
    ct = edsgpu.CoarseTracker(ctx, len(pb["levels"]))
    for lvl, L in enumerate(pb["levels"]):
        ct.set_level(lvl, L["w"], L["h"], L["fx"], L["fy"], L["cx"], L["cy"], L["Ki"])
        ct.set_reference(lvl, L["pc_u"], L["pc_v"], L["pc_idepth"], L["pc_color"])
        ct.set_new_frame(lvl, L["dI_new"])
    out = []
    for lvl, L in enumerate(pb["levels"]):
        args = (lvl, pb["R"], pb["t"], pb["affLL"], pb["b0"], pb["cutoffTH"])
        for _ in range(10):
            ct.calc_res_gs(*args)
        t0 = time.perf_counter()
        for _ in range(reps):
            ct.calc_res_gs(*args)
        ms = 1e3 * (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        for _ in range(5):
            O.coarse_calc_res_gs(lvl, L["dI_new"], L["fx"], L["fy"], L["cx"], L["cy"], L["Ki"], pb["R"], pb["t"], pb["affLL"], pb["b0"],
                                 pb["cutoffTH"], L["pc_u"], L["pc_v"], L["pc_idepth"], L["pc_color"])
        cpu_ms = 1e3 * (time.perf_counter() - t0) / 5
        n = len(L["pc_u"])
        out.append({"level": lvl, "points": n, "ms_per_evaluation": ms, "points_per_s": n / (ms * 1e-3), "cpu_port_ms_1_thread": cpu_ms,
                    "algorithmic_bytes": n * (16 + 4 * 16)})  # point record + 4 bilinear taps of {I,dx,dy,0}
    # the whole of trackNewestCoarse from a wrong start (coarsest level 3)
    for _ in range(3):
        tr = ct.track(3, pb["R"], pb["t"])
    t0 = time.perf_counter()
    for _ in range(20):
        tr = ct.track(3, pb["R"], pb["t"])
    track_ms = 1e3 * (time.perf_counter() - t0) / 20
    t0 = time.perf_counter()
    for _ in range(3):
        tr_cpu = O.coarse_track(pb, 3, pb["R"], pb["t"])
    track_cpu_ms = 1e3 * (time.perf_counter() - t0) / 3
    ct.close()
    return {"workload": "coarse tracker: calcRes + calcGSSSE per level, 640x480, host-synchronous call", "levels": out,
            "track_newest_coarse": {"ms": track_ms, "evaluations": tr["evaluations"], "ok": tr["ok"], "cpu_port_ms_1_thread": track_cpu_ms,
                                    "cpu_evaluations": tr_cpu["evaluations"]}}


def bench_depth(ctx, stream, n=10240, reps=100):
    """DepthPoints::update for the key frame's points (config 2: 10 240 points), host
    coordinates in, state resident on the device; wall clock of the host-synchronous call."""
    import edsgpu
    from oracle import oracle as O  # CPU baseline leg of this side line only
    rng = np.random.default_rng(0)
    fx = fy = 520.0
    cx, cy = 320.0, 240.0
    kf = np.stack([rng.uniform(20, 620, n), rng.uniform(20, 460, n)], 1)
    depth = rng.uniform(1.0, 6.0, n)
    T = np.eye(4); T[:3, 3] = [0.25, -0.08, 0.05]
    Pk = np.stack([(kf[:, 0] - cx) / fx * depth, (kf[:, 1] - cy) / fy * depth, depth], 1) - T[:3, 3]
    ef = np.stack([fx * Pk[:, 0] / Pk[:, 2] + cx, fy * Pk[:, 1] / Pk[:, 2] + cy], 1)
    dp = edsgpu.DepthPoints(ctx, n, fx, fy, cx, cy, 0.5, 5.5, inv_depth=1.0 / depth)
    for _ in range(5):
        dp.update(T, kf, ef)
    t0 = time.perf_counter()
    for _ in range(reps):
        dp.update(T, kf, ef)
    ms = 1e3 * (time.perf_counter() - t0) / reps
    st = dp.get()
    t0 = time.perf_counter()
    for _ in range(10):
        O.depth_update(fx, fy, cx, cy, 5.0, np.arctan(3.0 / (2 * fx)) + np.arctan(3.0 / (2 * fy)), T, kf, ef, st)
    cpu_ms = 1e3 * (time.perf_counter() - t0) / 10
    dp.close()
    return {"workload": "depth filter: DepthPoints::update, %d points, coordinates from the host, state on the device" % n,
            "ms_per_update": ms, "points_per_s": n / (ms * 1e-3), "cpu_port_ms_1_thread": cpu_ms, "h2d_bytes": 32 * n}


# --------------------------------------------------------------------------------------- GPU arm
class TrackRun:
    """`S` sequences of one sensor configuration on this rank: device objects, event buffers (pinned host + HBM-resident
    copies) and the two timed loops.  Sequence i follows data[i % len(data)] (the headline workload has one entry per sequence)."""

    def __init__(self, ctx, torch, dev, stream, config, data, S):
        import edsgpu
        from edsgpu import synth
        self.edsgpu, self.torch, self.ctx, self.dev, self.stream = edsgpu, torch, ctx, dev, stream
        self.c = c = synth.CONFIGS[config]
        self.S, self.data, self.n_win = S, data, len(data[0][1])
        self.kfs = [edsgpu.KeyFrame(ctx, kf, NUM_BLOCKS) for kf, _ in data]
        # two banks of S frame slots: the library builds event frames on its own stream, so the frames of window
        # k+1 (bank (k+1)&1) are built while window k (bank k&1) is being solved
        self.frames = edsgpu.Frames(ctx, c["H"], c["W"], 2 * S)
        self.build_stream = torch.cuda.ExternalStream(self.frames.build_stream(), device=dev)
        self.trackers = [edsgpu.Tracker(ctx, num_blocks=NUM_BLOCKS, loss_type=edsgpu.LOSS_HUBER, loss_param=TAU0, max_iterations=MAX_ITER,
                                        function_tolerance=1e-6, loss_param_method=edsgpu.LOSS_PARAM_MAD) for _ in range(S)]
        self.banks = [edsgpu.TrackerBatch(ctx, self.trackers, [self.kfs[s % len(data)] for s in range(S)], self.frames, bank * S)
                      for bank in (0, 1)]
        # event windows of step k (window k % n_win of every sequence): pinned host copies and device-resident copies
        self.host_ev, self.dev_ev = [], []
        for k in range(self.n_win):
            arrs = []
            for key in ("x", "y", "pol"):
                a = np.concatenate([data[s % len(data)][1][k][key] for s in range(S)])
                t = torch.from_numpy(a.view(np.int16).copy() if a.dtype == np.uint16 else a.copy()).pin_memory()
                arrs.append(t)
            self.host_ev.append(tuple(arrs))
            self.dev_ev.append(tuple(t.to(dev) for t in arrs))
        self.states_dev = torch.zeros(S, 14, dtype=torch.float64, device=dev)
        self.states_host = [torch.zeros(S, 14, dtype=torch.float64).pin_memory() for _ in range(2)]
        self.done_evt = [torch.cuda.Event(), torch.cuda.Event()]

    def reset_states(self):
        for s, t in enumerate(self.trackers):
            x0 = self.data[s % len(self.data)][1][0]["x_init"]
            t.set_state(x0[:3], x0[3:7], x0[7:], TAU0)

    def _create_device(self, k, events=None):
        dx, dy, dp = self.dev_ev[k % self.n_win]
        if events is not None:
            a, b = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
            a.record(self.build_stream)
        self.edsgpu.event_frames_batch_dev(self.ctx, self.frames, (k & 1) * self.S, self.S, dx.data_ptr(), dy.data_ptr(), dp.data_ptr(), self.c["E"])
        if events is not None:
            b.record(self.build_stream)
            events.append((a, b))

    def run_device(self, first, n, lm_events=None, ef_events=None):
        """n steps, inputs resident in HBM; one step = event frames of window k + batched LM solve + MAD.
        Software-pipelined like the streaming front end: window k+1's frames are queued before window k's solve."""
        self._create_device(first, ef_events)
        for i in range(n):
            k = first + i
            if i + 1 < n:
                self._create_device(k + 1, ef_events)
            if lm_events is not None:
                a, b = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
                a.record(self.stream)
            self.ctx.check(self.ctx.lib.edsgpu_batch_optimize(self.banks[k & 1].h))  # track_lm_kernel + mad_kernel
            if lm_events is not None:
                b.record(self.stream)
                lm_events.append((a, b))

    def time_event_frames(self, n):
        """Mean CUDA-event time of one frame build (S windows: accumulator clear + scatter + blur/norm) with nothing else
        on the device: the kernel time behind the event_frame roofline."""
        self.ctx.synchronize()
        evs = []
        for k in range(n + 2):
            self._create_device(k, evs)
        self.ctx.synchronize()
        return float(np.mean([a.elapsed_time(b) for a, b in evs[2:]]))

    def time_lm_alone(self, n):
        """The batched solve (track_lm_kernel + mad_kernel) with nothing else on the device: every launch restarts window 0 from
        its initial state, frames already built.  Returns (mean launch ms, mean evaluations per launch)."""
        self._create_device(0)
        ms, evals = [], []
        for r in range(n + 1):
            self.reset_states()
            self.ctx.synchronize()
            a, b = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
            a.record(self.stream)
            self.ctx.check(self.ctx.lib.edsgpu_batch_optimize(self.banks[0].h))
            b.record(self.stream)
            self.ctx.synchronize()
            if r:  # the first one warms up
                ms.append(a.elapsed_time(b))
                evals.append(float(sum(i["evaluations"] for i in self.banks[0].gather()[1])))
        return float(np.mean(ms)), float(np.mean(evals))

    def _issue_create(self, k):
        # host-facing C ABI with HOST (pinned) buffers: asynchronous, H2D on the library's copy stream
        C = self.edsgpu.C
        hx, hy, hp = self.host_ev[k % self.n_win]
        self.ctx.check(self.ctx.lib.edsgpu_event_frame_create_batch(
            self.ctx.h, self.frames.h, (k & 1) * self.S, self.S, None, C.c_void_p(hx.data_ptr()), C.c_void_p(hy.data_ptr()),
            C.c_void_p(hp.data_ptr()), self.c["E"], self.edsgpu.DRAW_BILINEAR, 1, C.c_float(0.5), None))

    def run_e2e(self, first, n):
        """n steps; every step copies its events H2D from pinned host memory and reads its S x 14 state records back.
        Software-pipelined like a streaming front end: the next window's events are handed to the library while the
        current window is being solved (their H2D copy and frame build overlap the solve), and the host picks up step k's
        states while step k+1 is already queued (the poses of a window are consumed one window later; every step's
        read-back completes inside the timed region)."""
        self._issue_create(first)
        for i in range(n):
            k = first + i
            if i + 1 < n:
                self._issue_create(k + 1)
            self.banks[k & 1].optimize()
            self.banks[k & 1].pack_states_dev(self.states_dev.data_ptr())
            self.states_host[i & 1].copy_(self.states_dev, non_blocking=True)
            self.done_evt[i & 1].record(self.stream)
            if i > 0:
                self.done_evt[(i - 1) & 1].synchronize()  # step k-1's result is on the host
        self.done_evt[(n - 1) & 1].synchronize()

    def close(self):
        self.ctx.synchronize()
        for b in self.banks:
            b.close()
        for t in self.trackers:
            t.close()
        self.frames.close()
        for k in self.kfs:
            k.close()


def measure(run, steps, warmup, barrier, max_over_ranks, peak, traffic):
    """value / e2e / roofline of one TrackRun (all ranks call this together)."""
    torch = run.torch
    c, S = run.c, run.S
    run.reset_states()
    run.run_device(0, warmup)
    barrier()
    launches0 = run.ctx.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lm_events = []
    ev0.record(run.stream)
    run.run_device(warmup, steps, lm_events)
    ev1.record(run.stream)
    barrier()
    launches = run.ctx.launches - launches0
    t_ms = max_over_ranks(ev0.elapsed_time(ev1))
    lm_ms = float(np.mean([a.elapsed_time(b) for a, b in lm_events]))
    ef_ms = run.time_event_frames(min(steps, 10))
    states, infos = run.banks[(warmup + steps - 1) & 1].gather()
    alone_ms, alone_evals = run.time_lm_alone(min(steps, 10))
    alone_alg = algorithmic_bytes_lm(c["N"], c["H"], c["W"], NUM_BLOCKS, alone_evals) + 4 * c["N"] * S
    evals = float(sum(i["evaluations"] for i in infos))  # last step's launch
    usable = sum(i["usable"] for i in infos)
    iters = float(np.mean([i["iterations"] for i in infos]))
    run.reset_states()
    run.run_e2e(0, max(1, warmup))
    barrier()
    t0 = time.perf_counter()
    run.run_e2e(warmup, steps)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    alg = algorithmic_bytes_lm(c["N"], c["H"], c["W"], NUM_BLOCKS, evals) + 4 * c["N"] * S
    achieved = alg / (lm_ms * 1e-3) / 1e9
    ef_alg = S * algorithmic_bytes_ef(c["E"], c["H"], c["W"])
    shape = run.banks[0].launch_shape()
    return {
        "t_ms": t_ms, "e2e_s": e2e_s, "launches": int(launches), "iters": iters, "usable": usable, "states": states,
        "launch_shape": "%d evaluator CTAs + %d leader CTAs of 768 threads, %d problems in flight" % shape,
        "roofline": {"kernel": "track_lm_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (traffic or {}).get("track_lm_kernel_dram_bytes_per_launch"),
                     "l2_bytes": (traffic or {}).get("track_lm_kernel_lts_bytes_per_launch"),
                     "algorithmic_bytes_per_launch": alg, "launch_ms": lm_ms, "evaluations_per_launch": evals,
                     "alone": {"launch_ms": alone_ms, "evaluations_per_launch": alone_evals, "achieved": alone_alg / (alone_ms * 1e-3) / 1e9,
                               "frac": alone_alg / (alone_ms * 1e-3) / 1e9 / peak,
                               "note": "the same launch with nothing else on the device (window 0 of every sequence from its initial state): "
                                       "what the concurrent frame build of the next window costs the solve is the difference"},
                     "note": "launch_ms covers track_lm_kernel + mad_kernel (CUDA events on the launching stream, frame builds of the next window running "
                             "beside it); traffic / l2_bytes are from the ncu capture in profiles/ (config 2, 64 sequences): the working set is "
                             "L2-resident, the kernel is bound by per-SM latency / issue, see DESIGN.md 4.2"},
        "event_frame": {"kernels": "scatter_events_kernel + blur_norm_kernel (+ accumulator clear)", "bound": "hbm", "ms": ef_ms,
                        "algorithmic_bytes": ef_alg, "achieved": ef_alg / (ef_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": ef_alg / (ef_ms * 1e-3) / 1e9 / peak,
                        "note": "CUDA events on the library's build stream around one build of all S windows, measured alone (inside a step the build of "
                                "window k+1 shares the device with the solve of window k)"},
    }


def run_native(args):
    import torch
    import torch.distributed as dist
    import edsgpu
    from edsgpu import shard, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the CUDA path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    os.environ.setdefault("EDSGPU_RESERVE_SMS", str(RESERVE_SMS))
    c = synth.CONFIGS[CONFIG]
    # one explicit stream shared by torch (events, copies, NCCL ordering) and the library
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = edsgpu.Context(local_rank, stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        return shard.max_over_ranks(v, dev)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except (OSError, ValueError):
        pass

    total = args.sequences
    procs = max(1, (os.cpu_count() or 1) // world)
    clocks = ClockSampler(local_rank)
    # ---- headline: strong scaling -- the `total` sequences dealt round-robin to the ranks --------------------------
    strong = args.scaling == "strong"
    n_global = total if strong else total * world
    ids = shard.local_sequences(n_global, world, rank)  # sequence s runs on rank s mod world
    data = make_sequences(CONFIG, ids, WINDOWS, procs)
    run = TrackRun(ctx, torch, dev, stream, CONFIG, data, len(ids))
    barrier()
    if rank == 0:
        clocks.start()
    m = measure(run, args.steps, args.warmup, barrier, max_over_ranks, peak, traffic)
    clk = clocks.stop() if rank == 0 else None
    # ---- the one collective of the path: gather the final states with NCCL, from the C++ host layer (libedsgpu_nccl.so) ----
    uid = [edsgpu.Comm.unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(uid, src=0)
    comm = edsgpu.Comm(ctx, world, rank, uid[0])
    final_states = comm.gather_batch(run.banks[(args.warmup + args.steps - 1) & 1], n_global)
    comm.close()
    run.close()
    weak = None
    if world > 1 and strong and not args.no_weak:
        wdata = make_sequences(CONFIG, shard.local_sequences(total * world, world, rank), WINDOWS, procs)
        wrun = TrackRun(ctx, torch, dev, stream, CONFIG, wdata, total)
        barrier()
        wm = measure(wrun, args.steps, args.warmup, barrier, max_over_ranks, peak, traffic)
        weak = {"scaling": "weak", "sequences_per_gpu": total, "value": world * total * args.steps / (wm["t_ms"] * 1e-3),
                "ms_per_step": wm["t_ms"] / args.steps, "e2e_value": world * total * args.steps / wm["e2e_s"], "roofline_frac": wm["roofline"]["frac"]}
        wrun.close()
    side = {}
    if world == 1 and not args.no_side:
        # the other two sensor shapes named in BASELINE.json (configs[0] and configs[2]): same step, own roofline
        for key, cfg, S_side, steps in (("config1", "davis240c", 64, args.steps), ("config3", "gen4_hd", 32, max(5, args.steps // 5))):
            sdata = make_sequences(cfg, range(SIDE_SCENES), SIDE_WINDOWS, procs)
            srun = TrackRun(ctx, torch, dev, stream, cfg, sdata, S_side)
            sm = measure(srun, steps, args.warmup, barrier, max_over_ranks, peak, None)
            sc = synth.CONFIGS[cfg]
            val = S_side * steps / (sm["t_ms"] * 1e-3)
            side[key] = {"workload": "%s %dx%d, %d points, %d events/window, %d sequences (%d distinct scenes x %d windows), B=%d, Huber+MAD, max_iter %d"
                                     % (cfg, sc["W"], sc["H"], sc["N"], sc["E"], S_side, SIDE_SCENES, SIDE_WINDOWS, NUM_BLOCKS, MAX_ITER),
                         "value": val, "unit": "windows/s", "mevents_per_s": val * sc["E"] / 1e6, "ms_per_step": sm["t_ms"] / steps, "steps": steps,
                         "e2e": {"value": S_side * steps / sm["e2e_s"], "unit": "windows/s", "h2d_bytes_per_step": int(S_side * sc["E"] * 5),
                                 "d2h_bytes_per_step": int(S_side * 14 * 8)},
                         "mean_lm_iterations": sm["iters"], "usable": "%d/%d" % (sm["usable"], S_side), "launch_shape": sm["launch_shape"],
                         "roofline": sm["roofline"], "event_frame": sm["event_frame"]}
            srun.close()
    ba_line = bench_ba(ctx, stream) if (rank == 0 and not args.no_ba) else None
    coarse_line = bench_coarse(ctx, stream) if (rank == 0 and not args.no_ba) else None
    depth_line = bench_depth(ctx, stream) if (rank == 0 and not args.no_ba) else None

    if rank == 0:
        S_local = len(ids)
        windows = n_global * args.steps
        value = windows / (m["t_ms"] * 1e-3)
        e2e_value = windows / m["e2e_s"]
        H, W, N, E = c["H"], c["W"], c["N"], c["E"]
        ws = S_local * H * W * 12 + S_local * E * 5 + S_local * N * 48
        out = {
            "metric": "tracking_windows_per_s_640x480", "value": value, "unit": "windows/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": m["t_ms"] / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32 (fp64 warp/projection, reductions and LM solve)", "data": "synthetic",
            "mevents_per_s": value * E / 1e6,
            "config": {"workload": workload_string(c), "sequences": n_global, "sequences_per_gpu": S_local,
                       "windows_per_sequence": "%d consecutive windows per sequence, then round again; every sequence has its own scene and key frame" % WINDOWS,
                       "l2": "per-step working set ~%d MB per GPU (event accumulators + frames + key frames) %s 126 MB L2, no explicit flush"
                             % (int(ws / 1e6), ">" if ws > 126e6 else "<= (the working set of a sharded batch fits L2: stated, not flushed)"),
                       "mean_lm_iterations": m["iters"], "launch_shape": m["launch_shape"], "reserved_sms": int(os.environ["EDSGPU_RESERVE_SMS"]),
                       "usable": "%d/%d" % (m["usable"], S_local), "parity": PARITY_NOTE,
                       "parallelism": "%d sequences dealt round-robin to %d rank(s), no collective on the data path, one ncclAllGather of the state records at the end (edsgpu_batch_gather_states_nccl)" % (n_global, world)},
            "e2e": {"value": e2e_value, "unit": "windows/s", "h2d_bytes_per_step": int(S_local * E * 5), "d2h_bytes_per_step": int(S_local * 14 * 8),
                    "ms_per_step": 1e3 * m["e2e_s"] / args.steps,
                    "how": "edsgpu_event_frame_create_batch (pinned host events) + edsgpu_batch_optimize + state read-back every step; "
                           "next window's H2D copy and frame build (other bank of slots, build stream) overlap the current solve; the host reads step k's states while step k+1 is queued; bytes are per GPU"},
            "gpu_launches": m["launches"],
            "clocks": clk,
            "roofline": m["roofline"],
            "event_frame": m["event_frame"],
            "final_state_checksum": float(np.abs(final_states).sum()),
        }
        if weak:
            out["weak"] = weak
        out.update(side)
        if ba_line:
            out["ba"] = ba_line
        if coarse_line:
            out["coarse_tracker"] = coarse_line
        if depth_line:
            out["depth_filter"] = depth_line
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as O
            cores = os.cpu_count() or 1
            th = min(NUM_BLOCKS, cores)
            n_seq, per_seq = min(8, len(data)), 25  # ~10-20 s of CPU work: 200 windows
            cdata = data[:n_seq]
            wps, dt, n, it = cpu_windows_per_s(cdata, per_seq, 1, th)
            wps_a, dt_a, n_a, it_a = cpu_windows_per_s(cdata, per_seq, 1, th, jacobian_mode=0)
            out["cpu_baseline"] = {"value": wps, "unit": "windows/s", "cores": th, "kind": "port",
                                   "sample": "%d windows of the same workload (%d sequences x %d consecutive windows, warm-started, tau carried), one sequence at a time, "
                                             "%d threads (one per residual block), dual-number Jacobians, %.1f s" % (n, n_seq, per_seq, th, dt),
                                   "build": O.BUILD_FLAGS, "mean_lm_iterations": it,
                                   "analytic_jacobian": {"value": wps_a, "unit": "windows/s", "seconds": dt_a, "mean_lm_iterations": it_a,
                                                         "note": "C-analytic variant (hand Jacobian, same solver): the fairer best-CPU number, BASELINE.md section 3"}}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--sequences", type=int, default=64, help="independent sequences: in total (strong, configs[4]) or per GPU (weak)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the weak-scaling extra measurement")
    ap.add_argument("--no-side", action="store_true", help="skip the config1 / config3 measurements")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ba", action="store_true", help="skip the BA / coarse tracker / depth filter side measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not args.no_cpu_baseline and int(os.environ.get("WORLD_SIZE", "1")) == 1:
            from oracle import oracle as O  # cpu_baseline leg: the timing build of the checker, chosen before anything loads it
            O.use_native_build()
        run_native(args)


if __name__ == "__main__":
    main()
