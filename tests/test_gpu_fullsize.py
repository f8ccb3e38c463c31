"""GPU tests at BASELINE.json's full sizes (configs 2 and 3): parity with the oracle where it finishes
in seconds, plus size-independent properties of the domain (additivity and polarity antisymmetry of
the event accumulation, unit L2 norm of the normalised frame, determinism, warm-start consistency)."""
import numpy as np
import pytest

import drift
import edsgpu
from edsgpu import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["gen3_vga", "gen4_hd"])
def prob(request):
    scene, kf, wins = synth.make_problem(request.param, 3, 2)
    return request.param, kf, wins


def test_event_accumulation_is_additive_and_antisymmetric(gpu_ctx, prob):
    name, kf, wins = prob
    H, W = kf["H"], kf["W"]
    a, b = wins
    ef = edsgpu.EventFrame(gpu_ctx, H, W)

    def acc(x, y, p):
        ef.create(x, y, p, None, mode=edsgpu.DRAW_NN, use_exp_weights=False, sigma=0.0)
        return ef.frames.read_accumulator(0)

    A, B = acc(a["x"], a["y"], a["pol"]), acc(b["x"], b["y"], b["pol"])
    AB = acc(np.concatenate([a["x"], b["x"]]), np.concatenate([a["y"], b["y"]]), np.concatenate([a["pol"], b["pol"]]))
    assert np.array_equal(AB, A + B)                       # integer accumulation is exactly additive
    assert np.array_equal(acc(a["x"], a["y"], 1 - a["pol"]), -A)  # flipping every polarity negates the image
    assert int(np.abs(A).sum() >> edsgpu.ACC_FRACTION_BITS) <= len(a["x"])
    o = O.event_frame(a["x"], a["y"], a["pol"], a["ts"], H, W, method="nn", use_exp=False, sigma=0.0)
    assert np.array_equal(A, np.rint(o["img"]).astype(np.int64) << edsgpu.ACC_FRACTION_BITS)


def test_normalised_frame_has_unit_norm_and_matches_oracle(gpu_ctx, prob):
    name, kf, wins = prob
    H, W = kf["H"], kf["W"]
    w = wins[0]
    ef = edsgpu.EventFrame(gpu_ctx, H, W).create(w["x"], w["y"], w["pol"], w["ts"], want_host_frame=True)
    assert abs(np.linalg.norm(ef.event_frame) - 1.0) < 1e-12
    o = O.event_frame(w["x"], w["y"], w["pol"], w["ts"], H, W)
    assert abs(ef.norm - o["norm"]) < 1e-11 * o["norm"] and np.abs(ef.event_frame - o["frame"]).max() < 1e-11


def test_full_size_evaluate_and_solve_match_oracle(gpu_ctx, prob):
    name, kf, wins = prob
    H, W = kf["H"], kf["W"]
    w = wins[0]
    ef = edsgpu.EventFrame(gpu_ctx, H, W).create(w["x"], w["y"], w["pol"], w["ts"])
    o = O.event_frame(w["x"], w["y"], w["pol"], w["ts"], H, W)
    kfd = edsgpu.KeyFrame(gpu_ctx, kf, 8)
    g = edsgpu.tracker_evaluate(gpu_ctx, kfd, ef.frames, 0, w["x_init"])
    e = O.tracker_evaluate(kf, o["frame"], w["x_init"], 8)
    assert np.abs(g["residuals"] - e["residuals"]).max() <= 1e-5 * np.abs(e["residuals"]).max()
    assert np.all(np.abs(g["jacobian"] - e["jacobian"]).max(0) <= 1e-5 * np.abs(e["jacobian"]).max(0))
    assert abs(g["cost"] - e["cost"]) <= 1e-6 * e["cost"]
    # Strict pose gate (1e-4 rad, 1e-4 x scene depth) at 20 iterations, where the reference algorithm is
    # deterministic at its input precision (tests/test_oracle_tracking.py::test_oracle_self_drift_bounds_the_pose_gate).
    x0 = w["x_init"]
    tr20 = edsgpu.Tracker(gpu_ctx, num_blocks=8, max_iterations=20)
    tr20.set_state(x0[:3], x0[3:7], x0[7:], 0.05)
    r20 = tr20.optimize(kfd, ef.frames, 0)
    s20 = O.tracker_solve(kf, o["frame"], x0, num_blocks=8, max_iterations=20, threads=8)
    assert r20["usable"] and r20["info"]["iterations"] == s20["info"]["iterations"]
    assert r20["info"]["successful_steps"] == s20["info"]["successful_steps"]
    assert synth.quat_angle(r20["qx"], s20["x"][3:7]) < drift.ANGLE_GATE and np.linalg.norm(r20["px"] - s20["x"][:3]) < drift.DEPTH_GATE
    assert abs(r20["info"]["final_cost"] - s20["info"]["final_cost"]) < 1e-6 * s20["info"]["final_cost"]
    tr20.close()
    # At the benchmark's cap of 30 the oracle's own pose moves by more than the gate under one-ulp perturbations of its
    # inputs (same oracle test): the GPU must agree with the oracle to within DRIFT_FACTOR x that measured self-drift
    # (or the gate, whichever is larger), take the same number of steps and reach the same cost.
    tr = edsgpu.Tracker(gpu_ctx, num_blocks=8, max_iterations=30)
    tr.set_state(x0[:3], x0[3:7], x0[7:], 0.05)
    r = tr.optimize(kfd, ef.frames, 0)
    bound_a, bound_t, (da, dt, _), s = drift.pose_bounds(kf, o["frame"], x0, 30)
    assert r["usable"] and r["info"]["iterations"] == s["info"]["iterations"]
    assert synth.quat_angle(r["qx"], s["x"][3:7]) < bound_a and np.linalg.norm(r["px"] - s["x"][:3]) < bound_t, (da, dt)
    assert bound_a < 2e-3 and bound_t < 2e-3 * synth.Z0  # the bound itself stays far below anything a tracker would notice
    assert abs(r["info"]["final_cost"] - s["info"]["final_cost"]) < 2e-4 * s["info"]["final_cost"]
    assert r["info"]["final_cost"] <= r20["info"]["final_cost"] <= r["info"]["initial_cost"]
    # determinism + odd batch sizes: three trackers on the same problem give bit-identical states
    fr3 = edsgpu.Frames(gpu_ctx, H, W, 3)
    E = len(w["x"])
    edsgpu.event_frames_batch(gpu_ctx, fr3, 0, 3, np.tile(w["x"], 3), np.tile(w["y"], 3), np.tile(w["pol"], 3), E)
    trs = [edsgpu.Tracker(gpu_ctx, num_blocks=8, max_iterations=30) for _ in range(3)]
    for t in trs:
        t.set_state(x0[:3], x0[3:7], x0[7:], 0.05)
    b = edsgpu.TrackerBatch(gpu_ctx, trs, [kfd] * 3, fr3, 0)
    b.optimize()
    st, infos = b.gather()
    assert np.array_equal(st[0], st[1]) and np.array_equal(st[1], st[2])
    assert np.array_equal(st[0, :13], r["x"]) and st[0, 13] == r["next_loss_param"]
    b.close()
    for t in trs + [tr]:
        t.close()
    kfd.close()


def test_sequence_of_full_size_windows_carries_pose_and_mad_tau(gpu_ctx):
    """Four consecutive 640x480 windows of one sequence, solved the way bench.py solves them: warm start from the previous
    window's (px, qx, vx) and the MAD-updated loss parameter (Tracker.cpp:233), everything carried on the device.  Every
    window is checked against the oracle started from the state the device started from: strict gate at 20 iterations,
    self-drift bound at the benchmark's 30; tau within 2e-4 (one rank of the median) whenever the poses agree."""
    scene, kf, wins = synth.make_problem("gen3_vga", 5, 4)
    H, W = kf["H"], kf["W"]
    kfd = edsgpu.KeyFrame(gpu_ctx, kf, 8)
    ef = edsgpu.EventFrame(gpu_ctx, H, W)
    for cap in (20, 30):
        tr = edsgpu.Tracker(gpu_ctx, num_blocks=8, max_iterations=cap, loss_param_method=edsgpu.LOSS_PARAM_MAD)
        x0 = wins[0]["x_init"]
        tr.set_state(x0[:3], x0[3:7], x0[7:], 0.05)
        x_prev, tau_prev = x0.copy(), 0.05
        taus = []
        for w in wins:
            ef.create(w["x"], w["y"], w["pol"], w["ts"])
            frame = O.event_frame(w["x"], w["y"], w["pol"], w["ts"], H, W)["frame"]
            r = tr.optimize(kfd, ef.frames, 0)
            if cap == 20:
                s = O.tracker_solve(kf, frame, x_prev, num_blocks=8, loss_param=tau_prev, max_iterations=cap, threads=8)
                bound_a, bound_t = drift.ANGLE_GATE, drift.DEPTH_GATE
            else:
                bound_a, bound_t, _, s = drift.pose_bounds(kf, frame, x_prev, cap, loss_param=tau_prev)
            assert r["usable"] and s["info"]["usable"] and r["info"]["iterations"] == s["info"]["iterations"]
            da, dt = drift.pose_diff(r["x"], s["x"])
            assert da < bound_a and dt < bound_t, (cap, da, dt)
            assert abs(r["info"]["final_cost"] - s["info"]["final_cost"]) < 2e-4 * s["info"]["final_cost"]
            # tau is an order statistic (1.345 x 1.4826 x the median of |r - median r|, Tracker.cpp:281-317) of 10 240 residuals
            # spaced ~2e-5 apart: fp32 residuals can move the pick by one rank, i.e. by ~1e-4 relative
            if da < 1e-6 and dt < 1e-6:
                assert abs(r["next_loss_param"] - s["next_loss_param"]) < 2e-4 * s["next_loss_param"]
            assert abs(r["next_loss_param"] - s["next_loss_param"]) < 2e-2 * s["next_loss_param"]
            taus.append(r["next_loss_param"])
            x_prev, tau_prev = r["x"].copy(), r["next_loss_param"]  # the device's own state feeds the next window
        assert len(set(np.round(taus, 12))) == len(taus) and all(0.0 < t < 1.0 for t in taus)  # tau really is updated per window
        tr.close()
    kfd.close()
