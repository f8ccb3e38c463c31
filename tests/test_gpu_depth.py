"""GPU parity test for the per-point depth filter (run with -m gpu): DepthPoints::update through the C ABI
against the CPU oracle, fp64 on both sides (libm vs CUDA math: a few ulp), gate 1e-10 relative."""
import numpy as np
import pytest

import edsgpu
from oracle import oracle as O
from test_oracle_depth import CX, CY, FX, FY, px_angle, scene

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 333, 20000])
def test_depth_filter_matches_oracle_over_several_updates(gpu_ctx, n):
    T, kf, ef, depth = scene(n=n, seed=n, noise=0.25)
    idp0 = (1.0 / depth) * (1.0 + np.random.default_rng(1).normal(scale=0.2, size=n))
    dp = edsgpu.DepthPoints(gpu_ctx, n, FX, FY, CX, CY, 0.5, 5.5, inv_depth=idp0, init_a=10.0, init_b=10.0)
    st = np.stack([idp0, np.full(n, 25.0 / 36.0), np.full(n, 10.0), np.full(n, 10.0)], 1)
    assert np.array_equal(dp.get(), st)
    rng = np.random.default_rng(2)
    for k in range(4):
        meas = ef + rng.normal(scale=0.2, size=ef.shape)
        if k == 2:
            meas[::7] += 30.0                            # outliers
        tracks = k % 2 == 1
        ok = dp.update(T, kf, meas - kf if tracks else meas, coords_are_tracks=tracks)
        st, ok_ref = O.depth_update(FX, FY, CX, CY, 5.0, px_angle(), T, kf, meas, st)
        got = dp.get()
        assert np.array_equal(ok, ok_ref)
        assert np.allclose(got, st, rtol=1e-10, atol=0)
    dp.close()


def test_depth_filter_initialisation_and_validation(gpu_ctx):
    dp = edsgpu.DepthPoints(gpu_ctx, 5, FX, FY, CX, CY, 1.0, 5.0)        # no prior: mean depth, sigma2 = range^2
    st = dp.get()
    assert np.allclose(st, np.tile([1.0 / 2.0, 16.0, 10.0, 10.0], (5, 1)))
    dp.close()
    with pytest.raises(edsgpu.EdsGpuError):
        edsgpu.DepthPoints(gpu_ctx, 5, FX, FY, CX, CY, 5.0, 1.0)


def test_tracking_loop_with_the_depth_filter_stays_on_the_device(gpu_ctx):
    """The per-window loop of the reference: optimize -> getCoord -> DepthPoints::update -> the next optimize
    uses the filter's inverse depths.  On the device the last three steps are one call and nothing is re-uploaded;
    the oracle runs the same loop on the host."""
    from edsgpu import synth
    scene_, kf, wins = synth.make_problem("davis240c", 3, 3)
    N = len(kf["idp"])
    rng = np.random.default_rng(9)
    idp0 = kf["idp"] * (1.0 + rng.normal(scale=0.05, size=N))          # the map's depths are a bit off
    kf_coord = np.stack([kf["fx"] * kf["norm_coord"][:, 0] + kf["cx"], kf["fy"] * kf["norm_coord"][:, 1] + kf["cy"]], 1)
    kf_gpu = dict(kf); kf_gpu["idp"] = idp0
    kfd = edsgpu.KeyFrame(gpu_ctx, kf_gpu, 8)
    dp = edsgpu.DepthPoints(gpu_ctx, N, kf["fx"], kf["fy"], kf["cx"], kf["cy"], 0.5, 5.5, inv_depth=idp0)
    tr = edsgpu.Tracker(gpu_ctx, num_blocks=8, max_iterations=12)
    ef = edsgpu.EventFrame(gpu_ctx, kf["H"], kf["W"])
    x = wins[0]["x_init"].copy()
    tau = 0.05
    tr.set_state(x[:3], x[3:7], x[7:], tau)
    st = np.stack([idp0, np.full(N, 25.0 / 36.0), np.full(N, 10.0), np.full(N, 10.0)], 1)
    kf_cpu = dict(kf); kf_cpu["idp"] = idp0.copy()
    px_ang = np.arctan(3.0 / (2 * kf["fx"])) + np.arctan(3.0 / (2 * kf["fy"]))
    with pytest.raises(edsgpu.EdsGpuError):
        dp.update_from_tracker(tr, kfd, None)            # KeyFrame::coord was never given
    for k, w in enumerate(wins):
        ef.create(w["x"], w["y"], w["pol"], w["ts"])
        r = tr.optimize(kfd, ef.frames, 0)
        coord_gpu, out_gpu = dp.get_coord(tr, kfd)
        dp.update_from_tracker(tr, kfd, kf_coord if k == 0 else None, refresh_keyframe=True)
        # oracle: the same loop on the host
        o = O.event_frame(w["x"], w["y"], w["pol"], w["ts"], kf["H"], kf["W"])
        s = O.tracker_solve(kf_cpu, o["frame"], x, num_blocks=8, loss_param=tau, max_iterations=12)
        x, tau = s["x"], s["next_loss_param"]
        coord, outl = O.tracker_get_coord(kf_cpu, st[:, 0], x[:3], x[3:7])
        R = synth.quat_to_rot(x[3:7]) if hasattr(synth, "quat_to_rot") else None
        T_ef_kf = np.eye(4)
        xq, yq, zq, wq = x[3:7]
        T_ef_kf[:3, :3] = np.array([[1 - 2 * (yq * yq + zq * zq), 2 * (xq * yq - zq * wq), 2 * (xq * zq + yq * wq)],
                                    [2 * (xq * yq + zq * wq), 1 - 2 * (xq * xq + zq * zq), 2 * (yq * zq - xq * wq)],
                                    [2 * (xq * zq - yq * wq), 2 * (yq * zq + xq * wq), 1 - 2 * (xq * xq + yq * yq)]])
        T_ef_kf[:3, 3] = x[:3]
        st_new, _ = O.depth_update(kf["fx"], kf["fy"], kf["cx"], kf["cy"], 5.0, px_ang, np.linalg.inv(T_ef_kf), kf_coord, coord, st)
        # points that left the event frame are erased by the reference before the update (Tracker.cpp:356-372); the device
        # keeps them in place with their state untouched
        st = np.where(np.asarray(outl, bool)[:, None], st, st_new)
        kf_cpu["idp"] = st[:, 0].copy()
        # pose of this window, warped coordinates, filter state
        assert synth.quat_angle(r["qx"], x[3:7]) < 1e-4 and np.linalg.norm(r["px"] - x[:3]) < 1e-4 * 2.0
        assert np.array_equal(out_gpu, outl) and np.abs(coord_gpu - coord).max() < 5e-3
        got = dp.get()
        assert np.allclose(got[:, 0], st[:, 0], rtol=2e-3) and np.median(np.abs(got[:, 0] - st[:, 0]) / st[:, 0]) < 1e-5
        assert np.allclose(got[:, 1], st[:, 1], rtol=5e-2)
    # (getCoord warps with the filter's own depths, so these measurements carry no new depth information: in the
    # reference the KLT refinement of Tracker::trackPoints supplies it; the loop above checks the plumbing.)
    # the key frame on the device follows the filter: refresh == re-upload, bit for bit
    ev = edsgpu.tracker_evaluate(gpu_ctx, kfd, ef.frames, 0, x)
    kf_ref = dict(kf); kf_ref["idp"] = dp.get()[:, 0]
    kfd2 = edsgpu.KeyFrame(gpu_ctx, kf_ref, 8)
    ev2 = edsgpu.tracker_evaluate(gpu_ctx, kfd2, ef.frames, 0, x)
    assert np.array_equal(ev["residuals"], ev2["residuals"]) and np.array_equal(ev["H"], ev2["H"])
    kfd.close(); kfd2.close(); dp.close(); tr.close()
