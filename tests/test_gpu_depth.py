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
