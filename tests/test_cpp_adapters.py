"""The C++ host layer (slam-eds_b200/host/edsgpu_adapters.hpp) over the C ABI: compiles everywhere,
runs on the GPU box and must agree with the oracle like the Python mirror does."""
import os
import subprocess

import numpy as np
import pytest

from edsgpu import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "adapter_demo.cpp")
LIBDIR = os.path.join(ROOT, "slam-eds_b200")


def _build(tmp_path):
    exe = str(tmp_path / "adapter_demo")
    if not os.path.exists(os.path.join(LIBDIR, "libedsgpu.so")):
        subprocess.run(["make", "-C", LIBDIR], check=True, capture_output=True)
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-o", exe, SRC, "-L" + LIBDIR, "-ledsgpu", "-Wl,-rpath," + LIBDIR,
           "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_adapters_compile_and_link(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


@pytest.mark.gpu
def test_adapters_match_oracle(tmp_path):
    from oracle import oracle as O
    exe = _build(tmp_path)
    scene, kf, wins = synth.make_problem("davis240c", 0, 1)
    w = wins[0]
    B, iters = 8, 20
    path = tmp_path / "in.bin"
    with open(path, "wb") as f:
        np.array([kf["H"], kf["W"], len(kf["idp"]), len(w["x"]), B, iters], np.int32).tofile(f)
        np.array([kf["fx"], kf["fy"], kf["cx"], kf["cy"]], np.float64).tofile(f)
        for a in (kf["grad"], kf["norm_coord"], kf["idp"], kf["weights"], w["x_init"]):
            np.ascontiguousarray(a, np.float64).tofile(f)
        w["x"].astype(np.uint16).tofile(f); w["y"].astype(np.uint16).tofile(f); w["pol"].astype(np.uint8).tofile(f)
        w["ts"].astype(np.int64).tofile(f)
    r = subprocess.run([exe, str(path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = {l.split()[0]: [float(v) for v in l.split()[1:]] for l in r.stdout.strip().splitlines()}
    ef = O.event_frame(w["x"], w["y"], w["pol"], w["ts"], kf["H"], kf["W"])
    s = O.tracker_solve(kf, ef["frame"], w["x_init"], num_blocks=B, max_iterations=iters)
    assert out["ok"] == [1.0] and out["throw"] == [1.0]
    assert abs(out["norm"][0] - ef["norm"]) < 1e-11 * ef["norm"] and out["time"][0] == ef["time"]
    assert synth.quat_angle(np.array(out["qx"]), s["x"][3:7]) < 1e-4 and np.linalg.norm(np.array(out["px"]) - s["x"][:3]) < 2e-4
    assert out["iterations"][0] == s["info"]["iterations"]
    assert abs(out["tau"][0] - s["next_loss_param"]) < 2e-3 * s["next_loss_param"]
    # T_kf_ef is the inverse of (qx, px) (Tracker.cpp:220)
    R = synth.quat_to_rot(np.array(out["qx"]))
    np.testing.assert_allclose(np.array(out["T_kf_ef"][:3]), -R.T @ np.array(out["px"]), atol=1e-12)
    assert abs(out["res0"][0] - s["residuals"][0]) < 2e-3 * np.abs(s["residuals"]).max()
