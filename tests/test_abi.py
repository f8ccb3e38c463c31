"""CPU: the C-ABI library loads, exports every symbol include/edsgpu.h declares, and fails loudly
without a GPU (no CPU fallback).  No compute call is made here."""
import ctypes
import os
import re
import subprocess

import pytest

import edsgpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(edsgpu.LIB_PATH):
        subprocess.run(["make", "-C", os.path.join(ROOT, "slam-eds_b200")], check=True, capture_output=True)
    return edsgpu.load()


def header_symbols(name="edsgpu.h"):
    txt = open(os.path.join(ROOT, "include", name)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(edsgpu_[a-z_0-9]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(lib):
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "libedsgpu.so does not export %s" % s
    assert sorted(edsgpu.SYMBOLS) == syms, "python binding and header disagree"


def test_nccl_gather_library_exports_its_header(lib):
    """libedsgpu_nccl.so (the final gather of a sharded batch, in the C++ host layer) loads next to libedsgpu.so and
    exports every symbol include/edsgpu_nccl.h declares; nothing is called (no GPU, no communicator)."""
    nl = edsgpu.load_nccl()
    syms = header_symbols("edsgpu_nccl.h")
    assert sorted(edsgpu.NCCL_SYMBOLS) == syms
    for s in syms:
        assert hasattr(nl, s), "libedsgpu_nccl.so does not export %s" % s
    needed = subprocess.run(["ldd", edsgpu.NCCL_LIB_PATH], capture_output=True, text=True).stdout
    assert "libnccl.so" in needed and "libedsgpu.so" in needed
    assert "libnccl" not in subprocess.run(["ldd", edsgpu.LIB_PATH], capture_output=True, text=True).stdout  # the core library stays NCCL-free


def test_version_and_struct_sizes(lib):
    assert b"sm_100a" in lib.edsgpu_version()
    assert ctypes.sizeof(edsgpu.TrackerConfig) == 40
    assert ctypes.sizeof(edsgpu.TrackerInfo) == 56


def test_library_is_built_for_sm_100a():
    out = subprocess.run(["cuobjdump", "-lelf", edsgpu.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(edsgpu.EdsGpuError):
        edsgpu.Context(0)


def test_product_does_not_touch_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may reference oracle/."""
    pkg = os.path.join(ROOT, "slam-eds_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.lower(), os.path.join(dp, f)
