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


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "edsgpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(edsgpu_[a-z_0-9]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(lib):
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "libedsgpu.so does not export %s" % s
    assert sorted(edsgpu.SYMBOLS) == syms, "python binding and header disagree"


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
