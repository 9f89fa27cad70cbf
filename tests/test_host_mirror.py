"""The C++ host mirror of the reference's `sampler` API (grample_b200/host/grample.hpp) is
exercised by tests/host_mirror_test.cpp, a restatement of the reference's own sampler tests
against that mirror.  Here: build it (CPU), and run it on the GPU."""
import os
import subprocess

import pytest

import grample_b200 as gb
from conftest import RES, ROOT

EXE = os.path.join(ROOT, "tests", "host_mirror_test.bin")


def build_exe():
    src = os.path.join(ROOT, "tests", "host_mirror_test.cpp")
    libdir = os.path.join(ROOT, "grample_b200")
    deps = [src, os.path.join(libdir, "host", "grample.hpp"), os.path.join(ROOT, "include", "grample_b200.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-o", EXE, src, "-L" + libdir, "-lgrample_b200",
                        "-Wl,-rpath," + libdir], check=True, cwd=ROOT)
    return EXE


def test_host_mirror_builds_against_the_c_abi():
    exe = build_exe()
    assert os.path.exists(exe)
    if gb.device_count() == 0:  # and fails loudly, test by test, when there is no GPU
        r = subprocess.run([exe, RES], capture_output=True, text=True)
        assert r.returncode != 0 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_host_mirror_reference_tests_pass_on_gpu():
    r = subprocess.run([build_exe(), RES], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout
    for name in ("TestWorkingGibbsSimple", "TestSingleStepSample", "TestWorkingGibbsCollapsed", "TestFullGibbsCollapsed", "TestMergeChains",
                 "TestMainLoopSimple", "TestMainLoopAdaptive"):
        assert f"PASS {name}" in r.stdout
