"""C++ programs that use ONLY the C ABI (include/grample_b200.h) the way a compiled host in the place of cmd/root.go
would: tests/threads_test.cpp (16 pthreads on one handle, SURVEY 8b thread safety) and tests/fleet_test.cpp (one
process driving every visible GPU through a gb_fleet, compared with one GPU holding all chains).  Built on the CPU,
run on the GPU."""
import os
import subprocess

import pytest

import grample_b200 as gb
from conftest import RES, ROOT

LIBDIR = os.path.join(ROOT, "grample_b200")


def build_exe(name):
    src = os.path.join(ROOT, "tests", name + ".cpp")
    exe = os.path.join(ROOT, "tests", name + ".bin")
    deps = [src, os.path.join(ROOT, "include", "grample_b200.h")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-o", exe, src, "-L" + LIBDIR, "-lgrample_b200",
                        "-Wl,-rpath," + LIBDIR, "-lpthread"], check=True, cwd=ROOT)
    return exe


@pytest.mark.parametrize("name", ["threads_test", "fleet_test"])
def test_native_hosts_build_against_the_c_abi(name):
    exe = build_exe(name)
    if gb.device_count() == 0:  # and refuse to run without a GPU
        r = subprocess.run([exe, RES], capture_output=True, text=True)
        assert r.returncode == 3 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_sixteen_threads_on_one_handle():
    r = subprocess.run([build_exe("threads_test"), RES], capture_output=True, text=True, timeout=600)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout
    for name in ("TestThreadsF64", "TestThreadsTable", "TestThreadsF32HighCard"):
        assert f"PASS {name}" in r.stdout


@pytest.mark.gpu
def test_fleet_of_all_visible_gpus_equals_one_gpu():
    """with one visible GPU this still drives the fleet entry points (world of 1, no NCCL); with N it is the multi-GPU test"""
    r = subprocess.run([build_exe("fleet_test"), RES], capture_output=True, text=True, timeout=900)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout
    for name in ("TestFleetSimpleF64", "TestFleetSimpleTable", "TestFleetAdaptiveHybrid"):
        assert f"PASS {name}" in r.stdout
