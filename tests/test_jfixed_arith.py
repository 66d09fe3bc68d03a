"""Arithmetic of the NIX_J_FIXED experiment (nix_b200/csrc/jfixed.cuh: k_deposit's J tile as 64-bit fixed-point words
added to with two native 32-bit atomics), checked on the host: tests/native/jfixed_arith.cpp includes the header AS IS
with one-line stand-ins for the CUDA intrinsics.  The experiment itself is OFF in the built library (DESIGN.md section
6) and has never run on a device; this test pins only what can be pinned without one."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fixed_point_tile_arithmetic(tmp_path):
    exe = str(tmp_path / "jfixed_arith")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "nix_b200", "csrc"), "-o", exe,
                    os.path.join(ROOT, "tests", "native", "jfixed_arith.cpp")], check=True, capture_output=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stdout + out.stderr


def test_the_experiment_is_off_by_default():
    src = open(os.path.join(ROOT, "nix_b200", "csrc", "push_deposit.cu")).read()
    assert "#define NIX_J_FIXED 0" in src
    from nix_b200 import build
    assert not any("NIX_J_FIXED" in f for f in build.NVCC_FLAGS)
