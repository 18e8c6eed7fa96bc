"""The C++ side of the drop-in boundary: a flowz user's program (tests/cpp/dropin.cpp, reference spelling:
placeholders, `|=`, `~`, `_1[_n]`, std::ref, compile(), operator()) built against include/flowz/flowz.hpp and
libzignal_b200.so.  Host ticks run anywhere (BASELINE configs[0]); with a GPU the same graphs run as blocks
through on_device() and must reproduce the per-sample ticks bit for bit."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, zg):
    exe = str(tmp_path / "dropin")
    libdir = os.path.dirname(zg.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "dropin.cpp"), "-L", libdir, "-lzignal_b200",
                           f"-Wl,-rpath,{libdir}", "-o", exe])
    return exe


def test_cpp_program_host_ticks(zg, tmp_path):
    out = subprocess.run([_build(tmp_path, zg)], capture_output=True, text=True)
    assert out.returncode == 0 and "dropin ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_program_blocks_on_device(zg, tmp_path):
    out = subprocess.run([_build(tmp_path, zg), "gpu"], capture_output=True, text=True)
    assert out.returncode == 0 and "dropin ok" in out.stdout, out.stdout + out.stderr
