"""host/plonky2_api.hpp (the C++ mirror of the plonky2 interface above the C ABI): compiles everywhere, runs on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "build", "test_host_api")


def _build():
    import intmax_zkp_core_b200 as z
    z.lib()   # make sure libb200zkp.so exists
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    libdir = os.path.join(ROOT, "intmax_zkp_core_b200")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-o", BIN, os.path.join(ROOT, "host", "test_host_api.cpp"),
                           "-L" + libdir, "-lb200zkp", "-Wl,-rpath," + libdir])


def test_host_cpp_api_compiles_and_links():
    _build()
    assert os.path.exists(BIN)


@pytest.mark.gpu
def test_host_cpp_api_runs_on_gpu():
    _build()
    out = subprocess.run([BIN], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host C++ API ok" in out.stdout
