"""Host-only checks of the layout arithmetic the packer, the search kernels and the parity hooks share (tri layout
offsets and tails, combination indices, shared-memory maps): tests/native/layout_check.cu is built with nvcc as a plain
host program -- no kernel is launched, so it runs on the CPU."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc not available")
def test_layout_arithmetic(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "layout_check")
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.run([nvcc, "-std=c++17", "-O1", "-ccbin", "g++", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                    os.path.join(HERE, "native", "layout_check.cu")], check=True, env=env)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "layout checks ok" in r.stdout
