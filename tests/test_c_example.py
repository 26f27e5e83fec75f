"""The C ABI from plain C: examples/search_demo.c must compile as C11 against include/vlite_b200.h
(so the header is valid C, not only C++), link against the shared library and -- without a GPU --
fail loudly with the 'no CPU fallback' message; on a B200 it must find its needle."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "vlite_b200", "lib")


def _build(tmp_path):
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    from vlite_b200 import build
    build.build()                                                        # no-op when the library is current
    exe = str(tmp_path / "search_demo")
    cmd = ["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "search_demo.c"), "-L", LIBDIR, "-lvlite_b200",
           f"-Wl,-rpath,{LIBDIR}", "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_c_example_builds_and_refuses_to_run_without_a_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu test")
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 2 and "no CPU fallback" in res.stderr


@pytest.mark.gpu
def test_c_example_finds_its_needle(gpu, tmp_path):
    exe = _build(tmp_path)
    res = subprocess.run([exe, "300000"], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    lines = res.stdout.splitlines()
    assert lines[1].split() == ["#0", "row", "100000", "distance", "2"]
    assert "after deleting row 100000" in lines[-1]
