"""tests/c_abi_harness.c: a plain-C caller that performs the call sequence of the cgo binding (go/frontier/frontier.go) through
the same flat wrappers (go/frontier/shim.h).  Compiling it is a CPU test (the public header and the shim are valid C, warning-free);
running it needs a B200."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bobrapet_b200", "lib")


def _build(tmp_path):
    exe = str(tmp_path / "c_abi_harness")
    cmd = ["gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "go", "frontier"),
           os.path.join(ROOT, "tests", "c_abi_harness.c"), "-L" + LIB, "-lbobrafrontier", "-Wl,-rpath," + LIB, "-o", exe]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


def test_harness_compiles_as_c_against_the_public_header(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True)       # without a GPU bf_create fails cleanly with a status
    assert out.returncode in (0, 1) and "Segmentation" not in out.stderr


@pytest.mark.gpu
def test_harness_runs_the_go_call_sequence(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "c_abi_harness: ok" in out.stdout, (out.stdout, out.stderr)
