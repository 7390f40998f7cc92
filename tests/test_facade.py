"""The C++ facade (semantic_slam_b200/host/...) mirrors ps_graph_slam::GraphSLAM; it must compile
against include/ssb.h and link libssb.so (CPU), and run end to end on a GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "facade_check")
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "semantic_slam_b200", "host"),
           os.path.join(ROOT, "tests", "facade_check.cpp"), "-o", exe, "-L", os.path.join(ROOT, "semantic_slam_b200"), "-lssb",
           "-Wl,-rpath," + os.path.join(ROOT, "semantic_slam_b200"), "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
    subprocess.check_call(cmd)
    return exe


def test_facade_compiles_and_links(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "sm_100a" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_facade_runs_on_gpu(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe, "run"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
