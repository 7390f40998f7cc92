"""The C++ facade (semantic_slam_b200/host/...) mirrors ps_graph_slam::GraphSLAM; it must compile
against include/ssb.h and link libssb.so (CPU), and run end to end on a GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, eigen=False):
    exe = str(tmp_path / ("facade_check_eigen" if eigen else "facade_check"))
    # eigen=True: the `#if SSB_HAVE_EIGEN` branch of the facade against tests/mock_eigen (Eigen's call syntax and its
    # compile-time restrictions; the image has no Eigen) — the configuration the reference is built in
    extra = ["-DSSB_USE_EIGEN", "-I", os.path.join(ROOT, "tests", "mock_eigen")] if eigen else []
    cmd = ["g++", "-std=c++17", "-O1", "-Wall"] + extra + ["-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "semantic_slam_b200", "host"),
           os.path.join(ROOT, "tests", "facade_check.cpp"), "-o", exe, "-L", os.path.join(ROOT, "semantic_slam_b200"), "-lssb",
           "-Wl,-rpath," + os.path.join(ROOT, "semantic_slam_b200"), "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
    subprocess.check_call(cmd)
    return exe


@pytest.mark.parametrize("eigen", [False, True])
def test_facade_compiles_and_links(tmp_path, eigen):
    exe = _build(tmp_path, eigen)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "sm_100a" in out.stdout, out.stdout + out.stderr


def test_mock_eigen_rejects_what_eigen_rejects(tmp_path):
    """the mock must refuse the one-argument size constructor of a dynamic matrix (`MatrixXd M(3)` compiled
    with the POD twin only), otherwise the Eigen leg above proves nothing"""
    src = tmp_path / "bad.cpp"
    src.write_text("#include <Eigen/Core>\nint main() { Eigen::MatrixXd M(3); return M.rows(); }\n")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "tests", "mock_eigen"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "one-argument size constructor" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("eigen", [False, True])
def test_facade_runs_on_gpu(tmp_path, eigen):
    exe = _build(tmp_path, eigen)
    out = subprocess.run([exe, "run"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert ("Eigen call syntax" if eigen else "POD twins") in out.stdout
