"""The C-ABI library loads without a GPU and exports every symbol include/ssb.h declares; host-only
entry points behave; GPU entry points fail loudly (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "ssb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ssb_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from semantic_slam_b200 import _lib
    L = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 40
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(set(_lib.SYMBOLS)) == declared, set(declared) ^ set(_lib.SYMBOLS)
    assert b"sm_100a" in L.ssb_build_info()


def test_struct_layouts_match_python_mirrors():
    import ctypes as C
    from semantic_slam_b200 import _lib
    from semantic_slam_b200.segmentation import PLANE_RESULT_DTYPE
    import oracle
    assert PLANE_RESULT_DTYPE.itemsize == 72 and oracle.PLANE_RESULT_DTYPE == PLANE_RESULT_DTYPE
    assert C.sizeof(_lib.CloudLayoutC) == 32 and C.sizeof(_lib.RansacOpts) == 48
    from semantic_slam_b200.segmentation import PLANE_CLUSTER_DTYPE
    assert PLANE_CLUSTER_DTYPE.itemsize == 56 and oracle.PLANE_CLUSTER_DTYPE == PLANE_CLUSTER_DTYPE
    co = _lib.ClusterOpts()
    _lib.lib().ssb_cluster_default_opts(C.byref(co))
    assert (co.num_centroids_normals, co.num_centroids_distance, co.kmeans_attempts, co.kmeans_max_count) == (4, 2, 10, 10)
    assert co.kmeans_epsilon == 0.01 and co.min_cluster_points == 500 and co.ransac_seed == 12345
    o = _lib.GraphOpts()
    _lib.lib().ssb_graph_default_opts(C.byref(o))
    assert o.max_pcg_iters == 20000 and o.pcg_tol == 1e-8 and o.device == -1


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from semantic_slam_b200 import GraphSLAM, PlaneSegmentation, SsbError
    with pytest.raises(SsbError):
        GraphSLAM()
    with pytest.raises(SsbError):
        PlaneSegmentation()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "semantic_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f
