"""semantic_slam_b200 — B200-native (sm_100a) backend for the two numeric hot paths of
hridaybavle/semantic_slam: the ps_graph_slam Levenberg-Marquardt optimiser and the
planar_segmentation RANSAC plane fit.  The product is the C-ABI shared library
``semantic_slam_b200/libssb.so`` (include/ssb.h); this package is the thin Python mirror of the
reference's C++ call surface used by the tests and the benchmark.  There is no CPU fallback:
importing the bindings fails loudly when the CUDA extension is missing.
"""
from ._lib import lib, build, SsbError  # noqa: F401
from .graph_slam import GraphSLAM  # noqa: F401
from .segmentation import PlaneSegmentation, OrganizedSegmentation, PlaneClustering, CloudLayout, pcl_sample_stream  # noqa: F401
from .association import DataAssociation  # noqa: F401
from .semantic_graph_slam import SemanticGraphSLAM  # noqa: F401
