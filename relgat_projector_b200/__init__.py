"""relgat_projector_b200 — B200-native (sm_100a) RelGAT message-passing hot path.

Drop-in replacements for the reference's ``RelGATLayer``, ``DistMultScorer`` / ``TransEScorer``
and ``RelGATModel`` (radlab-dev-group/relgat-projector, ``relgat_projector/core``), running on
hand-written CUDA kernels behind the C ABI of ``include/relgat_b200.h``.
"""
from .layer import RelGATLayer
from .scorer import DistMultScorer, TransEScorer
from .projection import ProjectionHead
from .model import RelGATModel
from .graph import GraphIndex, get_graph_index

__all__ = ["RelGATLayer", "DistMultScorer", "TransEScorer", "ProjectionHead", "RelGATModel", "GraphIndex",
           "get_graph_index"]
__version__ = "0.1.0"
