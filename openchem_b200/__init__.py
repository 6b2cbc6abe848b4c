"""openchem_b200 — B200-native GraphCNN hot path for OpenChem (GraphConvolution stack + readout).

Public surface mirrors the reference modules on this path:
    openchem_b200.layers.gcn.GraphConvolution
    openchem_b200.modules.encoders.gcn_encoder.GraphCNNEncoder
The arithmetic lives in csrc/libocgcn.so (hand-written sm_100a CUDA behind the C ABI of include/ocgcn.h).
"""
from ._lib import build, load  # noqa: F401

__all__ = ["build", "load", "GraphConvolution", "GraphCNNEncoder"]


def __getattr__(name):
    if name == "GraphConvolution":
        from .layers.gcn import GraphConvolution
        return GraphConvolution
    if name == "GraphCNNEncoder":
        from .modules.encoders.gcn_encoder import GraphCNNEncoder
        return GraphCNNEncoder
    raise AttributeError(name)
