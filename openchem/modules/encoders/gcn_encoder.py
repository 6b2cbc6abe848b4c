"""Namespace-package shadow of openchem.modules.encoders.gcn_encoder (drop-in route 2)."""
from openchem_b200.modules.encoders.gcn_encoder import GraphCNNEncoder  # noqa: F401
