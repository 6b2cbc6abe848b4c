"""Namespace-package shadow of openchem.modules.encoders.edge_attention_encoder (drop-in route 2)."""
from openchem_b200.modules.encoders.edge_attention_encoder import GraphEdgeAttentionEncoder  # noqa: F401
