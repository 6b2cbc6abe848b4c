"""Namespace-package shadow of openchem.layers.gcn (drop-in route 2)."""
from openchem_b200.layers.gcn import GraphConvolution  # noqa: F401
