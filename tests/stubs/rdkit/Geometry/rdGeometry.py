"""rdkit.Geometry.rdGeometry import stub (see tests/stubs/rdkit/__init__.py)."""
from .. import _Missing

Point3D = _Missing("rdkit.Geometry.rdGeometry.Point3D")
