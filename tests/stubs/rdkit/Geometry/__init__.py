"""rdkit.Geometry import stub (see tests/stubs/rdkit/__init__.py)."""
