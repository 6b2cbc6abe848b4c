"""rdkit.Chem import stub (see tests/stubs/rdkit/__init__.py)."""
from .. import _Missing

QED = _Missing("rdkit.Chem.QED")
Descriptors = _Missing("rdkit.Chem.Descriptors")
rdMolDescriptors = _Missing("rdkit.Chem.rdMolDescriptors")
AllChem = _Missing("rdkit.Chem.AllChem")


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Missing("rdkit.Chem." + name)
