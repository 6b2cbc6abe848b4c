"""TEST FIXTURE, not product code: an import stub for RDKit.

RDKit (un-vendored, unpinned third-party dependency of the reference: README.md:57) is absent from this image, and the
reference's run.py imports it at module scope through openchem.data.utils / openchem.utils.metrics (run.py:23,28-29) even when
the config's data layer never parses a SMILES string. The drop-in tests run the UNMODIFIED run.py with a synthetic graph
data layer, so only the imports have to resolve; any attempt to actually use RDKit raises.
"""


class _Missing:
    def __init__(self, name):
        self._name = name

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return _Missing(self._name + "." + item)

    def __call__(self, *a, **k):
        raise RuntimeError("rdkit stub: %s was called, but RDKit is not installed in this image" % self._name)


class _RdBase:
    @staticmethod
    def DisableLog(_what):      # openchem/data/utils.py:20 calls this at import time
        return None


rdBase = _RdBase()
DataStructs = _Missing("rdkit.DataStructs")
RDConfig = _Missing("rdkit.RDConfig")
from . import Chem  # noqa: E402,F401
