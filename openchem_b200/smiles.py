"""RDKit-free SMILES -> padded graph tensors in the layout of OpenChem's graph data layer (input path for tests / benchmarks).

The reference featurises with RDKit (`openchem/utils/graph.py:57-63,77-86`, user featuriser `example_configs/logp_gcnn_config.py:22-48`),
an un-vendored third-party dependency that is not part of this image. This module parses the SMILES TOPOLOGY itself (atoms in
SMILES order - RDKit's atom order for a parsed molecule -, bonds, branches, ring closures, bracket atoms with charge / hydrogen
count) and derives the logP config's five categorical atom attributes from it:

    valence (6) | formal charge (6) | hybridisation (8) | aromatic (2) | element (11)        -> 33 one-hot features, same order

Element, charge, aromaticity, bonds and hence the adjacency (bonds + self loops, `graph.py:80-84`) are exact for valid SMILES;
total valence and hybridisation are APPROXIMATIONS of RDKit's perception (default-valence hydrogen filling, bond-order counting;
no kekulisation, no sanitisation). The tensors have exactly the shapes, dtypes and value set ({0,1}) the reference produces, which is
what the hot path sees; featurisation parity with RDKit is not claimed.
"""
from __future__ import annotations

import re

import numpy as np

ELEMENTS = {5: 0, 7: 1, 6: 2, 8: 3, 9: 4, 15: 5, 16: 6, 17: 7, 35: 8, 53: 9}        # logp_gcnn_config.py:26-31 (else -> 10)
ATOMIC_NUM = {"H": 1, "B": 5, "C": 6, "N": 7, "O": 8, "F": 9, "Na": 11, "Si": 14, "P": 15, "S": 16, "Cl": 17, "K": 19, "Se": 34,
              "Br": 35, "I": 53, "Li": 3, "Mg": 12, "Al": 13, "Ca": 20, "Fe": 26, "Zn": 30, "As": 33, "Sn": 50, "Hg": 80, "Cu": 29}
DEFAULT_VALENCE = {5: (3,), 6: (4,), 7: (3, 5), 8: (2,), 9: (1,), 15: (3, 5), 16: (2, 4, 6), 17: (1,), 35: (1,), 53: (1,)}
_BRACKET = re.compile(r"^(\d+)?([A-Za-z][a-z]?)(@{0,2})(H(\d+)?)?((\+{1,4}|-{1,4})|([+-]\d+))?(:\d+)?$")
_BOND_ORDER = {"-": 1.0, "=": 2.0, "#": 3.0, ":": 1.5, "/": 1.0, "\\": 1.0}


def parse_smiles(smiles):
    """-> (atoms, bonds): atoms = list of dict(z, aromatic, charge, hcount (None = implicit)), bonds = list of (i, j, order)."""
    atoms, bonds, stack, ring = [], [], [], {}
    prev, pending = None, None
    i, n = 0, len(smiles)

    def add_atom(sym, aromatic, charge=0, hcount=None):
        nonlocal prev, pending
        z = ATOMIC_NUM.get(sym)
        if z is None:
            z = ATOMIC_NUM.get(sym.capitalize(), 0)
        atoms.append(dict(z=z, aromatic=aromatic, charge=charge, hcount=hcount))
        k = len(atoms) - 1
        if prev is not None:
            order = _BOND_ORDER[pending] if pending else (1.5 if aromatic and atoms[prev]["aromatic"] else 1.0)
            bonds.append((prev, k, order))
        prev, pending = k, None

    while i < n:
        c = smiles[i]
        if c == "[":
            j = smiles.index("]", i)
            m = _BRACKET.match(smiles[i + 1:j])
            if not m:
                raise ValueError("unsupported bracket atom %r in %r" % (smiles[i:j + 1], smiles))
            sym, hc, hn, chg = m.group(2), m.group(4), m.group(5), m.group(6)
            charge = 0
            if chg:
                charge = int(chg) if chg[-1].isdigit() else (len(chg) if chg[0] == "+" else -len(chg))
            add_atom(sym, sym[0].islower(), charge, (int(hn) if hn else 1) if hc else 0)
            i = j + 1
        elif c in "-=#:/\\":
            pending = c
            i += 1
        elif c == "(":
            stack.append(prev)
            i += 1
        elif c == ")":
            prev = stack.pop()
            i += 1
        elif c == ".":
            prev, pending = None, None
            i += 1
        elif c.isdigit() or c == "%":
            if c == "%":
                label, i = smiles[i + 1:i + 3], i + 3
            else:
                label, i = c, i + 1
            if label in ring:
                other, bsym = ring.pop(label)
                sym = pending or bsym
                order = _BOND_ORDER[sym] if sym else (1.5 if atoms[prev]["aromatic"] and atoms[other]["aromatic"] else 1.0)
                bonds.append((other, prev, order))
            else:
                ring[label] = (prev, pending)
            pending = None
        elif smiles.startswith(("Cl", "Br"), i):
            add_atom(smiles[i:i + 2], False)
            i += 2
        elif c in "BCNOPSFI":
            add_atom(c, False)
            i += 1
        elif c in "bcnops":
            add_atom(c, True)
            i += 1
        else:
            raise ValueError("unsupported SMILES character %r in %r" % (c, smiles))
    if ring or stack:
        raise ValueError("unbalanced ring closures / branches in %r" % smiles)
    return atoms, bonds


def atom_attributes(atoms, bonds):
    """Approximate (valence, charge, hybridisation code, aromatic, element class) per atom; see the module docstring."""
    nb = len(atoms)
    order_sum, n_double, n_triple, degree = [0.0] * nb, [0] * nb, [0] * nb, [0] * nb
    for a, b, o in bonds:
        for k in (a, b):
            order_sum[k] += 1.0 if o == 1.5 else o
            degree[k] += 1
            n_double[k] += o == 2.0
            n_triple[k] += o == 3.0
    out = []
    for k, at in enumerate(atoms):
        bo = order_sum[k] + (1.0 if at["aromatic"] and at["z"] in (6, 7) and degree[k] <= 2 + (at["z"] == 6) else 0.0)
        if at["hcount"] is None:
            targets = [v + at["charge"] * (1 if at["z"] in (7, 8, 15, 16) else -1) for v in DEFAULT_VALENCE.get(at["z"], (int(round(bo)),))]
            h = next((t - bo for t in targets if t >= bo), 0.0)
        else:
            h = float(at["hcount"])
        valence = int(round(bo + h))
        if degree[k] == 0 and h == 0:
            hyb = 1                                      # S (isolated atom / ion)
        elif n_triple[k] or n_double[k] >= 2:
            hyb = 2                                      # SP
        elif n_double[k] or at["aromatic"]:
            hyb = 3                                      # SP2
        else:
            hyb = 4                                      # SP3
        out.append(dict(valence=valence, charge=at["charge"], hybridization=hyb, aromatic=int(at["aromatic"]), atom_element=ELEMENTS.get(at["z"], 10)))
    return out


# attribute order and value lists of logp_gcnn_config.py:39-48 (dict insertion order = feature order)
LOGP_ATTRIBUTES = (("valence", [1, 2, 3, 4, 5, 6]), ("charge", [-1, 0, 1, 2, 3, 4]), ("hybridization", [0, 1, 2, 3, 4, 5, 6, 7]),
                   ("aromatic", [0, 1]), ("atom_element", list(range(11))))


def smiles_to_graph(smiles, n_max, attributes=LOGP_ATTRIBUTES):
    """-> (node_feature_matrix (n_max, F0) fp32, adj_matrix (n_max, n_max) fp32, n_atoms): one-hot attribute groups (a value outside
    an attribute's list leaves its group all-zero, as the reference's Attribute one-hot does), bonds + self loops, zero padding."""
    atoms, bonds = parse_smiles(smiles)
    nb = len(atoms)
    if nb > n_max:
        raise ValueError("molecule has %d atoms > n_max %d" % (nb, n_max))
    f0 = sum(len(v) for _, v in attributes)
    x = np.zeros((n_max, f0), dtype=np.float32)
    adj = np.zeros((n_max, n_max), dtype=np.float32)
    for k, at in enumerate(atom_attributes(atoms, bonds)):
        off = 0
        for name, values in attributes:
            if at[name] in values:
                x[k, off + values.index(at[name])] = 1.0
            off += len(values)
        adj[k, k] = 1.0
    for a, b, _ in bonds:
        adj[a, b] = adj[b, a] = 1.0
    return x, adj, nb


def smiles_batch(smiles_list, n_max, attributes=LOGP_ATTRIBUTES):
    """-> dict(x (B,N,F0), adj (B,N,N), n_atoms (B,), kept (indices of the molecules that parsed and fit))."""
    xs, adjs, ns, kept = [], [], [], []
    for idx, s in enumerate(smiles_list):
        try:
            x, adj, nb = smiles_to_graph(s, n_max, attributes)
        except ValueError:
            continue
        xs.append(x); adjs.append(adj); ns.append(nb); kept.append(idx)
    return dict(x=np.stack(xs), adj=np.stack(adjs), n_atoms=np.asarray(ns, dtype=np.int64), kept=np.asarray(kept, dtype=np.int64))


class SmilesGraphDataset:
    """Map-style data set over SMILES strings with the sample layout of the reference's GraphDataset.__getitem__
    (openchem/data/graph_data_layer.py:116-156), featurised ONCE at construction by this module (no RDKit): dicts of fp32
    `adj_matrix` (N,N), `node_feature_matrix` (N,33) and `labels` (T,). Molecules that do not parse or exceed n_max are dropped
    (the reference drops what RDKit cannot sanitise, data/utils.py:150-161). Used by configs/*.py with OCB200_DATA=logp_csv so the
    unmodified run.py trains on the reference's own logP_labels.csv."""

    def __init__(self, smiles, labels, n_max=None):
        labels = np.asarray(labels, dtype=np.float32).reshape(len(smiles), -1)
        if n_max is None:
            n_max = max(len(parse_smiles(s)[0]) for s in smiles)
        g = smiles_batch(list(smiles), int(n_max))
        self.x, self.adj, self.n_atoms = g["x"], g["adj"], g["n_atoms"]
        self.target = labels[g["kept"]]
        self.smiles = [smiles[i] for i in g["kept"]]
        self.max_size = int(n_max)

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, index):
        return {"adj_matrix": self.adj[index], "node_feature_matrix": self.x[index], "labels": self.target[index]}


def logp_csv_datasets(csv_path, test_size=0.2, random_state=42):
    """(train, test) SmilesGraphDataset of logP_labels.csv, split like example_configs/logp_gcnn_config.py:51-58 (sklearn
    train_test_split, 20 % test, random_state 42); both padded to the largest molecule of the file (N = 85)."""
    import csv
    from sklearn.model_selection import train_test_split
    rows = list(csv.reader(open(csv_path)))[1:]
    smiles, labels = [r[1] for r in rows], np.asarray([float(r[2]) for r in rows], dtype=np.float32).reshape(-1, 1)
    n_max = max(len(parse_smiles(s)[0]) for s in smiles)
    s_tr, s_te, y_tr, y_te = train_test_split(smiles, labels, test_size=test_size, random_state=random_state)
    return SmilesGraphDataset(s_tr, y_tr, n_max), SmilesGraphDataset(s_te, y_te, n_max)
