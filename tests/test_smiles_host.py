"""RDKit-free SMILES topology featuriser (openchem_b200/smiles.py): layout contract of the reference's graph data layer."""
import numpy as np

from openchem_b200.smiles import LOGP_ATTRIBUTES, atom_attributes, parse_smiles, smiles_batch, smiles_to_graph


def test_topology_of_known_molecules():
    atoms, bonds = parse_smiles("CCC1(c2ccccc2)C(=O)NC(=O)NC1=O")           # phenobarbital-like, logP_labels.csv row 4
    assert len(atoms) == 17 and len(bonds) == 18                             # 17 heavy atoms, 16 + 2 ring closures
    assert [a["z"] for a in atoms[:4]] == [6, 6, 6, 6] and sum(a["aromatic"] for a in atoms) == 6
    atoms, bonds = parse_smiles("C[N+](C)(C)CC(=O)[O-]")                      # charges in brackets
    assert [a["charge"] for a in atoms] == [0, 1, 0, 0, 0, 0, 0, -1]
    atoms, bonds = parse_smiles("ClC1=CC=C(Br)C=C1.[Na+]")                    # two-letter elements, dot, ring closure with '='
    assert [a["z"] for a in atoms] == [17, 6, 6, 6, 6, 35, 6, 6, 11] and len(bonds) == 8
    att = atom_attributes(*parse_smiles("CC(=O)O"))                           # acetic acid: valences 4,4,2,2; sp3, sp2, sp2, sp3
    assert [a["valence"] for a in att] == [4, 4, 2, 2] and [a["hybridization"] for a in att] == [4, 3, 3, 4]
    att = atom_attributes(*parse_smiles("c1ccccc1C#N"))                       # aromatic carbons valence 4, nitrile sp
    assert all(a["valence"] == 4 for a in att[:7]) and att[6]["hybridization"] == 2 and att[7]["hybridization"] == 2


def test_graph_layout_matches_the_data_layer_contract():
    x, adj, n = smiles_to_graph("CC1CC2C3CCC4=CC(=O)C=CC4(C)C3(F)C(O)CC2(C)C1(O)C(=O)CO", 40)      # logP_labels.csv row 1
    assert x.shape == (40, 33) and adj.shape == (40, 40) and x.dtype == np.float32 and adj.dtype == np.float32
    assert sum(len(v) for _, v in LOGP_ATTRIBUTES) == 33
    assert set(np.unique(x)) <= {0.0, 1.0} and set(np.unique(adj)) <= {0.0, 1.0}
    assert np.array_equal(adj, adj.T) and np.array_equal(np.diag(adj)[:n], np.ones(n)) and adj[n:].sum() == 0 and adj[:, n:].sum() == 0
    assert np.all(x[:n].sum(axis=1) == 5) and x[n:].sum() == 0               # one entry per attribute group, zero padding rows
    assert adj.sum(axis=1).max() <= 5                                        # <= 4 bonds + self loop
    g = smiles_batch(["CCO", "not a smiles", "C" * 50, "c1ccccc1"], 10)
    assert list(g["kept"]) == [0, 3] and g["x"].shape == (2, 10, 33)
