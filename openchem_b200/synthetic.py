"""Synthetic padded molecular graphs in the layout OpenChem's graph data layer produces.

Layout contract (reference: openchem/utils/graph.py:77-86,133-152 and
openchem/data/graph_data_layer.py:116-122): per molecule a dense fp32 ``adj_matrix``
(N_max, N_max) that is 0/1, symmetric, WITH self loops and zero padded, and a
``node_feature_matrix`` (N_max, F0) of concatenated one-hot groups, zero rows for padding.

Generator: identity + random spanning tree (node k bonds to a uniformly
chosen earlier node of degree < 4) + floor(n/6) ring-closing bonds between degree < 4 nodes.
Pure numpy so the very same arrays feed the oracle (CPU) and the CUDA path.
"""
from __future__ import annotations

import numpy as np

# one-hot group sizes: logP featuriser (example_configs/logp_gcnn_config.py:39-48) and our 64-wide choice
ONE_HOT_GROUPS = {33: (6, 6, 8, 2, 11), 64: (8, 8, 16, 8, 24)}


def _groups_for(f0):
    if f0 in ONE_HOT_GROUPS:
        return ONE_HOT_GROUPS[f0]
    groups, left = [], f0
    while left > 0:
        g = min(left, 6)
        groups.append(g)
        left -= g
    return tuple(groups)


def make_graph_batch(batch, n_max, f0, occupancy="molecular", seed=1234, max_degree=4):
    """Returns dict(x (B,N,F0) fp32, adj (B,N,N) fp32, n_atoms (B,) int64).

    occupancy: "full" -> every molecule has n_max atoms; "molecular" -> n ~ U{ceil(N/4)..N}.
    """
    rng = np.random.RandomState(seed)
    B, N = int(batch), int(n_max)
    if occupancy == "full":
        n_atoms = np.full(B, N, dtype=np.int64)
    elif occupancy == "molecular":
        n_atoms = rng.randint((N + 3) // 4, N + 1, size=B).astype(np.int64)
    else:
        raise ValueError("occupancy must be 'full' or 'molecular'")
    adj = np.zeros((B, N, N), dtype=np.float32)
    deg = np.zeros((B, N), dtype=np.int64)
    bidx = np.arange(B)
    idx = np.arange(N)
    adj[:, idx, idx] = (idx[None, :] < n_atoms[:, None]).astype(np.float32)      # self loops (np.eye in graph.py:80)
    for k in range(1, N):
        alive = k < n_atoms
        if not alive.any():
            break
        score = rng.rand(B, k)
        score[deg[:, :k] >= max_degree] = -1.0
        parent = score.argmax(axis=1)
        ok = alive & (score[bidx, parent] >= 0)
        b = bidx[ok]
        p = parent[ok]
        adj[b, k, p] = 1.0
        adj[b, p, k] = 1.0
        deg[b, k] += 1
        deg[b, p] += 1
    n_rings = int(n_atoms.max()) // 6
    for r in range(n_rings):
        want = (n_atoms // 6) > r
        u = (rng.rand(B) * n_atoms).astype(np.int64)
        v = (rng.rand(B) * n_atoms).astype(np.int64)
        ok = want & (u != v) & (deg[bidx, u] < max_degree) & (deg[bidx, v] < max_degree) & (adj[bidx, u, v] == 0)
        b = bidx[ok]
        adj[b, u[ok], v[ok]] = 1.0
        adj[b, v[ok], u[ok]] = 1.0
        deg[b, u[ok]] += 1
        deg[b, v[ok]] += 1
    x = np.zeros((B, N, f0), dtype=np.float32)
    off = 0
    valid = idx[None, :] < n_atoms[:, None]
    bb, nn_ = np.nonzero(valid)
    for g in _groups_for(f0):
        cat = rng.randint(0, g, size=bb.shape[0])
        x[bb, nn_, off + cat] = 1.0
        off += g
    return dict(x=x, adj=adj, n_atoms=n_atoms)


def make_labels(batch, n_tasks=1, kind="regression", seed=4321, missing=0.15):
    """Regression: randn (B,T). 'multitask': 0/1 labels with `missing` fraction set to 9 (ignore_index)."""
    rng = np.random.RandomState(seed)
    if kind == "regression":
        return rng.randn(batch, n_tasks).astype(np.float32)
    y = (rng.rand(batch, n_tasks) < 0.3).astype(np.float32)
    y[rng.rand(batch, n_tasks) < missing] = 9.0
    return y


def random_params(widths, encoder_dim, seed):
    """Reference-style initialisation (U(+-1/sqrt(F_out)), gcn.py:27-31; nn.Linear-like dense) as numpy arrays, with the
    BatchNorm affine parameters and running statistics randomised so parity tests exercise them."""
    rng = np.random.RandomState(seed)
    p = dict(W=[], b=[], gamma=[], beta=[], running_mean=[], running_var=[])
    for l in range(len(widths) - 1):
        fin, fout = widths[l], widths[l + 1]
        s = 1.0 / np.sqrt(fout)
        p["W"].append(rng.uniform(-s, s, size=(fin, fout)).astype(np.float32))
        p["b"].append(rng.uniform(-s, s, size=(fout,)).astype(np.float32))
        p["gamma"].append(rng.uniform(0.5, 1.5, size=(fout,)).astype(np.float32))
        p["beta"].append(rng.uniform(-0.3, 0.3, size=(fout,)).astype(np.float32))
        p["running_mean"].append(rng.uniform(-0.2, 0.2, size=(fout,)).astype(np.float32))
        p["running_var"].append(rng.uniform(0.5, 1.5, size=(fout,)).astype(np.float32))
    s = 1.0 / np.sqrt(widths[-1])
    p["Wd"] = rng.uniform(-s, s, size=(encoder_dim, widths[-1])).astype(np.float32)
    p["bd"] = rng.uniform(-s, s, size=(encoder_dim,)).astype(np.float32)
    return p


# Named workloads of BASELINE.json `configs`: per-GPU batch, N, F0, hidden, E, tasks
WORKLOADS = {
    "c1_logp":  dict(batch=256, n_max=85, f0=33, hidden=[128] * 5, encoder_dim=128, mlp_hidden=[128, 1], task="regression"),
    "c2_tox21": dict(batch=256, n_max=50, f0=33, hidden=[128] * 5, encoder_dim=128, mlp_hidden=[128, 12], task="multitask"),
    "c3_synth": dict(batch=4096, n_max=64, f0=64, hidden=[128] * 4, encoder_dim=128, mlp_hidden=[128, 1], task="regression"),
    "c4_ddp":   dict(batch=2048, n_max=64, f0=64, hidden=[128] * 4, encoder_dim=128, mlp_hidden=[128, 1], task="regression"),
    "c5_large": dict(batch=1024, n_max=128, f0=64, hidden=[256] * 6, encoder_dim=256, mlp_hidden=[256, 1], task="regression"),
}


class SyntheticGraphDataset:
    """A map-style data set with the sample layout of the reference's GraphDataset.__getitem__
    (openchem/data/graph_data_layer.py:116-156): dicts of fp32 numpy arrays `adj_matrix` (N,N), `node_feature_matrix` (N,F0)
    and `labels` (T,), which the default collate of run.py's DataLoader batches and `Graph2Label.cast_inputs`
    (Graph2Label.py:39-64) uploads. Used by configs/*.py so the unmodified run.py / launch.py can be driven without RDKit."""

    def __init__(self, size, n_max, f0, n_tasks=1, kind="regression", occupancy="molecular", seed=1234):
        g = make_graph_batch(size, n_max, f0, occupancy=occupancy, seed=seed)
        self.adj, self.x, self.n_atoms = g["adj"], g["x"], g["n_atoms"]
        if kind == "regression":     # a learnable target: a fixed random linear readout of the atom-type counts, plus noise
            rng = np.random.RandomState(seed + 1)
            w = rng.randn(f0, n_tasks).astype(np.float32)
            y = self.x.sum(axis=1) @ w / np.sqrt(n_max)
            self.target = (y + 0.05 * rng.randn(size, n_tasks)).astype(np.float32)
        else:
            self.target = make_labels(size, n_tasks, kind=kind, seed=seed + 1)
        self.max_size = int(n_max)

    def __len__(self):
        return self.adj.shape[0]

    def __getitem__(self, index):
        return {"adj_matrix": self.adj[index], "node_feature_matrix": self.x[index], "labels": self.target[index]}
