"""CPU torch restatement of the reference's eager op sequence (TEST / BASELINE INFRASTRUCTURE ONLY).

Purpose: the `cpu_baseline` and `--impl reference` legs of bench.py need "the reference's CPU path" on the GPU
box, where /root/reference does not exist. This module replays the SAME ATen op sequence the reference modules
issue (bmm -> mm -> +bias -> transpose/contiguous -> batch_norm -> transpose -> tanh -> repeat/view/expand/mul/max
-> ... -> addmm/tanh/sum/tanh), including the materialised (B,N,N,D) pooling temporaries that dominate its run
time, and lets autograd produce the backward exactly as the reference does. It is validated against the real
reference modules by tests/test_oracle.py::test_torch_restatement_* via the golden fixtures.

Follows: openchem/layers/gcn.py:33-42, openchem/modules/encoders/gcn_encoder.py:38-53,
openchem/modules/mlp/openchem_mlp.py:51-61 (head used by example_configs/logp_gcnn_config.py:111-116).
Never imported by the product package.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def graph_conv(x, adj, W, b, gamma, beta, running_mean, running_var, training, eps=1e-5, momentum=0.1):
    n_atoms = adj.shape[1]
    agg = torch.bmm(adj, x)                                              # gcn.py:34
    z = torch.mm(agg.view(-1, W.shape[0]), W).view(-1, n_atoms, W.shape[1])   # gcn.py:35-36
    if b is not None:
        z = z + b                                                        # gcn.py:37-38
    z = z.transpose(1, 2).contiguous()                                   # gcn.py:39
    z = F.batch_norm(z, running_mean, running_var, gamma, beta, training, momentum, eps)   # gcn.py:40
    return z.transpose(1, 2)                                             # gcn.py:41


def neighbour_pool_materialised(h, adj):
    n, d = adj.size(1), h.size(-1)
    tiled = h.repeat(1, n, 1).view(-1, n, n, d)                           # gcn_encoder.py:48
    masked = tiled * adj.unsqueeze(3).expand(-1, n, n, d)                 # gcn_encoder.py:46-49
    return masked.max(dim=2)[0]                                           # gcn_encoder.py:50


def encoder(x, adj, p, training):
    """p: dict of torch tensors W[l], b[l], gamma[l], beta[l], running_mean[l], running_var[l], Wd, bd."""
    h = x
    for l in range(len(p["W"])):
        h = graph_conv(h, adj, p["W"][l], p["b"][l], p["gamma"][l], p["beta"][l], p["running_mean"][l], p["running_var"][l], training)
        h = neighbour_pool_materialised(torch.tanh(h), adj)               # gcn_encoder.py:43-50
    h = torch.tanh(F.linear(h, p["Wd"], p["bd"]))                         # gcn_encoder.py:51
    return torch.tanh(h.sum(dim=1))                                       # gcn_encoder.py:52-53


def mlp_head(h, hp, training):
    """OpenChemMLP with n_layers=2, activation [relu, identity], dropout 0 (openchem_mlp.py:51-61)."""
    h = F.linear(h, hp["W1"], hp["b1"])
    h = F.batch_norm(h, hp["rm"], hp["rv"], hp["g"], hp["be"], training, 0.1, 1e-5)
    h = F.relu(h)
    return F.linear(h, hp["W2"], hp["b2"])


def params_to_torch(p_np, dtype=torch.float32, requires_grad=True):
    out = {}
    for k, v in p_np.items():
        if isinstance(v, (list, tuple)):
            out[k] = [None if a is None else torch.tensor(a, dtype=dtype) for a in v]
        else:
            out[k] = torch.tensor(v, dtype=dtype)
    if requires_grad:
        for k in ("W", "b", "gamma", "beta"):
            for t in out[k]:
                if t is not None:
                    t.requires_grad_(True)
        out["Wd"].requires_grad_(True)
        out["bd"].requires_grad_(True)
    return out
