"""CPU oracle for GraphEdgeAttentionEncoder (TEST INFRASTRUCTURE ONLY; same rules as oracle/gcn_oracle.py).

numpy restatement of openchem/modules/encoders/edge_attention_encoder.py:44-67 and of its autograd, built on the pinned
building blocks of gcn_oracle.py (graph_conv_forward/backward = gcn.py:33-42, neighbour_max_pool = the x.repeat/view/expand/max
block). Pinned by oracle/make_golden_edge.py against the unmodified reference run in fp64 (tests/golden/edge_*.npz,
re-checked by tests/test_oracle.py).

params: dict(W, b, gamma, beta, running_mean, running_var: lists of n_layers*A arrays in `graph_conv` order (layer-major,
edge channel minor, :28-30), Wd (E,F_L), bd (E,)).
"""
from __future__ import annotations

import numpy as np

from . import gcn_oracle as orc


def softmax(a, axis):
    m = a.max(axis=axis, keepdims=True)
    e = np.exp(a - m)
    return e / e.sum(axis=axis, keepdims=True)


def forward(x, edge_attr, params, n_layers, training=True):
    """-> (out (B,E), cache, new_running_means, new_running_vars)."""
    A = edge_attr.shape[3]
    hard = np.where(edge_attr[:, :, :, 0] > 0, 1.0, 0.0).astype(x.dtype)         # :47-49
    cur, caches, rms, rvs = x, [], [], []
    for i in range(n_layers):
        acc = None
        for j in range(A):
            k = i * A + j
            soft = softmax(edge_attr[:, :, :, j], axis=2)                        # :52
            y, c, rm, rv = orc.graph_conv_forward(cur, soft, params["W"][k], params["b"][k], params["gamma"][k], params["beta"][k],
                                                  params["running_mean"][k], params["running_var"][k], training)   # :53
            T = np.tanh(y)                                                       # :54
            P, arg = orc.neighbour_max_pool(T, hard)                             # :55-61
            c.update(T=T, arg=arg)
            caches.append(c)
            rms.append(rm)
            rvs.append(rv)
            acc = P if acc is None else acc + P                                  # :62-63
        cur = acc
    D = np.tanh(np.matmul(cur, params["Wd"].T) + params["bd"])                    # :65
    O = np.tanh(D.sum(axis=1))                                                   # :66
    return O, dict(layers=caches, H_L=cur, D=D, O=O, hard=hard, A=A, n_layers=n_layers, Wd=params["Wd"]), rms, rvs


def backward(dO, cache):
    """Parameter gradients (x and edge_attr are data)."""
    A, L, hard = cache["A"], cache["n_layers"], cache["hard"]
    D, O, H_L, Wd = cache["D"], cache["O"], cache["H_L"], cache["Wd"]
    B, N, E = D.shape
    dU = (dO * (1 - O * O))[:, None, :] * (1 - D * D)
    dU2 = dU.reshape(B * N, E)
    n = L * A
    g = dict(dWd=np.matmul(dU2.T, H_L.reshape(B * N, -1)), dbd=dU2.sum(axis=0), dW=[None] * n, db=[None] * n, dgamma=[None] * n, dbeta=[None] * n)
    dcur = np.matmul(dU, Wd)
    for i in reversed(range(L)):
        dprev = None
        for j in range(A):
            k = i * A + j
            c = cache["layers"][k]
            dT = orc.neighbour_max_pool_backward(dcur, hard, c["arg"])
            dY = dT * (1 - c["T"] * c["T"])
            dx, g["dW"][k], g["db"][k], g["dgamma"][k], g["dbeta"][k] = orc.graph_conv_backward(dY, c, need_dx=(i > 0))
            if i > 0:
                dprev = dx if dprev is None else dprev + dx
        dcur = dprev
    return g
