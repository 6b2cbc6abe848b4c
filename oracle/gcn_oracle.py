"""CPU oracle for the OpenChem GraphCNN hot path (TEST INFRASTRUCTURE ONLY).

This file is a numpy restatement of the reference algorithm.  It is *not* a
product path: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product (``openchem_b200``) never falls back to it; it raises when the CUDA
library is missing.

Parity pinning: the reference ships NO golden vectors / known-answer tests for
this path (only constructor smoke tests, openchem/tests/module_instantiation_test.py:36-43).
The oracle is therefore pinned against outputs of the reference itself, run in
the authoring container by ``oracle/make_golden.py`` (imports the unmodified
reference from /root/reference, asserts this restatement agrees with reference
forward + autograd in fp64 to <=1e-10 and writes ``tests/golden/*.npz``).

Reference lines each function follows (paths relative to /root/reference):
  graph_conv_forward   openchem/layers/gcn.py:33-42  (+ torch.nn.BatchNorm1d semantics, gcn.py:25)
  neighbour_max_pool   openchem/modules/encoders/gcn_encoder.py:44-50
  encoder_forward      openchem/modules/encoders/gcn_encoder.py:38-53
  *_backward           autograd of the above (closed forms)

All functions are dtype-generic: pass float32 arrays for the "fp32 reference",
float64 arrays for the "fp64-promoted reference".
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

EPS = 1e-5
MOMENTUM = 0.1

# Optional plain-C scan of the max-pool (oracle/pool_scan.c: same arithmetic and scan order, OpenMP over molecules),
# built by __graft_entry__.build() into oracle/_build/. Used only for large batches; the numpy loops below are the
# definition and tests/test_oracle.py checks the two agree bit for bit.
_POOL_LIB = None
POOL_C_MIN_ELEMS = 1 << 22      # B*N*D above which the C scan is used (when the library is built)


def pool_lib():
    global _POOL_LIB
    if _POOL_LIB is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libpool_scan.so")
        _POOL_LIB = ctypes.CDLL(path) if os.path.exists(path) else False
    return _POOL_LIB or None


def _pool_c(name, T, adj, arg=None):
    lib = pool_lib()
    dt = T.dtype
    if lib is None or dt not in (np.float64, np.float32) or adj.dtype != dt:
        return None
    B, N, D = T.shape
    T = np.ascontiguousarray(T)
    adj = np.ascontiguousarray(adj)
    suffix = "f64" if dt == np.float64 else "f32"
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    dims = (ctypes.c_long(B), ctypes.c_long(N), ctypes.c_long(D))
    if name == "fwd":
        P = np.empty_like(T)
        arg = np.empty((B, N, D), dtype=np.int64)
        getattr(lib, "pool_fwd_" + suffix)(ptr(T), ptr(adj), *dims, ptr(P), ptr(arg))
        return P, arg
    arg = np.ascontiguousarray(arg, dtype=np.int64)
    dT = np.empty_like(T)
    getattr(lib, "pool_bwd_" + suffix)(ptr(T), ptr(adj), ptr(arg), *dims, ptr(dT))
    return dT


# --------------------------------------------------------------------------- #
# GraphConvolution  (openchem/layers/gcn.py:33-42)
# --------------------------------------------------------------------------- #
def graph_conv_forward(x, adj, W, b, gamma, beta, running_mean, running_var,
                       training=True, eps=EPS, momentum=MOMENTUM):
    """y = BatchNorm1d_{over B*N rows}((adj @ x) @ W + b).

    x (B,N,Fin), adj (B,N,N), W (Fin,Fout) [NOT nn.Linear layout], b (Fout,) or None.
    Returns (y, cache, new_running_mean, new_running_var).
    gcn.py:34 bmm -> :35 mm -> :37-38 +bias -> :39-41 BN over the (B,Fout,N) transpose,
    i.e. per-channel statistics over all B*N rows, padded rows included.
    """
    B, N, _ = x.shape
    S = np.matmul(adj, x)                       # gcn.py:34
    Z = np.matmul(S.reshape(B * N, -1), W)      # gcn.py:35
    if b is not None:
        Z = Z + b                               # gcn.py:37-38
    n = B * N
    if training:
        mu = Z.mean(axis=0)
        var = ((Z - mu) ** 2).mean(axis=0)      # biased, used for normalisation
        unbiased = var * (n / max(n - 1, 1))
        new_rm = (1 - momentum) * running_mean + momentum * mu
        new_rv = (1 - momentum) * running_var + momentum * unbiased
    else:
        mu, var = running_mean, running_var
        new_rm, new_rv = running_mean, running_var
    invstd = 1.0 / np.sqrt(var + np.asarray(eps, dtype=Z.dtype))
    Yhat = (Z - mu) * invstd
    Y = Yhat * gamma + beta                     # gcn.py:40
    cache = dict(x=x, adj=adj, W=W, S=S, Yhat=Yhat, invstd=invstd, gamma=gamma,
                 training=training, has_bias=b is not None)
    return Y.reshape(B, N, -1), cache, new_rm, new_rv


def graph_conv_backward(dY, cache, need_dx=True):
    """Gradients of graph_conv_forward w.r.t. x, W, b, gamma, beta."""
    x, adj, W, S = cache["x"], cache["adj"], cache["W"], cache["S"]
    Yhat, invstd, gamma = cache["Yhat"], cache["invstd"], cache["gamma"]
    B, N, _ = x.shape
    n = B * N
    dY2 = dY.reshape(n, -1)
    dbeta = dY2.sum(axis=0)
    dgamma = (dY2 * Yhat).sum(axis=0)
    if cache["training"]:
        dZ = (gamma * invstd) * (dY2 - dbeta / n - Yhat * (dgamma / n))
    else:
        dZ = (gamma * invstd) * dY2
    db = dZ.sum(axis=0) if cache["has_bias"] else None
    dW = np.matmul(S.reshape(n, -1).T, dZ)
    dx = None
    if need_dx:
        dS = np.matmul(dZ, W.T).reshape(B, N, -1)
        dx = np.matmul(np.swapaxes(adj, 1, 2), dS)
    return dx, dW, db, dgamma, dbeta


# --------------------------------------------------------------------------- #
# Neighbour max-pool  (gcn_encoder.py:44-50)
# --------------------------------------------------------------------------- #
def neighbour_max_pool(T, adj, use_c=None):
    """P[b,i,c] = max_j adj[b,i,j] * T[b,j,c]; returns (P, first-argmax j*).

    The reference materialises a (B,N,N,D) tensor (x.repeat -> view -> * adj.expand
    -> max(dim=2)); torch's CPU max keeps the FIRST index on ties (strict '>').
    We scan j ascending with a strict '>' instead of materialising.
    """
    B, N, D = T.shape
    if use_c is None:
        use_c = B * N * D >= POOL_C_MIN_ELEMS
    if use_c:
        r = _pool_c("fwd", T, adj)
        if r is not None:
            return r
    best = adj[:, :, 0, None] * T[:, 0, None, :]
    arg = np.zeros((B, N, D), dtype=np.int64)
    for j in range(1, N):
        cand = adj[:, :, j, None] * T[:, j, None, :]
        upd = cand > best
        best = np.where(upd, cand, best)
        arg[upd] = j
    return best, arg


def neighbour_max_pool_backward(dP, adj, arg, use_c=None):
    """dT[b,j,c] = sum_{i: j*[b,i,c]==j} adj[b,i,j] * dP[b,i,c]."""
    B, N, D = dP.shape
    if use_c is None:
        use_c = B * N * D >= POOL_C_MIN_ELEMS
    if use_c:
        r = _pool_c("bwd", dP, adj, arg)
        if r is not None:
            return r
    dT = np.zeros_like(dP)
    for j in range(N):
        sel = (arg == j)
        dT[:, j, :] = (sel * adj[:, :, j, None] * dP).sum(axis=1)
    return dT


# --------------------------------------------------------------------------- #
# GraphCNNEncoder  (gcn_encoder.py:38-53)
# --------------------------------------------------------------------------- #
def encoder_forward(x, adj, params, training=True, eps=EPS, momentum=MOMENTUM):
    """params: dict with lists W,b,gamma,beta,running_mean,running_var (len L) and Wd (E,F_L), bd (E,).

    Returns (out (B,E), cache, new_running_means, new_running_vars).
    """
    L = len(params["W"])
    H = x
    layer_caches, new_rms, new_rvs = [], [], []
    for l in range(L):
        b = params["b"][l] if params.get("b") is not None else None
        Y, c, rm, rv = graph_conv_forward(H, adj, params["W"][l], b, params["gamma"][l], params["beta"][l],
                                          params["running_mean"][l], params["running_var"][l],
                                          training, eps, momentum)           # gcn_encoder.py:42
        T = np.tanh(Y)                                                       # :43
        P, arg = neighbour_max_pool(T, adj)                                  # :44-50
        c.update(T=T, arg=arg)
        layer_caches.append(c)
        new_rms.append(rm)
        new_rvs.append(rv)
        H = P
    Wd, bd = params["Wd"], params["bd"]
    D = np.tanh(np.matmul(H, Wd.T) + bd)                                     # :51
    R = D.sum(axis=1)                                                        # :52 (UNMASKED: all N rows)
    O = np.tanh(R)
    cache = dict(layers=layer_caches, H_L=H, D=D, O=O, adj=adj, Wd=Wd)
    return O, cache, new_rms, new_rvs


def encoder_backward(dO, cache, arg_override=None):
    """Parameter gradients of encoder_forward (x and adj are data).

    arg_override (list of (B,N,F_l) int arrays, optional): route the pool gradient through these arg-max indices
    instead of the oracle's own. Used by the GPU parity tests when the implementation under test resolved a
    NEAR-TIE (top-2 candidates closer than fp32 resolution) differently from the fp64 oracle: the routing is
    then checked separately (must be a near-tie) and everything else at the full tolerance."""
    adj, Wd, D, O, H_L = cache["adj"], cache["Wd"], cache["D"], cache["O"], cache["H_L"]
    B, N, E = D.shape
    dR = dO * (1 - O * O)
    dU = dR[:, None, :] * (1 - D * D)
    dU2 = dU.reshape(B * N, E)
    grads = dict(dWd=np.matmul(dU2.T, H_L.reshape(B * N, -1)), dbd=dU2.sum(axis=0),
                 dW=[], db=[], dgamma=[], dbeta=[])
    dH = np.matmul(dU, Wd)
    L = len(cache["layers"])
    for l in reversed(range(L)):
        c = cache["layers"][l]
        dT = neighbour_max_pool_backward(dH, adj, c["arg"] if arg_override is None else arg_override[l])
        dY = dT * (1 - c["T"] * c["T"])
        dH, dW, db, dgamma, dbeta = graph_conv_backward(dY, c, need_dx=(l > 0))
        grads["dW"].insert(0, dW)
        grads["db"].insert(0, db)
        grads["dgamma"].insert(0, dgamma)
        grads["dbeta"].insert(0, dbeta)
    return grads


# --------------------------------------------------------------------------- #
# helpers shared by tests / golden generation
# --------------------------------------------------------------------------- #
def rel_l2(a, b):
    """||a-b||_2 / ||b||_2 (b = reference); the per-tensor parity metric used everywhere."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


def cast_params(params, dtype):
    out = {}
    for k, v in params.items():
        if isinstance(v, (list, tuple)):
            out[k] = [None if a is None else np.asarray(a, dtype=dtype) for a in v]
        elif v is None:
            out[k] = None
        else:
            out[k] = np.asarray(v, dtype=dtype)
    return out
