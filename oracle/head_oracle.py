"""CPU oracle for the MLP head + criteria of Graph2Label (TEST INFRASTRUCTURE ONLY; see oracle/gcn_oracle.py's header for
the rules: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it, the product never falls back to it).

Parity pinning: the reference ships no golden vectors for this path; `oracle/make_golden_head.py` runs the unmodified
`OpenChemMLP`, `nn.MSELoss` and `MultitaskLoss` from /root/reference (forward + autograd, fp32 and fp64), asserts this
restatement reproduces the fp64 run to <= 1e-10 and writes `tests/golden/head_*.npz`.

Reference lines (relative to /root/reference):
  mlp_forward      openchem/modules/mlp/openchem_mlp.py:51-61   (Dropout inactive) + torch.nn.BatchNorm1d / nn.Linear semantics
  mse_loss         torch.nn.MSELoss (reduction='mean'), the criterion of example_configs/logp_gcnn_config.py:91
  multitask_loss   openchem/criterion/multitask_loss.py:40-49
  *_backward       autograd of the above (closed forms)
dtype-generic: float32 arrays give the "fp32 reference", float64 the "fp64-promoted reference".
"""
from __future__ import annotations

import numpy as np

EPS = 1e-5
MOMENTUM = 0.1
ACTS = ("identity", "relu", "sigmoid", "tanh")


def act_forward(u, name):
    if name == "relu":
        return np.maximum(u, 0)
    if name == "sigmoid":
        return 1 / (1 + np.exp(-u))
    if name == "tanh":
        return np.tanh(u)
    return u


def act_grad_from_output(a, name):
    if name == "relu":
        return (a > 0).astype(a.dtype)
    if name == "sigmoid":
        return a * (1 - a)
    if name == "tanh":
        return 1 - a * a
    return np.ones_like(a)


def mlp_forward(inp, p, acts, training=True, eps=EPS, momentum=MOMENTUM):
    """openchem_mlp.py:51-61. p: dict W[i] (out,in), b[i], gamma[i], beta[i], running_mean[i], running_var[i] (i < L-1).
    Returns (pred, cache, new_running) — new_running = updated (running_mean, running_var) lists when training."""
    L = len(p["W"])
    a = inp
    cache = dict(A=[inp], U=[], yhat=[], invstd=[])
    rm, rv = [np.array(m, copy=True) for m in p["running_mean"]], [np.array(v, copy=True) for v in p["running_var"]]
    for i in range(L):
        u = a @ p["W"][i].T + p["b"][i]                       # nn.Linear (:54/:59)
        cache["U"].append(u)
        if i < L - 1:                                          # BatchNorm1d over the B rows (:55)
            if training:
                n = u.shape[0]
                mean = u.mean(axis=0)
                var = ((u - mean) ** 2).mean(axis=0)
                rm[i] = (1 - momentum) * rm[i] + momentum * mean
                rv[i] = (1 - momentum) * rv[i] + momentum * var * n / max(n - 1, 1)
            else:
                mean, var = p["running_mean"][i], p["running_var"][i]
            invstd = 1 / np.sqrt(var + eps)
            yhat = (u - mean) * invstd
            cache["yhat"].append(yhat)
            cache["invstd"].append(invstd)
            u = yhat * p["gamma"][i] + p["beta"][i]
        a = act_forward(u, acts[i])                            # (:56/:60)
        cache["A"].append(a)
    return a, cache, (rm, rv)


def mlp_backward(grad_pred, p, acts, cache, training=True):
    """-> (grads dict dW, db, dgamma, dbeta, grad_inp)."""
    L = len(p["W"])
    g = dict(dW=[None] * L, db=[None] * L, dgamma=[None] * (L - 1), dbeta=[None] * (L - 1))
    d = grad_pred * act_grad_from_output(cache["A"][L], acts[L - 1])
    for i in range(L - 1, -1, -1):
        if i < L - 1:
            dy = d * act_grad_from_output(cache["A"][i + 1], acts[i])
            yhat, invstd = cache["yhat"][i], cache["invstd"][i]
            g["dgamma"][i] = (dy * yhat).sum(axis=0)
            g["dbeta"][i] = dy.sum(axis=0)
            n = dy.shape[0]
            if training:
                d = p["gamma"][i] * invstd * (dy - g["dbeta"][i] / n - yhat * g["dgamma"][i] / n)
            else:
                d = p["gamma"][i] * invstd * dy
        g["dW"][i] = d.T @ cache["A"][i]
        g["db"][i] = d.sum(axis=0)
        d = d @ p["W"][i]
    return g, d


def mse_loss(pred, target):
    diff = pred - target
    return (diff * diff).mean(), 2 * diff / diff.size


def multitask_loss(pred, target, ignore_index):
    """multitask_loss.py:40-49, as executed: `F.binary_cross_entropy(input, mask * target, weight)` (:46) is called with
    its DEFAULT reduction ('mean' — the constructor's reduction='none' only sets an attribute the call never passes), so it
    returns the scalar mean BCE over all B*T entries, ignored labels counted as class 0. `loss * mask` (:47) broadcasts that
    scalar, `loss.sum(0) / n_samples` (:48-49) gives it back per task (n_j / n_j), and the task mean returns it once more:
        loss = mean_{i,j} BCE(p_ij, mask_ij * y_ij) * mean_j(n_j / n_j)        (nan when a task has no valid sample)
    torch clamps the logs at -100 and, in the backward, p(1-p) at 1e-12. Returns (loss, dloss/dpred)."""
    mask = (target != ignore_index).astype(pred.dtype)
    t = mask * target
    bce = -(t * np.maximum(np.log(pred), -100) + (1 - t) * np.maximum(np.log(1 - pred), -100))
    s = bce.mean()
    n = mask.sum(axis=0)
    with np.errstate(invalid="ignore", divide="ignore"):
        coef = ((s * mask).sum(axis=0) / n).mean() / s if s != 0 else (n / n).mean()
    grad = coef * (pred - t) / np.maximum((1 - pred) * pred, 1e-12) / pred.size
    return s * coef, grad
