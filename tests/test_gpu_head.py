"""MLP head + criteria on the B200 (ocgcn_head_forward / ocgcn_head_backward through the module surface) against fixtures
of the unmodified reference (fp64-promoted run) and against the oracle at sizes the fixtures do not cover.
Tolerance: 1e-5 relative (exact-fp32 class), written below."""
import os

import numpy as np
import pytest
import torch

from oracle import head_oracle as ho
from test_head_host import HEAD_CASES, load_head

pytestmark = pytest.mark.gpu
TOL = 1e-5
ACT = {"identity": "identity", "relu": torch.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh}


def _dev():
    return torch.device("cuda:0")


def _rel(a, b, scale=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), scale, 1e-30)


def _build(widths, acts, p, dev, training):
    from openchem_b200.modules.mlp.openchem_mlp import OpenChemMLP
    mlp = OpenChemMLP({"input_size": widths[0], "n_layers": len(widths) - 1, "hidden_size": list(widths[1:]),
                       "activation": [ACT[a] for a in acts]})
    sd = {}
    for i in range(len(widths) - 1):
        sd[f"layers.{i}.weight"] = torch.from_numpy(np.asarray(p["W"][i], dtype=np.float32))
        sd[f"layers.{i}.bias"] = torch.from_numpy(np.asarray(p["b"][i], dtype=np.float32))
    for i in range(len(widths) - 2):
        sd[f"bn.{i}.weight"] = torch.from_numpy(np.asarray(p["gamma"][i], dtype=np.float32))
        sd[f"bn.{i}.bias"] = torch.from_numpy(np.asarray(p["beta"][i], dtype=np.float32))
        sd[f"bn.{i}.running_mean"] = torch.from_numpy(np.asarray(p["running_mean"][i], dtype=np.float32))
        sd[f"bn.{i}.running_var"] = torch.from_numpy(np.asarray(p["running_var"][i], dtype=np.float32))
        sd[f"bn.{i}.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    mlp.load_state_dict(sd, strict=True)
    return mlp.to(dev).train(training)


def _criterion(kind, T):
    from openchem_b200.criterion.multitask_loss import MultitaskLoss
    return torch.nn.MSELoss() if kind == "mse" else MultitaskLoss(ignore_index=9, n_tasks=T)


def _run(mlp, inp, target, gpred, loss_kind, fused):
    from openchem_b200.modules.mlp.openchem_mlp import fused_head_loss
    dev = _dev()
    x = torch.from_numpy(inp).to(dev).requires_grad_(True)
    if loss_kind == "none":
        pred = mlp(x)
        loss = (pred * torch.from_numpy(gpred).to(dev)).sum()
    elif fused:
        loss, pred = fused_head_loss(mlp, x, torch.from_numpy(target).to(dev), _criterion(loss_kind, target.shape[1]))
    else:
        pred = mlp(x)
        loss = _criterion(loss_kind, target.shape[1])(pred, torch.from_numpy(target).to(dev))
    loss.backward()
    res = dict(pred=pred.detach().cpu().numpy(), loss=float(loss.detach()), dx=x.grad.cpu().numpy())
    for i, l in enumerate(mlp.layers):
        res[f"dW{i}"], res[f"db{i}"] = l.weight.grad.cpu().numpy(), l.bias.grad.cpu().numpy()
    for i, b in enumerate(mlp.bn):
        res[f"dgamma{i}"], res[f"dbeta{i}"] = b.weight.grad.cpu().numpy(), b.bias.grad.cpu().numpy()
        res[f"rm{i}"], res[f"rv{i}"], res[f"nbt{i}"] = b.running_mean.cpu().numpy(), b.running_var.cpu().numpy(), int(b.num_batches_tracked)
    return res


def _compare(res, ref, L, training, label, tol=TOL):
    assert _rel(res["pred"], ref["pred"]) < tol, label
    assert abs(res["loss"] - float(ref["loss"])) <= tol * max(abs(float(ref["loss"])), 1.0), label
    assert _rel(res["dx"], ref["dx"]) < tol, label
    for i in range(L):
        assert _rel(res[f"dW{i}"], ref[f"dW{i}"]) < tol, (label, "dW", i)
        # hidden-layer biases feed a BatchNorm: their gradient is analytically zero -> judged on the weight-gradient scale
        assert _rel(res[f"db{i}"], ref[f"db{i}"], np.abs(ref[f"dW{i}"]).max()) < tol, (label, "db", i)
    for i in range(L - 1):
        assert _rel(res[f"dgamma{i}"], ref[f"dgamma{i}"]) < tol and _rel(res[f"dbeta{i}"], ref[f"dbeta{i}"]) < tol, (label, "bn", i)
        if training:
            assert _rel(res[f"rm{i}"], ref[f"rm{i}"]) < tol and _rel(res[f"rv{i}"], ref[f"rv{i}"]) < tol, (label, "running", i)
            assert res[f"nbt{i}"] == int(ref[f"nbt{i}"])


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("name", HEAD_CASES)
def test_head_vs_reference_golden(name, fused):
    z, widths, acts, loss_kind, training, p = load_head(name, np.float32)
    mlp = _build(widths, acts, p, _dev(), training)
    res = _run(mlp, z["inp"], z["target"], z["gpred"], loss_kind, fused)
    ref = {k[6:]: z[k] for k in z.files if k.startswith("ref64_")}
    _compare(res, ref, len(widths) - 1, training, name)


def _oracle_case(B, widths, acts, loss_kind, seed, training=True):
    rng = np.random.RandomState(seed)
    L = len(widths) - 1
    p = dict(W=[], b=[], gamma=[], beta=[], running_mean=[], running_var=[])
    for i in range(L):
        s = 1 / np.sqrt(widths[i])
        p["W"].append(rng.uniform(-s, s, (widths[i + 1], widths[i])).astype(np.float32))
        p["b"].append(rng.uniform(-s, s, widths[i + 1]).astype(np.float32))
    for i in range(L - 1):
        p["gamma"].append(rng.uniform(0.5, 1.5, widths[i + 1]).astype(np.float32))
        p["beta"].append(rng.uniform(-0.3, 0.3, widths[i + 1]).astype(np.float32))
        p["running_mean"].append(rng.uniform(-0.2, 0.2, widths[i + 1]).astype(np.float32))
        p["running_var"].append(rng.uniform(0.5, 1.5, widths[i + 1]).astype(np.float32))
    inp = np.tanh(rng.randn(B, widths[0])).astype(np.float32)
    T = widths[-1]
    if loss_kind == "multitask":
        target = (rng.rand(B, T) < 0.3).astype(np.float32)
        target[rng.rand(B, T) < 0.15] = 9.0
    else:
        target = rng.randn(B, T).astype(np.float32)
    gpred = rng.randn(B, T).astype(np.float32)
    P = {k: [a.astype(np.float64) for a in v] for k, v in p.items()}
    pred, cache, (rm, rv) = ho.mlp_forward(inp.astype(np.float64), P, acts, training)
    if loss_kind == "mse":
        loss, gp = ho.mse_loss(pred, target.astype(np.float64))
    elif loss_kind == "multitask":
        loss, gp = ho.multitask_loss(pred, target.astype(np.float64), 9.0)
    else:
        loss, gp = (pred * gpred).sum(), gpred.astype(np.float64)
    g, dx = ho.mlp_backward(gp, P, acts, cache, training)
    ref = dict(pred=pred, loss=loss, dx=dx)
    for i in range(L):
        ref[f"dW{i}"], ref[f"db{i}"] = g["dW"][i], g["db"][i]
    for i in range(L - 1):
        ref[f"dgamma{i}"], ref[f"dbeta{i}"], ref[f"rm{i}"], ref[f"rv{i}"], ref[f"nbt{i}"] = g["dgamma"][i], g["dbeta"][i], rm[i], rv[i], 1
    return p, inp, target, gpred, ref


@pytest.mark.parametrize("B,widths,acts,loss_kind", [
    (4096, [128, 128, 1], ["relu", "identity"], "mse"),            # C3 head (bench shape)
    (256, [128, 128, 12], ["relu", "sigmoid"], "multitask"),       # C2 head
    (1024, [256, 256, 1], ["relu", "identity"], "mse"),            # C5 head
    (10001, [24, 40, 3], ["tanh", "identity"], "mse"),             # more tiles than CTAs: partials accumulate across a CTA's tiles
    (2, [8, 8, 2], ["relu", "identity"], "mse"),                   # smallest batch BatchNorm accepts
    (33, [5, 3, 7, 2], ["sigmoid", "relu", "tanh"], "none"),
])
def test_head_vs_oracle(B, widths, acts, loss_kind):
    p, inp, target, gpred, ref = _oracle_case(B, widths, acts, loss_kind, seed=B)
    mlp = _build(widths, acts, p, _dev(), True)
    res = _run(mlp, inp, target, gpred, loss_kind, True)
    # B = 2: BatchNorm over two rows has a tiny variance, 1/sigma amplifies fp32 rounding in ANY fp32 implementation (the
    # reference's own fp32 run differs from its fp64 run by as much); B = 10001: sums over more terms
    _compare(res, ref, len(widths) - 1, True, (B, widths), tol=1e-4 if B == 2 else 2e-5 if B > 4000 else TOL)


@pytest.mark.parametrize("name", ["mlp_simple_xavier", "mlp_simple_default", "mlp_simple_single"])
def test_mlp_simple_vs_reference_golden(name):
    """OpenChemMLPSimple (openchem_mlp.py:64-111: Linear -> activation per layer) on the head kernels against the unmodified
    reference's fp64 run: outputs, input gradient and every parameter gradient within the fp32 class (1e-5 rel-L2)."""
    from test_head_host import build_simple
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    mlp, w = build_simple(z)
    mlp = mlp.to(_dev()).train()
    x = torch.from_numpy(z["inp"]).to(_dev()).requires_grad_(True)
    y = mlp(x)
    (y * torch.from_numpy(z["gout"]).to(_dev())).sum().backward()

    def rel(a, b):
        b = np.asarray(b, dtype=np.float64)
        return float(np.linalg.norm(a.detach().double().cpu().numpy() - b) / max(np.linalg.norm(b), 1e-30))
    assert rel(y, z["out_f64"]) < 1e-5 and rel(x.grad, z["dx_f64"]) < 1e-5
    for i, l in enumerate(mlp.layers):
        assert rel(l.weight.grad, z[f"dW{i}_f64"]) < 1e-5, i
        assert rel(l.bias.grad, z[f"db{i}_f64"]) < 1e-5, i
    # eval mode: same function (no BatchNorm, no dropout)
    assert torch.equal(mlp.eval()(x.detach()), y.detach())


def test_head_is_deterministic_and_rejects_bad_input():
    from openchem_b200.modules.mlp.openchem_mlp import fused_head_loss
    p, inp, target, gpred, _ = _oracle_case(700, [64, 96, 4], ["relu", "identity"], "mse", seed=3)
    runs = []
    for _ in range(2):
        mlp = _build([64, 96, 4], ["relu", "identity"], p, _dev(), True)
        runs.append(_run(mlp, inp, target, gpred, "mse", True))
    for k in runs[0]:
        assert np.array_equal(np.asarray(runs[0][k]), np.asarray(runs[1][k])), k      # fixed-order reductions, no atomics
    mlp = _build([64, 96, 4], ["relu", "identity"], p, _dev(), True)
    with pytest.raises(RuntimeError):
        mlp(torch.zeros(5, 63, device=_dev()))
    with pytest.raises(ValueError):
        mlp(torch.zeros(1, 64, device=_dev()))                     # BatchNorm1d: more than one value per channel in training
    with pytest.raises(NotImplementedError):
        fused_head_loss(mlp, torch.zeros(4, 64, device=_dev()), torch.zeros(4, 4, device=_dev()), torch.nn.L1Loss())
    mlp.eval()
    assert mlp(torch.zeros(1, 64, device=_dev())).shape == (1, 4)  # eval mode accepts a single row
