"""Generate tests/golden/head_*.npz by running the UNMODIFIED reference MLP head and criteria (TEST INFRASTRUCTURE ONLY).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_head.py

Imports openchem.modules.mlp.openchem_mlp.OpenChemMLP and openchem.criterion.multitask_loss.MultitaskLoss from
/root/reference, runs forward + autograd backward in fp32 and fp64 on seeded inputs, asserts oracle/head_oracle.py agrees
with the fp64 run to 1e-10 and stores inputs, weights and reference results. MultitaskLoss hard-codes `.cuda()` on two
temporaries (multitask_loss.py:43-44); in this GPU-less container `torch.Tensor.cuda` is patched to the identity for the
duration of the run — the reference source is not modified.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("OPENCHEM_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.append(ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from openchem.criterion.multitask_loss import MultitaskLoss  # noqa: E402
from openchem.modules.mlp.openchem_mlp import OpenChemMLP  # noqa: E402
from openchem.utils.utils import identity  # noqa: E402

assert os.path.abspath(sys.modules["openchem.modules.mlp.openchem_mlp"].__file__).startswith(REF), "reference not imported"
torch.Tensor.cuda = lambda self, *a, **k: self      # see the module docstring

from oracle import head_oracle as ho  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
ACT_FN = {"identity": identity, "relu": F.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh}

CASES = {
    # logp_gcnn_config.py:112-117 shape (128 -> [128, 1], relu/identity) + MSELoss
    "head_logp":      dict(B=70, widths=[128, 128, 1], acts=["relu", "identity"], loss="mse", training=True),
    # tox21-style multitask head: sigmoid output, 12 tasks, ignore_index 9
    "head_tox21":     dict(B=45, widths=[128, 128, 12], acts=["relu", "sigmoid"], loss="multitask", training=True),
    "head_eval":      dict(B=33, widths=[20, 24, 3], acts=["tanh", "identity"], loss="mse", training=False),
    "head_deep":      dict(B=67, widths=[40, 36, 50, 7], acts=["relu", "tanh", "sigmoid"], loss="none", training=True),
    "head_single":    dict(B=9, widths=[16, 5], acts=["identity"], loss="mse", training=True),
    "head_wide":      dict(B=130, widths=[256, 160, 2], acts=["relu", "identity"], loss="none", training=True),
}


def build(case, seed, dtype):
    torch.manual_seed(seed)
    w = case["widths"]
    mlp = OpenChemMLP({"input_size": w[0], "n_layers": len(w) - 1, "hidden_size": list(w[1:]),
                       "activation": [ACT_FN[a] for a in case["acts"]]})
    g = torch.Generator().manual_seed(seed + 1)
    for bn in mlp.bn:
        with torch.no_grad():
            bn.weight.copy_(torch.rand(bn.weight.shape, generator=g) + 0.5)
            bn.bias.copy_(torch.rand(bn.bias.shape, generator=g) * 0.6 - 0.3)
            bn.running_mean.copy_(torch.rand(bn.bias.shape, generator=g) * 0.4 - 0.2)
            bn.running_var.copy_(torch.rand(bn.bias.shape, generator=g) + 0.5)
    return mlp.to(dtype)


def params_of(mlp):
    p = dict(W=[l.weight.detach().numpy().copy() for l in mlp.layers], b=[l.bias.detach().numpy().copy() for l in mlp.layers],
             gamma=[b.weight.detach().numpy().copy() for b in mlp.bn], beta=[b.bias.detach().numpy().copy() for b in mlp.bn],
             running_mean=[b.running_mean.numpy().copy() for b in mlp.bn], running_var=[b.running_var.numpy().copy() for b in mlp.bn])
    return p


def run_reference(case, seed, dtype, inp, target, gpred):
    mlp = build(case, seed, dtype)
    p0 = params_of(mlp)
    mlp.train(case["training"])
    x = torch.from_numpy(inp).to(dtype).requires_grad_(True)
    pred = mlp(x)
    res = {}
    if case["loss"] == "mse":
        loss = torch.nn.MSELoss()(pred, torch.from_numpy(target).to(dtype))
    elif case["loss"] == "multitask":
        loss = MultitaskLoss(ignore_index=9, n_tasks=case["widths"][-1])(pred, torch.from_numpy(target).to(dtype))
    else:
        loss = (pred * torch.from_numpy(gpred).to(dtype)).sum()
    loss.backward()
    res["pred"] = pred.detach().numpy()
    res["loss"] = np.asarray(loss.detach().numpy())
    res["dx"] = x.grad.numpy()
    for i, l in enumerate(mlp.layers):
        res[f"dW{i}"], res[f"db{i}"] = l.weight.grad.numpy(), l.bias.grad.numpy()
    for i, b in enumerate(mlp.bn):
        res[f"dgamma{i}"], res[f"dbeta{i}"] = b.weight.grad.numpy(), b.bias.grad.numpy()
        res[f"rm{i}"], res[f"rv{i}"] = b.running_mean.numpy().copy(), b.running_var.numpy().copy()
        res[f"nbt{i}"] = np.asarray(b.num_batches_tracked.numpy())
    return p0, res


def main():
    os.makedirs(OUT, exist_ok=True)
    for k, (name, case) in enumerate(CASES.items()):
        seed = 100 + k
        rng = np.random.RandomState(seed)
        B, w = case["B"], case["widths"]
        inp = np.tanh(rng.randn(B, w[0])).astype(np.float32)           # encoder outputs are tanh values
        T = w[-1]
        if case["loss"] == "multitask":
            target = (rng.rand(B, T) < 0.3).astype(np.float32)
            target[rng.rand(B, T) < 0.15] = 9.0
        else:
            target = rng.randn(B, T).astype(np.float32)
        gpred = rng.randn(B, T).astype(np.float32)
        p32, r32 = run_reference(case, seed, torch.float32, inp, target, gpred)
        p64, r64 = run_reference(case, seed, torch.float64, inp, target, gpred)
        # the restatement must reproduce the fp64 reference
        P = {k2: [np.asarray(a, dtype=np.float64) for a in v] for k2, v in p64.items()}
        pred, cache, (rm, rv) = ho.mlp_forward(inp.astype(np.float64), P, case["acts"], case["training"])
        if case["loss"] == "mse":
            loss, gp = ho.mse_loss(pred, target.astype(np.float64))
        elif case["loss"] == "multitask":
            loss, gp = ho.multitask_loss(pred, target.astype(np.float64), 9.0)
        else:
            loss, gp = (pred * gpred).sum(), gpred.astype(np.float64)
        g, dx = ho.mlp_backward(gp, P, case["acts"], cache, case["training"])
        def chk(a, b, what, scale=0.0):
            # scale: hidden-layer biases feed a BatchNorm, their gradient is analytically 0 (rounding noise in any
            # implementation) -> compared on the scale of the layer's weight gradient
            err = np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), scale, 1e-30)
            # MultitaskLoss builds its mask with torch.zeros/ones (float32, multitask_loss.py:43-45); `loss * mask` (:47) is
            # 0-dim fp64 x fp32 tensor -> float32 under torch's promotion rules, so even the fp64-promoted reference rounds
            # the criterion (and the gradient flowing through it) to fp32 once
            assert err < (1e-6 if case["loss"] == "multitask" else 1e-10), (name, what, err)
        chk(pred, r64["pred"], "pred"); chk(loss, r64["loss"], "loss"); chk(dx, r64["dx"], "dx")
        for i in range(len(w) - 1):
            chk(g["dW"][i], r64[f"dW{i}"], f"dW{i}"); chk(g["db"][i], r64[f"db{i}"], f"db{i}", np.abs(r64[f"dW{i}"]).max())
        for i in range(len(w) - 2):
            chk(g["dgamma"][i], r64[f"dgamma{i}"], "dgamma"); chk(g["dbeta"][i], r64[f"dbeta{i}"], "dbeta")
            chk(rm[i], r64[f"rm{i}"], "rm"); chk(rv[i], r64[f"rv{i}"], "rv")
        out = dict(inp=inp, target=target, gpred=gpred, widths=np.asarray(w), acts=np.asarray(case["acts"]),
                   loss_kind=np.asarray(case["loss"]), training=np.asarray(case["training"]))
        for k2, v in p32.items():
            for i, a in enumerate(v):
                out[f"p_{k2}{i}"] = a
        for k2, v in r32.items():
            if k2 in ("pred", "loss", "dx"):          # the fp32 run only documents the reference's own rounding noise
                out[f"ref32_{k2}"] = v
        for k2, v in r64.items():
            out[f"ref64_{k2}"] = v
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print("wrote", name, "loss", float(r64["loss"]))


def init_fixture():
    """Parameter initialisation under a fixed seed: pins the RNG consumption order of OpenChemMLP.__init__ (:29-36)."""
    torch.manual_seed(11)
    mlp = OpenChemMLP({"input_size": 24, "n_layers": 3, "hidden_size": [20, 12, 3], "activation": [F.relu, F.relu, identity]})
    np.savez_compressed(os.path.join(OUT, "head_init_seed11.npz"), **{k: v.numpy() for k, v in mlp.state_dict().items()})
    print("wrote head_init_seed11")


if __name__ == "__main__":
    init_fixture()
    main()
