"""MLP head + criteria, CPU side: the oracle (oracle/head_oracle.py) against fixtures produced by the unmodified reference
(oracle/make_golden_head.py), the module surface (constructor, state_dict, init RNG order) and the criterion class."""
import glob
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import head_oracle as ho

GOLD = os.path.join(os.path.dirname(__file__), "golden")
HEAD_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "head_*.npz")) if "init" not in p)


def load_head(name, dtype=np.float64):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    widths = [int(w) for w in z["widths"]]
    L = len(widths) - 1
    p = dict(W=[z[f"p_W{i}"].astype(dtype) for i in range(L)], b=[z[f"p_b{i}"].astype(dtype) for i in range(L)],
             gamma=[z[f"p_gamma{i}"].astype(dtype) for i in range(L - 1)], beta=[z[f"p_beta{i}"].astype(dtype) for i in range(L - 1)],
             running_mean=[z[f"p_running_mean{i}"].astype(dtype) for i in range(L - 1)],
             running_var=[z[f"p_running_var{i}"].astype(dtype) for i in range(L - 1)])
    return z, widths, [str(a) for a in z["acts"]], str(z["loss_kind"]), bool(z["training"]), p


def test_fixtures_present():
    assert set(HEAD_CASES) >= {"head_logp", "head_tox21", "head_eval", "head_deep", "head_single", "head_wide"}


@pytest.mark.parametrize("name", HEAD_CASES)
def test_oracle_reproduces_reference(name):
    z, widths, acts, loss_kind, training, p = load_head(name)
    pred, cache, (rm, rv) = ho.mlp_forward(z["inp"].astype(np.float64), p, acts, training)
    if loss_kind == "mse":
        loss, gp = ho.mse_loss(pred, z["target"].astype(np.float64))
    elif loss_kind == "multitask":
        loss, gp = ho.multitask_loss(pred, z["target"].astype(np.float64), 9.0)
    else:
        loss, gp = (pred * z["gpred"]).sum(), z["gpred"].astype(np.float64)
    g, dx = ho.mlp_backward(gp, p, acts, cache, training)
    tol = 1e-6 if loss_kind == "multitask" else 1e-10      # the reference criterion rounds through fp32 once (make_golden_head.py)

    def close(a, b, scale=0.0):
        return np.abs(np.asarray(a) - np.asarray(b)).max() <= tol * max(np.abs(np.asarray(b)).max(), scale, 1e-30)

    assert close(pred, z["ref64_pred"]) and close(loss, z["ref64_loss"]) and close(dx, z["ref64_dx"])
    for i in range(len(widths) - 1):
        assert close(g["dW"][i], z[f"ref64_dW{i}"])
        assert close(g["db"][i], z[f"ref64_db{i}"], np.abs(z[f"ref64_dW{i}"]).max())
    for i in range(len(widths) - 2):
        assert close(g["dgamma"][i], z[f"ref64_dgamma{i}"]) and close(g["dbeta"][i], z[f"ref64_dbeta{i}"])
        if training:
            assert close(rm[i], z[f"ref64_rm{i}"]) and close(rv[i], z[f"ref64_rv{i}"])


def test_module_surface_and_init_order():
    from openchem_b200.modules.mlp.openchem_mlp import OpenChemMLP
    z = np.load(os.path.join(GOLD, "head_init_seed11.npz"))
    torch.manual_seed(11)
    mlp = OpenChemMLP({"input_size": 24, "n_layers": 3, "hidden_size": [20, 12, 3], "activation": [F.relu, F.relu, "identity"]})
    sd = mlp.state_dict()
    assert set(sd.keys()) == set(z.files)
    for k in z.files:                                   # same seed -> identical initial weights as the reference constructor
        assert np.array_equal(sd[k].numpy(), z[k]), k
    assert mlp.hidden_size == [20, 12, 3] and mlp.input_size == [24, 20, 12] and mlp.n_layers == 3 and mlp.dropout == 0
    assert len(mlp.layers) == 3 and len(mlp.bn) == 2 and len(mlp.dropouts) == 3
    with pytest.raises(ValueError):
        OpenChemMLP({"input_size": 24, "n_layers": 1, "hidden_size": [3]})          # 'activation' is required (check_params)
    with pytest.raises(NotImplementedError):
        OpenChemMLP({"input_size": 24, "n_layers": 1, "hidden_size": [3], "activation": F.gelu})
    with pytest.raises(RuntimeError):                    # no CPU path
        mlp(torch.zeros(4, 24))
    drop = OpenChemMLP({"input_size": 8, "n_layers": 1, "hidden_size": [3], "activation": F.relu, "dropout": 0.5})
    with pytest.raises(NotImplementedError):
        drop.train()(torch.zeros(4, 8))


SIMPLE_CASES = ["mlp_simple_xavier", "mlp_simple_default", "mlp_simple_single"]
ACT_FN = {"identity": "identity", "relu": F.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh}


def build_simple(z):
    """openchem_b200's OpenChemMLPSimple built under the fixture's seed, the way oracle/make_golden_mlp_simple.py built the reference's."""
    from openchem_b200.modules.mlp.openchem_mlp import OpenChemMLPSimple
    w = [int(v) for v in z["widths"]]
    params = {"input_size": w[0], "n_layers": len(w) - 1, "hidden_size": w[1:], "activation": [ACT_FN[str(a)] for a in z["acts"]]}
    if str(z["init"]):
        params["init"] = str(z["init"])
    torch.manual_seed(int(z["seed"]))
    return OpenChemMLPSimple(params), w


@pytest.mark.parametrize("name", SIMPLE_CASES)
def test_mlp_simple_surface_and_init_order(name):
    """Same constructor arguments and seed as the reference's OpenChemMLPSimple (openchem_mlp.py:64-94) -> bit-identical initial
    weights (default nn.Linear init and the xavier_uniform branch), same state_dict keys and attributes; no CPU path."""
    z = np.load(os.path.join(GOLD, name + ".npz"))
    mlp, w = build_simple(z)
    L = len(w) - 1
    assert set(mlp.state_dict().keys()) == {"layers.%d.%s" % (i, k) for i in range(L) for k in ("weight", "bias")}
    for i in range(L):
        assert np.array_equal(mlp.layers[i].weight.detach().numpy(), z[f"W{i}"]), i
        assert np.array_equal(mlp.layers[i].bias.detach().numpy(), z[f"b{i}"]), i
    assert mlp.hidden_size == w[1:] and mlp.input_size == w[:-1] and mlp.n_layers == L and len(mlp.activation) == L
    with pytest.raises(RuntimeError):
        mlp(torch.zeros(4, w[0]))
    from openchem_b200.modules.mlp.openchem_mlp import OpenChemMLPSimple
    with pytest.raises(NotImplementedError):
        OpenChemMLPSimple({"input_size": 4, "n_layers": 1, "hidden_size": [3], "activation": F.relu, "init": "kaiming"})
    with pytest.raises(ValueError):
        OpenChemMLPSimple({"input_size": 4, "n_layers": 1, "hidden_size": [3]})      # 'activation' is required (check_params)


def test_multitask_criterion_class_matches_oracle():
    from openchem_b200.criterion.multitask_loss import MultitaskLoss
    rng = np.random.RandomState(0)
    pred = 1 / (1 + np.exp(-rng.randn(40, 5)))
    target = (rng.rand(40, 5) < 0.4).astype(np.float64)
    target[rng.rand(40, 5) < 0.2] = 9.0
    pt = torch.tensor(pred, requires_grad=True)
    loss = MultitaskLoss(ignore_index=9, n_tasks=5)(pt, torch.tensor(target))
    loss.backward()
    lo, go = ho.multitask_loss(pred, target, 9.0)
    assert abs(float(loss.detach()) - lo) < 1e-12 and np.abs(pt.grad.numpy() - go).max() < 1e-12
    z = np.load(os.path.join(GOLD, "head_tox21.npz"))     # and against the reference's own criterion on the reference's predictions
    l2 = MultitaskLoss(ignore_index=9, n_tasks=12)(torch.tensor(z["ref64_pred"]), torch.tensor(z["target"].astype(np.float64)))
    assert abs(float(l2) - float(z["ref64_loss"])) < 1e-6
