"""Generate tests/golden/g2l_*.npz by running the UNMODIFIED reference Graph2Label model (TEST INFRASTRUCTURE ONLY):
openchem.models.Graph2Label.Graph2Label with the reference GraphCNNEncoder + OpenChemMLP and its criterion, one training
step's forward + autograd backward on the CPU in fp32 and fp64. The GPU tests rebuild the same model from this package's
modules (same state_dict) and compare predictions, loss and every parameter gradient.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_model.py
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
from openchem.models.Graph2Label import Graph2Label  # noqa: E402
from openchem.modules.encoders.gcn_encoder import GraphCNNEncoder  # noqa: E402
from openchem.modules.mlp.openchem_mlp import OpenChemMLP  # noqa: E402
from openchem.utils.utils import identity  # noqa: E402

assert os.path.abspath(sys.modules["openchem.models.Graph2Label"].__file__).startswith(REF), "reference not imported"
torch.Tensor.cuda = lambda self, *a, **k: self      # MultitaskLoss hard-codes .cuda() (multitask_loss.py:43-44); no GPU here

from openchem_b200.synthetic import make_graph_batch, make_labels  # noqa: E402

from openchem_b200.smiles import smiles_batch  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def csv_molecules(kind, B, n_max):
    """The first B molecules (that fit n_max atoms) of the reference's own benchmark CSVs, featurised without RDKit
    (openchem_b200/smiles.py: exact topology / element / charge / aromaticity, approximate valence and hybridisation) in the logP
    config's 33-feature layout, with the CSV's labels (Tox21: empty cells -> 9 = ignore_index, tox21_rnn_config.py pattern)."""
    import csv
    if kind == "logp":
        rows = list(csv.reader(open(os.path.join(REF, "benchmark_datasets", "logp_dataset", "logP_labels.csv"))))[1:]
        smiles, labels = [r[1] for r in rows], np.asarray([[float(r[2])] for r in rows], dtype=np.float32)
    else:
        rows = list(csv.reader(open(os.path.join(REF, "benchmark_datasets", "tox21", "tox21.csv"))))[1:]
        smiles = [r[13] for r in rows]
        labels = np.asarray([[9.0 if v == "" else float(v) for v in r[:12]] for r in rows], dtype=np.float32)
    g = smiles_batch(smiles[:4 * B], n_max)
    keep = g["kept"][:B]
    assert len(keep) == B
    return dict(x=g["x"][:B], adj=g["adj"][:B], n_atoms=g["n_atoms"][:B]), labels[keep]


CASES = {
    # the named configurations at their real widths and batch size, on the first 256 molecules of the reference's own data sets:
    # BASELINE config 1 (logp_gcnn_config.py: 33 -> 5x128, E 128, MLP [128,1], MSELoss, N = 85 = the data set's largest molecule)
    "g2l_logp_c1": dict(B=256, N=85, F0=33, hidden=[128] * 5, E=128, mlp=[128, 1], task="regression", data="logp"),
    # BASELINE config 2 (Tox21-shaped: 12 tasks, max_atoms 50, sigmoid head, MultitaskLoss(ignore_index=9))
    "g2l_tox21_c2": dict(B=256, N=50, F0=33, hidden=[128] * 5, E=128, mlp=[128, 12], task="multitask", data="tox21"),
    # logp_gcnn_config.py shape, narrower: regression head + MSELoss
    "g2l_logp": dict(B=20, N=30, F0=33, hidden=[64, 64, 64], E=64, mlp=[64, 1], task="regression"),
    # Tox21-style: 12 tasks, sigmoid head, MultitaskLoss(ignore_index=9)
    "g2l_tox21": dict(B=24, N=25, F0=33, hidden=[48, 48], E=40, mlp=[32, 12], task="multitask"),
}


def build(case, dtype):
    torch.manual_seed(5)
    params = {
        "task": case["task"], "batch_size": case["B"], "num_epochs": 1, "train_data_layer": None, "val_data_layer": None,
        "use_cuda": False, "eval_metrics": None, "logdir": "/tmp", "random_seed": 5, "print_every": 1, "save_every": 1,
        "encoder": GraphCNNEncoder,
        "encoder_params": {"input_size": case["F0"], "encoder_dim": case["E"], "n_layers": len(case["hidden"]), "hidden_size": list(case["hidden"])},
        "mlp": OpenChemMLP,
        "mlp_params": {"input_size": case["E"], "n_layers": 2, "hidden_size": list(case["mlp"]),
                       "activation": [F.relu, torch.sigmoid if case["task"] == "multitask" else identity]},
    }
    return Graph2Label(params).to(dtype)


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, case in CASES.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        T = case["mlp"][-1]
        if "data" in case:
            g, y = csv_molecules(case["data"], case["B"], case["N"])
        else:
            g = make_graph_batch(case["B"], case["N"], case["F0"], occupancy="molecular", seed=21)
            y = make_labels(case["B"], T, "multitask" if case["task"] == "multitask" else "regression", seed=22)
        out = dict(x=g["x"], adj=g["adj"], y=y, hidden=np.asarray(case["hidden"]), E=np.asarray(case["E"]), mlp=np.asarray(case["mlp"]),
                   task=np.asarray(case["task"]))
        for tag, dtype in (("ref32", torch.float32), ("ref64", torch.float64)):
            model = build(case, dtype)
            if tag == "ref32":
                for k, v in model.state_dict().items():
                    out["sd_" + k] = v.numpy().copy()
            crit = MultitaskLoss(ignore_index=9, n_tasks=T) if case["task"] == "multitask" else torch.nn.MSELoss()
            pred = model.forward((torch.from_numpy(g["x"]).to(dtype), torch.from_numpy(g["adj"]).to(dtype)), eval=False)   # Graph2Label.py:30-37
            loss = crit(pred, torch.from_numpy(y).to(dtype))
            loss.backward()
            out[tag + "_pred"], out[tag + "_loss"] = pred.detach().numpy(), np.asarray(loss.detach().numpy())
            for k, p in model.named_parameters():
                out[tag + "_grad_" + k] = p.grad.numpy().copy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print("wrote", name, "loss", float(out["ref64_loss"]))


if __name__ == "__main__":
    main()
