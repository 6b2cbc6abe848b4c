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

OUT = os.path.join(ROOT, "tests", "golden")
CASES = {
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
        g = make_graph_batch(case["B"], case["N"], case["F0"], occupancy="molecular", seed=21)
        T = case["mlp"][-1]
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
