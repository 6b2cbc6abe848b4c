"""Whole-model parity: fixtures written by the UNMODIFIED reference Graph2Label (reference GraphCNNEncoder + OpenChemMLP +
criterion, oracle/make_golden_model.py) against the same model assembled from this package's modules with the reference's
state_dict loaded (strict): predictions, loss and every parameter gradient of one training step.
Tolerances: 1e-5 class for precision='fp32' (2e-5 on relative L2 of whole tensors), 1e-2 for 'bf16'."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


class Graph2LabelLike(torch.nn.Module):
    """Attribute names of openchem/models/Graph2Label.py:21-37 (Encoder, MLP) so the reference state_dict loads unchanged."""

    def __init__(self, z, precision):
        super().__init__()
        from openchem_b200.modules.encoders.gcn_encoder import GraphCNNEncoder
        from openchem_b200.modules.mlp.openchem_mlp import OpenChemMLP
        hidden, E, mlp = [int(v) for v in z["hidden"]], int(z["E"]), [int(v) for v in z["mlp"]]
        self.Encoder = GraphCNNEncoder({"input_size": z["x"].shape[2], "encoder_dim": E, "n_layers": len(hidden), "hidden_size": hidden,
                                        "precision": precision}, True)
        self.MLP = OpenChemMLP({"input_size": E, "n_layers": 2, "hidden_size": mlp,
                                "activation": [torch.relu, torch.sigmoid if str(z["task"]) == "multitask" else "identity"]})

    def forward(self, inp):
        return self.MLP(self.Encoder(inp))


def _rel(a, b, scale=0.0):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), scale, 1e-30)


@pytest.mark.parametrize("fused_loss", [True, False])
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("bf16", 1e-2)])
# g2l_*_c1 / _c2: BASELINE configs 1 and 2 at their real widths (5x128, encoder 128, MLP [128,1] / [128,12]) and batch size 256, on the
# first 256 molecules of the reference's own logP / Tox21 CSVs (featurised without RDKit, openchem_b200/smiles.py); the others are
# small synthetic cases of the same two model shapes.
@pytest.mark.parametrize("name", ["g2l_logp_c1", "g2l_tox21_c2", "g2l_logp", "g2l_tox21"])
def test_model_step_matches_reference_graph2label(name, precision, tol, fused_loss):
    from openchem_b200.criterion.multitask_loss import MultitaskLoss
    from openchem_b200.modules.mlp.openchem_mlp import fused_head_loss
    z = np.load(os.path.join(GOLD, name + ".npz"))
    dev = torch.device("cuda:0")
    model = Graph2LabelLike(z, precision)
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd_")}
    model.load_state_dict(sd, strict=True)                       # the reference's keys and shapes, unchanged
    model = model.to(dev).train()
    x, adj, y = (torch.from_numpy(z[k]).to(dev) for k in ("x", "adj", "y"))
    T = y.shape[1]
    crit = MultitaskLoss(ignore_index=9, n_tasks=T) if str(z["task"]) == "multitask" else torch.nn.MSELoss()
    if fused_loss:
        loss, pred = fused_head_loss(model.MLP, model.Encoder((x, adj)), y, crit)
    else:
        pred = model((x, adj))
        loss = crit(pred, y)
    loss.backward()
    assert _rel(pred.detach().cpu().numpy(), z["ref64_pred"]) < tol
    assert abs(float(loss.detach()) - float(z["ref64_loss"])) < tol * max(1.0, abs(float(z["ref64_loss"])))
    grads = {k: p.grad.cpu().numpy() for k, p in model.named_parameters()}
    assert set(grads) == {k[len("ref64_grad_"):] for k in z.files if k.startswith("ref64_grad_")}
    worst = {}
    for k, gval in grads.items():
        ref = z["ref64_grad_" + k]
        scale = 0.0
        if k.endswith(".bias") and ".bn." not in k and "dense" not in k and not k.startswith("MLP.layers.1"):
            # biases that feed a BatchNorm: gradient analytically zero -> judged on the scale of the layer's weight gradient
            scale = np.linalg.norm(z["ref64_grad_" + k[:-4] + "weight"])
        worst[k] = _rel(gval, ref, scale)
    bad = {k: v for k, v in worst.items() if not v < tol}
    assert not bad, bad
