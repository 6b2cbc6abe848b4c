"""Runs the fused MLP head + criterion (forward + backward) a few times at a workload's head shape — the target of
`ncu --metrics gpu__time_duration.sum` launch lists for the head kernels (profiles/)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openchem_b200.criterion.multitask_loss import MultitaskLoss  # noqa: E402
from openchem_b200.modules.mlp.openchem_mlp import OpenChemMLP, fused_head_loss  # noqa: E402
from openchem_b200.synthetic import WORKLOADS, make_labels  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3_synth")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--head", default="fused", choices=["fused", "torch"])
a = ap.parse_args()
w = WORKLOADS[a.workload]
dev = torch.device("cuda:0")
B, E, hs = w["batch"], w["encoder_dim"], w["mlp_hidden"]
multitask = w["task"] == "multitask"
torch.manual_seed(0)
if a.head == "fused":
    mlp = OpenChemMLP({"input_size": E, "n_layers": 2, "hidden_size": list(hs),
                       "activation": [torch.relu, torch.sigmoid if multitask else "identity"]}).to(dev)
else:
    mlp = torch.nn.Sequential(torch.nn.Linear(E, hs[0]), torch.nn.BatchNorm1d(hs[0]), torch.nn.ReLU(), torch.nn.Linear(hs[0], hs[1]),
                              *([torch.nn.Sigmoid()] if multitask else [])).to(dev)
crit = MultitaskLoss(9, hs[-1]) if multitask else torch.nn.MSELoss()
x = torch.tanh(torch.randn(B, E, device=dev)).requires_grad_(True)
y = torch.from_numpy(make_labels(B, hs[-1], "multitask" if multitask else "regression")).to(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(a.steps + 1):
    if it == 1:
        torch.cuda.synchronize()
        e0.record()
    loss = fused_head_loss(mlp, x, y, crit)[0] if a.head == "fused" else crit(mlp(x), y)
    loss.backward()
e1.record()
torch.cuda.synchronize()
print("%s head %s: %.1f us per fwd+bwd (eager launches, includes host gaps)" % (a.head, a.workload, e0.elapsed_time(e1) / a.steps * 1e3))
