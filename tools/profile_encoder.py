#!/usr/bin/env python
"""Short encoder-only fwd+bwd loop for ncu captures (no torch head, no optimizer)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

from _util import build_encoder
from openchem_b200.synthetic import WORKLOADS, make_graph_batch, random_params

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3_synth")
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--precision", default="fp32")
a = ap.parse_args()
w = dict(WORKLOADS[a.workload])
if a.batch:
    w["batch"] = a.batch
dev = torch.device("cuda:0")
g = make_graph_batch(w["batch"], w["n_max"], w["f0"], occupancy="full", seed=1234)
p = random_params([w["f0"]] + w["hidden"], w["encoder_dim"], seed=0)
enc = build_encoder({"input_size": w["f0"], "encoder_dim": w["encoder_dim"], "n_layers": len(w["hidden"]), "hidden_size": w["hidden"]},
                    p, dev, precision=a.precision)
enc.train()
x, adj = torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["adj"]).to(dev)
dO = torch.randn(w["batch"], w["encoder_dim"], device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(a.steps):
    if it == a.steps - 1:
        e0.record()
    enc.zero_grad()
    out = enc((x, adj))
    out.backward(dO)
e1.record()
torch.cuda.synchronize()
print("last step fwd+bwd ms:", e0.elapsed_time(e1))
