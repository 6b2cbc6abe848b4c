#!/usr/bin/env python
"""Tiny forward+backward of both engines for compute-sanitizer runs:  compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

from _util import build_encoder
from openchem_b200.synthetic import make_graph_batch, random_params

dev = torch.device("cuda:0")
for (B, N, F0, hidden, E) in ((5, 37, 33, [16, 24], 12), (6, 64, 64, [128, 128], 128), (3, 100, 20, [200, 136], 40)):
    g = make_graph_batch(B, N, F0, occupancy="molecular", seed=3)
    p = random_params([F0] + hidden, E, seed=1)
    x, adj = torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["adj"]).to(dev)
    for precision in ("fp32", "bf16"):
        enc = build_encoder({"input_size": F0, "encoder_dim": E, "n_layers": len(hidden), "hidden_size": hidden}, p, dev, precision=precision)
        enc.train()
        out = enc((x, adj))
        out.sum().backward()
        torch.cuda.synchronize()
        print(B, N, F0, hidden, E, precision, float(out.abs().mean()))
print("done")
