#!/usr/bin/env python
"""Per-kernel-kind times of the training forward + backward fed with dense fp32 tensors vs the bit words of the SAME batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import torch
import bench
from openchem_b200 import _lib, compact
from openchem_b200.synthetic import WORKLOADS, make_graph_batch

ap = argparse.ArgumentParser(); ap.add_argument("--workload", default="c3_synth"); ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
w = WORKLOADS[a.workload]
dev = torch.device("cuda:0")
g = make_graph_batch(w["batch"], w["n_max"], w["f0"], occupancy="full", seed=1)
x, adj = torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["adj"]).to(dev)
cb = compact.pack_graph_batch(g["x"], g["adj"]).to(dev)
model = bench.build_model(w, "bf16", dev, "fused").train()
enc = model.Encoder
for name, inp in (("dense", (x, adj)), ("words", (cb.x_bits, cb.adj_bits)), ("dense", (x, adj)), ("words", (cb.x_bits, cb.adj_bits))):
    for _ in range(3):
        enc(inp).sum().backward()
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for _ in range(a.reps):
        enc(inp).sum().backward()
    torch.cuda.synchronize()
    k = _lib.profile_collect()
    _lib.profile_enable(False)
    print(name, {n: round(1e3 * ms / max(c, 1), 1) for n, (ms, c) in k.items() if n in ("adj_pack", "fwd_first", "fwd_mid", "wprep")})
