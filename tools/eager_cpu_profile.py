#!/usr/bin/env python
"""Where does the HOST spend its time in the eager route (the loop an unmodified run.py drives: zero_grad, forward, criterion,
backward, Adam, loss.item() under anomaly mode)? cProfile over N steps + wall time per step."""
import os, sys, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import torch
import bench
from openchem_b200.synthetic import WORKLOADS, make_graph_batch, make_labels

ap = argparse.ArgumentParser(); ap.add_argument("--workload", default="c3_synth"); ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--anomaly", type=int, default=1)
a = ap.parse_args()
w = WORKLOADS[a.workload]
dev = torch.device("cuda:0")
g = make_graph_batch(w["batch"], w["n_max"], w["f0"], occupancy="full", seed=1)
task = w["task"] if w["task"] == "multitask" else "regression"
y = torch.from_numpy(make_labels(w["batch"], w["mlp_hidden"][-1], task)).to(dev)
x, adj = torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["adj"]).to(dev)
torch.manual_seed(0)
model = bench.build_model(w, "bf16", dev, "fused")
crit = bench.make_criterion(w)
opt = torch.optim.Adam(model.parameters(), lr=5e-4)


def step():
    with torch.autograd.set_detect_anomaly(bool(a.anomaly)):
        opt.zero_grad()
        loss = crit(model((x, adj)), y)
        loss.backward()
        opt.step()
    return loss.item()


for _ in range(5):
    step()
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(a.steps):
    step()
torch.cuda.synchronize()
print("wall ms per eager step: %.3f (anomaly=%d)" % (1e3 * (time.perf_counter() - t) / a.steps, a.anomaly))
pr = cProfile.Profile()
pr.enable()
for _ in range(a.steps):
    step()
pr.disable()
for key in ("cumulative", "tottime"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(28)
    print("\n".join(l[:170] for l in s.getvalue().splitlines()[4:]))
