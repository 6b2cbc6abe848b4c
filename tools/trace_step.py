#!/usr/bin/env python
"""Which kernels does ONE eager train step launch, and from which ops? (torch.profiler; the step's CUDA graph replays the same list)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import torch
from torch.profiler import ProfilerActivity, profile
import bench
from openchem_b200.step import TrainStep
from openchem_b200.synthetic import WORKLOADS, make_graph_batch, make_labels

ap = argparse.ArgumentParser(); ap.add_argument("--workload", default="c2_tox21"); a = ap.parse_args()
w = WORKLOADS[a.workload]
dev = torch.device("cuda:0")
g = make_graph_batch(w["batch"], w["n_max"], w["f0"], occupancy="full", seed=1)
y = make_labels(w["batch"], w["mlp_hidden"][-1], w["task"] if w["task"] == "multitask" else "regression")
x, adj, yy = (torch.from_numpy(t).to(dev) for t in (g["x"], g["adj"], y))
args = argparse.Namespace(precision="bf16", head="fused", runtime="graph")
model = bench.build_model(w, "bf16", dev, "fused")
crit = bench.make_criterion(w)
st = TrainStep(model, crit, lambda ps: torch.optim.Adam(ps, lr=5e-4, fused=True, capturable=True), (x, adj, yy), use_graph=False, loss_fn=bench.head_loss_fn(args, crit))
st._body(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=False) as prof:
    st._body()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
