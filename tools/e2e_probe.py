#!/usr/bin/env python
"""Host-side timings of the compact feed on the GPU box (where does an end-to-end step spend its time?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from openchem_b200.compact import CompactGraphDataset, CompactGraphBatch
from openchem_b200.synthetic import make_graph_batch
B, n_set = 4096, 16384
g = make_graph_batch(n_set, 64, 64, occupancy="full", seed=1)
ds = CompactGraphDataset.from_dense(g["x"], g["adj"], np.zeros((n_set, 1), np.float32), pinned_slots=6)
rs = np.random.RandomState(0)
for name, fn in (("permutation", lambda: rs.permutation(n_set)[:B]), ("collate", lambda: ds.collate(rs.permutation(n_set)[:B]))):
    ts = []
    for _ in range(12):
        t = time.perf_counter(); fn(); ts.append(1e3 * (time.perf_counter() - t))
    print(name, "ms: first %.2f median %.2f" % (ts[0], float(np.median(ts[2:]))), "threads", torch.get_num_threads())
dev = torch.device("cuda:0")
cb = ds.collate(rs.permutation(n_set)[:B])
x = torch.empty(B, 64, 64, device=dev); adj = torch.empty(B, 64, 64, device=dev)
for _ in range(3):
    d = cb.to(dev, non_blocking=True); d.expand(out=(x, adj))
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(20):
    d = cb.to(dev, non_blocking=True); d.expand(out=(x, adj))
torch.cuda.synchronize()
print("h2d + expand per batch ms: %.3f" % (1e3 * (time.perf_counter() - t) / 20))
