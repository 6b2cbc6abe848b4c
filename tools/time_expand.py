"""Times ocgcn_expand_compact / ocgcn_pack_compact alone at a workload's batch shape (CUDA events, inputs larger than L2)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openchem_b200 import compact  # noqa: E402
from openchem_b200.synthetic import WORKLOADS, make_graph_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3_synth")
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
w = WORKLOADS[a.workload]
g = make_graph_batch(w["batch"], w["n_max"], w["f0"], occupancy="full")
dev = torch.device("cuda:0")
cb = compact.pack_graph_batch(g["x"], g["adj"]).to(dev)
x, adj = cb.expand()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("expand", lambda: cb.expand(out=(x, adj))), ("pack", lambda: compact.pack_graph_batch_device(x, adj))):
    fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    dense = (x.numel() + adj.numel()) * 4
    print("%s: %.1f us per batch, dense side %.1f MB -> %.0f GB/s" % (name, ms * 1e3, dense / 1e6, dense / ms / 1e6))
