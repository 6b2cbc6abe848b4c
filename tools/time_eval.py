#!/usr/bin/env python
"""Eval-mode forward of the encoder: single-kernel fused path (csrc/ocgcn_fused.cuh) vs the layer-by-layer kernels, per workload.
    python tools/time_eval.py [workload ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from _util import build_encoder
from openchem_b200 import functional
from openchem_b200.synthetic import WORKLOADS, make_graph_batch, random_params

dev = torch.device("cuda:0")
for name in (sys.argv[1:] or ["c2_tox21", "c1_logp", "c4_ddp", "c3_synth"]):
    w = WORKLOADS[name]
    g = make_graph_batch(w["batch"], w["n_max"], w["f0"], occupancy="molecular" if name in ("c1_logp", "c2_tox21") else "full", seed=3)
    p = random_params([w["f0"]] + w["hidden"], w["encoder_dim"], seed=0)
    enc = build_encoder({"input_size": w["f0"], "encoder_dim": w["encoder_dim"], "n_layers": len(w["hidden"]), "hidden_size": w["hidden"]},
                        p, dev, precision="bf16").eval()
    x, adj = torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["adj"]).to(dev)
    res = {}
    for mode, on in (("fused", True), ("layered", False)):
        functional.fused_eval.update(on=on, max_tiles=1 << 30)
        with torch.no_grad():
            for _ in range(5):
                out = enc((x, adj))
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                out = enc((x, adj))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(30):
                gr.replay()
            e1.record()
            torch.cuda.synchronize()
        res[mode] = (e0.elapsed_time(e1) / 30 * 1e3, out.clone())
    print("%-9s B %5d N %3d tiles %5d: fused %8.1f us   layered %8.1f us   (%.2fx)   bit-identical %s" % (
        name, w["batch"], w["n_max"], functional.fused_eval_tiles(type("D", (), dict(B=w["batch"], N=w["n_max"]))()), res["fused"][0],
        res["layered"][0], res["layered"][0] / res["fused"][0], bool(torch.equal(res["fused"][1], res["layered"][1]))), flush=True)
