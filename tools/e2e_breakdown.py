#!/usr/bin/env python
"""Where does the end-to-end step (bench.py `e2e`) lose time against the device-resident step? Adds the pieces one at a time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import numpy as np
import torch
import bench
from openchem_b200.compact import CompactGraphBatch, CompactGraphDataset
from openchem_b200.loader import BackgroundIterator, DevicePrefetcher
from openchem_b200.step import DelayedLoss, TrainStep
from openchem_b200.synthetic import WORKLOADS, make_graph_batch, make_labels

ap = argparse.ArgumentParser(); ap.add_argument("--workload", default="c3_synth"); ap.add_argument("--steps", type=int, default=40)
a = ap.parse_args()
w = WORKLOADS[a.workload]; B = w["batch"]; T = w["mlp_hidden"][-1]
dev = torch.device("cuda:0")
task = w["task"] if w["task"] == "multitask" else "regression"
g = make_graph_batch(B, w["n_max"], w["f0"], occupancy="full", seed=1)
y = make_labels(B, T, task)
x, adj, yy = (torch.from_numpy(t).to(dev) for t in (g["x"], g["adj"], y))
args = argparse.Namespace(precision="bf16", head="fused", runtime="graph")
model = bench.build_model(w, "bf16", dev, "fused")
crit = bench.make_criterion(w)
st = TrainStep(model, crit, lambda ps: torch.optim.Adam(ps, lr=5e-4, fused=True, capturable=True), (x, adj, yy), loss_fn=bench.head_loss_fn(args, crit))
n_set = 4 * B
gs = make_graph_batch(n_set, w["n_max"], w["f0"], occupancy="full", seed=2)
ys = make_labels(n_set, T, task, seed=3)
ds = CompactGraphDataset.from_dense(gs["x"], gs["adj"], ys, pinned_slots=6)
rs = np.random.RandomState(0)
fixed_host = ds.collate(rs.permutation(n_set)[:B]).tensors()
fixed_dev = tuple(t.to(dev) for t in fixed_host)


def timed(fn, n=a.steps):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n


def with_compact(t3):
    st.load_compact(CompactGraphBatch(t3[0], t3[1], w["f0"], t3[2]))
    return st()


print("1 resident step                          %.1f us" % timed(lambda: st()))
print("2 + expand from a device compact batch   %.1f us" % timed(lambda: with_compact(fixed_dev)))
rd = DelayedLoss(dev)
print("3 + delayed loss read                    %.1f us" % timed(lambda: rd.push(with_compact(fixed_dev))))
n = a.steps + 8
feed = DevicePrefetcher([fixed_host] * n, dev)
print("4 + H2D prefetch of a fixed host batch   %.1f us" % timed(lambda: rd.push(with_compact(next(feed)))))


def batches(k):
    for _ in range(k):
        yield ds.collate(rs.permutation(n_set)[:B]).tensors()


feed = DevicePrefetcher(BackgroundIterator(batches(n), depth=2), dev)
print("5 + per-step collate in a thread (= e2e) %.1f us" % timed(lambda: rd.push(with_compact(next(feed)))))
feed = DevicePrefetcher(BackgroundIterator(batches(n), depth=2), dev)
print("6 same, no loss read                     %.1f us" % timed(lambda: with_compact(next(feed))))

# ---- the same region with a step whose static inputs are the bit words themselves (ocgcn_encoder_forward_compact) ----
torch.manual_seed(0)
model_c = bench.build_model(w, "bf16", dev, "fused")
ex = (fixed_dev[0].clone(), fixed_dev[1].clone(), fixed_dev[2].clone())
stc = TrainStep(model_c, crit, lambda ps: torch.optim.Adam(ps, lr=5e-4, fused=True, capturable=True), ex, loss_fn=bench.head_loss_fn(args, crit))


def with_words(t3):
    stc.load_compact(CompactGraphBatch(t3[0], t3[1], w["f0"], t3[2]))
    return stc()


print("7 resident step on the bit words          %.1f us" % timed(lambda: stc()))
print("8 + copies from a device compact batch    %.1f us" % timed(lambda: with_words(fixed_dev)))
feed = DevicePrefetcher(BackgroundIterator(batches(n), depth=2), dev)
rd2 = DelayedLoss(dev)
print("9 + collate thread, H2D, loss read (= e2e) %.1f us" % timed(lambda: rd2.push(with_words(next(feed)))))

# ---- per-kernel-kind times (CUDA events around every launch) of one eager body: dense inputs vs bit words ----
from openchem_b200 import _lib
for name, s in (("dense", st), ("words", stc)):
    _lib.profile_enable(True)
    for _ in range(3):
        s._body()
    torch.cuda.synchronize()
    k = _lib.profile_collect()
    _lib.profile_enable(False)
    print(name, {n: (round(1e3 * ms / max(c, 1), 1), c // 3) for n, (ms, c) in k.items()})   # (us per launch, launches per step)
