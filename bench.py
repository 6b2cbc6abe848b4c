#!/usr/bin/env python
"""GraphCNN train-step benchmark (BASELINE.json metric: "GraphCNN train molecules/sec at 1/2/4/8 B200; % of
HBM/tensor roofline").

    python bench.py --gpus 1 --steps K --warmup W                 # this repo's CUDA path (default)
    python bench.py --impl reference --steps K --warmup W         # the reference's CPU op sequence (oracle port)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One step = Graph2Label-shaped train step on one synthetic padded-graph batch: encoder (libocgcn.so) -> MLP head
(Linear-BN-ReLU-Linear, plain torch: SURVEY section 8(f) "next" row) -> loss -> backward -> (DDP grad
all-reduce over NCCL when N > 1) -> Adam step. Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from openchem_b200.synthetic import WORKLOADS, make_graph_batch, make_labels, random_params  # noqa: E402

METRIC = "graphcnn_train_molecules_per_sec"
UNIT = "molecules/s"


# --------------------------------------------------------------------------------------------------------------
# algorithmic work per molecule (DESIGN.md "Roofline arithmetic"): dense GEMM flops, compulsory HBM bytes
# --------------------------------------------------------------------------------------------------------------
def algorithmic_work(w):
    N, F, E = w["n_max"], [w["f0"]] + list(w["hidden"]), w["encoder_dim"]
    head = 0
    prev = E
    for h in w["mlp_hidden"]:
        head += 2 * prev * h
        prev = h
    L = len(F) - 1
    fwd = sum(2 * N * N * F[l] + 2 * N * F[l] * F[l + 1] for l in range(L)) + 2 * N * F[L] * E + head
    bwd = sum(2 * N * F[l] * F[l + 1] for l in range(L))                                    # dW_l = S^T dZ
    bwd += sum(2 * N * F[l] * F[l + 1] + 2 * N * N * F[l] for l in range(1, L))              # dS = dZ W^T, dH = A^T dS (l >= 2)
    bwd += 2 * (2 * N * F[L] * E) + 2 * head                                                # readout + head
    bytes_ = 4 * N * N + 4 * N * F[0] + 8 * E
    per_layer_fwd = [2 * N * N * F[l] + 2 * N * F[l] * F[l + 1] for l in range(L)]
    return dict(flops_fwd=fwd, flops_bwd=bwd, flops=fwd + bwd, bytes=bytes_, per_layer_fwd=per_layer_fwd)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], bf16_burst=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# --------------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# --------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.samples, self.stop_flag, self.thread = index, [], threading.Event(), None

    def _loop(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            self.stop_flag.wait(0.1)

    def start(self):
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag.set()
        if self.thread:
            self.thread.join(timeout=6)
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 7 for i in range(4) if s[3 + i].lower() == "active"})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons, samples=len(sm))


# --------------------------------------------------------------------------------------------------------------
# the model of the step
# --------------------------------------------------------------------------------------------------------------
def build_model(w, precision, device, head="fused"):
    import torch
    import torch.nn as nn
    from openchem_b200.modules.encoders.gcn_encoder import GraphCNNEncoder
    from openchem_b200.modules.mlp.openchem_mlp import OpenChemMLP

    class Graph2LabelStep(nn.Module):
        """Encoder -> MLP head, the shape of openchem/models/Graph2Label.py:21-37 with OpenChemMLP
        (Linear -> BatchNorm1d -> ReLU -> Linear, example_configs/logp_gcnn_config.py:111-116)."""

        def __init__(self):
            super().__init__()
            self.Encoder = GraphCNNEncoder({"input_size": w["f0"], "encoder_dim": w["encoder_dim"], "n_layers": len(w["hidden"]),
                                            "hidden_size": list(w["hidden"]), "precision": precision}, True)
            hs = w["mlp_hidden"]
            self.sigmoid_out = w["task"] == "multitask"
            if head == "fused":      # the package's OpenChemMLP (libocgcn head kernels)
                self.MLP = OpenChemMLP({"input_size": w["encoder_dim"], "n_layers": 2, "hidden_size": list(hs),
                                        "activation": [torch.relu, torch.sigmoid if self.sigmoid_out else "identity"]})
            else:                    # plain torch modules (eager kernels), for comparison
                self.MLP = nn.Sequential(nn.Linear(w["encoder_dim"], hs[0]), nn.BatchNorm1d(hs[0]), nn.ReLU(), nn.Linear(hs[0], hs[1]))
            self.head = head

        def forward(self, inp):
            out = self.MLP(self.Encoder(inp))
            return torch.sigmoid(out) if (self.sigmoid_out and self.head != "fused") else out

    torch.manual_seed(0)
    return Graph2LabelStep().to(device)


def head_loss_fn(args, crit):
    """loss_fn for TrainStep: criterion evaluated inside the head's last kernel (None with --head torch)."""
    if args.head != "fused":
        return None
    from openchem_b200.modules.mlp.openchem_mlp import fused_head_loss

    def loss_fn(model, inp, y):
        return fused_head_loss(model.MLP, model.Encoder(inp), y, crit)[0]

    return loss_fn


def make_criterion(w):
    import torch
    if w["task"] != "multitask":
        return torch.nn.MSELoss()
    from openchem_b200.criterion.multitask_loss import MultitaskLoss     # openchem/criterion/multitask_loss.py:40-49 as executed
    return MultitaskLoss(ignore_index=9, n_tasks=w["mlp_hidden"][-1])


# --------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's CPU op sequence (oracle/torch_restatement.py) on host cores
# --------------------------------------------------------------------------------------------------------------
def cpu_reference_rate(w, steps, warmup, chunk=256, seed=1234):
    """molecules/s of the reference's eager CPU path (fwd + loss + bwd + Adam) on a bounded sample: chunks of `chunk`
    molecules of the same workload (the reference materialises (B,N,N,D) temporaries, ~8.6 GB each at B=4096)."""
    import torch
    from oracle import torch_restatement as tr
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = min(chunk, w["batch"])
    g = make_graph_batch(B, w["n_max"], w["f0"], occupancy="full", seed=seed)
    x, adj = torch.from_numpy(g["x"]), torch.from_numpy(g["adj"])
    T = w["mlp_hidden"][-1]
    y = torch.from_numpy(make_labels(B, T, "regression"))
    p = tr.params_to_torch(random_params([w["f0"]] + list(w["hidden"]), w["encoder_dim"], seed=0))
    H1 = w["mlp_hidden"][0]
    hp = dict(W1=torch.randn(H1, w["encoder_dim"]) * 0.05, b1=torch.zeros(H1), g=torch.ones(H1), be=torch.zeros(H1),
              rm=torch.zeros(H1), rv=torch.ones(H1), W2=torch.randn(T, H1) * 0.05, b2=torch.zeros(T))
    leaves = [t for k in ("W", "b", "gamma", "beta") for t in p[k]] + [p["Wd"], p["bd"]]
    for k in ("W1", "b1", "g", "be", "W2", "b2"):
        hp[k].requires_grad_(True)
        leaves.append(hp[k])
    opt = torch.optim.Adam(leaves, lr=5e-4)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        out = tr.mlp_head(tr.encoder(x, adj, p, True), hp, True)
        loss = torch.nn.functional.mse_loss(out, y)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    return dict(value=B * len(times) / total, ms_per_step=1e3 * total / len(times), cores=cores, chunk=B, steps=len(times))


def run_reference_arm(args, w, wname):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_rate(w, args.steps, args.warmup)
    sample = "%d steps of a %d-molecule chunk of %s (full occupancy), fwd+loss+bwd+Adam, torch CPU eager, %d threads" % (
        r["steps"], r["chunk"], wname, r["cores"])
    line = dict(impl="reference", metric=METRIC, value=r["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=wname, chunk=r["chunk"], n_max=w["n_max"], f0=w["f0"], hidden=w["hidden"], encoder_dim=w["encoder_dim"]),
                cpu_baseline=dict(value=r["value"], unit=UNIT, cores=r["cores"], kind="port", sample=sample),
                e2e=dict(value=r["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------
# native arm
# --------------------------------------------------------------------------------------------------------------
def finish(world):
    """Multi-rank runs end with a hard exit after a final device sync: tearing down an NCCL communicator that captured CUDA
    graphs still reference was observed to hang in destroy_process_group (r1, 2 GPUs) after the result had been printed."""
    import torch
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        os._exit(0)


def measure_other(wname, args, dev, steps, warmup):
    """Device-resident train-step rate of another named workload with the same runtime (one GPU, rank 0 only)."""
    import torch
    from openchem_b200.step import TrainStep
    w = WORKLOADS[wname]
    B, T = w["batch"], w["mlp_hidden"][-1]
    g = make_graph_batch(B, w["n_max"], w["f0"], occupancy="molecular" if wname in ("c1_logp", "c2_tox21") else "full", seed=99)
    labels = make_labels(B, T, w["task"] if w["task"] == "multitask" else "regression", seed=98)
    x, adj, y = (torch.from_numpy(a).to(dev) for a in (g["x"], g["adj"], labels))
    model = build_model(w, args.precision, dev, args.head)
    crit = make_criterion(w)
    st = TrainStep(model, crit, lambda ps: torch.optim.Adam(ps, lr=5e-4, fused=True, capturable=True), (x, adj, y),
                   use_graph=(args.runtime == "graph"), loss_fn=head_loss_fn(args, crit))
    for _ in range(warmup):
        st()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        st()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return dict(value=B / (ms * 1e-3), unit=UNIT, ms_per_step=ms, per_gpu_batch=B, n_max=w["n_max"], f0=w["f0"], hidden=w["hidden"],
                occupancy="molecular", steps=steps, graph=st.graph is not None)


def run_native(args, w, wname):
    import torch
    import torch.distributed as dist
    from openchem_b200 import _lib, functional

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # rank 0 must print exactly ONE JSON line on stdout: keep library banners (NCCL prints its version there) off it by
    # pointing fd 1 at stderr for the run and writing the JSON line to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group(backend="nccl", init_method="env://", device_id=dev)
    _lib.load()

    B = w["batch"]
    from openchem_b200 import dist as ocd
    g = make_graph_batch(B, w["n_max"], w["f0"], occupancy=args.occupancy, seed=ocd.shard_seed(1234, rank))
    T = w["mlp_hidden"][-1]
    labels = make_labels(B, T, w["task"] if w["task"] == "multitask" else "regression", seed=ocd.shard_seed(4321, rank))
    hx, hadj, hy = (torch.from_numpy(a).pin_memory() for a in (g["x"], g["adj"], labels))
    x, adj, y = hx.to(dev), hadj.to(dev), hy.to(dev)

    model = build_model(w, args.precision, dev, args.head)
    crit = make_criterion(w)
    stepper, runtime = None, "eager (torch autograd + DistributedDataParallel, as run.py drives it)"
    if args.runtime == "graph":
        # the package's step runtime: fwd + loss + bwd + NCCL grad all-reduce + Adam captured in one CUDA graph
        from openchem_b200.step import TrainStep
        stepper = TrainStep(model, crit, lambda ps: torch.optim.Adam(ps, lr=5e-4, fused=True, capturable=True), (x, adj, y),
                            loss_fn=head_loss_fn(args, crit))
        runtime = "openchem_b200.step.TrainStep: one CUDA graph per step (flat-gradient NCCL all-reduce inside)" if stepper.graph is not None \
            else "openchem_b200.step.TrainStep eager (graph capture failed: %s)" % stepper.graph_error
        opt = stepper.opt

        def step(xb, adjb, yb):
            return stepper(xb, adjb, yb)
    else:
        step_model = model
        if world > 1:
            step_model = ocd.wrap_ddp(model, local)
        opt = torch.optim.Adam(model.parameters(), lr=5e-4, fused=True)

        def step(xb, adjb, yb):
            opt.zero_grad(set_to_none=True)
            loss = crit(step_model((xb, adjb)), yb)
            loss.backward()
            opt.step()
            return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        return ocd.max_over_ranks(ms, dev)

    launches_per_body = 0
    if stepper is not None:
        x, adj, y = stepper.inputs          # device-resident static inputs: a step is one graph replay
        functional.launch_counter["n"] = 0
        stepper._body()                      # one eager body: counts the library launches a replay re-issues
        launches_per_body = functional.launch_counter["n"]

    for _ in range(args.warmup):
        step(x, adj, y)
    barrier()

    # ---- device-resident timed region: exactly K steps, CUDA events, max over ranks ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    functional.launch_counter["n"] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(x, adj, y)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = functional.launch_counter["n"]
    if stepper is not None and stepper.graph is not None:
        launches = launches_per_body * args.steps     # replays re-issue the captured launches; count them from one eager body

    # ---- end-to-end region: every step's inputs come from pinned HOST memory (reference layout: dense fp32 adj / features /
    # labels) through the package's feeder (openchem_b200.loader.DevicePrefetcher: upload of batch t+1 overlaps step t), and
    # the loss is read back to the host every step ----
    from openchem_b200.loader import DevicePrefetcher
    h2d = hx.numel() * 4 + hadj.numel() * 4 + hy.numel() * 4
    host_batches = [(hx, hadj, hy)] * (2 + args.steps)
    feed = DevicePrefetcher(host_batches, dev)
    from openchem_b200.step import DelayedLoss
    for _ in range(2):
        xb, adjb, yb = next(feed)
        float(step(xb, adjb, yb).item())
    barrier()
    reader, host_losses = DelayedLoss(dev), []
    e0.record()
    for _ in range(args.steps):
        xb, adjb, yb = next(feed)
        host_losses.append(reader.push(step(xb, adjb, yb)))     # step t's loss reaches the host while step t+1 runs
    host_losses.append(reader.flush())                           # ... and the last one before the clock stops
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    assert len([v for v in host_losses if v is not None]) == args.steps and all(math.isfinite(v) for v in host_losses[1:])

    # ---- the same end-to-end region fed in the package's compact batch format (openchem_b200.compact: bit-packed adjacency
    # and features, 32x fewer host bytes; expanded bit-exactly on the device into the step's static inputs). Reported
    # beside `e2e`, which stays on the reference's dense fp32 host format ----
    from openchem_b200.compact import CompactGraphBatch, pack_graph_batch
    cb_host = pack_graph_batch(g["x"], g["adj"], hy).pin_memory()
    feed_c = DevicePrefetcher([cb_host.tensors()] * (2 + args.steps), dev)

    def compact_step():
        xbits, abits, yb = next(feed_c)
        cb = CompactGraphBatch(xbits, abits, w["f0"], yb)
        if stepper is not None:
            stepper.load_compact(cb)
            return stepper()
        xe, ae = cb.expand()
        return step(xe, ae, yb)

    for _ in range(2):
        float(compact_step().item())
    barrier()
    reader, host_losses = DelayedLoss(dev), []
    e0.record()
    for _ in range(args.steps):
        host_losses.append(reader.push(compact_step()))
    host_losses.append(reader.flush())
    e1.record()
    barrier()
    assert len([v for v in host_losses if v is not None]) == args.steps and all(math.isfinite(v) for v in host_losses[1:])
    ms_e2e_compact = max_over_ranks(e0.elapsed_time(e1))
    # the clock / throttle sampler (nvidia-smi, ~20 Hz at best) ran through the device-resident AND the two end-to-end timed
    # regions: the same load, enough samples for a median even when K steps last only tens of milliseconds
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "device-resident + end-to-end timed regions"

    # ---- per-kernel timing (CUDA events around every launch, on the launching stream) ----
    # (every rank runs the steps - they contain the DDP all-reduce - but only rank 0 records events)
    kern = {}
    if args.profile_steps > 0:
        if rank == 0:
            _lib.profile_enable(True)
        for _ in range(args.profile_steps):
            if stepper is not None:
                stepper._body()
            else:
                step(x, adj, y)
        torch.cuda.synchronize()
        if rank == 0:
            kern = _lib.profile_collect()
            _lib.profile_enable(False)
    barrier()

    if rank != 0:
        finish(world)
        return

    peaks = measured_peaks()
    work = algorithmic_work(w)
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
    tensor_peak = peaks["bf16_sustained"] * (0.5 if args.precision == "tf32" else 1.0)
    t_roof_mol = max(work["bytes"] / (peaks["hbm_gbs"] * 1e9), work["flops"] / (tensor_peak * 1e12))
    step_frac = (B * t_roof_mol) / (ms_step * 1e-3)

    # dominant kernel: the one with the largest share of the profiled steps. Its roofline is the slower of
    # (algorithmic bytes / measured HBM bandwidth) and (algorithmic flops / measured tensor peak) for ONE launch
    # (DESIGN.md section 5 states the per-launch figures: boundary tensors only, dense GEMM flops only).
    roofline, shares = None, {}
    tot_ms = sum(v[0] for v in kern.values()) or 1.0
    for k, (ms, n) in sorted(kern.items(), key=lambda kv: -kv[1][0]):
        shares[k] = dict(ms_per_step=ms / max(args.profile_steps, 1), launches_per_step=n / max(args.profile_steps, 1), share=ms / tot_ms)
    if kern:
        top = max(kern, key=lambda k: kern[k][0])
        ms_launch = kern[top][0] / kern[top][1]
        N, F, E = w["n_max"], [w["f0"]] + list(w["hidden"]), w["encoder_dim"]
        mid, R = F[1], B * w["n_max"]
        tc = args.precision != "fp32"
        simg = 2 if tc else 4                      # bytes/element of the saved S / dZ operand (bf16 image vs fp32)
        fused = tc and max(F + [E]) <= 128            # weight gradients accumulated inside the backward tile kernels
        alg = {   # kernel -> (dense GEMM flops, boundary-tensor bytes) per launch, mid layers (F_in = F_out = mid)
            "fwd_mid": (B * (2 * N * N * mid + 2 * N * mid * mid), R * mid * (4 + 4 + 1 + simg)),            # Z in; Z, arg, S out
            "fwd_first": (B * (2 * N * N * F[0] + 2 * N * F[0] * F[1]), R * (F[0] * (4 + simg) + F[1] * 4)),
            # dY, Z, Z_prev, arg (+ S image when dW is fused) in; dY_prev (+ dZ image when it is not) out
            "bwd_dz": (B * (2 * N * mid * mid + 2 * N * N * mid + (2 * N * mid * mid if fused else 0)), R * mid * (4 + 4 + 4 + 1 + 4 + simg)),
            "gemm_tn": (B * (2 * N * mid * mid), R * mid * 8),
            "dw_mma": (B * (2 * N * mid * mid), R * mid * 4),
            "fwd_readout": (B * (2 * N * F[-1] * E), R * (F[-1] * (4 + 1 + simg) + E * 4)),
            "bwd_readout": (B * (2 * N * F[-1] * E * (2 if fused else 1)), R * (E * 4 + F[-1] * (4 + 1 + 4 + simg))),
            "bwd_dy": (0, R * mid * (4 + 4 + 1 + 4)),
        }.get(top)
        if alg is not None:
            flops_launch, bytes_launch = alg
            if top == "bwd_dz":
                # its L launches are not alike: the one for layer 0 has no dS product, no fused epilogue and a narrower S image.
                # Use the mean algorithmic work per launch (sum over the L launches / L), matching the mean launch time below.
                nL = len(F) - 1
                light_b = R * (F[1] * 8 + (F[0] * simg if fused else F[1] * simg))
                light_f = B * (2 * N * F[0] * F[1]) if fused else 0
                bytes_launch = ((nL - 1) * bytes_launch + light_b) / nL
                flops_launch = ((nL - 1) * flops_launch + light_f) / nL
            t_hbm = bytes_launch / (peaks["hbm_gbs"] * 1e9)
            t_tensor = flops_launch / (peaks["bf16_burst"] * 1e12)
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
            if os.path.exists(tpath):
                traffic = json.load(open(tpath)).get("%s/%s/%s" % (wname, args.precision, top))
            tnote = "dram__bytes_read+write of ONE launch from the committed ncu --set full capture (profiles/r1_traffic.json); for bwd_dz that " \
                    "is a full-layer launch, while achieved/algorithmic figures are means over the L launches (the layer-0 one is lighter)"
            if t_hbm >= t_tensor:
                ach = bytes_launch / (ms_launch * 1e-3) / 1e9
                roofline = dict(bound="hbm", kernel=top, traffic_note=tnote, achieved=ach, peak=peaks["hbm_gbs"], unit="GB/s", frac=ach / peaks["hbm_gbs"],
                                traffic=traffic, ms_per_launch=ms_launch, algorithmic_bytes_per_launch=bytes_launch,
                                algorithmic_flops_per_launch=flops_launch, peak_source=peaks["source"] + ", STREAM-style copy")
            else:
                ach = flops_launch / (ms_launch * 1e-3) / 1e12
                roofline = dict(bound="tensor", kernel=top, traffic_note=tnote, achieved=ach, peak=peaks["bf16_burst"], unit="TFLOP/s", frac=ach / peaks["bf16_burst"],
                                traffic=traffic, ms_per_launch=ms_launch, algorithmic_bytes_per_launch=bytes_launch,
                                algorithmic_flops_per_launch=flops_launch, peak_source=peaks["source"] + ", burst bf16")

    # ---- secondary device-resident numbers for the other named configurations (BASELINE configs[0], [1], [4] shapes at
    # their per-GPU batch) with the same runtime ----
    others = {}
    if world == 1 and not args.no_extra:
        for other, n_steps in (("c1_logp", 100), ("c2_tox21", 100), ("c5_large", 20)):
            if other == wname:
                continue
            try:
                others[other] = measure_other(other, args, dev, steps=n_steps, warmup=10)
            except Exception as exc:  # noqa: BLE001 - secondary numbers must never cost the headline line
                others[other] = dict(error=repr(exc)[:200])
            torch.cuda.empty_cache()

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        r = cpu_reference_rate(w, steps=args.cpu_steps, warmup=1)
        cpu = dict(value=r["value"], unit=UNIT, cores=r["cores"], kind="port",
                   sample="%d steps of a %d-molecule chunk of %s, fwd+loss+bwd+Adam, reference op sequence on torch CPU (%d threads)" % (
                       r["steps"], r["chunk"], wname, r["cores"]))

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_step,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype={"fp32": "f32", "tf32": "tf32", "bf16": "bf16"}[args.precision],
                data="synthetic",
                config=dict(workload=wname, per_gpu_batch=B, global_batch=B * world, n_max=w["n_max"], f0=w["f0"], hidden=w["hidden"],
                            encoder_dim=w["encoder_dim"], occupancy=args.occupancy, precision=args.precision, parallelism="dp%d" % world,
                            optimizer="Adam(fused)", runtime=runtime,
                            head="openchem_b200 OpenChemMLP + criterion (libocgcn head kernels)" if args.head == "fused" else "torch modules", l2_policy="inputs+saved state per step (>1 GB at c3) exceed the 126 MB L2; no explicit flush"),
                e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=4, ms_per_step=ms_e2e / args.steps,
                         feeder="openchem_b200.loader.DevicePrefetcher (pinned host -> device on a copy stream, one batch ahead)",
                         loss_readback="every step's loss is copied to pinned host memory and read there, one step delayed "
                                       "(openchem_b200.step.DelayedLoss); the last read is inside the timed region"),
                e2e_compact=dict(value=world * B * args.steps / (ms_e2e_compact * 1e-3), unit=UNIT, h2d_bytes_per_step=cb_host.nbytes(),
                                 d2h_bytes_per_step=4, ms_per_step=ms_e2e_compact / args.steps,
                                 format="openchem_b200.compact.CompactGraphBatch: bit-packed adj/features from pinned host memory, "
                                        "expanded on the device (ocgcn_expand_compact) inside the timed region"),
                gpu_launches=launches, clocks=clocks, roofline=roofline,
                step_roofline=dict(frac=step_frac, t_roof_us=B * t_roof_mol * 1e6, flops_per_molecule=work["flops"],
                                   bytes_per_molecule=work["bytes"], tensor_peak_tflops=tensor_peak, hbm_gbs=peaks["hbm_gbs"],
                                   peak_source=peaks["source"] + ", sustained bf16"),
                kernels=shares, cpu_baseline=cpu, other_workloads=others)
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    finish(world)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["fp32", "tf32", "bf16"])
    ap.add_argument("--occupancy", default="full", choices=["full", "molecular"])
    ap.add_argument("--runtime", default="graph", choices=["graph", "eager"])
    ap.add_argument("--head", default="fused", choices=["fused", "torch"], help="MLP head + criterion: libocgcn kernels or torch modules")
    ap.add_argument("--profile-steps", type=int, default=3)
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary c2_tox21 measurement")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # Headline: C3 (batch 4096 per B200). N>1 keeps the SAME per-GPU workload (weak scaling: global batch 4096*N), so the
    # per-N values the driver compares are like for like; C4 (per-GPU batch 2048) is `--workload c4_ddp`.
    wname = args.workload or "c3_synth"
    w = WORKLOADS[wname]
    if args.impl == "reference":
        run_reference_arm(args, w, wname)
    else:
        run_native(args, w, wname)


if __name__ == "__main__":
    main()
