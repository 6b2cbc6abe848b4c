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
import contextlib
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
    """SM clock and throttle reasons DURING the timed regions (B200_PROFILING.md's clocks line). The samples come from NVML, the
    library nvidia-smi itself reads, polled every few milliseconds by a thread of this process and only inside `window()`, i.e.
    while a timed region runs: the regions last tens of milliseconds, shorter than one `nvidia-smi` process start, and a fork
    per sample stalls the launching thread for a millisecond or more each time, which the end-to-end region (at most one step
    ahead of the device) cannot hide. Without NVML bindings: one `nvidia-smi -lms 100` process for the whole span."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.samples, self.proc, self.log = index, [], None, None
        self.thread, self.stop_flag, self.active, self.source = None, threading.Event(), threading.Event(), None

    def _nvml_handle(self):
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            uuid = "GPU-" + str(torch.cuda.get_device_properties(self.index).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        except Exception:  # noqa: BLE001
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def _loop(self, nv, h):
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        while not self.stop_flag.is_set():
            if self.active.is_set():
                try:
                    sm, r = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), int(reasons_fn(h))
                    self.samples.append([str(sm), str(mx), ""] + ["Active" if r & b else "Not Active" for _, b in bits])
                except Exception:  # noqa: BLE001
                    pass
                self.stop_flag.wait(0.004)
            else:
                self.stop_flag.wait(0.001)

    def start(self):
        try:
            nv, h = self._nvml_handle()
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            self.source = "NVML (pynvml), polled every 4 ms inside the timed regions only"
            self.thread = threading.Thread(target=self._loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:  # noqa: BLE001
            pass
        import tempfile
        self.source = "nvidia-smi -lms 100, one process over the whole span of the timed regions"
        self.log = tempfile.TemporaryFile(mode="w+")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=self.log, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    @contextlib.contextmanager
    def window(self):
        self.active.set()
        try:
            yield
        finally:
            self.active.clear()

    def stop(self):
        self.stop_flag.set()
        if self.thread is not None:
            self.thread.join(timeout=6)
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=6)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.log.seek(0)
            self.samples = [[t.strip() for t in line.split(",")] for line in self.log.read().splitlines() if line.strip()]
            self.log.close()
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 7 for i in range(4) if s[3 + i].lower() == "active"})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons, samples=len(sm),
                    source=self.source)


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
def _reference_stepper(w, B, g, y):
    """The UNMODIFIED reference (staged under baseline/_ref by oracle/stage_reference.py) on the host cores: its own Graph2Label
    (reference GraphCNNEncoder + OpenChemMLP), criterion, OpenChemOptimizer(Adam) and `train_step` (anomaly mode included,
    openchem_model.py:102-121) through its public API. Returns a zero-argument step function, or None if it is not staged."""
    import torch
    from oracle import stage_reference
    ref = stage_reference.staged_path()
    if ref is None:
        return None
    stubs = os.path.join(ROOT, "tests", "stubs")          # rdkit import stub (RDKit is not in this image; never called here)
    for pth in (stubs, ref):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    for k in [k for k in sys.modules if k == "openchem" or k.startswith("openchem.")]:
        sys.modules.pop(k)                                   # (this repo's namespace shadow must not win: reference first)
    from openchem.models.Graph2Label import Graph2Label
    from openchem.models.openchem_model import build_training, train_step
    from openchem.modules.encoders.gcn_encoder import GraphCNNEncoder
    from openchem.modules.mlp.openchem_mlp import OpenChemMLP
    from openchem.utils.utils import identity
    assert GraphCNNEncoder.__module__ == "openchem.modules.encoders.gcn_encoder" and os.path.realpath(
        sys.modules[GraphCNNEncoder.__module__].__file__).startswith(os.path.realpath(ref)), "reference arm must run the reference's own encoder"
    multitask = w["task"] == "multitask"
    if multitask:     # the reference's MultitaskLoss hard-codes .cuda() (multitask_loss.py:43-44): unusable on the host; BCE as executed
        crit = torch.nn.BCELoss()
    else:
        crit = torch.nn.MSELoss()
    params = {
        "task": "regression", "batch_size": B, "num_epochs": 1, "train_data_layer": None, "val_data_layer": None, "use_cuda": False,
        "eval_metrics": None, "logdir": "/tmp/ocb200_ref_arm", "random_seed": 0, "print_every": 1, "save_every": 1,
        "criterion": crit, "optimizer": torch.optim.Adam, "optimizer_params": {"lr": 5e-4},
        "encoder": GraphCNNEncoder,
        "encoder_params": {"input_size": w["f0"], "encoder_dim": w["encoder_dim"], "n_layers": len(w["hidden"]), "hidden_size": list(w["hidden"])},
        "mlp": OpenChemMLP,
        "mlp_params": {"input_size": w["encoder_dim"], "n_layers": 2, "hidden_size": list(w["mlp_hidden"]),
                       "activation": [torch.nn.functional.relu, torch.sigmoid if multitask else identity]},
    }
    torch.manual_seed(0)
    model = Graph2Label(params)
    criterion, optimizer, _ = build_training(model, params)
    sample = {"adj_matrix": torch.from_numpy(g["adj"]), "node_feature_matrix": torch.from_numpy(g["x"]), "labels": torch.from_numpy(y)}
    if multitask:
        sample["labels"] = (sample["labels"] == 1).float()

    def step():
        inp, target = Graph2Label.cast_inputs(sample, "regression", False)       # Graph2Label.py:39-64, as fit() calls it every batch
        return float(train_step(model, optimizer, criterion, inp, target).item())

    return step


def cpu_reference_rate(w, steps, warmup, chunk=256, seed=1234):
    """molecules/s of the reference's CPU path (fwd + loss + bwd + Adam) on a bounded sample: chunks of `chunk` molecules of the
    same workload (the reference materialises (B,N,N,D) temporaries, ~8.6 GB each at B=4096). kind "reference": the unmodified
    reference itself (baseline/_ref); kind "port" (only when it is not staged): oracle/torch_restatement.py, the same ATen op
    sequence."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = min(chunk, w["batch"])
    g = make_graph_batch(B, w["n_max"], w["f0"], occupancy="full", seed=seed)
    T = w["mlp_hidden"][-1]
    ynp = make_labels(B, T, w["task"] if w["task"] == "multitask" else "regression")
    step, kind = _reference_stepper(w, B, g, ynp), "reference"
    if step is None:
        kind = "port"
        from oracle import torch_restatement as tr
        x, adj, y = torch.from_numpy(g["x"]), torch.from_numpy(g["adj"]), torch.from_numpy(make_labels(B, T, "regression"))
        p = tr.params_to_torch(random_params([w["f0"]] + list(w["hidden"]), w["encoder_dim"], seed=0))
        H1 = w["mlp_hidden"][0]
        hp = dict(W1=torch.randn(H1, w["encoder_dim"]) * 0.05, b1=torch.zeros(H1), g=torch.ones(H1), be=torch.zeros(H1),
                  rm=torch.zeros(H1), rv=torch.ones(H1), W2=torch.randn(T, H1) * 0.05, b2=torch.zeros(T))
        leaves = [t for k in ("W", "b", "gamma", "beta") for t in p[k]] + [p["Wd"], p["bd"]]
        for k in ("W1", "b1", "g", "be", "W2", "b2"):
            hp[k].requires_grad_(True)
            leaves.append(hp[k])
        opt = torch.optim.Adam(leaves, lr=5e-4)

        def step():
            opt.zero_grad()
            loss = torch.nn.functional.mse_loss(tr.mlp_head(tr.encoder(x, adj, p, True), hp, True), y)
            loss.backward()
            opt.step()
            return float(loss)

    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    what = ("the unmodified reference (baseline/_ref): Graph2Label + train_step (anomaly mode) through its own API" if kind == "reference"
            else "oracle/torch_restatement.py (the reference's ATen op sequence; baseline/_ref not staged)")
    sample = "%d steps of a %d-molecule chunk of %%s (full occupancy), fwd+loss+bwd+Adam, %s, torch CPU, %d threads" % (len(times), B, what, cores)
    return dict(value=B * len(times) / total, ms_per_step=1e3 * total / len(times), cores=cores, chunk=B, steps=len(times), kind=kind, sample=sample)


def run_reference_arm(args, w, wname):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_rate(w, args.steps, args.warmup)
    line = dict(impl="reference", metric=METRIC, value=r["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=wname, chunk=r["chunk"], n_max=w["n_max"], f0=w["f0"], hidden=w["hidden"], encoder_dim=w["encoder_dim"]),
                cpu_baseline=dict(value=r["value"], unit=UNIT, cores=r["cores"], kind=r["kind"], sample=r["sample"] % wname),
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


def measure_other(wname, args, dev, steps, warmup, rank=0, world=1):
    """Device-resident train-step rate of another named workload with the same runtime. world == 1: one GPU; world > 1: EVERY rank
    calls this (the captured step contains the NCCL gradient all-reduce), per-rank shards, barrier + max over ranks."""
    import torch
    import torch.distributed as dist
    from openchem_b200 import dist as ocd
    from openchem_b200.step import TrainStep
    w = WORKLOADS[wname]
    B, T = w["batch"], w["mlp_hidden"][-1]
    g = make_graph_batch(B, w["n_max"], w["f0"], occupancy="molecular" if wname in ("c1_logp", "c2_tox21") else "full", seed=99 + rank)
    labels = make_labels(B, T, w["task"] if w["task"] == "multitask" else "regression", seed=98 + rank)
    x, adj, y = (torch.from_numpy(a).to(dev) for a in (g["x"], g["adj"], labels))
    model = build_model(w, args.precision, dev, args.head)
    crit = make_criterion(w)
    st = TrainStep(model, crit, lambda ps: torch.optim.Adam(ps, lr=5e-4, fused=True, capturable=True), (x, adj, y),
                   use_graph=(args.runtime == "graph"), loss_fn=head_loss_fn(args, crit))
    for _ in range(warmup):
        st()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        st()
    e1.record()
    barrier()
    ms = ocd.max_over_ranks(e0.elapsed_time(e1), dev) / steps
    return dict(value=world * B / (ms * 1e-3), unit=UNIT, ms_per_step=ms, per_gpu_batch=B, global_batch=B * world, n_gpus=world, n_max=w["n_max"],
                f0=w["f0"], hidden=w["hidden"], occupancy="molecular" if wname in ("c1_logp", "c2_tox21") else "full", steps=steps,
                graph=st.graph is not None)


def kernel_algorithmic_flops(w, B, fused_dw):
    """Dense-GEMM flops of ONE launch of each tile-kernel kind (mean over that kind's launches in a step), SURVEY 8(d) counting:
    products of the reference's algorithm only - split-operand passes, recomputation and padding are not counted."""
    N, F, E = w["n_max"], [w["f0"]] + list(w["hidden"]), w["encoder_dim"]
    L = len(F) - 1
    fwd = [B * (2 * N * N * F[l] + 2 * N * F[l] * F[l + 1]) for l in range(L)]
    dz = [B * ((2 * N * F[l] * F[l + 1] + 2 * N * N * F[l]) if l > 0 else 0) + (B * 2 * N * F[l] * F[l + 1] if fused_dw else 0) for l in range(L)]
    dw = [B * 2 * N * F[l] * F[l + 1] for l in range(L)] + [B * 2 * N * F[L] * E]
    return {
        "fwd_first": fwd[0], "fwd_mid": sum(fwd[1:]) / max(L - 1, 1), "fwd_readout": B * 2 * N * F[L] * E,
        "bwd_readout": B * 2 * N * F[L] * E * (2 if fused_dw else 1), "bwd_dz": sum(dz) / L,
        "dw_mma": sum(dw) / len(dw), "gemm_tn": sum(dw) / len(dw),
    }


def kernel_boundary_bytes(w, B, tc, fused_dw):
    """Bytes of the boundary tensors one launch of each kind reads + writes in THIS implementation (saved activations, arg-max,
    operand images, gradient ping-pong): the numerator of `hbm_achieved_gbs`, NOT the roofline's compulsory bytes."""
    N, F, E = w["n_max"], [w["f0"]] + list(w["hidden"]), w["encoder_dim"]
    L, R = len(F) - 1, B * w["n_max"]
    simg = 2 if tc else 4
    mid = [R * (F[l] * (4 + 1 + simg) + F[l + 1] * 4) for l in range(1, L)]
    dzb = [R * (F[l + 1] * 8 + F[l] * simg + (F[l] * (4 + 1 + 4) if l > 0 else 0)) for l in range(L)]
    return {
        "fwd_first": R * (F[0] * (4 + simg) + F[1] * 4), "fwd_mid": sum(mid) / max(L - 1, 1),
        "fwd_readout": R * (F[L] * (4 + 1 + simg) + E * 4), "bwd_readout": R * (E * 4 + F[L] * (4 + 1 + 4 + simg)),
        "bwd_dz": sum(dzb) / L,
    }


def parity_block(w, g, precision, dev):
    """Parity of the TIMED configuration (same batch, same widths), computed outside every timed region: one train-mode
    forward + backward of the encoder against the fp64 oracle (oracle/ is used here strictly as the checker). Reports the
    rel-L2 of the output and the worst parameter-gradient rel-L2, raw (oracle's own pool routing) and with the kernel's
    routing, plus how many pool routings differ (all of them must be near-ties)."""
    import torch
    from oracle import gcn_oracle as orc
    from openchem_b200 import functional
    from openchem_b200.modules.encoders.gcn_encoder import GraphCNNEncoder
    t0 = time.perf_counter()
    F, E = [w["f0"]] + list(w["hidden"]), w["encoder_dim"]
    L, B, N = len(F) - 1, g["x"].shape[0], w["n_max"]
    if B * N * max(F) >= orc.POOL_C_MIN_ELEMS and orc.pool_lib() is None:
        return dict(skipped="oracle/_build/libpool_scan.so not built (run __graft_entry__.build())")
    p = random_params(F, E, seed=0)
    enc = GraphCNNEncoder({"input_size": F[0], "encoder_dim": E, "n_layers": L, "hidden_size": list(w["hidden"]), "precision": precision}, True)
    with torch.no_grad():
        for l, gc in enumerate(enc.graph_convolutions):
            gc.weight.copy_(torch.from_numpy(p["W"][l])); gc.bias.copy_(torch.from_numpy(p["b"][l]))
            gc.bn.weight.copy_(torch.from_numpy(p["gamma"][l])); gc.bn.bias.copy_(torch.from_numpy(p["beta"][l]))
            gc.bn.running_mean.copy_(torch.from_numpy(p["running_mean"][l])); gc.bn.running_var.copy_(torch.from_numpy(p["running_var"][l]))
        enc.dense.weight.copy_(torch.from_numpy(p["Wd"])); enc.dense.bias.copy_(torch.from_numpy(p["bd"]))
    enc = enc.to(dev).train()
    dO = np.random.RandomState(7).randn(B, E).astype(np.float32)
    x, adj = torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["adj"]).to(dev)
    functional.debug_keep["on"] = True
    try:
        out = enc((x, adj))
        args = [functional.saved_tensor("argmax", l).cpu().numpy().reshape(B, N, -1).astype(np.int64) for l in range(L)]
    finally:
        functional.debug_keep.update(on=False, saved=None, desc=None)
    out.backward(torch.from_numpy(dO).to(dev))
    torch.cuda.synchronize()
    p64 = orc.cast_params(p, np.float64)
    O, cache, _, _ = orc.encoder_forward(g["x"].astype(np.float64), g["adj"].astype(np.float64), p64, training=True)
    flips = int(sum((args[l] != cache["layers"][l]["arg"]).sum() for l in range(L)))
    res = {"dWd": enc.dense.weight.grad, "dbd": enc.dense.bias.grad}
    for l, gc in enumerate(enc.graph_convolutions):
        res.update({"dW%d" % l: gc.weight.grad, "dgamma%d" % l: gc.bn.weight.grad, "dbeta%d" % l: gc.bn.bias.grad})
    res = {k: v.detach().cpu().numpy() for k, v in res.items()}

    def worst(routing):
        gr = orc.encoder_backward(dO.astype(np.float64), cache, arg_override=routing)
        ref = {"dWd": gr["dWd"], "dbd": gr["dbd"]}
        for l in range(L):
            ref.update({"dW%d" % l: gr["dW"][l], "dgamma%d" % l: gr["dgamma"][l], "dbeta%d" % l: gr["dbeta"][l]})
        errs = {k: orc.rel_l2(res[k], ref[k]) for k in ref}
        k = max(errs, key=errs.get)
        return k, errs[k]

    k_raw, e_raw = worst(None)
    k_rt, e_rt = worst(args) if flips else (k_raw, e_raw)
    tol = 1e-5 if precision == "fp32" else 1e-2
    return dict(oracle="oracle/gcn_oracle.py fp64 (pinned to the unmodified reference by tests/golden)", batch=B, tolerance=tol,
                out_rel_l2=orc.rel_l2(out.detach().cpu().numpy(), O), worst_grad_rel_l2=e_raw, worst_grad=k_raw,
                worst_grad_rel_l2_kernel_routing=e_rt, worst_grad_kernel_routing=k_rt,
                pool_routings_differing=flips, pool_routings=int(sum(a.size for a in args)),
                ok=bool(orc.rel_l2(out.detach().cpu().numpy(), O) < tol and e_rt < tol), seconds=time.perf_counter() - t0)


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
    adam = lambda ps: torch.optim.Adam(ps, lr=5e-4, fused=True, capturable=True)  # noqa: E731
    stepper, runtime = None, "eager (torch autograd + DistributedDataParallel, as run.py drives it)"
    if args.runtime == "graph":
        # the package's step runtime: fwd + loss + bwd + NCCL grad all-reduce + Adam captured in one CUDA graph
        from openchem_b200.step import TrainStep
        stepper = TrainStep(model, crit, adam, (x, adj, y), loss_fn=head_loss_fn(args, crit))
        runtime = ("openchem_b200.step.TrainStep: one CUDA graph per step (flat-gradient all-reduce inside: %s)" % stepper.exchange_kind) if stepper.graph is not None \
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

    def timed(fn, n):
        """exactly n calls of fn bracketed by barrier + synchronize, CUDA events on the current stream, max over ranks -> ms"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

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
    with sampler.window():
        ms_total = timed(lambda: step(x, adj, y), args.steps)
    launches = functional.launch_counter["n"]
    if stepper is not None and stepper.graph is not None:
        launches = launches_per_body * args.steps     # replays re-issue the captured launches; count them from one eager body

    # ---- replicas stayed identical (N > 1): bitwise checksum of every parameter after the timed steps, gathered on all ranks ----
    replica_check = None
    if world > 1:
        flat = torch.cat([q.detach().reshape(-1).double() for q in model.parameters()])
        mine = torch.stack([flat.sum(), (flat * flat).sum(), flat.abs().max()])
        allv = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        replica_check = dict(identical=bool(all(torch.equal(v, allv[0]) for v in allv)), checksum=[float(v) for v in allv[0]],
                             what="sum, sum of squares and max |p| (fp64) of all parameters on every rank after the timed steps")
        assert replica_check["identical"], "data-parallel replicas diverged: %s" % ([v.tolist() for v in allv],)

    from openchem_b200.loader import DevicePrefetcher
    from openchem_b200.step import DelayedLoss

    def e2e_region(next_batch_step):
        """K steps whose inputs start in HOST memory and whose loss is read on the host every step, inside the timed region."""
        for _ in range(2):
            float(next_batch_step().item())
        reader, host_losses = DelayedLoss(dev), []

        def one():
            host_losses.append(reader.push(next_batch_step()))     # step t's loss reaches the host while step t+1 runs

        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            one()
        host_losses.append(reader.flush())                          # ... and the last one before the clock stops
        e1.record()
        barrier()
        assert len([v for v in host_losses if v is not None]) == args.steps and all(math.isfinite(v) for v in host_losses[1:])
        return max_over_ranks(e0.elapsed_time(e1))

    # ---- e2e (headline): the package's data path. A featurised data set cached on the host in the compact format
    # (openchem_b200.compact.CompactGraphDataset: packed ONCE, like any featurisation cache); EVERY step inside the timed region a
    # sampler's index batch is collated from it into pinned memory (row gather), uploaded on the copy stream, expanded on the
    # device into the step's static inputs, the step runs and its loss is read back on the host ----
    from openchem_b200.compact import CompactGraphBatch, CompactGraphDataset
    n_set = 4 * B
    gset = make_graph_batch(n_set, w["n_max"], w["f0"], occupancy=args.occupancy, seed=ocd.shard_seed(777, rank))
    yset = make_labels(n_set, T, w["task"] if w["task"] == "multitask" else "regression", seed=ocd.shard_seed(778, rank))
    host_set = CompactGraphDataset.from_dense(gset["x"], gset["adj"], yset, pinned_slots=6)
    del gset
    perm = np.random.RandomState(5 + rank)

    def compact_batches(n):
        for _ in range(n):
            yield host_set.collate(perm.permutation(n_set)[:B]).tensors()      # per-step host work: the sampler + the collate

    from openchem_b200.loader import BackgroundIterator
    e2e_h2d = [0]

    # The step of the compact route is captured ON the bit words: its static inputs are (x_bits, adj_bits, labels) and the captured
    # kernels read the words themselves (ocgcn_encoder_forward_compact / _backward_compact; GraphCNNEncoder.forward given int32 words).
    # Same model class, same initial weights, its own optimizer state; under N > 1 it carries the same in-graph gradient all-reduce.
    cstepper = None
    if stepper is not None:
        from openchem_b200.step import TrainStep
        torch.manual_seed(0)
        cmodel = build_model(w, args.precision, dev, args.head)
        c0 = host_set.collate(perm.permutation(n_set)[:B]).to(dev)
        cstepper = TrainStep(cmodel, crit, adam, (c0.x_bits, c0.adj_bits, c0.labels), loss_fn=head_loss_fn(args, crit))
        assert cstepper.graph is not None, getattr(cstepper, "graph_error", None)

    def compact_step_via(st_words, st_dense, feed):
        def one():
            xbits, abits, yb = next(feed)
            e2e_h2d[0] = (xbits.numel() + abits.numel()) * 4 + yb.numel() * 4
            cb = CompactGraphBatch(xbits, abits, w["f0"], yb)
            if st_words is not None:
                st_words.load_compact(cb)          # three small device copies into the static word buffers
                return st_words()
            if st_dense is not None:
                st_dense.load_compact(cb)          # ocgcn_expand_compact into the static dense inputs
                return st_dense()
            xe, ae = cb.expand()
            return step(xe, ae, yb)
        return one

    feed_c = DevicePrefetcher(BackgroundIterator(compact_batches(2 + args.steps), depth=2), dev)
    with sampler.window():
        ms_e2e = e2e_region(compact_step_via(cstepper, stepper, feed_c))
    if world > 1 and cstepper is not None:
        flat = torch.cat([q.detach().reshape(-1).double() for q in cmodel.parameters()])
        mine = torch.stack([flat.sum(), (flat * flat).sum(), flat.abs().max()])
        allv = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        assert all(torch.equal(v, allv[0]) for v in allv), "data-parallel replicas of the compact-route step diverged"
    # the earlier form of the compact route, kept for comparison: the words are expanded on the device into the dense step's inputs
    ms_e2e_expand = None
    if cstepper is not None and not args.no_extra:
        feed_x = DevicePrefetcher(BackgroundIterator(compact_batches(2 + args.steps), depth=2), dev)
        ms_e2e_expand = e2e_region(compact_step_via(None, stepper, feed_x))

    # ---- e2e_dense: the same region fed in the REFERENCE's host format (dense fp32 adjacency / features / labels from pinned
    # memory, Graph2Label.cast_inputs' layout) through the same feeder ----
    h2d_dense = hx.numel() * 4 + hadj.numel() * 4 + hy.numel() * 4
    feed = DevicePrefetcher([(hx, hadj, hy)] * (2 + args.steps), dev)
    with sampler.window():
        ms_e2e_dense = e2e_region(lambda: step(*next(feed)))

    # ---- the same per-GPU workload without the gradient exchange (N > 1 only): every rank steps a purely local replica.
    # This is the single-GPU number the N-GPU value should be read against (BASELINE config 4 is quoted at 2048 per GPU) ----
    local_only = None
    if world > 1 and stepper is not None:
        from openchem_b200.step import TrainStep
        torch.manual_seed(0)
        lmodel = build_model(w, args.precision, dev, args.head)
        lstep = TrainStep(lmodel, crit, adam, (x, adj, y), loss_fn=head_loss_fn(args, crit), exchange=False)
        for _ in range(args.warmup):
            lstep()
        ms_local = timed(lstep, args.steps)
        local_only = dict(value=B * args.steps / (ms_local * 1e-3), unit=UNIT, ms_per_step=ms_local / args.steps,
                          what="one GPU, same per-GPU batch, no gradient exchange (slowest rank)")
        del lstep, lmodel

    # ---- BASELINE config 4 as stated (per-GPU batch 2048, N ranks, gradient all-reduce inside the captured step), measured in the
    # same job next to the headline so that every N > 1 line carries it together with its own single-GPU figure ----
    c4_line = None
    if world > 1 and wname != "c4_ddp" and stepper is not None:
        from openchem_b200.step import TrainStep
        c4_line = measure_other("c4_ddp", args, dev, steps=max(args.steps, 20), warmup=max(args.warmup, 5), rank=rank, world=world)
        w4 = WORKLOADS["c4_ddp"]
        g4 = make_graph_batch(w4["batch"], w4["n_max"], w4["f0"], occupancy="full", seed=99 + rank)
        y4 = make_labels(w4["batch"], w4["mlp_hidden"][-1], "regression", seed=98 + rank)
        torch.manual_seed(0)
        m4 = build_model(w4, args.precision, dev, args.head)
        s4 = TrainStep(m4, crit, adam, tuple(torch.from_numpy(a).to(dev) for a in (g4["x"], g4["adj"], y4)), loss_fn=head_loss_fn(args, crit), exchange=False)
        for _ in range(5):
            s4()
        ms4 = timed(s4, 20) / 20
        c4_line["single_gpu_same_workload"] = dict(value=w4["batch"] / (ms4 * 1e-3), unit=UNIT, ms_per_step=ms4,
                                                   what="one GPU, per-GPU batch 2048, no gradient exchange (slowest rank)")
        del s4, m4

    # the clock / throttle sampler polled inside the device-resident AND the end-to-end timed regions (ClockSampler.window):
    # the same load, enough samples for a median even when K steps last only tens of milliseconds
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "inside the device-resident + end-to-end timed regions"

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

    # ---- the route an unmodified run.py takes: eager autograd (+ DDP) + per-step .item(), same model class ----
    eager = None
    if args.runtime == "graph" and not args.no_extra:
        torch.manual_seed(0)
        emodel = build_model(w, args.precision, dev, args.head)
        estep_model = ocd.wrap_ddp(emodel, local) if world > 1 else emodel
        eopt = torch.optim.Adam(emodel.parameters(), lr=5e-4)
        eloss = head_loss_fn(args, crit)

        def eager_step(anomaly=True):
            with torch.autograd.set_detect_anomaly(anomaly):     # openchem_model.py:103
                eopt.zero_grad()
                loss = crit(estep_model((x, adj)), y)
                loss.backward()
                eopt.step()
            return loss.item()                                   # openchem_model.py:159-163

        for _ in range(3):
            eager_step()
        n_eager = max(args.steps // 2, 5)
        ms_eager = timed(eager_step, n_eager)
        ms_plain = timed(lambda: eager_step(False), n_eager)
        eager = dict(value=world * B * n_eager / (ms_eager * 1e-3), unit=UNIT, ms_per_step=ms_eager / n_eager,
                     what="device-resident batch; eager autograd + torch Adam + anomaly mode + loss.item() every step, as the reference's "
                          "train_step/fit drive the model (DistributedDataParallel when N > 1)",
                     without_anomaly_mode=dict(ms_per_step=ms_plain / n_eager, value=world * B * n_eager / (ms_plain * 1e-3),
                                               what="same loop without torch.autograd.set_detect_anomaly: the difference is the reference's own "
                                                    "train_step setting (a Python stack capture per autograd node and a device sync per "
                                                    "gradient for the NaN check), not this package"))
        del eloss, estep_model, emodel

    if rank != 0:
        finish(world)
        return

    peaks = measured_peaks()
    work = algorithmic_work(w)
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)
    tensor_scale = 0.5 if args.precision == "tf32" else 1.0
    tensor_peak = peaks["bf16_sustained"] * tensor_scale
    t_roof_mol = max(work["bytes"] / (peaks["hbm_gbs"] * 1e9), work["flops"] / (tensor_peak * 1e12))
    step_frac = (B * t_roof_mol) / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel, SURVEY 8(d): ALGORITHMIC work of one launch (dense-GEMM flops of the reference's
    # products; compulsory bytes) / its mean duration, against the measured peak. The tile kernels are tensor-bound on paper
    # (flops at tensor peak take longer than their share of the compulsory bytes at HBM peak), so `bound` is "tensor";
    # what the implementation actually moves is reported separately (hbm_achieved_gbs, traffic, traffic_ratio) ----
    roofline, shares = None, {}
    tot_ms = sum(v[0] for v in kern.values()) or 1.0
    for k, (ms, n) in sorted(kern.items(), key=lambda kv: -kv[1][0]):
        shares[k] = dict(ms_per_step=ms / max(args.profile_steps, 1), launches_per_step=n / max(args.profile_steps, 1), share=ms / tot_ms)
    tc = args.precision != "fp32"
    Fall = [w["f0"]] + list(w["hidden"]) + [w["encoder_dim"]]
    fused_dw = tc and max(Fall) <= 128 and w["n_max"] <= 128
    traffic_tab = {}
    tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(tpath):
        traffic_tab = json.load(open(tpath))
    if kern:
        flops_tab, bytes_tab = kernel_algorithmic_flops(w, B, fused_dw), kernel_boundary_bytes(w, B, tc, fused_dw)
        top = max((k for k in kern if k in flops_tab), key=lambda k: kern[k][0], default=None)
        if top is not None:
            ms_launch = kern[top][0] / kern[top][1]
            flops_launch = flops_tab[top]
            comp_bytes_launch = B * work["bytes"] * flops_launch / (B * work["flops"])     # this launch's share of the step's compulsory bytes
            t_tensor = flops_launch / (tensor_peak * 1e12)
            t_hbm = comp_bytes_launch / (peaks["hbm_gbs"] * 1e9)
            traffic = traffic_tab.get("%s/%s/%s" % (wname, args.precision, top))
            if t_tensor >= t_hbm:
                ach, peak, unit, bound = flops_launch / (ms_launch * 1e-3) / 1e12, tensor_peak, "TFLOP/s", "tensor"
            else:
                ach, peak, unit, bound = comp_bytes_launch / (ms_launch * 1e-3) / 1e9, peaks["hbm_gbs"], "GB/s", "hbm"
            # the same two fractions for EVERY tile-kernel kind of the step (the dominant one is `roofline` itself)
            for k in kern:
                if k in flops_tab and kern[k][1]:
                    ms_k = kern[k][0] / kern[k][1]
                    shares[k]["tensor_frac"] = flops_tab[k] / (ms_k * 1e-3) / 1e12 / tensor_peak
                    if k in bytes_tab:
                        shares[k]["implementation_hbm_frac"] = bytes_tab[k] / (ms_k * 1e-3) / 1e9 / peaks["hbm_gbs"]
            roofline = dict(bound=bound, kernel=top, achieved=ach, peak=peak, unit=unit, frac=ach / peak, traffic=traffic,
                            ms_per_launch=ms_launch, algorithmic_flops_per_launch=flops_launch,
                            compulsory_bytes_per_launch=comp_bytes_launch,
                            peak_source=peaks["source"] + (", sustained bf16 (kernel timed inside the step)" if bound == "tensor" else ", STREAM-style copy"),
                            hbm_achieved_gbs=(bytes_tab[top] / (ms_launch * 1e-3) / 1e9) if top in bytes_tab else None,
                            implementation_bytes_per_launch=bytes_tab.get(top),
                            traffic_ratio=(traffic / comp_bytes_launch) if traffic else None,
                            note="frac = algorithmic dense-GEMM flops of one launch (mean over this kind's launches; split-operand passes and "
                                 "recomputation not counted) / mean launch time / measured tensor peak. hbm_achieved_gbs = bytes of the boundary "
                                 "tensors this implementation moves per launch / time; traffic = dram__bytes_read+write of one launch from "
                                 "ncu --set full (profiles/r2_traffic.json); traffic_ratio = traffic / this launch's share of the compulsory bytes")
    step_traffic = traffic_tab.get("%s/%s/step_total" % (wname, args.precision))

    # ---- secondary device-resident numbers for the other named configurations (BASELINE configs[0], [1], [4] shapes at
    # their per-GPU batch) with the same runtime ----
    others = {}
    if world == 1 and not args.no_extra:
        for other, n_steps in (("c1_logp", 100), ("c2_tox21", 100), ("c4_ddp", 30), ("c5_large", 20)):
            if other == wname:
                continue
            try:
                others[other] = measure_other(other, args, dev, steps=n_steps, warmup=10)
                ow = algorithmic_work(WORKLOADS[other])
                t_o = max(ow["bytes"] / (peaks["hbm_gbs"] * 1e9), ow["flops"] / (tensor_peak * 1e12))
                others[other]["step_roofline"] = dict(frac=WORKLOADS[other]["batch"] * t_o / (others[other]["ms_per_step"] * 1e-3),
                                                      t_roof_us=WORKLOADS[other]["batch"] * t_o * 1e6)
            except Exception as exc:  # noqa: BLE001 - secondary numbers must never cost the headline line
                others[other] = dict(error=repr(exc)[:200])
            torch.cuda.empty_cache()

    # ---- eval-mode forward of the encoder (inference: Graph2Label.forward(eval=True) / evaluate / predict): the single-kernel path
    # (activations stay on the SM across all layers, csrc/ocgcn_fused.cuh) next to the layer-by-layer kernels, device-resident ----
    eval_fwd = {}
    if world == 1 and not args.no_extra and args.precision != "fp32":
        from openchem_b200 import functional as fn_
        for other in (wname, "c2_tox21"):
            try:
                wo = WORKLOADS[other]
                go = make_graph_batch(wo["batch"], wo["n_max"], wo["f0"], occupancy="molecular" if other in ("c1_logp", "c2_tox21") else "full", seed=3)
                enc = build_model(wo, args.precision, dev, args.head).Encoder.eval()
                xo, ao = torch.from_numpy(go["x"]).to(dev), torch.from_numpy(go["adj"]).to(dev)
                res = {}
                for mode, on in (("fused", True), ("layer_by_layer", False)):
                    fn_.fused_eval["on"] = on
                    with torch.no_grad():
                        for _ in range(3):
                            o_ = enc((xo, ao))
                        gr_ = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(gr_):
                            o_ = enc((xo, ao))
                        ms_ = timed(gr_.replay, 30) / 30
                    res[mode] = (ms_, o_.clone())
                fn_.fused_eval["on"] = True
                eval_fwd[other] = dict(fused_us=res["fused"][0] * 1e3, layer_by_layer_us=res["layer_by_layer"][0] * 1e3,
                                       molecules_per_s=wo["batch"] / (res["fused"][0] * 1e-3), bit_identical=bool(torch.equal(res["fused"][1], res["layer_by_layer"][1])))
                del enc, gr_
            except Exception as exc:  # noqa: BLE001
                eval_fwd[other] = dict(error=repr(exc)[:200])
            finally:
                fn_.fused_eval["on"] = True
            torch.cuda.empty_cache()

    parity = None
    if not args.no_parity:
        try:
            parity = parity_block(w, g, args.precision, dev)
        except Exception as exc:  # noqa: BLE001
            parity = dict(error=repr(exc)[:300])

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        r = cpu_reference_rate(w, steps=args.cpu_steps, warmup=1)
        cpu = dict(value=r["value"], unit=UNIT, cores=r["cores"], kind=r["kind"], sample=r["sample"] % wname)

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_step,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype={"fp32": "f32", "tf32": "tf32", "bf16": "bf16"}[args.precision],
                data="synthetic",
                config=dict(workload=wname, per_gpu_batch=B, global_batch=B * world, n_max=w["n_max"], f0=w["f0"], hidden=w["hidden"],
                            encoder_dim=w["encoder_dim"], occupancy=args.occupancy, precision=args.precision,
                            precision_engine={"fp32": "exact fp32 FFMA", "bf16": "tcgen05 kind::f16, bf16 hi+lo split operands (4 passes forward, 3 backward, 1 for dW)",
                                              "tf32": "same bf16-split tcgen05 engine as 'bf16' (no kind::tf32 MMA is issued)"}[args.precision],
                            parallelism="dp%d" % world,
                            optimizer="Adam(fused)", runtime=runtime,
                            head="openchem_b200 OpenChemMLP + criterion (libocgcn head kernels)" if args.head == "fused" else "torch modules",
                            l2_policy="inputs+saved state per step (>1 GB at c3) exceed the 126 MB L2; no explicit flush",
                            workload_note="BASELINE config 3 (batch 4096 per B200) at every N: weak scaling of the headline workload, so the per-N "
                                          "values are like for like; BASELINE config 4 (per-GPU batch 2048 at 2/4/8 GPUs) is measured in the same job "
                                          "under `c4_ddp` (N > 1) and under other_workloads.c4_ddp (N = 1)"),
                e2e=dict(value=world * B * args.steps / (ms_e2e * 1e-3), unit=UNIT, h2d_bytes_per_step=e2e_h2d[0], d2h_bytes_per_step=4,
                         ms_per_step=ms_e2e / args.steps,
                         feeder="openchem_b200.compact.CompactGraphDataset (host cache, packed once) -> per-step sampler + collate into pinned "
                                "memory (loader.BackgroundIterator thread, 2 ahead) -> loader.DevicePrefetcher (copy stream, one batch ahead) -> "
                                + ("TrainStep captured on the bit words: ocgcn_encoder_forward_compact reads them directly (no dense x / adj on the "
                                   "device)" if cstepper is not None else "ocgcn_expand_compact on the device") + "; all inside the timed region",
                         via_dense_expand=(dict(value=world * B * args.steps / (ms_e2e_expand * 1e-3), ms_per_step=ms_e2e_expand / args.steps,
                                                what="same feed, words expanded by ocgcn_expand_compact into the dense step's inputs")
                                           if ms_e2e_expand else None),
                         loss_readback="every step's loss is copied to pinned host memory and read there, one step delayed "
                                       "(openchem_b200.step.DelayedLoss); the last read is inside the timed region"),
                e2e_dense=dict(value=world * B * args.steps / (ms_e2e_dense * 1e-3), unit=UNIT, h2d_bytes_per_step=h2d_dense, d2h_bytes_per_step=4,
                               ms_per_step=ms_e2e_dense / args.steps,
                               format="the reference's host format: dense fp32 adj / features / labels from pinned memory (Graph2Label.cast_inputs layout)"),
                gpu_launches=launches, clocks=clocks, roofline=roofline,
                step_roofline=dict(frac=step_frac, t_roof_us=B * t_roof_mol * 1e6, flops_per_molecule=work["flops"],
                                   bytes_per_molecule=work["bytes"], tensor_peak_tflops=tensor_peak, hbm_gbs=peaks["hbm_gbs"],
                                   peak_source=peaks["source"] + ", sustained bf16",
                                   step_dram_bytes=step_traffic,
                                   traffic_ratio=(step_traffic / (B * work["bytes"])) if step_traffic else None),
                kernels=shares, eval_forward=eval_fwd, parity=parity, replica_check=replica_check, single_gpu_same_workload=local_only, c4_ddp=c4_line, runtime_eager=eager,
                cpu_baseline=cpu, other_workloads=others)
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
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workloads and the eager-runtime measurement")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity block (fp64 oracle on the timed configuration)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # Headline: BASELINE config 3 (batch 4096 per B200) at every N - weak scaling, so the driver's per-N values compare the same
    # per-GPU work. BASELINE config 4 as stated (per-GPU batch 2048 at 2/4/8 GPUs, gradient all-reduce) is measured in the same job and
    # reported under `c4_ddp` with its own single-GPU figure; `--workload c4_ddp` makes it the headline.
    wname = args.workload or "c3_synth"
    w = WORKLOADS[wname]
    if args.impl == "reference":
        run_reference_arm(args, w, wname)
    else:
        run_native(args, w, wname)


if __name__ == "__main__":
    main()
