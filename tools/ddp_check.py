#!/usr/bin/env python
"""N-rank correctness of the data-parallel step runtime (openchem_b200.step.TrainStep: in-graph NCCL all-reduce(AVG) of the flat
gradient). Launch with torchrun (one rank per GPU). Every rank:
  1. runs K captured train steps on ITS shard of a seeded synthetic batch;
  2. checks that all replicas hold bit-identical parameters afterwards (all_gather of the flat parameters);
  3. replays the same K steps WITHOUT a process group on one GPU: per step the gradient of every shard is computed with the same
     kernels, the shard gradients are averaged in fp32 in rank order and one Adam step is taken - the N-rank parameters must
     agree with that single-GPU emulation (the reference semantics: DistributedDataParallel averages gradients, BatchNorm
     statistics stay per rank, run.py:233-236).
Rank 0 prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

from _util import build_encoder, random_params
from openchem_b200.step import TrainStep
from openchem_b200.synthetic import make_graph_batch

K, B, N, F0, HID, E = 3, 192, 50, 33, [128, 128, 128], 64


class Model(torch.nn.Module):
    def __init__(self, dev, precision):
        super().__init__()
        cfg = {"input_size": F0, "encoder_dim": E, "n_layers": len(HID), "hidden_size": HID}
        self.enc = build_encoder(cfg, random_params([F0] + HID, E, seed=3), dev, precision=precision)
        torch.manual_seed(1)
        self.head = torch.nn.Linear(E, 1).to(dev)

    def forward(self, inp):
        return self.head(self.enc(inp))


def shard(rank, dev):
    g = make_graph_batch(B, N, F0, occupancy="molecular", seed=100 + rank)
    y = np.random.RandomState(200 + rank).randn(B, 1).astype(np.float32)
    return tuple(torch.from_numpy(a).to(dev) for a in (g["x"], g["adj"], y))


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", init_method="env://", device_id=dev)
    precision = os.environ.get("OCB200_PRECISION", "bf16")
    crit = torch.nn.MSELoss()
    mk = lambda ps: torch.optim.Adam(ps, lr=1e-3, fused=True, capturable=True)  # noqa: E731
    model = Model(dev, precision).train()
    step = TrainStep(model, crit, mk, shard(rank, dev), use_graph=True, warmup=2)
    losses = [float(step()) for _ in range(K)]
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    allp = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(allp, flat)
    identical = all(torch.equal(a, allp[0]) for a in allp)

    # single-GPU emulation of the same K steps (no process group involved)
    emu = Model(dev, precision).train()
    replicas = [emu] + [Model(dev, precision).train() for _ in range(world - 1)]      # per-rank BatchNorm buffers
    opt = mk([p for p in emu.parameters()])
    shards = [shard(r, dev) for r in range(world)]
    for _ in range(K):
        grads = []
        for r in range(world):
            m = replicas[r]
            with torch.no_grad():
                for dst, src in zip(m.parameters(), emu.parameters()):
                    if dst is not src:
                        dst.copy_(src)
            m.zero_grad()
            x, adj, y = shards[r]
            crit(m((x, adj)), y).backward()
            grads.append([p.grad.detach().clone() for p in m.parameters()])
        for i, p in enumerate(emu.parameters()):
            acc = grads[0][i].clone()
            for r in range(1, world):
                acc += grads[r][i]
            p.grad = acc / world
        opt.step()
    flat_emu = torch.cat([p.detach().reshape(-1) for p in emu.parameters()])
    rel = float(torch.linalg.norm((flat - flat_emu).double()) / torch.linalg.norm(flat_emu.double()))
    # the analytically-zero GraphConvolution.bias gradients make Adam amplify round-off there (tests/test_gpu_dropin.py): compare
    # everything, and separately without those biases
    keep = torch.cat([torch.full((p.numel(),), 0.0 if (n.startswith("enc.graph_convolutions") and n.endswith(".bias") and ".bn." not in n) else 1.0)
                      for n, p in model.named_parameters()]).to(dev)
    rel_nb = float(torch.linalg.norm(((flat - flat_emu) * keep).double()) / torch.linalg.norm((flat_emu * keep).double()))
    ok = identical and rel_nb < 1e-4
    if rank == 0:
        print(json.dumps(dict(world=world, precision=precision, steps=K, graph=step.graph is not None, exchange=step.exchange_kind, replicas_bit_identical=identical,
                              rel_l2_vs_single_gpu_emulation=rel, rel_l2_without_zero_gradient_biases=rel_nb, losses_rank0=losses, ok=ok)), flush=True)
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
