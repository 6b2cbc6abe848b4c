"""N > 1 host logic on CPU: world_size-2 gloo. The hot path has no collective of its own; what must hold across ranks is
(1) shards are disjoint and cover the batch, (2) DDP's averaged gradients over the encoder's parameter layout equal the mean
of the per-shard reference gradients, (3) BatchNorm statistics stay per-rank while rank 0's running buffers are broadcast
(run.py:235-236 semantics), (4) timings reduce as max over ranks. The forward used here is the CPU torch restatement of the
reference (oracle/, test infrastructure) bound to OUR module's parameters — the CUDA path itself cannot run without a GPU."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _CpuEncoder(torch.nn.Module):
    """Our GraphCNNEncoder's parameters/buffers, forward = reference op sequence on CPU."""

    def __init__(self, enc):
        super().__init__()
        self.enc = enc

    def forward(self, x, adj):
        from oracle import torch_restatement as tr
        gcs = list(self.enc.graph_convolutions)
        p = dict(W=[g.weight for g in gcs], b=[g.bias for g in gcs], gamma=[g.bn.weight for g in gcs], beta=[g.bn.bias for g in gcs],
                 running_mean=[g.bn.running_mean for g in gcs], running_var=[g.bn.running_var for g in gcs],
                 Wd=self.enc.dense.weight, bd=self.enc.dense.bias)
        return tr.encoder(x, adj, p, self.training)


def _worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method="env://", world_size=world, rank=rank)
    from openchem_b200 import dist as ocd
    from openchem_b200.modules.encoders.gcn_encoder import GraphCNNEncoder
    from openchem_b200.synthetic import make_graph_batch
    assert ocd.env_world() == (world, rank, rank)
    B, N, F0 = 6, 7, 12
    g = make_graph_batch(B, N, F0, occupancy="molecular", seed=3)          # the same global batch on every rank
    lo, hi = ocd.shard_bounds(B, world, rank)
    x, adj = torch.from_numpy(g["x"][lo:hi]), torch.from_numpy(g["adj"][lo:hi])
    torch.manual_seed(100 + rank)                                           # different init per rank: DDP must broadcast rank 0's
    enc = GraphCNNEncoder({"input_size": F0, "encoder_dim": 5, "n_layers": 2, "hidden_size": [8, 6]}, False)
    model = ocd.wrap_ddp(_CpuEncoder(enc))
    model.train()
    out = model(x, adj)
    out.sum().backward()
    res = {k: v.grad.detach().numpy().copy() for k, v in enc.named_parameters()}
    res["w0"] = enc.graph_convolutions[0].weight.detach().numpy().copy()
    res["rm0_after"] = enc.graph_convolutions[0].bn.running_mean.numpy().copy()
    model(x, adj)                                                           # 2nd forward: buffers re-broadcast from rank 0 first
    res["shard"] = np.array([lo, hi])
    res["tmax"] = np.array([ocd.max_over_ranks(10.0 + rank)])
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **res)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_ddp_semantics(tmp_path):
    from oracle import torch_restatement as tr
    from openchem_b200 import dist as ocd
    from openchem_b200.modules.encoders.gcn_encoder import GraphCNNEncoder
    from openchem_b200.synthetic import make_graph_batch
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [dict(np.load(os.path.join(str(tmp_path), "rank%d.npz" % k))) for k in range(world)]
    # shards: disjoint, ordered, covering
    assert r[0]["shard"].tolist() == [0, 3] and r[1]["shard"].tolist() == [3, 6]
    assert [ocd.shard_bounds(7, 3, k) for k in range(3)] == [(0, 3), (3, 5), (5, 7)]
    assert ocd.shard_seed(1234, 3) == 1237
    # DDP broadcast rank 0's initial weights; averaged grads identical on both ranks
    np.testing.assert_array_equal(r[0]["w0"], r[1]["w0"])
    keys = [k for k in r[0] if k not in ("w0", "rm0_after", "shard", "tmax")]
    for k in keys:
        np.testing.assert_allclose(r[0][k], r[1][k], rtol=0, atol=0)
    # ... and equal to the mean of the per-shard reference gradients (single process, same weights)
    torch.manual_seed(100)
    enc = GraphCNNEncoder({"input_size": 12, "encoder_dim": 5, "n_layers": 2, "hidden_size": [8, 6]}, False)
    g = make_graph_batch(6, 7, 12, occupancy="molecular", seed=3)
    ref = {k: 0.0 for k in keys}
    rms = []
    for lo, hi in ((0, 3), (3, 6)):
        e = _CpuEncoder(enc)
        e.train()
        enc.zero_grad()
        for gc in enc.graph_convolutions:
            gc.bn.reset_running_stats()
        e(torch.from_numpy(g["x"][lo:hi]), torch.from_numpy(g["adj"][lo:hi])).sum().backward()
        for k, v in enc.named_parameters():
            ref[k] = ref[k] + v.grad.detach().numpy() / 2
        rms.append(enc.graph_convolutions[0].bn.running_mean.numpy().copy())
    for k in keys:
        np.testing.assert_allclose(r[0][k], ref[k], rtol=1e-5, atol=1e-6)
    # BatchNorm statistics are per rank (not SyncBN): each rank's running mean comes from its own shard
    np.testing.assert_allclose(r[0]["rm0_after"], rms[0], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(r[1]["rm0_after"], rms[1], rtol=1e-5, atol=1e-7)
    assert not np.allclose(rms[0], rms[1])
    # timings reduce as max over ranks
    assert r[0]["tmax"][0] == 11.0 and r[1]["tmax"][0] == 11.0
    assert tr is not None
