"""Compact batch format on the device (ocgcn_expand_compact / ocgcn_pack_compact through the C ABI): bit-exact against the
dense reference-format tensors for ragged sizes, gathers, strided inputs; the encoder gives identical results from either
format."""
import numpy as np
import pytest
import torch

from openchem_b200 import compact
from _util import build_encoder, random_params
from openchem_b200.synthetic import make_graph_batch

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def _case(B, N, F0, hidden, E, seed):
    g = make_graph_batch(B, N, F0, occupancy="molecular", seed=seed)
    p = random_params([F0] + list(hidden), E, seed=seed + 1)
    cfg = {"input_size": F0, "encoder_dim": E, "n_layers": len(hidden), "hidden_size": list(hidden)}
    return g, p, cfg


def _rand_binary(B, N, F0, seed):
    rng = np.random.RandomState(seed)
    return (rng.rand(B, N, F0) < 0.2).astype(np.float32), (rng.rand(B, N, N) < 0.1).astype(np.float32)


@pytest.mark.parametrize("B,N,F0", [(1, 1, 1), (3, 37, 33), (5, 64, 64), (2, 85, 75), (2, 132, 33), (7, 50, 130), (4, 256, 64)])
def test_expand_is_bit_exact(B, N, F0):
    x, adj = _rand_binary(B, N, F0, B * 1000 + N)
    cb = compact.pack_graph_batch(x, adj).to(_dev())
    xe, ae = cb.expand()
    assert xe.dtype == torch.float32 and tuple(xe.shape) == (B, N, F0) and tuple(ae.shape) == (B, N, N)
    assert np.array_equal(xe.cpu().numpy(), x) and np.array_equal(ae.cpu().numpy(), adj)
    # into preallocated outputs (poisoned first)
    xo = torch.full((B, N, F0), 7.0, device=_dev())
    ao = torch.full((B, N, N), 7.0, device=_dev())
    cb.expand(out=(xo, ao))
    assert torch.equal(xo, xe) and torch.equal(ao, ae)


def test_expand_gather_from_resident_dataset():
    x, adj = _rand_binary(40, 30, 33, 11)
    y = torch.arange(40, dtype=torch.float32).reshape(40, 1)
    ds = compact.ResidentCompactDataset(compact.pack_graph_batch(x, adj, y), _dev())
    assert len(ds) == 40
    idx = torch.tensor([39, 0, 7, 7, 12], dtype=torch.int64)
    xb, ab, yb = ds.batch(idx)
    assert np.array_equal(xb.cpu().numpy(), x[idx.numpy()]) and np.array_equal(ab.cpu().numpy(), adj[idx.numpy()])
    assert torch.equal(yb.cpu(), y[idx])


@pytest.mark.parametrize("B,N,F0", [(3, 37, 33), (4, 64, 64), (2, 132, 70)])
def test_device_pack_matches_host_pack_and_handles_strides(B, N, F0):
    x, adj = _rand_binary(B, N, F0, 5)
    host = compact.pack_graph_batch(x, adj)
    xd, ad = torch.from_numpy(x).to(_dev()), torch.from_numpy(adj).to(_dev())
    dev = compact.pack_graph_batch_device(xd, ad)
    assert torch.equal(dev.x_bits.cpu(), host.x_bits) and torch.equal(dev.adj_bits.cpu(), host.adj_bits)
    # transposed (non-contiguous) adjacency view
    dev_t = compact.pack_graph_batch_device(xd, ad.transpose(1, 2))
    assert torch.equal(dev_t.adj_bits.cpu(), compact.pack_graph_batch(x, adj.transpose(0, 2, 1)).adj_bits)


def test_device_pack_rejects_non_binary():
    x, adj = _rand_binary(2, 16, 8, 1)
    adj[1, 3, 4] = 0.25
    with pytest.raises(ValueError):
        compact.pack_graph_batch_device(torch.from_numpy(x).to(_dev()), torch.from_numpy(adj).to(_dev()))


def _star_batch(B, N, F0, seed):
    """Molecules whose atom 0 is bonded to every other atom: more neighbours than a 16-byte record holds, so those rows take the
    kernels' mask / dense fallback paths (which read the bit words themselves on the compact route)."""
    g = make_graph_batch(B, N, F0, occupancy="full", seed=seed)
    g["adj"][:, 0, :] = 1.0
    g["adj"][:, :, 0] = 1.0
    return g


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("shape", ["small", "odd_words", "many_tiles", "star", "warp_pack"])
@pytest.mark.parametrize("training", [True, False])
def test_encoder_reads_compact_words_with_identical_results(precision, shape, training):
    """GraphCNNEncoder.forward(CompactGraphBatch) feeds the bit words to ocgcn_encoder_*_compact (no dense tensors): outputs and
    every gradient are bit-identical to the dense route, for the cooperative fused forward (few tiles), the layer-by-layer kernels
    (> 148 tiles), the single-kernel eval forward, word counts that are odd, and rows that overflow a neighbour record."""
    from openchem_b200 import functional
    B, N, F0, hidden, E = {"small": (12, 30, 33, (64, 64), 32), "odd_words": (9, 70, 75, (64, 128), 64),
                           "many_tiles": (330, 64, 64, (128, 128), 128), "star": (10, 40, 33, (64, 64), 32),
                           "warp_pack": (1030, 20, 16, (32,), 16)}[shape]      # B >= 1024: the warp-per-molecule adjacency packing
    g, p, cfg = _case(B=B, N=N, F0=F0, hidden=hidden, E=E, seed=2)
    if shape == "star":
        g = _star_batch(B, N, F0, 3)
    if shape == "warp_pack":        # a few hub atoms with more neighbours than a record holds, inside the large batch
        g["adj"][:3, 0, :] = 1.0
        g["adj"][:3, :, 0] = 1.0
    dev = _dev()
    enc = build_encoder(cfg, p, dev, precision=precision)
    enc.train(training)
    x, adj = torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["adj"]).to(dev)
    state = {k: v.clone() for k, v in enc.state_dict().items()}
    out_d = enc((x, adj))
    out_d.sum().backward()
    gd = [q.grad.clone() for q in enc.parameters()]
    stats_d = {k: v.clone() for k, v in enc.state_dict().items() if "running" in k}
    cb = compact.pack_graph_batch(g["x"], g["adj"]).to(dev)
    for direct, inp in ((True, cb), (True, (cb.x_bits, cb.adj_bits)), (False, cb)):
        enc.load_state_dict(state)
        enc.zero_grad()
        functional.compact_direct["on"] = direct
        try:
            n0 = functional.launch_counter["n"]
            out_c = enc(inp)
            out_c.sum().backward()
        finally:
            functional.compact_direct["on"] = True
        assert functional.launch_counter["n"] > n0
        assert torch.equal(out_c, out_d)
        for a, b in zip(gd, [q.grad for q in enc.parameters()]):
            assert torch.equal(a, b)
        for k, v in stats_d.items():
            assert torch.equal(enc.state_dict()[k], v), k


def test_compact_route_rejects_mismatched_words():
    g, p, cfg = _case(B=4, N=30, F0=33, hidden=(32,), E=16, seed=2)
    enc = build_encoder(cfg, p, _dev(), precision="bf16")
    cb = compact.pack_graph_batch(g["x"], g["adj"]).to(_dev())
    with pytest.raises(RuntimeError, match="compact batch needs"):
        enc((cb.x_bits[:, :, :1], cb.adj_bits))          # F0 = 33 needs two words per row
    with pytest.raises(RuntimeError, match="compact batch needs"):
        enc((cb.x_bits, cb.adj_bits[:, :-1]))


class _Model(torch.nn.Module):
    def __init__(self, enc, E):
        super().__init__()
        self.enc, self.head = enc, torch.nn.Linear(E, 1)

    def forward(self, inp):
        return self.head(self.enc(inp))


def test_train_step_load_compact_equals_dense_load():
    from openchem_b200.step import TrainStep
    g, p, cfg = _case(16, 20, 16, (32,), 16, seed=4)
    g2 = make_graph_batch(16, 20, 16, seed=9)
    dev = _dev()
    y = torch.from_numpy(np.random.RandomState(1).randn(16, 1).astype(np.float32))
    losses = []
    for use_compact in (False, True):
        torch.manual_seed(0)
        model = _Model(build_encoder(cfg, p, dev, precision="bf16"), 16).to(dev).train()
        ex = (torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["adj"]).to(dev), y.to(dev))
        step = TrainStep(model, torch.nn.MSELoss(), lambda ps: torch.optim.Adam(ps, lr=1e-3, fused=True, capturable=True), ex,
                         use_graph=True, warmup=1)
        assert step.graph is not None, step.graph_error
        run = []
        for data in (g2, g):
            if use_compact:
                step.load_compact(compact.pack_graph_batch(data["x"], data["adj"], y).to(dev))
                run.append(float(step()))
            else:
                run.append(float(step(torch.from_numpy(data["x"]).to(dev), torch.from_numpy(data["adj"]).to(dev), y.to(dev))))
        losses.append(run)
    assert losses[0] == losses[1], losses


def test_train_step_built_on_compact_words_equals_dense_step():
    """A TrainStep whose static inputs ARE the bit words (captured ocgcn_encoder_forward_compact / _backward_compact) takes the same
    steps, bit for bit, as the dense one."""
    from openchem_b200.step import TrainStep
    g, p, cfg = _case(16, 20, 16, (32,), 16, seed=4)
    g2 = make_graph_batch(16, 20, 16, seed=9)
    dev = _dev()
    y = torch.from_numpy(np.random.RandomState(1).randn(16, 1).astype(np.float32))
    losses = []
    for use_compact in (False, True):
        torch.manual_seed(0)
        model = _Model(build_encoder(cfg, p, dev, precision="bf16"), 16).to(dev).train()
        if use_compact:
            c0 = compact.pack_graph_batch(g["x"], g["adj"], y).to(dev)
            ex = (c0.x_bits, c0.adj_bits, c0.labels)
        else:
            ex = (torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["adj"]).to(dev), y.to(dev))
        step = TrainStep(model, torch.nn.MSELoss(), lambda ps: torch.optim.Adam(ps, lr=1e-3, fused=True, capturable=True), ex,
                         use_graph=True, warmup=1)
        assert step.graph is not None, step.graph_error
        assert (step.inputs[0].dtype == torch.int32) == use_compact
        run = []
        for data in (g2, g, g2):
            if use_compact:
                step.load_compact(compact.pack_graph_batch(data["x"], data["adj"], y).to(dev))
                run.append(float(step()))
            else:
                run.append(float(step(torch.from_numpy(data["x"]).to(dev), torch.from_numpy(data["adj"]).to(dev), y.to(dev))))
        losses.append(run)
    assert losses[0] == losses[1], losses


def test_gather_index_is_bounds_checked():
    """Host indices are validated before the upload; a CUDA index out of range is never dereferenced: the kernel writes zeros
    for that molecule and raises the batch's device flag (no out-of-bounds read, no wrapped int64)."""
    x, adj = _rand_binary(6, 12, 8, 3)
    ds = compact.ResidentCompactDataset(compact.pack_graph_batch(x, adj), _dev())
    xs, As, _ = ds.batch([4, 0, 5])
    assert torch.equal(xs.cpu(), torch.from_numpy(x[[4, 0, 5]])) and torch.equal(As.cpu(), torch.from_numpy(adj[[4, 0, 5]]))
    assert int(ds.data.last_index_oob.item()) == 0
    for bad in ([0, 6], [-1, 2], np.array([1, 2 ** 32 + 1], dtype=np.int64)):
        with pytest.raises(IndexError):
            ds.batch(bad)
    bad_dev = torch.tensor([1, 2 ** 32 + 2, 3], dtype=torch.int64, device=_dev())     # would wrap to 2 as int32
    xs, As, _ = ds.batch(bad_dev)
    assert int(ds.data.last_index_oob.item()) == 1
    assert torch.equal(xs[0].cpu(), torch.from_numpy(x[1])) and torch.equal(xs[2].cpu(), torch.from_numpy(x[3]))
    assert float(xs[1].abs().sum()) == 0.0 and float(As[1].abs().sum()) == 0.0
