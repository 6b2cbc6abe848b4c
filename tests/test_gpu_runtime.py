"""Behaviour around the kernels on a B200: strided inputs, anomaly mode, threads, determinism, varying batch sizes, the
CUDA-graph step runtime and the prefetching feeder. Numerical parity proper lives in test_gpu_parity.py."""
import threading

import numpy as np
import pytest

from _util import build_encoder, random_params, rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _case(B=24, N=30, F0=16, hidden=(32, 24), E=12, seed=2, occ="molecular"):
    from openchem_b200.synthetic import make_graph_batch
    g = make_graph_batch(B, N, F0, occupancy=occ, seed=seed)
    p = random_params([F0] + list(hidden), E, seed=seed + 1)
    cfg = {"input_size": F0, "encoder_dim": E, "n_layers": len(hidden), "hidden_size": list(hidden)}
    return g, p, cfg


def _grads(enc):
    return [q.grad.detach().clone() for q in enc.parameters()]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_strided_inputs_equal_contiguous(precision):
    dev = _dev()
    g, p, cfg = _case()
    enc = build_encoder(cfg, p, dev, precision=precision).train()
    x, adj = torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["adj"]).to(dev)
    out_c = enc((x, adj))
    out_c.sum().backward()
    gc = _grads(enc)
    enc2 = build_encoder(cfg, p, dev, precision=precision).train()
    big = torch.zeros(x.shape[0], x.shape[1], x.shape[2] + 5, device=dev)
    big[:, :, 2:2 + x.shape[2]] = x
    xs = big[:, :, 2:2 + x.shape[2]]                       # inner stride 1, row pitch != F0
    adjs = adj.transpose(1, 2).contiguous().transpose(1, 2)   # same values, swapped strides
    assert not xs.is_contiguous() and not adjs.is_contiguous()
    out_s = enc2((xs, adjs))
    out_s.sum().backward()
    assert torch.equal(out_c, out_s)
    for a, b in zip(gc, _grads(enc2)):
        assert torch.equal(a, b)


def test_bitwise_reproducible_and_varying_batch():
    dev = _dev()
    g, p, cfg = _case(B=40)
    x, adj = torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["adj"]).to(dev)
    runs = []
    for _ in range(2):
        enc = build_encoder(cfg, p, dev, precision="bf16").train()
        o = enc((x, adj))
        o.square().sum().backward()
        runs.append((o.detach().clone(), _grads(enc)))
    assert torch.equal(runs[0][0], runs[1][0])
    for a, b in zip(runs[0][1], runs[1][1]):
        assert torch.equal(a, b)                            # fixed-order reductions: no floating-point atomics anywhere
    # the last batch of an epoch is smaller (data/utils.py:101-109 has no drop_last): B changes call to call
    enc = build_encoder(cfg, p, dev, precision="bf16").train()
    for b in (40, 7, 1, 33):
        o = enc((x[:b], adj[:b]))
        assert o.shape == (b, cfg["encoder_dim"]) and torch.isfinite(o).all()
        o.sum().backward()
    assert int(enc.graph_convolutions[0].bn.num_batches_tracked) == 4


def test_anomaly_mode_and_eval_without_no_grad():
    dev = _dev()
    g, p, cfg = _case()
    enc = build_encoder(cfg, p, dev, precision="bf16")
    x, adj = torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["adj"]).to(dev)
    with torch.autograd.set_detect_anomaly(True):          # openchem_model.py:103 wraps every train step in it
        enc.train()
        enc((x, adj)).sum().backward()
    enc.eval()                                             # evaluate()/predict() run without no_grad (openchem_model.py:218-348)
    rm = enc.graph_convolutions[0].bn.running_mean.clone()
    o = enc((x, adj))
    assert o.requires_grad and torch.equal(rm, enc.graph_convolutions[0].bn.running_mean)
    o.sum().backward()                                     # eval-mode backward (running statistics) is defined too


def test_two_host_threads_two_streams():
    """DataParallel drives replicas from Python threads (run.py:238): the library must be re-entrant."""
    dev = _dev()
    g, p, cfg = _case(B=48, N=50, F0=33, hidden=(128, 128), E=64)
    x, adj = torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["adj"]).to(dev)
    ref_enc = build_encoder(cfg, p, dev, precision="bf16").train()
    ref = ref_enc((x, adj)).detach().clone()
    outs, errs = [None, None], []

    def work(k):
        try:
            s = torch.cuda.Stream(device=dev)
            enc = build_encoder(cfg, p, dev, precision="bf16").train()
            with torch.cuda.stream(s):
                for _ in range(5):
                    o = enc((x, adj))
                    o.sum().backward()
                s.synchronize()
            outs[k] = o.detach().clone()
        except Exception as exc:  # noqa: BLE001
            errs.append(exc)

    ts = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    assert torch.equal(outs[0], ref) and torch.equal(outs[1], ref)


def test_graph_step_matches_eager_steps():
    from openchem_b200.step import TrainStep
    dev = _dev()
    g, p, cfg = _case(B=64, N=50, F0=33, hidden=(128, 128), E=64)
    x, adj = torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["adj"]).to(dev)
    y = torch.randn(64, 1, generator=torch.Generator().manual_seed(0)).to(dev)

    def make():
        torch.manual_seed(0)
        enc = build_encoder(cfg, p, dev, precision="bf16")
        head = torch.nn.Linear(64, 1).to(dev)
        return torch.nn.Sequential(enc, head).train()

    class Wrap(torch.nn.Module):
        def __init__(self, seq):
            super().__init__()
            self.seq = seq

        def forward(self, inp):
            return self.seq[1](self.seq[0](inp))

    crit = torch.nn.MSELoss()
    mk = lambda ps: torch.optim.Adam(ps, lr=1e-3, fused=True, capturable=True)  # noqa: E731
    m_graph, m_eager = Wrap(make()), Wrap(make())
    before = {k: v.clone() for k, v in m_graph.state_dict().items()}
    st_g = TrainStep(m_graph, crit, mk, (x, adj, y), use_graph=True, warmup=2)
    st_e = TrainStep(m_eager, crit, mk, (x, adj, y), use_graph=False, warmup=2)
    assert st_g.graph is not None, st_g.graph_error
    # construction (warm-up steps + capture) leaves the model and the optimizer exactly as they were handed over
    for k, v in m_graph.state_dict().items():
        assert torch.equal(v, before[k]), k
    assert all(float(t.abs().sum()) == 0.0 for st in st_g.opt.state.values() for t in st.values() if isinstance(t, torch.Tensor))
    # ... so the first call is training step 1: it equals one plain eager torch step from the same state
    m_plain = Wrap(make())
    opt = mk([q for q in m_plain.parameters()])
    loss = crit(m_plain((x, adj)), y)
    loss.backward()
    opt.step()
    for it in range(3):
        lg, le = float(st_g(x, adj, y)), float(st_e(x, adj, y))
        assert lg == le
        if it == 0:
            assert lg == float(loss)
            for a, b in zip(m_graph.parameters(), m_plain.parameters()):
                assert torch.equal(a, b)
    for a, b in zip(m_graph.parameters(), m_eager.parameters()):
        assert torch.equal(a, b)


def test_prefetcher_feeds_batches_in_order():
    from openchem_b200.loader import DevicePrefetcher
    dev = _dev()
    host = []
    for k in range(5):
        host.append(tuple(torch.full(s, float(k)).pin_memory() for s in ((4, 6, 3), (4, 6, 6), (4, 1))))
    seen = []
    for xb, adjb, yb in DevicePrefetcher(host, dev):
        assert xb.is_cuda and xb.shape == (4, 6, 3)
        seen.append((float(xb[0, 0, 0]), float(adjb[-1, -1, -1]), float(yb[0, 0])))
    assert seen == [(float(k),) * 3 for k in range(5)]


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("bf16", 1e-2)])
def test_training_trajectory_matches_reference_op_sequence(precision, tol):
    """A short Adam training run (encoder + linear head, MSE) on the CUDA path follows the same loss trajectory as the
    reference's eager op sequence (oracle/torch_restatement.py, fp32 on the CPU) from identical weights and batches:
    parameters, BatchNorm running statistics and optimizer state all evolve through the kernels' outputs."""
    from oracle import torch_restatement as tr
    dev = _dev()
    g, p, cfg = _case(B=48, N=30, F0=16, hidden=(32, 32), E=24, seed=5)
    rng = np.random.RandomState(0)
    y_np = rng.randn(48, 1).astype(np.float32)
    Wh = (rng.randn(1, 24) * 0.2).astype(np.float32)

    # CUDA path
    enc = build_encoder(cfg, p, dev, precision=precision).train()
    head = torch.nn.Linear(24, 1).to(dev)
    with torch.no_grad():
        head.weight.copy_(torch.from_numpy(Wh))
        head.bias.zero_()
    opt = torch.optim.Adam(list(enc.parameters()) + list(head.parameters()), lr=2e-3)
    x, adj, y = (torch.from_numpy(a).to(dev) for a in (g["x"], g["adj"], y_np))
    losses_gpu = []
    for _ in range(8):
        opt.zero_grad()
        loss = torch.nn.functional.mse_loss(head(enc((x, adj))), y)
        loss.backward()
        opt.step()
        losses_gpu.append(float(loss.detach()))

    # reference op sequence on the CPU (fp32), same initial state
    pt = tr.params_to_torch(p)
    hw = torch.tensor(Wh, requires_grad=True)
    hb = torch.zeros(1, requires_grad=True)
    leaves = [t for k in ("W", "b", "gamma", "beta") for t in pt[k]] + [pt["Wd"], pt["bd"], hw, hb]
    opt_c = torch.optim.Adam(leaves, lr=2e-3)
    xc, ac, yc = torch.from_numpy(g["x"]), torch.from_numpy(g["adj"]), torch.from_numpy(y_np)
    losses_cpu = []
    for _ in range(8):
        opt_c.zero_grad()
        loss = torch.nn.functional.mse_loss(torch.nn.functional.linear(tr.encoder(xc, ac, pt, True), hw, hb), yc)
        loss.backward()
        opt_c.step()
        losses_cpu.append(float(loss.detach()))
    rel = max(abs(a - b) / max(abs(b), 1e-12) for a, b in zip(losses_gpu, losses_cpu))
    assert losses_cpu[-1] < losses_cpu[0]                      # it trains
    assert rel < tol, (losses_gpu, losses_cpu)
    # running_var, not running_mean: the conv bias is a null direction of the loss (BatchNorm removes it), its gradient is
    # rounding noise that Adam normalises to +-lr steps, so bias and batch mean random-walk differently on any two
    # implementations (the reference on two machines included) while everything the loss sees agrees
    rv_gpu = enc.graph_convolutions[0].bn.running_var.cpu().numpy()
    assert rel_l2(rv_gpu, pt["running_var"][0].numpy()) < 10 * tol


def test_delayed_loss_returns_every_value_in_order():
    from openchem_b200.step import DelayedLoss
    dev = _dev()
    reader = DelayedLoss(dev)
    static = torch.zeros((), device=dev)          # like TrainStep.loss: one device scalar overwritten every step
    got = []
    for t in range(7):
        static.fill_(float(t) + 0.5)
        got.append(reader.push(static))
    got.append(reader.flush())
    assert got[0] is None and got[1:] == [t + 0.5 for t in range(7)]


def test_module_left_on_the_cpu_or_cast_raises_cleanly():
    """CUDA inputs with a module whose parameters are on the host / in another dtype: a clear RuntimeError before any kernel
    runs (the kernels take raw pointers), like the reference's ATen device / dtype errors - and the context stays usable."""
    from _util import build_encoder, random_params
    from openchem_b200.layers.gcn import GraphConvolution
    from openchem_b200.synthetic import make_graph_batch
    dev = torch.device("cuda:0")
    g = make_graph_batch(4, 10, 8, seed=1)
    x, adj = torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["adj"]).to(dev)
    cfg = {"input_size": 8, "encoder_dim": 6, "n_layers": 2, "hidden_size": [12, 12]}
    enc = build_encoder(cfg, random_params([8, 12, 12], 6, seed=0), "cpu")
    with pytest.raises(RuntimeError, match="is on cpu"):
        enc((x, adj))
    enc = enc.to(dev)
    enc.graph_convolutions[1].bn.running_var.data = enc.graph_convolutions[1].bn.running_var.data.cpu()   # one stray buffer
    with pytest.raises(RuntimeError, match="running_var"):
        enc((x, adj))
    with pytest.raises(RuntimeError, match="dtype"):
        build_encoder(cfg, random_params([8, 12, 12], 6, seed=0), dev).double()((x, adj))
    with pytest.raises(RuntimeError, match="is on cpu"):
        GraphConvolution(8, 5)(x, adj)
    out = build_encoder(cfg, random_params([8, 12, 12], 6, seed=0), dev)((x, adj))      # the context is still healthy
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
