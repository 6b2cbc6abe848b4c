"""The oracle (oracle/gcn_oracle.py) against the fixtures produced by the unmodified reference, plus
reference-independent checks (hand-computed tie rule, finite differences). CPU only."""
import os
import numpy as np
import pytest

from _util import ENC_CASES, LAYER_CASES, golden_params, load_golden, rel_l2
from oracle import gcn_oracle as orc


def _oracle_encoder(g, dtype):
    p = golden_params(g, dtype)
    x, adj, dO = g["x"].astype(dtype), g["adj"].astype(dtype), g["dO"].astype(dtype)
    O, cache, rms, rvs = orc.encoder_forward(x, adj, p, training=True)
    gr = orc.encoder_backward(dO, cache)
    Oe = orc.encoder_forward(x, adj, dict(p, running_mean=rms, running_var=rvs), training=False)[0]
    res = dict(out_train=O, out_eval=Oe, dWd=gr["dWd"], dbd=gr["dbd"])
    for l in range(len(g["hidden"])):
        res.update({f"dW{l}": gr["dW"][l], f"db{l}": gr["db"][l], f"dgamma{l}": gr["dgamma"][l], f"dbeta{l}": gr["dbeta"][l],
                    f"rm{l}": rms[l], f"rv{l}": rvs[l]})
    return res


@pytest.mark.parametrize("case", ENC_CASES)
def test_oracle_encoder_fp64_matches_reference(case):
    g = load_golden(case)
    res = _oracle_encoder(g, np.float64)
    for k, v in res.items():
        ref = g["ref64_" + k]
        if k.startswith("db") and k != "dbd":
            assert np.abs(v - ref).max() < 1e-9      # analytically zero, never a relative tolerance
        else:
            assert rel_l2(v, ref) < 1e-10, k


@pytest.mark.parametrize("case", ENC_CASES)
def test_oracle_encoder_fp32_close_to_reference_fp32(case):
    g = load_golden(case)
    res = _oracle_encoder(g, np.float32)
    for k, v in res.items():
        if k.startswith("db") and k != "dbd":
            continue
        assert rel_l2(v, g["ref32_" + k]) < 2e-5, k
        assert v.dtype == np.float32


@pytest.mark.parametrize("case", LAYER_CASES)
def test_oracle_layer_matches_reference(case):
    g = load_golden(case)
    f = lambda k: g[k].astype(np.float64) if k in g else None
    Y, cache, rm, rv = orc.graph_conv_forward(f("x"), f("adj"), f("W"), f("b"), f("gamma"), f("beta"), f("running_mean"), f("running_var"), True)
    dx, dW, db, dg, dbt = orc.graph_conv_backward(f("dY"), cache)
    for v, k in ((Y, "y_train"), (dx, "dx"), (dW, "dW"), (dg, "dgamma"), (dbt, "dbeta"), (rm, "rm"), (rv, "rv")):
        assert rel_l2(v, g["ref64_" + k]) < 1e-10, k
    Ye, cache_e, _, _ = orc.graph_conv_forward(f("x"), f("adj"), f("W"), f("b"), f("gamma"), f("beta"), rm, rv, False)
    assert rel_l2(Ye, g["ref64_y_eval"]) < 1e-10
    assert rel_l2(orc.graph_conv_backward(f("dY"), cache_e)[0], g["ref64_dx_eval"]) < 1e-10


def test_pool_first_index_tie_rule_and_zero_candidate():
    # row 0: neighbours {0,1}, entry (0,2) is zero -> candidates T0, T1, 0*T2
    T = np.array([[[0.5, -0.5], [0.5, -0.25], [0.25, -0.75]]])
    adj = np.ones((1, 3, 3))
    adj[0, 0, 2] = 0.0
    P, arg = orc.neighbour_max_pool(T, adj)
    assert P[0, 0, 0] == 0.5 and arg[0, 0, 0] == 0          # exact tie -> first index
    assert P[0, 0, 1] == 0.0 and arg[0, 0, 1] == 2          # all neighbours negative -> the zero entry wins (ReLU-like clamp)
    assert P[0, 1, 1] == -0.25 and arg[0, 1, 1] == 1        # full row: no zero candidate, negative max survives
    dP = np.ones_like(P)
    dT = orc.neighbour_max_pool_backward(dP, adj, arg)
    assert dT[0, 2, 1] == 0.0                                # routed through a zero adjacency entry -> no gradient
    assert dT.sum() == pytest.approx(3 * 2 - 1)


def test_padded_rows_are_analytic():
    """a padded row's pre-BN value is exactly the bias, its pooled value 0."""
    from openchem_b200.synthetic import make_graph_batch
    g = make_graph_batch(3, 8, 12, occupancy="molecular", seed=3)
    n = g["n_atoms"]
    assert n.min() < 8
    rng = np.random.RandomState(0)
    W, b = rng.randn(12, 5), rng.randn(5)
    x, adj = g["x"].astype(np.float64), g["adj"].astype(np.float64)
    Z = np.matmul(np.matmul(adj, x), W) + b
    P, _ = orc.neighbour_max_pool(np.tanh(Z), adj)
    for m in range(3):
        assert np.array_equal(Z[m, n[m]:], np.broadcast_to(b, Z[m, n[m]:].shape))
        assert np.all(P[m, n[m]:] == 0)


def test_oracle_backward_matches_finite_differences():
    from openchem_b200.synthetic import make_graph_batch
    g = make_graph_batch(3, 6, 7, occupancy="molecular", seed=1)
    rng = np.random.RandomState(2)
    widths, E = [7, 5, 4], 3
    p = dict(W=[rng.randn(widths[i], widths[i + 1]) * 0.4 for i in range(2)], b=[rng.randn(widths[i + 1]) * 0.1 for i in range(2)],
             gamma=[rng.uniform(0.5, 1.5, widths[i + 1]) for i in range(2)], beta=[rng.uniform(-0.3, 0.3, widths[i + 1]) for i in range(2)],
             running_mean=[np.zeros(widths[i + 1]) for i in range(2)], running_var=[np.ones(widths[i + 1]) for i in range(2)],
             Wd=rng.randn(E, widths[-1]) * 0.4, bd=rng.randn(E) * 0.1)
    x, adj = g["x"].astype(np.float64), g["adj"].astype(np.float64)
    dO = rng.randn(3, E)

    def loss(pp):
        return float((orc.encoder_forward(x, adj, pp, True)[0] * dO).sum())

    O, cache, _, _ = orc.encoder_forward(x, adj, p, True)
    gr = orc.encoder_backward(dO, cache)
    h = 1e-6
    for key, grad in (("W", gr["dW"]), ("gamma", gr["dgamma"]), ("beta", gr["dbeta"])):
        for l in range(2):
            flat = p[key][l].reshape(-1)
            for idx in rng.choice(flat.size, size=min(4, flat.size), replace=False):
                old = flat[idx]
                flat[idx] = old + h
                lp = loss(p)
                flat[idx] = old - h
                lm = loss(p)
                flat[idx] = old
                assert (lp - lm) / (2 * h) == pytest.approx(grad[l].reshape(-1)[idx], rel=1e-4, abs=1e-7), (key, l, idx)
    for key, grad in (("Wd", gr["dWd"]), ("bd", gr["dbd"])):
        flat = p[key].reshape(-1)
        for idx in range(min(3, flat.size)):
            old = flat[idx]
            flat[idx] = old + h
            lp = loss(p)
            flat[idx] = old - h
            lm = loss(p)
            flat[idx] = old
            assert (lp - lm) / (2 * h) == pytest.approx(grad.reshape(-1)[idx], rel=1e-4, abs=1e-7)


def test_synthetic_generator_layout():
    from openchem_b200.synthetic import make_graph_batch
    g = make_graph_batch(16, 20, 33, occupancy="molecular", seed=5)
    adj, x, n = g["adj"], g["x"], g["n_atoms"]
    assert adj.dtype == np.float32 and x.dtype == np.float32
    assert set(np.unique(adj)) <= {0.0, 1.0}
    assert np.array_equal(adj, np.swapaxes(adj, 1, 2))
    for m in range(16):
        assert np.all(np.diag(adj[m])[: n[m]] == 1) and np.all(adj[m, n[m]:] == 0) and np.all(adj[m, :, n[m]:] == 0)
        assert np.all(x[m, : n[m]].sum(1) == 5) and np.all(x[m, n[m]:] == 0)          # 5 one-hot groups
        deg = adj[m].sum(1) - np.diag(adj[m])
        assert deg.max() <= 4
    g2 = make_graph_batch(16, 20, 33, occupancy="molecular", seed=5)
    assert np.array_equal(g2["adj"], adj) and np.array_equal(g2["x"], x)
    assert np.all(make_graph_batch(4, 9, 64, occupancy="full", seed=1)["n_atoms"] == 9)


@pytest.mark.parametrize("case", ["enc_reftest", "enc_odd", "enc_nonbin"])
def test_torch_restatement_matches_reference_fixture(case):
    """The eager-op replay used as bench.py's CPU baseline reproduces the reference's own fp32 results."""
    torch = pytest.importorskip("torch")
    from oracle import torch_restatement as tr
    g = load_golden(case)
    p = tr.params_to_torch(golden_params(g, np.float32))
    x, adj = torch.from_numpy(g["x"]), torch.from_numpy(g["adj"])
    out = tr.encoder(x, adj, p, True)
    out.backward(torch.from_numpy(g["dO"]))
    assert rel_l2(out.detach().numpy(), g["ref32_out_train"]) < 1e-6
    for l in range(len(g["hidden"])):
        assert rel_l2(p["W"][l].grad.numpy(), g[f"ref32_dW{l}"]) < 1e-5
        assert rel_l2(p["running_var"][l].numpy(), g[f"ref32_rv{l}"]) < 1e-6
    assert rel_l2(p["Wd"].grad.numpy(), g["ref32_dWd"]) < 1e-5
    assert rel_l2(tr.encoder(x, adj, p, False).detach().numpy(), g["ref32_out_eval"]) < 1e-6


def test_c_pool_scan_is_bit_identical_to_the_numpy_definition():
    """oracle/pool_scan.c (used for benchmark-sized parity cases) vs the numpy loops that define the oracle: ties, signed
    zeros, float adjacency, fully connected and empty rows, fp32 and fp64."""
    import subprocess
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import __graft_entry__ as ge
    if orc.pool_lib() is None:
        try:
            ge.build_oracle()
        except (OSError, subprocess.CalledProcessError) as e:   # no gcc: the numpy definition is used everywhere
            pytest.skip("cannot build oracle/pool_scan.c: %s" % e)
        orc._POOL_LIB = None
    assert orc.pool_lib() is not None
    rs = np.random.RandomState(0)
    for dt in (np.float64, np.float32):
        for B, N, D in ((5, 17, 9), (3, 64, 33), (2, 1, 4)):
            T = rs.randn(B, N, D).astype(dt)
            if N > 5:
                T[0, 3] = T[0, 5]                      # exact ties -> first index
            T[-1] = -np.abs(T[-1])                     # all negative: the 0 * T candidates (signed zeros) win
            adj = (rs.rand(B, N, N) < 0.2).astype(dt)
            adj[0, 0] = 0                              # empty row
            if B > 2:
                adj[1] = rs.rand(N, N).astype(dt)      # float adjacency
                adj[2] = 1                             # fully connected: no zero candidate
            P0, a0 = orc.neighbour_max_pool(T, adj, use_c=False)
            P1, a1 = orc.neighbour_max_pool(T, adj, use_c=True)
            assert np.array_equal(a0, a1) and np.array_equal(P0, P1)
            dP = rs.randn(B, N, D).astype(dt)
            d0 = orc.neighbour_max_pool_backward(dP, adj, a0, use_c=False)
            d1 = orc.neighbour_max_pool_backward(dP, adj, a0, use_c=True)
            assert np.array_equal(d0, d1)


@pytest.mark.parametrize("case", ["edge_small", "edge_wide"])
def test_edge_attention_oracle_matches_reference_fixture(case):
    """oracle/edge_attention_oracle.py vs the unmodified reference's fp64 run of GraphEdgeAttentionEncoder (make_golden_edge.py)."""
    from oracle import edge_attention_oracle as eo
    g = load_golden(case)
    L, A = len(g["hidden"]), int(g["sizes"].sum())
    n = L * A
    f64 = lambda k: g["sd_" + k].astype(np.float64)  # noqa: E731
    p = dict(W=[f64("graph_conv.%d.weight" % k) for k in range(n)], b=[f64("graph_conv.%d.bias" % k) for k in range(n)],
             gamma=[f64("graph_conv.%d.bn.weight" % k) for k in range(n)], beta=[f64("graph_conv.%d.bn.bias" % k) for k in range(n)],
             running_mean=[f64("graph_conv.%d.bn.running_mean" % k) for k in range(n)],
             running_var=[f64("graph_conv.%d.bn.running_var" % k) for k in range(n)], Wd=f64("dense.weight"), bd=f64("dense.bias"))
    O, cache, rms, rvs = eo.forward(g["x"].astype(np.float64), g["edge_attr"].astype(np.float64), p, L, training=True)
    gr = eo.backward(g["dO"].astype(np.float64), cache)
    assert rel_l2(O, g["ref64_out_train"]) < 1e-10
    assert rel_l2(gr["dWd"], g["ref64_grad_dense.weight"]) < 1e-10 and rel_l2(gr["dbd"], g["ref64_grad_dense.bias"]) < 1e-10
    for k in range(n):
        assert rel_l2(gr["dW"][k], g["ref64_grad_graph_conv.%d.weight" % k]) < 1e-10
        assert rel_l2(gr["dgamma"][k], g["ref64_grad_graph_conv.%d.bn.weight" % k]) < 1e-10
        assert rel_l2(gr["dbeta"][k], g["ref64_grad_graph_conv.%d.bn.bias" % k]) < 1e-10
        assert rel_l2(rvs[k], g["ref64_buf_graph_conv.%d.bn.running_var" % k]) < 1e-10
    Oe = eo.forward(g["x"].astype(np.float64), g["edge_attr"].astype(np.float64), dict(p, running_mean=rms, running_var=rvs), L, training=False)[0]
    assert rel_l2(Oe, g["ref64_out_eval"]) < 1e-10
