"""Generate tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).

Run in the authoring container (the reference cannot travel to the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

It imports openchem.layers.gcn.GraphConvolution and
openchem.modules.encoders.gcn_encoder.GraphCNNEncoder from /root/reference, runs
forward + autograd backward in fp32 and in fp64 ("fp64-promoted reference") on seeded
inputs/weights, asserts that oracle/gcn_oracle.py reproduces the fp64 run to 1e-10,
and stores inputs, weights and the reference results as small fixtures.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("OPENCHEM_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
# the reference must win over the repo's `openchem/` shadow package here
sys.path.insert(0, REF)
sys.path.append(ROOT)

import torch  # noqa: E402

from openchem.layers.gcn import GraphConvolution  # noqa: E402
from openchem.modules.encoders.gcn_encoder import GraphCNNEncoder  # noqa: E402

assert os.path.abspath(sys.modules["openchem.layers.gcn"].__file__).startswith(REF), "reference not imported"

from oracle import gcn_oracle as orc  # noqa: E402
from openchem_b200.synthetic import make_graph_batch  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

ENCODER_CASES = {
    # name: B, N, F0, hidden, E, occupancy, adj kind
    "enc_reftest":  dict(B=5, N=9, F0=12, hidden=[10, 10, 10], E=10, occ="molecular", adj="binary"),   # module_instantiation_test.py:36-43 shape
    "enc_odd":      dict(B=4, N=37, F0=33, hidden=[16, 24, 8], E=12, occ="molecular", adj="binary"),
    "enc_full":     dict(B=3, N=16, F0=64, hidden=[32, 32], E=20, occ="full", adj="binary"),
    "enc_nonbin":   dict(B=3, N=10, F0=12, hidden=[10, 14], E=6, occ="molecular", adj="float"),         # edge-attention style float adjacency
    "enc_c2small":  dict(B=6, N=50, F0=33, hidden=[128] * 5, E=128, occ="molecular", adj="binary"),    # C2 shape, small batch
    "enc_wide":     dict(B=2, N=70, F0=64, hidden=[136, 72], E=40, occ="molecular", adj="binary"),     # N>64 (two mask words), F>128
}
LAYER_CASES = {
    "layer_bias":    dict(B=4, N=11, Fin=12, Fout=10, bias=True, adj="binary"),
    "layer_nobias":  dict(B=3, N=20, Fin=33, Fout=40, bias=False, adj="binary"),
    "layer_float":   dict(B=3, N=9, Fin=10, Fout=130, bias=True, adj="float"),
}


def _adj_of(kind, g, rng):
    adj = g["adj"].copy()
    if kind == "float":
        # GraphEdgeAttentionEncoder feeds softmax'd edge channels (edge_attention_encoder.py:42-53): dense non-binary
        w = rng.rand(*adj.shape).astype(np.float32)
        adj = adj * w + 0.05 * rng.rand(*adj.shape).astype(np.float32) * (adj == 0)
        n = g["n_atoms"]
        idx = np.arange(adj.shape[1])
        valid = (idx[None, :] < n[:, None])
        adj = adj * valid[:, :, None] * valid[:, None, :]
    return adj.astype(np.float32)


def _params_from_encoder(enc):
    p = dict(W=[], b=[], gamma=[], beta=[], running_mean=[], running_var=[])
    for gc in enc.graph_convolutions:
        p["W"].append(gc.weight.detach().numpy().copy())
        p["b"].append(gc.bias.detach().numpy().copy())
        p["gamma"].append(gc.bn.weight.detach().numpy().copy())
        p["beta"].append(gc.bn.bias.detach().numpy().copy())
        p["running_mean"].append(gc.bn.running_mean.numpy().copy())
        p["running_var"].append(gc.bn.running_var.numpy().copy())
    p["Wd"] = enc.dense.weight.detach().numpy().copy()
    p["bd"] = enc.dense.bias.detach().numpy().copy()
    return p


def _run_reference_encoder(enc, x, adj, dO, dtype):
    """train step (fwd+bwd, running stats updated) followed by an eval forward."""
    enc = enc.to(dtype)
    enc.train()
    enc.zero_grad()
    tx, ta = torch.from_numpy(x).to(dtype), torch.from_numpy(adj).to(dtype)
    out = enc((tx, ta))
    out.backward(torch.from_numpy(dO).to(dtype))
    res = dict(out_train=out.detach().numpy().copy())
    for l, gc in enumerate(enc.graph_convolutions):
        res[f"dW{l}"] = gc.weight.grad.numpy().copy()
        res[f"db{l}"] = gc.bias.grad.numpy().copy()
        res[f"dgamma{l}"] = gc.bn.weight.grad.numpy().copy()
        res[f"dbeta{l}"] = gc.bn.bias.grad.numpy().copy()
        res[f"rm{l}"] = gc.bn.running_mean.numpy().copy()
        res[f"rv{l}"] = gc.bn.running_var.numpy().copy()
    res["dWd"] = enc.dense.weight.grad.numpy().copy()
    res["dbd"] = enc.dense.bias.grad.numpy().copy()
    enc.eval()
    res["out_eval"] = enc((tx, ta)).detach().numpy().copy()
    return res


def make_encoder_case(name, c, seed):
    rng = np.random.RandomState(seed)
    g = make_graph_batch(c["B"], c["N"], c["F0"], occupancy=c["occ"], seed=seed)
    x, adj = g["x"], _adj_of(c["adj"], g, rng)
    if c["adj"] == "float":
        x = x + 0.1 * rng.randn(*x.shape).astype(np.float32) * (x.sum(-1, keepdims=True) > 0)
    torch.manual_seed(seed)
    cfg = {"input_size": c["F0"], "encoder_dim": c["E"], "n_layers": len(c["hidden"]), "hidden_size": list(c["hidden"])}
    enc = GraphCNNEncoder(cfg, False)
    with torch.no_grad():          # randomise BN affine + running stats so they are exercised
        for gc in enc.graph_convolutions:
            gc.bn.weight.uniform_(0.5, 1.5)
            gc.bn.bias.uniform_(-0.3, 0.3)
            gc.bn.running_mean.uniform_(-0.2, 0.2)
            gc.bn.running_var.uniform_(0.5, 1.5)
    params = _params_from_encoder(enc)
    dO = rng.randn(c["B"], c["E"]).astype(np.float32)

    import copy
    ref32 = _run_reference_encoder(copy.deepcopy(enc), x, adj, dO, torch.float32)
    ref64 = _run_reference_encoder(copy.deepcopy(enc), x, adj, dO, torch.float64)

    # pin the oracle against the reference (fp64, so summation order does not matter)
    p64 = orc.cast_params(params, np.float64)
    O, cache, rms, rvs = orc.encoder_forward(x.astype(np.float64), adj.astype(np.float64), p64, training=True)
    grads = orc.encoder_backward(dO.astype(np.float64), cache)
    worst = orc.rel_l2(O, ref64["out_train"])
    for l in range(len(c["hidden"])):
        worst = max(worst, orc.rel_l2(grads["dW"][l], ref64[f"dW{l}"]), orc.rel_l2(grads["dgamma"][l], ref64[f"dgamma{l}"]),
                    orc.rel_l2(grads["dbeta"][l], ref64[f"dbeta{l}"]), orc.rel_l2(rms[l], ref64[f"rm{l}"]),
                    orc.rel_l2(rvs[l], ref64[f"rv{l}"]))
        assert np.abs(grads["db"][l] - ref64[f"db{l}"]).max() < 1e-9, "bias grad (analytically 0)"
    worst = max(worst, orc.rel_l2(grads["dWd"], ref64["dWd"]), orc.rel_l2(grads["dbd"], ref64["dbd"]))
    p64e = dict(p64, running_mean=rms, running_var=rvs)
    Oe, _, _, _ = orc.encoder_forward(x.astype(np.float64), adj.astype(np.float64), p64e, training=False)
    worst = max(worst, orc.rel_l2(Oe, ref64["out_eval"]))
    assert worst < 1e-10, f"{name}: oracle deviates from reference fp64: {worst}"
    gap32 = orc.rel_l2(ref32["out_train"], ref64["out_train"])
    print(f"{name:14s} oracle-vs-ref64 worst rel-L2 {worst:.2e};  ref32-vs-ref64 out {gap32:.2e}")

    blob = dict(x=x, adj=adj, dO=dO, hidden=np.asarray(c["hidden"]), Wd=params["Wd"], bd=params["bd"])
    for l in range(len(c["hidden"])):
        for k in ("W", "b", "gamma", "beta", "running_mean", "running_var"):
            blob[f"{k}{l}"] = params[k][l]
    for k, v in ref32.items():
        blob["ref32_" + k] = v
    for k, v in ref64.items():
        blob["ref64_" + k] = v
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **blob)


def make_layer_case(name, c, seed):
    rng = np.random.RandomState(seed)
    g = make_graph_batch(c["B"], c["N"], c["Fin"], occupancy="molecular", seed=seed)
    adj = _adj_of(c["adj"], g, rng)
    x = (g["x"] + 0.3 * rng.randn(*g["x"].shape)).astype(np.float32)      # layers deeper in a stack see dense activations
    torch.manual_seed(seed)
    gc = GraphConvolution(c["Fin"], c["Fout"], bias=c["bias"])
    with torch.no_grad():
        gc.bn.weight.uniform_(0.5, 1.5)
        gc.bn.bias.uniform_(-0.3, 0.3)
        gc.bn.running_mean.uniform_(-0.2, 0.2)
        gc.bn.running_var.uniform_(0.5, 1.5)
    W = gc.weight.detach().numpy().copy()
    b = gc.bias.detach().numpy().copy() if c["bias"] else None
    gamma, beta = gc.bn.weight.detach().numpy().copy(), gc.bn.bias.detach().numpy().copy()
    rm, rv = gc.bn.running_mean.numpy().copy(), gc.bn.running_var.numpy().copy()
    dY = rng.randn(c["B"], c["N"], c["Fout"]).astype(np.float32)
    blob = dict(x=x, adj=adj, dY=dY, W=W, gamma=gamma, beta=beta, running_mean=rm, running_var=rv)
    if b is not None:
        blob["b"] = b
    import copy
    for tag, dt in (("ref32", torch.float32), ("ref64", torch.float64)):
        m = copy.deepcopy(gc).to(dt)
        m.train()
        tx = torch.from_numpy(x).to(dt).requires_grad_(True)
        ta = torch.from_numpy(adj).to(dt)
        y = m(tx, ta)
        y.backward(torch.from_numpy(dY).to(dt))
        blob[tag + "_y_train"] = y.detach().numpy().copy()
        blob[tag + "_dx"] = tx.grad.numpy().copy()
        blob[tag + "_dW"] = m.weight.grad.numpy().copy()
        if c["bias"]:
            blob[tag + "_db"] = m.bias.grad.numpy().copy()
        blob[tag + "_dgamma"] = m.bn.weight.grad.numpy().copy()
        blob[tag + "_dbeta"] = m.bn.bias.grad.numpy().copy()
        blob[tag + "_rm"] = m.bn.running_mean.numpy().copy()
        blob[tag + "_rv"] = m.bn.running_var.numpy().copy()
        m.eval()
        tx2 = torch.from_numpy(x).to(dt).requires_grad_(True)
        ye = m(tx2, ta)
        ye.backward(torch.from_numpy(dY).to(dt))
        blob[tag + "_y_eval"] = ye.detach().numpy().copy()
        blob[tag + "_dx_eval"] = tx2.grad.numpy().copy()
    f64 = lambda a: None if a is None else a.astype(np.float64)
    Y, cache, nrm, nrv = orc.graph_conv_forward(f64(x), f64(adj), f64(W), f64(b), f64(gamma), f64(beta), f64(rm), f64(rv), True)
    dx, dW, db, dg, dbt = orc.graph_conv_backward(f64(dY), cache)
    worst = max(orc.rel_l2(Y, blob["ref64_y_train"]), orc.rel_l2(dx, blob["ref64_dx"]), orc.rel_l2(dW, blob["ref64_dW"]),
                orc.rel_l2(dg, blob["ref64_dgamma"]), orc.rel_l2(dbt, blob["ref64_dbeta"]),
                orc.rel_l2(nrm, blob["ref64_rm"]), orc.rel_l2(nrv, blob["ref64_rv"]))
    Ye, cache_e, _, _ = orc.graph_conv_forward(f64(x), f64(adj), f64(W), f64(b), f64(gamma), f64(beta), nrm, nrv, False)
    dxe = orc.graph_conv_backward(f64(dY), cache_e)[0]
    worst = max(worst, orc.rel_l2(Ye, blob["ref64_y_eval"]), orc.rel_l2(dxe, blob["ref64_dx_eval"]))
    assert worst < 1e-10, f"{name}: oracle deviates from reference fp64: {worst}"
    print(f"{name:14s} oracle-vs-ref64 worst rel-L2 {worst:.2e}")
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **blob)


def make_init_fixture():
    """Initial weights of the reference encoder under a fixed seed: our module must consume the RNG identically
    (dense first, then the layers; gcn_encoder.py:23-25, gcn.py:27-31) so --random_seed reproduces the same model."""
    cfg = {"input_size": 33, "encoder_dim": 24, "n_layers": 3, "hidden_size": [16, 40, 8]}
    torch.manual_seed(7)
    enc = GraphCNNEncoder(cfg, False)
    blob = {k: v.numpy().copy() for k, v in enc.state_dict().items()}
    blob["after_rand"] = torch.rand(4).numpy()          # RNG position after construction
    np.savez_compressed(os.path.join(OUT, "init_seed7.npz"), **blob)
    torch.manual_seed(9)
    gc = GraphConvolution(12, 20, bias=False)
    np.savez_compressed(os.path.join(OUT, "init_layer_seed9.npz"), weight=gc.weight.detach().numpy().copy(), after_rand=torch.rand(4).numpy())
    print("init fixtures written")


def check_tie_rule():
    """The reference's CPU max keeps the first index on exact ties; the oracle must as well."""
    T = np.zeros((1, 3, 2), dtype=np.float64)
    T[0, :, 0] = [0.5, 0.5, 0.25]
    T[0, :, 1] = [-0.5, -0.25, -0.75]
    adj = np.ones((1, 3, 3))
    adj[0, 0, 2] = 0.0
    tT, tA = torch.from_numpy(T), torch.from_numpy(adj)
    n, d = 3, 2
    res = tT.repeat(1, n, 1).view(-1, n, n, d) * tA.unsqueeze(3).expand(-1, n, n, d)     # gcn_encoder.py:46-49
    val, idx = res.max(dim=2)
    P, arg = orc.neighbour_max_pool(T, adj)
    assert np.array_equal(val.numpy(), P) and np.array_equal(idx.numpy(), arg), "tie rule mismatch"
    print("tie rule: first index, matches torch CPU max")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(4)
    check_tie_rule()
    make_init_fixture()
    for i, (name, c) in enumerate(ENCODER_CASES.items()):
        make_encoder_case(name, c, seed=100 + i)
    for i, (name, c) in enumerate(LAYER_CASES.items()):
        make_layer_case(name, c, seed=200 + i)
    print("fixtures written to", OUT)
