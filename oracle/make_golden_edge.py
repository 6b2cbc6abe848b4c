"""Generate tests/golden/edge_*.npz from the UNMODIFIED reference GraphEdgeAttentionEncoder (TEST INFRASTRUCTURE ONLY).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_edge.py        (authoring container: needs /root/reference)

The reference's forward hard-codes `.cuda()` on two temporaries (edge_attention_encoder.py:47-48); in this GPU-less container
`torch.Tensor.cuda` is patched to the identity for the duration of the run (same arithmetic, host tensors). Runs forward +
autograd in fp32 and fp64, asserts oracle/edge_attention_oracle.py agrees with the fp64 run to 1e-10, stores inputs, weights
and the reference's results.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("OPENCHEM_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.append(ROOT)

import torch  # noqa: E402

from openchem.modules.encoders.edge_attention_encoder import GraphEdgeAttentionEncoder  # noqa: E402

assert os.path.abspath(sys.modules["openchem.modules.encoders.edge_attention_encoder"].__file__).startswith(REF), "reference not imported"
torch.Tensor.cuda = lambda self, *a, **k: self

from oracle import edge_attention_oracle as eo  # noqa: E402
from oracle import gcn_oracle as orc  # noqa: E402
from openchem_b200.synthetic import make_graph_batch  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CASES = {
    "edge_small": dict(B=5, N=12, F0=10, hidden=[16, 12], sizes=[3, 2], E=8, seed=41),
    "edge_wide":  dict(B=3, N=40, F0=33, hidden=[128, 64], sizes=[4], E=32, seed=42),
}


def edge_attributes(g, A, rng):
    """(B,N,N,A): channel 0 is positive exactly on bonds + self loops (what utils/graph.py's one-hot bond attributes give);
    the others are arbitrary real scores (softmax'd by the encoder)."""
    adj = g["adj"]
    e = rng.randn(*adj.shape, A).astype(np.float32)
    e[..., 0] = np.where(adj > 0, 0.5 + rng.rand(*adj.shape), -rng.rand(*adj.shape)).astype(np.float32)
    return e


def run_case(name, c):
    rng = np.random.RandomState(c["seed"])
    g = make_graph_batch(c["B"], c["N"], c["F0"], occupancy="molecular", seed=17)
    A, L = sum(c["sizes"]), len(c["hidden"])
    edge = edge_attributes(g, A, rng)
    dO = rng.randn(c["B"], c["E"]).astype(np.float32)
    torch.manual_seed(5)
    enc = GraphEdgeAttentionEncoder({"input_size": c["F0"], "encoder_dim": c["E"], "n_layers": L, "hidden_size": list(c["hidden"]),
                                     "edge_attr_sizes": list(c["sizes"])}, False)
    with torch.no_grad():       # exercise the BatchNorm affine parameters and running statistics
        for gc in enc.graph_conv:
            gc.bn.weight.uniform_(0.5, 1.5)
            gc.bn.bias.uniform_(-0.3, 0.3)
            gc.bn.running_mean.uniform_(-0.2, 0.2)
            gc.bn.running_var.uniform_(0.5, 1.5)
    state0 = {k: v.clone() for k, v in enc.state_dict().items()}
    out = dict(x=g["x"], edge_attr=edge, dO=dO, hidden=np.asarray(c["hidden"]), sizes=np.asarray(c["sizes"]))
    for k, v in state0.items():
        out["sd_" + k] = v.numpy()
    for tag, dt in (("ref32", torch.float32), ("ref64", torch.float64)):
        enc.load_state_dict(state0)
        enc = enc.to(dt).train()
        enc.zero_grad()
        o = enc((torch.from_numpy(g["x"]).to(dt), torch.from_numpy(edge).to(dt)))
        o.backward(torch.from_numpy(dO).to(dt))
        out[tag + "_out_train"] = o.detach().numpy()
        for k, p in enc.named_parameters():
            out[tag + "_grad_" + k] = p.grad.numpy().copy()
        for k, b in enc.named_buffers():
            out[tag + "_buf_" + k] = b.detach().numpy().copy()
        enc.eval()
        out[tag + "_out_eval"] = enc((torch.from_numpy(g["x"]).to(dt), torch.from_numpy(edge).to(dt))).detach().numpy()
    # the oracle must reproduce the reference's fp64 run
    n = L * A
    p = dict(W=[state0["graph_conv.%d.weight" % k].double().numpy() for k in range(n)], b=[state0["graph_conv.%d.bias" % k].double().numpy() for k in range(n)],
             gamma=[state0["graph_conv.%d.bn.weight" % k].double().numpy() for k in range(n)], beta=[state0["graph_conv.%d.bn.bias" % k].double().numpy() for k in range(n)],
             running_mean=[state0["graph_conv.%d.bn.running_mean" % k].double().numpy() for k in range(n)],
             running_var=[state0["graph_conv.%d.bn.running_var" % k].double().numpy() for k in range(n)],
             Wd=state0["dense.weight"].double().numpy(), bd=state0["dense.bias"].double().numpy())
    O, cache, rms, rvs = eo.forward(g["x"].astype(np.float64), edge.astype(np.float64), p, L, training=True)
    gr = eo.backward(dO.astype(np.float64), cache)
    errs = [orc.rel_l2(O, out["ref64_out_train"]), orc.rel_l2(gr["dWd"], out["ref64_grad_dense.weight"])]
    for k in range(n):
        errs += [orc.rel_l2(gr["dW"][k], out["ref64_grad_graph_conv.%d.weight" % k]), orc.rel_l2(gr["dgamma"][k], out["ref64_grad_graph_conv.%d.bn.weight" % k]),
                 orc.rel_l2(rms[k], out["ref64_buf_graph_conv.%d.bn.running_mean" % k])]
    assert max(errs) < 1e-10, (name, errs)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "ok: oracle vs reference fp64 max rel-L2 %.1e" % max(errs))


if __name__ == "__main__":
    for name, c in CASES.items():
        run_case(name, c)
