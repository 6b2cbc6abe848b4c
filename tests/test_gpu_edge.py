"""GraphEdgeAttentionEncoder and the stand-alone tanh + neighbour max-pool op on the GPU, against
fixtures written by the UNMODIFIED reference encoder (oracle/make_golden_edge.py: fp32 and fp64 runs) and the numpy oracle."""
import numpy as np
import pytest

from _util import load_golden, rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.parametrize("strided", [False, True])
def test_pool_op_vs_oracle(strided):
    from oracle import gcn_oracle as orc
    from openchem_b200.functional import tanh_neighbour_max_pool
    dev = _dev()
    rs = np.random.RandomState(2)
    B, N, D = 5, 37, 150
    y = rs.randn(B, N, D).astype(np.float32)
    y[0, 3] = y[0, 5]                                   # exact ties -> first index
    adj = (rs.rand(B, N, N) < 0.15).astype(np.float32)
    adj[1] = rs.rand(N, N).astype(np.float32)           # float adjacency
    adj[2] = 1.0                                        # fully connected (no zero candidate)
    adj[3, 4] = 0.0                                     # an empty row
    dP = rs.randn(B, N, D).astype(np.float32)
    T = np.tanh(y.astype(np.float64))
    P, arg = orc.neighbour_max_pool(T, adj.astype(np.float64))
    dT = orc.neighbour_max_pool_backward(dP.astype(np.float64), adj.astype(np.float64), arg)
    ty = torch.from_numpy(y).to(dev).requires_grad_(True)
    ta = torch.from_numpy(adj).to(dev)
    if strided:
        ta = ta.transpose(1, 2).contiguous().transpose(1, 2)        # same values, column-major molecules
    out = tanh_neighbour_max_pool(ty, ta)
    out.backward(torch.from_numpy(dP).to(dev))
    assert rel_l2(out.detach().cpu().numpy(), P) < 1e-6
    assert rel_l2(ty.grad.cpu().numpy(), dT * (1 - T * T)) < 1e-5
    with pytest.raises(RuntimeError):
        tanh_neighbour_max_pool(ty, ta.requires_grad_(True))


@pytest.mark.parametrize("case", ["edge_small", "edge_wide"])
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 1e-2)])
def test_edge_attention_encoder_vs_reference_golden(case, precision, tol):
    from openchem_b200.modules.encoders.edge_attention_encoder import GraphEdgeAttentionEncoder
    dev = _dev()
    g = load_golden(case)
    hidden, sizes = [int(h) for h in g["hidden"]], [int(s) for s in g["sizes"]]
    enc = GraphEdgeAttentionEncoder({"input_size": int(g["x"].shape[2]), "encoder_dim": int(g["dO"].shape[1]), "n_layers": len(hidden),
                                     "hidden_size": hidden, "edge_attr_sizes": sizes, "precision": precision}, True)
    sd = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd_")}
    enc.load_state_dict(sd, strict=True)                 # the reference's state_dict, key for key
    enc = enc.to(dev).train()
    x, e = torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["edge_attr"]).to(dev)
    out = enc((x, e))
    out.backward(torch.from_numpy(g["dO"]).to(dev))
    torch.cuda.synchronize()
    errs = {"out_train": rel_l2(out.detach().cpu().numpy(), g["ref64_out_train"])}
    for k, p in enc.named_parameters():
        ref = g["ref64_grad_" + k]
        if ".bias" in k and ".bn." not in k and k.startswith("graph_conv"):      # analytically zero (BatchNorm removes the mean)
            assert np.abs(p.grad.cpu().numpy()).max() < 1e-4 * (np.abs(g["ref64_grad_" + k.replace("bias", "weight")]).max() + 1e-30) + 1e-6
            continue
        errs["grad_" + k] = rel_l2(p.grad.cpu().numpy(), ref)
    for k, b in enc.named_buffers():
        if b.dtype.is_floating_point:
            errs["buf_" + k] = rel_l2(b.cpu().numpy(), g["ref64_buf_" + k])
        else:
            assert int(b) == int(g["ref64_buf_" + k])
    enc.eval()
    errs["out_eval"] = rel_l2(enc((x, e)).detach().cpu().numpy(), g["ref64_out_eval"])
    worst = max(errs, key=errs.get)
    assert errs[worst] < tol, (worst, errs[worst], sorted(errs.items(), key=lambda kv: -kv[1])[:4])
    if precision == "fp32":      # and the reference's own fp32 run
        assert rel_l2(out.detach().cpu().numpy(), g["ref32_out_train"]) < 2e-5
