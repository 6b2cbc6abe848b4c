"""Single-kernel eval forward (csrc/ocgcn_fused.cuh: the whole layer stack of a tile inside one CTA, activations on the SM): results
must equal the layer-by-layer eval forward BIT FOR BIT (same arithmetic in the same order), match the fp64 oracle at the class
tolerance, and a backward pass requested after it must give the gradients of the layered path."""
import numpy as np
import pytest

from _util import build_encoder, random_params, rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

SHAPES = [
    dict(B=256, N=50, F0=33, hidden=[128] * 5, E=128, occ="molecular"),    # C2 (Tox21 shape): 128 tiles, one per CTA
    dict(B=256, N=85, F0=33, hidden=[128] * 5, E=128, occ="molecular"),    # C1 (logP shape): one molecule per tile, 256 tiles on 148 CTAs
    dict(B=700, N=64, F0=64, hidden=[128] * 4, E=128, occ="full"),         # C3 widths, 350 tiles: several tiles per CTA (weight ring wraps)
    dict(B=9, N=40, F0=9, hidden=[7, 13, 100], E=5, occ="molecular"),      # odd widths, partial chunks, ragged last tile
    dict(B=3, N=128, F0=20, hidden=[64, 128], E=36, occ="molecular"),      # N = 128
]


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.parametrize("shape", SHAPES)
def test_fused_eval_equals_layered_eval_and_oracle(shape):
    from oracle import gcn_oracle as orc
    from openchem_b200 import functional
    from openchem_b200.synthetic import make_graph_batch
    dev = _dev()
    g = make_graph_batch(shape["B"], shape["N"], shape["F0"], occupancy=shape["occ"], seed=13)
    p = random_params([shape["F0"]] + shape["hidden"], shape["E"], seed=4)
    cfg = {"input_size": shape["F0"], "encoder_dim": shape["E"], "n_layers": len(shape["hidden"]), "hidden_size": list(shape["hidden"])}
    x, adj = torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["adj"]).to(dev)
    enc = build_encoder(cfg, p, dev, precision="bf16").eval()
    assert functional.fused_eval["on"]
    n0 = functional.launch_counter["n"]
    out_f = enc((x, adj))
    assert functional.launch_counter["n"] - n0 == 3          # adj_pack, wprep, the fused kernel
    functional.fused_eval["on"] = False
    try:
        out_l = enc((x, adj))
    finally:
        functional.fused_eval["on"] = True
    torch.cuda.synchronize()
    assert torch.equal(out_f, out_l), float((out_f - out_l).abs().max())
    O = orc.encoder_forward(g["x"].astype(np.float64), g["adj"].astype(np.float64), orc.cast_params(p, np.float64), training=False)[0]
    assert rel_l2(out_f.detach().cpu().numpy(), O) < 1e-3
    # buffers untouched in eval mode
    assert int(enc.graph_convolutions[0].bn.num_batches_tracked) == 0


def test_fused_eval_with_hub_atoms_and_float_adjacency():
    """Rows that do not fit a neighbour record take the dense scan inside the fused kernel."""
    from openchem_b200 import functional
    from openchem_b200.synthetic import make_graph_batch
    dev = _dev()
    g = make_graph_batch(8, 30, 12, occupancy="full", seed=4)
    adj = g["adj"].copy()
    adj[:, 3, :20] = 1.0
    adj[:, :20, 3] = 1.0                                   # hub atom with 20 neighbours
    adj[1] = 1.0                                           # fully connected molecule
    adj[2] = np.random.RandomState(0).rand(30, 30).astype(np.float32)     # float adjacency
    p = random_params([12, 64, 32], 16, seed=2)
    cfg = {"input_size": 12, "encoder_dim": 16, "n_layers": 2, "hidden_size": [64, 32]}
    enc = build_encoder(cfg, p, dev, precision="bf16").eval()
    x, a = torch.from_numpy(g["x"]).to(dev), torch.from_numpy(adj).to(dev)
    out_f = enc((x, a))
    functional.fused_eval["on"] = False
    try:
        out_l = enc((x, a))
    finally:
        functional.fused_eval["on"] = True
    assert torch.equal(out_f, out_l), float((out_f - out_l).abs().max())


def test_backward_after_fused_eval_forward_rebuilds_the_saved_state():
    """evaluate() in the reference runs without no_grad: a backward after the eval forward must still work (gcn_encoder in eval mode)."""
    from openchem_b200 import functional
    from openchem_b200.synthetic import make_graph_batch
    dev = _dev()
    g = make_graph_batch(40, 50, 33, occupancy="molecular", seed=6)
    p = random_params([33, 128, 128], 64, seed=1)
    cfg = {"input_size": 33, "encoder_dim": 64, "n_layers": 2, "hidden_size": [128, 128]}
    x, adj = torch.from_numpy(g["x"]).to(dev), torch.from_numpy(g["adj"]).to(dev)
    dO = torch.randn(40, 64, device=dev)
    grads = []
    for fused in (True, False):
        enc = build_encoder(cfg, p, dev, precision="bf16").eval()
        functional.fused_eval["on"] = fused
        try:
            out = enc((x, adj))
            out.backward(dO)
        finally:
            functional.fused_eval["on"] = True
        grads.append([q.grad.clone() for q in enc.parameters()])
    for a, b in zip(*grads):
        assert torch.equal(a, b)
