"""Not a test: localise a parity failure by comparing saved intermediates with the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from _util import build_encoder, rel_l2
from oracle import gcn_oracle as orc
from openchem_b200 import functional
from openchem_b200.synthetic import make_graph_batch, random_params

def run(B, seed=11):
    shape = dict(B=B, N=64, F0=64, hidden=[128] * 4, E=128, occ="full")
    dev = torch.device("cuda:0")
    gb = make_graph_batch(shape["B"], shape["N"], shape["F0"], occupancy=shape["occ"], seed=seed)
    widths = [shape["F0"]] + shape["hidden"]
    p = random_params(widths, shape["E"], seed=5)
    dO = np.random.RandomState(3).randn(shape["B"], shape["E"]).astype(np.float32)
    p64 = orc.cast_params(p, np.float64)
    x64, a64 = gb["x"].astype(np.float64), gb["adj"].astype(np.float64)
    O, cache, rms, rvs = orc.encoder_forward(x64, a64, p64, training=True)
    gr = orc.encoder_backward(dO.astype(np.float64), cache)
    cfg = {"input_size": shape["F0"], "encoder_dim": shape["E"], "n_layers": 4, "hidden_size": list(shape["hidden"])}
    functional.debug_keep["on"] = True
    res = []
    for rep in range(2):
        enc = build_encoder(cfg, p, dev)
        enc.train()
        tx, ta = torch.from_numpy(gb["x"]).to(dev), torch.from_numpy(gb["adj"]).to(dev)
        out = enc((tx, ta))
        out.backward(torch.from_numpy(dO).to(dev))
        torch.cuda.synchronize()
        res.append([g.weight.grad.cpu().numpy() for g in enc.graph_convolutions])
    print("B", B, "run-to-run bitwise equal:", [bool(np.array_equal(a, b)) for a, b in zip(*res)])
    for l in range(4):
        Z = functional.saved_tensor("Z", l).cpu().numpy().reshape(B, 64, -1)
        arg = functional.saved_tensor("argmax", l).cpu().numpy().reshape(B, 64, -1)
        S = functional.saved_tensor("S", l).cpu().numpy().reshape(B, 64, -1)
        c = cache["layers"][l]
        Zo = (c["Yhat"] / c["invstd"]).reshape(B, 64, -1)   # Z - mu
        mism = np.argwhere(arg != c["arg"])
        # a mismatch only matters where the routed adjacency entry is non-zero for either choice
        real = [(b, i, ch) for b, i, ch in mism if a64[b, i, arg[b, i, ch]] != 0 or a64[b, i, c["arg"][b, i, ch]] != 0]
        print(f"layer {l}: S err {rel_l2(S, c['S']):.2e}  Z-mu err {rel_l2(Z - Z.reshape(-1, Z.shape[-1]).mean(0), Zo):.2e}  argmax mismatches {len(mism)} (gradient-carrying {len(real)})  dW err {rel_l2(res[0][l], gr['dW'][l]):.2e}")
        for b, i, ch in real[:5]:
            T = c["T"]
            print("   ", (b, i, ch), "ours", arg[b, i, ch], "oracle", c["arg"][b, i, ch], "T", T[b, arg[b, i, ch], ch], T[b, c["arg"][b, i, ch], ch])

for B in (8, 96):
    run(B)
