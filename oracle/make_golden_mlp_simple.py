"""Generate tests/golden/mlp_simple_*.npz from the UNMODIFIED reference OpenChemMLPSimple (TEST INFRASTRUCTURE ONLY).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_mlp_simple.py

Imports openchem.modules.mlp.openchem_mlp.OpenChemMLPSimple (openchem_mlp.py:64-111) from /root/reference, builds it under a
fixed torch seed (so the fixture also pins the parameter-initialisation RNG order, with and without 'init': 'xavier_uniform'),
runs forward + autograd backward in fp32 and fp64 on seeded inputs and stores weights, inputs and results.
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

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from openchem.modules.mlp.openchem_mlp import OpenChemMLPSimple  # noqa: E402
from openchem.utils.utils import identity  # noqa: E402

assert os.path.abspath(sys.modules["openchem.modules.mlp.openchem_mlp"].__file__).startswith(REF), "reference not imported"

OUT = os.path.join(ROOT, "tests", "golden")
ACT_FN = {"identity": identity, "relu": F.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh}
CASES = {
    "mlp_simple_xavier": dict(seed=31, B=70, widths=[128, 128, 1], acts=["relu", "identity"], init="xavier_uniform"),
    "mlp_simple_default": dict(seed=32, B=45, widths=[40, 36, 50, 7], acts=["tanh", "relu", "sigmoid"], init=None),
    "mlp_simple_single": dict(seed=33, B=9, widths=[16, 5], acts=["tanh"], init=None),
}


def build(case):
    torch.manual_seed(case["seed"])
    w = case["widths"]
    params = {"input_size": w[0], "n_layers": len(w) - 1, "hidden_size": list(w[1:]), "activation": [ACT_FN[a] for a in case["acts"]]}
    if case["init"]:
        params["init"] = case["init"]
    return OpenChemMLPSimple(params)


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, case in CASES.items():
        rng = np.random.RandomState(case["seed"])
        w = case["widths"]
        inp = np.tanh(rng.randn(case["B"], w[0])).astype(np.float32)
        gout = rng.randn(case["B"], w[-1]).astype(np.float32)
        out = dict(inp=inp, gout=gout, widths=np.asarray(w), acts=np.asarray(case["acts"]), seed=np.asarray(case["seed"]),
                   init=np.asarray(case["init"] or ""))
        for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            mlp = build(case).to(dtype)
            if tag == "f32":
                for i, l in enumerate(mlp.layers):
                    out[f"W{i}"], out[f"b{i}"] = l.weight.detach().numpy().copy(), l.bias.detach().numpy().copy()
            x = torch.from_numpy(inp).to(dtype).requires_grad_(True)
            y = mlp(x)
            (y * torch.from_numpy(gout).to(dtype)).sum().backward()
            out[f"out_{tag}"], out[f"dx_{tag}"] = y.detach().numpy(), x.grad.numpy()
            for i, l in enumerate(mlp.layers):
                out[f"dW{i}_{tag}"], out[f"db{i}_{tag}"] = l.weight.grad.numpy(), l.bias.grad.numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, "ok", {k: v.shape for k, v in out.items() if k.startswith("out_")})


if __name__ == "__main__":
    main()
