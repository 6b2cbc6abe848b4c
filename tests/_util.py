"""Shared helpers for the parity tests: golden loading, oracle <-> module parameter plumbing."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

ENC_CASES = ["enc_reftest", "enc_odd", "enc_full", "enc_nonbin", "enc_c2small", "enc_wide"]
LAYER_CASES = ["layer_bias", "layer_nobias", "layer_float"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def golden_params(g, dtype=np.float64):
    L = len(g["hidden"])
    p = dict(W=[], b=[], gamma=[], beta=[], running_mean=[], running_var=[])
    for l in range(L):
        for k in p:
            p[k].append(g[f"{k}{l}"].astype(dtype))
    p["Wd"] = g["Wd"].astype(dtype)
    p["bd"] = g["bd"].astype(dtype)
    return p


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


def build_encoder(cfg, params, device, precision="fp32"):
    """Our GraphCNNEncoder with the given numpy parameters loaded (state_dict route, like a reference checkpoint)."""
    import torch
    from openchem_b200.modules.encoders.gcn_encoder import GraphCNNEncoder
    enc = GraphCNNEncoder(dict(cfg, precision=precision), True)
    sd = {}
    for l in range(len(params["W"])):
        sd[f"graph_convolutions.{l}.weight"] = torch.from_numpy(np.asarray(params["W"][l], dtype=np.float32))
        sd[f"graph_convolutions.{l}.bias"] = torch.from_numpy(np.asarray(params["b"][l], dtype=np.float32))
        sd[f"graph_convolutions.{l}.bn.weight"] = torch.from_numpy(np.asarray(params["gamma"][l], dtype=np.float32))
        sd[f"graph_convolutions.{l}.bn.bias"] = torch.from_numpy(np.asarray(params["beta"][l], dtype=np.float32))
        sd[f"graph_convolutions.{l}.bn.running_mean"] = torch.from_numpy(np.asarray(params["running_mean"][l], dtype=np.float32))
        sd[f"graph_convolutions.{l}.bn.running_var"] = torch.from_numpy(np.asarray(params["running_var"][l], dtype=np.float32))
        sd[f"graph_convolutions.{l}.bn.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    sd["dense.weight"] = torch.from_numpy(np.asarray(params["Wd"], dtype=np.float32))
    sd["dense.bias"] = torch.from_numpy(np.asarray(params["bd"], dtype=np.float32))
    enc.load_state_dict(sd, strict=True)
    return enc.to(device)


from openchem_b200.synthetic import random_params  # noqa: E402,F401
