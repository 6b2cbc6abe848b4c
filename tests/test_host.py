"""Host-side logic without a GPU: the C-ABI library loads and exports what include/ocgcn.h declares, the
nn.Module surface mirrors the reference (names, state_dict, RNG consumption, config validation, errors)."""
import ctypes
import os
import re

import numpy as np
import pytest

from _util import ROOT, load_golden

torch = pytest.importorskip("torch")


def test_library_exports_every_declared_symbol():
    from openchem_b200 import _lib
    header = open(os.path.join(ROOT, "include", "ocgcn.h")).read()
    declared = set(re.findall(r"\b(ocgcn_[a-z_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ocgcn_abi_version() == 5
    assert lib.ocgcn_status_string(0) == b"ok"
    assert b"unsupported" in lib.ocgcn_status_string(2)


def _desc(B=8, N=20, widths=(33, 16, 8), E=12):
    from openchem_b200 import _lib
    d = _lib.Desc()
    d.B, d.N, d.L, d.E = B, N, len(widths) - 1, E
    for i, f in enumerate(widths):
        d.F[i] = f
    d.training, d.has_bias, d.eps, d.momentum = 1, 1, 1e-5, 0.1
    return d


def test_query_workspace_and_error_codes_on_host():
    from openchem_b200 import _lib
    lib = _lib.load()
    sb, cb = ctypes.c_size_t(), ctypes.c_size_t()
    assert lib.ocgcn_query_workspace(ctypes.byref(_desc()), ctypes.byref(sb), ctypes.byref(cb)) == 0
    assert sb.value % 256 == 0 and cb.value % 256 == 0
    rows = 8 * 20
    assert sb.value >= rows * (16 + 8) * 4 + rows * (16 + 8)            # Z_l fp32 + u8 arg-max per layer
    s2 = ctypes.c_size_t()
    assert lib.ocgcn_query_workspace(ctypes.byref(_desc(B=16)), ctypes.byref(s2), ctypes.byref(cb)) == 0
    assert s2.value > sb.value
    assert lib.ocgcn_query_workspace(ctypes.byref(_desc(N=300)), ctypes.byref(sb), ctypes.byref(cb)) == 2   # N > 256 unsupported
    assert lib.ocgcn_query_workspace(ctypes.byref(_desc(B=0)), ctypes.byref(sb), ctypes.byref(cb)) == 1
    assert lib.ocgcn_query_workspace(ctypes.byref(_desc(widths=(33, 0, 8))), ctypes.byref(sb), ctypes.byref(cb)) == 1
    d = _desc()
    d.mode = 7
    assert lib.ocgcn_query_workspace(ctypes.byref(d), ctypes.byref(sb), ctypes.byref(cb)) == 1
    assert lib.ocgcn_query_workspace(None, ctypes.byref(sb), ctypes.byref(cb)) == 1
    # layer-only layout (E == 0, L == 1)
    assert lib.ocgcn_query_workspace(ctypes.byref(_desc(widths=(12, 10), E=0)), ctypes.byref(sb), ctypes.byref(cb)) == 0


def test_encoder_surface_matches_reference_contract():
    from openchem_b200.modules.encoders.gcn_encoder import GraphCNNEncoder
    from openchem_b200.modules.encoders.openchem_encoder import OpenChemEncoder
    cfg = {"input_size": 12, "encoder_dim": 10, "hidden_size": [10, 10, 10], "n_layers": 3}   # module_instantiation_test.py:36-43
    enc = GraphCNNEncoder(cfg, use_cuda=True)
    assert isinstance(enc, OpenChemEncoder)
    assert enc.n_layers == 3 and enc.hidden_size == [12, 10, 10, 10] and enc.dropout == 0
    assert enc.input_size == 12 and enc.encoder_dim == 10 and enc.params is cfg and enc.use_cuda is True
    assert isinstance(enc.dropout_layer, torch.nn.Dropout) and isinstance(enc.dense, torch.nn.Linear)
    assert GraphCNNEncoder.get_required_params() == {"n_layers": int, "hidden_size": list}
    assert "dropout" in GraphCNNEncoder.get_optional_params()
    sd = enc.state_dict()
    expect = {}
    widths = [12, 10, 10, 10]
    for l in range(3):
        expect[f"graph_convolutions.{l}.weight"] = (widths[l], widths[l + 1])
        for k in ("bias", "bn.weight", "bn.bias", "bn.running_mean", "bn.running_var"):
            expect[f"graph_convolutions.{l}.{k}"] = (widths[l + 1],)
        expect[f"graph_convolutions.{l}.bn.num_batches_tracked"] = ()
    expect["dense.weight"], expect["dense.bias"] = (10, 10), (10,)
    assert {k: tuple(v.shape) for k, v in sd.items()} == expect
    assert sd["graph_convolutions.0.bn.num_batches_tracked"].dtype == torch.int64
    assert repr(enc.graph_convolutions[0]).startswith("GraphConvolution (12 -> 10)")
    for p in enc.parameters():
        assert p.is_leaf and p.requires_grad


def test_config_validation_like_reference():
    from openchem_b200.modules.encoders.gcn_encoder import GraphCNNEncoder
    with pytest.raises(ValueError, match="n_layers parameter has to be specified"):
        GraphCNNEncoder({"input_size": 12, "encoder_dim": 10, "hidden_size": [10]}, False)
    with pytest.raises(ValueError, match="has to be of type"):
        GraphCNNEncoder({"input_size": 12, "encoder_dim": 10, "hidden_size": (10,), "n_layers": 1}, False)
    # like the reference, the base-class check resolves get_required_params() on the subclass
    # (openchem_encoder.py:11), so a missing input_size surfaces as the KeyError of openchem_encoder.py:16
    with pytest.raises(KeyError):
        GraphCNNEncoder({"encoder_dim": 10, "hidden_size": [10], "n_layers": 1}, False)
    with pytest.raises(AssertionError):
        GraphCNNEncoder({"input_size": 12, "encoder_dim": 10, "hidden_size": [10, 10], "n_layers": 1}, False)
    with pytest.raises(ValueError, match="has to be one of"):
        GraphCNNEncoder({"input_size": 12, "encoder_dim": 10, "hidden_size": [10], "n_layers": 1, "precision": "fp8"}, False)
    GraphCNNEncoder({"input_size": 12, "encoder_dim": 10, "hidden_size": [10], "n_layers": 1, "unknown_key": 3}, False)  # tolerated


def test_initialisation_consumes_rng_like_reference():
    from openchem_b200.layers.gcn import GraphConvolution
    from openchem_b200.modules.encoders.gcn_encoder import GraphCNNEncoder
    g = load_golden("init_seed7")
    torch.manual_seed(7)
    enc = GraphCNNEncoder({"input_size": 33, "encoder_dim": 24, "n_layers": 3, "hidden_size": [16, 40, 8]}, False)
    after = torch.rand(4).numpy()
    for k, v in enc.state_dict().items():
        assert np.array_equal(v.numpy(), g[k]), k
    assert np.array_equal(after, g["after_rand"])
    g = load_golden("init_layer_seed9")
    torch.manual_seed(9)
    gc = GraphConvolution(12, 20, bias=False)
    assert gc.bias is None and np.array_equal(gc.weight.detach().numpy(), g["weight"])
    assert np.array_equal(torch.rand(4).numpy(), g["after_rand"])


def test_reference_checkpoint_keys_load_with_module_prefix():
    """run.py:240-243 loads DataParallel/DDP checkpoints ('module.' prefix) with load_state_dict."""
    from openchem_b200.modules.encoders.gcn_encoder import GraphCNNEncoder
    g = load_golden("init_seed7")
    enc = GraphCNNEncoder({"input_size": 33, "encoder_dim": 24, "n_layers": 3, "hidden_size": [16, 40, 8]}, False)
    wrapped = torch.nn.Module()
    wrapped.module = torch.nn.Module()
    wrapped.module.Encoder = enc
    sd = {"module.Encoder." + k: torch.from_numpy(np.asarray(v)) for k, v in g.items() if k != "after_rand"}
    wrapped.load_state_dict(sd, strict=True)
    assert np.array_equal(enc.dense.weight.detach().numpy(), g["dense.weight"])


def test_cpu_tensors_raise_no_fallback():
    from openchem_b200.layers.gcn import GraphConvolution
    from openchem_b200.modules.encoders.gcn_encoder import GraphCNNEncoder
    enc = GraphCNNEncoder({"input_size": 12, "encoder_dim": 10, "hidden_size": [10], "n_layers": 1}, False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc((torch.zeros(2, 5, 12), torch.zeros(2, 5, 5)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        GraphConvolution(12, 4)(torch.zeros(2, 5, 12), torch.zeros(2, 5, 5))
    # the compact route has no CPU path either: bit words on the host are refused, as is a host CompactGraphBatch
    from openchem_b200 import compact, functional
    cb = compact.pack_graph_batch(np.zeros((2, 5, 12), np.float32), np.eye(5, dtype=np.float32)[None].repeat(2, 0))
    assert functional.is_compact_pair(cb.x_bits, cb.adj_bits) and not functional.is_compact_pair(torch.zeros(1), torch.zeros(1))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc((cb.x_bits, cb.adj_bits))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc(cb)
    with pytest.raises(RuntimeError, match="BOTH be int32 bit words"):       # half a compact batch is a caller error, not dense data
        enc((cb.x_bits, torch.zeros(2, 5, 5)))


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "openchem_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def test_namespace_shadow_resolves_to_this_package():
    import importlib
    import sys
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "openchem" or k.startswith("openchem.")}
    try:
        sys.path.insert(0, ROOT)
        m = importlib.import_module("openchem.layers.gcn")
        e = importlib.import_module("openchem.modules.encoders.gcn_encoder")
        assert m.GraphConvolution.__module__ == "openchem_b200.layers.gcn"
        assert e.GraphCNNEncoder.__module__ == "openchem_b200.modules.encoders.gcn_encoder"
    finally:
        sys.path.remove(ROOT)
        for k in [k for k in sys.modules if k == "openchem" or k.startswith("openchem.")]:
            sys.modules.pop(k)
        sys.modules.update(saved)


def test_header_compiles_as_plain_c_and_links(tmp_path):
    """include/ocgcn.h is a C header (no C++/torch types): a C translation unit including it must compile with gcc, link
    against libocgcn.so and get sane answers from the host-only entry points."""
    import shutil
    import subprocess
    from openchem_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    _lib.load()
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "ocgcn.h"
int main(void) {
  ocgcn_desc d;
  memset(&d, 0, sizeof d);
  d.B = 8; d.N = 20; d.L = 2; d.E = 12; d.F[0] = 33; d.F[1] = 16; d.F[2] = 8;
  d.mode = OCGCN_BF16; d.training = 1; d.has_bias = 1; d.eps = 1e-5f; d.momentum = 0.1f;
  size_t saved = 0, scratch = 0;
  int rc = ocgcn_query_workspace(&d, &saved, &scratch);
  printf("%d %d %zu %zu %s\n", ocgcn_abi_version(), rc, saved, scratch, ocgcn_status_string(OCGCN_EARCH));
  d.N = 1000;
  printf("%d\n", ocgcn_query_workspace(&d, &saved, &scratch));
  /* MLP head: struct sizes (the ctypes mirrors in openchem_b200/_lib.py must agree) and the host-only workspace query */
  ocgcn_head_desc h;
  memset(&h, 0, sizeof h);
  h.B = 64; h.n_layers = 2; h.width[0] = 128; h.width[1] = 128; h.width[2] = 1;
  h.act[0] = OCGCN_ACT_RELU; h.act[1] = OCGCN_ACT_IDENTITY; h.training = 1; h.loss = OCGCN_LOSS_MSE; h.eps = 1e-5f; h.momentum = 0.1f;
  rc = ocgcn_head_query_workspace(&h, &saved, &scratch);
  printf("%zu %zu %zu %zu %d %zu %zu\n", sizeof(ocgcn_desc), sizeof(ocgcn_head_desc), sizeof(ocgcn_head_params), sizeof(ocgcn_head_grads),
         rc, saved, scratch);
  h.width[1] = 300;
  printf("%d\n", ocgcn_head_query_workspace(&h, &saved, &scratch));
  return 0;
}
''')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-l:libocgcn.so", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    first = out[0].split(" ", 4)
    assert first[0] == str(_lib.ABI_VERSION) and first[1] == "0" and int(first[2]) > 0 and int(first[3]) > 0 and "sm_100" in first[4]
    assert out[1].strip() == "2"      # N > 256: OCGCN_EUNSUPPORTED
    import ctypes
    sizes = out[2].split()
    assert [int(v) for v in sizes[:4]] == [ctypes.sizeof(_lib.Desc), ctypes.sizeof(_lib.HeadDesc), ctypes.sizeof(_lib.HeadParams),
                                           ctypes.sizeof(_lib.HeadGrads)]
    assert sizes[4] == "0" and int(sizes[5]) > 0 and int(sizes[6]) > 0
    assert out[3].strip() == "2"      # head width > 256: OCGCN_EUNSUPPORTED


def test_edge_attention_encoder_surface_and_init_order():
    """Same constructor / attributes / state_dict keys as the reference's GraphEdgeAttentionEncoder, and the same initial weights
    under the same seed (fixture written by oracle/make_golden_edge.py from the unmodified reference: dense first, then the layers)."""
    import numpy as np
    import torch
    from _util import load_golden
    from openchem_b200.modules.encoders.edge_attention_encoder import GraphEdgeAttentionEncoder
    g = load_golden("edge_small")
    hidden, sizes = [int(h) for h in g["hidden"]], [int(s) for s in g["sizes"]]
    torch.manual_seed(5)
    enc = GraphEdgeAttentionEncoder({"input_size": int(g["x"].shape[2]), "encoder_dim": int(g["dO"].shape[1]), "n_layers": len(hidden),
                                     "hidden_size": hidden, "edge_attr_sizes": sizes}, False)
    sd = enc.state_dict()
    assert sorted(sd.keys()) == sorted(k[3:] for k in g if k.startswith("sd_"))
    for k, v in sd.items():
        if ".bn." not in k:          # (the fixture randomised the BatchNorm parameters after construction)
            assert np.array_equal(v.numpy(), g["sd_" + k]), k
    assert enc.attr_dim == sum(sizes) and len(enc.graph_conv) == len(hidden) * sum(sizes) and enc.hidden_size[0] == enc.input_size
    assert GraphEdgeAttentionEncoder.get_required_params() == {"n_layers": int, "hidden_size": list, "edge_attr_sizes": list}
    with pytest.raises(RuntimeError):        # no CPU fallback
        enc((torch.from_numpy(g["x"]), torch.from_numpy(g["edge_attr"])))
