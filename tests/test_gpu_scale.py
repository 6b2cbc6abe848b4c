"""Parity at the BENCHMARKED sizes (BASELINE configs C1-C5 at their stated batch sizes), against the fp64 oracle.

The small cases in test_gpu_parity.py never leave the first wave of CTAs. These run the code path bench.py times: more
tiles than persistent CTAs (several tiles per CTA, mbarrier phases carried across tiles, weight gradients accumulated in TMEM
across tiles), thousands of per-tile partials in the fixed-order reductions, BatchNorm statistics over up to 262 144 rows, and
the split weight-gradient GEMM with several tiles per CTA (C5 widths).

Oracle: oracle/gcn_oracle.py in fp64; its max-pool uses the plain-C scan (oracle/pool_scan.c, built by
__graft_entry__.build(), bit-identical to the numpy definition - tests/test_oracle.py) so a 4096-molecule case takes
well under a minute of host time. Same acceptance rules as test_gpu_parity.py:
  fp32 class   1e-5 rel-L2 (north star); routing differences only at near-ties below fp32 resolution
  bf16 / tf32  1e-2 rel-L2 on outputs and on every gradient (north star), see _tc_check
"""
import numpy as np
import pytest

from _util import random_params
from test_gpu_parity import _check_against_oracle, _tc_check

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

C1 = dict(N=85, F0=33, hidden=[128] * 5, E=128)
C2 = dict(N=50, F0=33, hidden=[128] * 5, E=128)
C3 = dict(N=64, F0=64, hidden=[128] * 4, E=128)
C5 = dict(N=128, F0=64, hidden=[256] * 6, E=256)

SCALE_CASES = [
    # (id, shape, batch, occupancy, precision)
    ("c3_b4096_bf16", C3, 4096, "full", "bf16"),        # the bench.py default: 2048 tiles over 296 persistent CTAs
    ("c3_b4096_tf32", C3, 4096, "molecular", "tf32"),
    ("c3_b1024_fp32", C3, 1024, "full", "fp32"),        # 512 tiles, exact-fp32 engine
    ("c4_b2048_bf16", C3, 2048, "molecular", "bf16"),   # per-GPU batch of BASELINE config 4
    ("c5_b1024_bf16", C5, 1024, "molecular", "bf16"),   # 1024 one-molecule tiles, 6x256: split dW GEMM with tiles_per_cta > 1
    ("c1_b256_fp32", C1, 256, "molecular", "fp32"),
    ("c1_b256_bf16", C1, 256, "molecular", "bf16"),
    ("c2_b256_fp32", C2, 256, "molecular", "fp32"),
    ("c2_b256_bf16", C2, 256, "molecular", "bf16"),
    ("c2_b257_bf16", C2, 257, "molecular", "bf16"),     # ragged last tile of a persistent CTA
]


@pytest.mark.parametrize("case", SCALE_CASES, ids=[c[0] for c in SCALE_CASES])
def test_encoder_at_benchmark_scale(case):
    from oracle import gcn_oracle as orc
    from openchem_b200.synthetic import make_graph_batch
    name, shape, B, occ, precision = case
    if B * shape["N"] * max(shape["hidden"]) >= orc.POOL_C_MIN_ELEMS and orc.pool_lib() is None:
        pytest.skip("oracle/_build/libpool_scan.so not built (python -c 'import __graft_entry__ as g; g.build()')")
    gb = make_graph_batch(B, shape["N"], shape["F0"], occupancy=occ, seed=31)
    p = random_params([shape["F0"]] + shape["hidden"], shape["E"], seed=8)
    dO = np.random.RandomState(9).randn(B, shape["E"]).astype(np.float32)
    cfg = {"input_size": shape["F0"], "encoder_dim": shape["E"], "n_layers": len(shape["hidden"]), "hidden_size": list(shape["hidden"])}
    if precision == "fp32":
        _check_against_oracle(gb["x"], gb["adj"], p, dO, cfg, name)
    else:
        worst = _tc_check(gb["x"], gb["adj"], p, dO, cfg, precision, name)
        print(f"{name}: worst rel-L2 {worst:.2e}")


def test_two_steps_identical_at_scale():
    """Determinism on the multi-tile persistent path: the same step twice gives bit-identical gradients (fixed-order merges,
    no floating-point atomics)."""
    from _util import build_encoder
    from openchem_b200.synthetic import make_graph_batch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    dev = torch.device("cuda:0")
    B = 1536
    gb = make_graph_batch(B, 64, 64, occupancy="molecular", seed=5)
    p = random_params([64, 128, 128, 128, 128], 128, seed=3)
    cfg = {"input_size": 64, "encoder_dim": 128, "n_layers": 4, "hidden_size": [128] * 4}
    tx, ta = torch.from_numpy(gb["x"]).to(dev), torch.from_numpy(gb["adj"]).to(dev)
    dO = torch.from_numpy(np.random.RandomState(1).randn(B, 128).astype(np.float32)).to(dev)
    runs = []
    for _ in range(2):
        enc = build_encoder(cfg, p, dev, precision="bf16")
        enc.train()
        out = enc((tx, ta))
        out.backward(dO)
        torch.cuda.synchronize()
        runs.append([out.detach().clone()] + [q.grad.clone() for q in enc.parameters()])
    for a, b in zip(*runs):
        assert torch.equal(a, b)
