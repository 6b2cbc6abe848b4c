"""CPU side of the drop-in proof: the unmodified reference run.py (staged copy of /root/reference) runs train_eval with the
repo's synthetic config and the RDKit import stub, and the checkpoint it writes (DataParallel `module.` prefix, run.py:237-243)
loads strictly into a Graph2Label assembled from THIS package's encoder class. The GPU side (tests/test_gpu_dropin.py) runs
the same drivers with the B200 classes doing the arithmetic."""
import os

import pytest

from _dropin import reference_root, run_reference_driver

torch = pytest.importorskip("torch")
SMALL = {"OCB200_TRAIN": "96", "OCB200_VAL": "32", "OCB200_EPOCHS": "2", "OCB200_BATCH": "32", "OCB200_NMAX": "20"}


@pytest.mark.skipif(reference_root() is None, reason="reference not staged (python oracle/stage_reference.py in the authoring container)")
def test_unmodified_run_py_trains_and_its_checkpoint_loads_into_the_b200_model(tmp_path):
    logdir = str(tmp_path / "logs")
    r = run_reference_driver("reference", "logp_gcnn_reference_imports.py", logdir, env_extra=dict(SMALL, CUDA_VISIBLE_DEVICES=""))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    ckpt = os.path.join(logdir, "checkpoint", "epoch_1")
    assert os.path.isfile(ckpt)
    sd = torch.load(ckpt, map_location="cpu")
    assert all(k.startswith("module.") for k in sd)
    assert "module.Encoder.graph_convolutions.4.bn.num_batches_tracked" in sd and "module.Encoder.dense.weight" in sd
    # same keys / shapes in a model built from this package's encoder (constructed on the CPU; only forward needs the GPU)
    from openchem_b200.modules.encoders.gcn_encoder import GraphCNNEncoder
    enc = GraphCNNEncoder({"input_size": 33, "encoder_dim": 128, "n_layers": 5, "hidden_size": [128] * 5}, False)
    enc.load_state_dict({k[len("module.Encoder."):]: v for k, v in sd.items() if k.startswith("module.Encoder.")}, strict=True)
    assert int(enc.graph_convolutions[0].bn.num_batches_tracked) == 2 * 3       # 2 epochs x 3 batches of 32


def test_trampoline_puts_the_shadow_first(tmp_path):
    """run_b200.py orders sys.path [repo, reference, ...]: the namespace shadow wins for the two replaced modules only."""
    import subprocess
    import sys
    ref = reference_root()
    if ref is None:
        pytest.skip("reference not staged")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, os; sys.argv=['x']; import run_b200 as t; ref=t.reference_root();"
            "sys.path[:] = [t.HERE, ref] + sys.path;"
            "import openchem.modules.encoders.gcn_encoder as g, openchem.layers.gcn as l, openchem.models.Graph2Label as m, openchem.modules.mlp.openchem_mlp as p;"
            "print(g.GraphCNNEncoder.__module__, l.GraphConvolution.__module__, os.path.dirname(m.__file__).startswith(ref), os.path.dirname(p.__file__).startswith(ref))")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([root, os.path.join(root, "tests", "stubs")]), OPENCHEM_ROOT=ref, PYTHONDONTWRITEBYTECODE="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, cwd=root, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.split() == ["openchem_b200.modules.encoders.gcn_encoder", "openchem_b200.layers.gcn", "True", "True"], out.stdout
