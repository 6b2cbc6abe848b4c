"""N > 1 correctness of the NCCL path (skipped on a single-GPU box): tools/ddp_check.py under torchrun - replicas bit-identical
after captured steps with the in-graph all-reduce, and equal to a single-GPU emulation of gradient averaging."""
import json
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("precision,allreduce", [("bf16", "nccl"), ("fp32", "nccl"), ("bf16", "one_shot"), ("bf16", "multimem")])
def test_two_rank_train_step_matches_single_gpu_emulation(precision, allreduce):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29741", os.path.join(ROOT, "tools", "ddp_check.py")]
    r = subprocess.run(cmd, env=dict(os.environ, OCB200_PRECISION=precision, OCB200_ALLREDUCE=allreduce), capture_output=True, text=True, timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-3000:]
    res = json.loads(lines[-1])
    assert res["ok"] and res["replicas_bit_identical"] and res["graph"], res
