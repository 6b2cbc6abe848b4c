"""Stage the UNMODIFIED reference under baseline/_ref (git-ignored, NOT gpurun-ignored) so it travels to the GPU box.

The reference is pure Python; /root/reference exists only in the authoring container. What is staged is a plain file copy of
the parts the GraphCNN path and its driver touch - nothing is edited, nothing is installed, and nothing under baseline/_ref is
ever committed:

    openchem/            the package (namespace package, no __init__.py)
    run.py, launch.py    the drivers the drop-in tests execute unchanged (tests/test_gpu_dropin.py)
    example_configs/logp_gcnn_config.py, tox21_rnn_config.py      cited by SURVEY section 8 (read-only context)
    benchmark_datasets/logp_dataset/logP_labels.csv, tox21/tox21.csv   inputs of the RDKit-free logP / Tox21-shaped batches

Users: `bench.py --impl reference` (times the reference's own Graph2Label on the host cores: kind "reference") and the drop-in
tests. Called by __graft_entry__.build(); a no-op when /root/reference is absent (GPU box: the staged copy is already there).
"""
from __future__ import annotations

import filecmp
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
ITEMS = ["openchem", "run.py", "launch.py", "example_configs/logp_gcnn_config.py", "example_configs/tox21_rnn_config.py",
         "benchmark_datasets/logp_dataset/logP_labels.csv", "benchmark_datasets/tox21/tox21.csv", "LICENSE"]


def staged_path():
    """baseline/_ref when staged, else /root/reference when present, else None."""
    if os.path.isfile(os.path.join(DST, "run.py")):
        return DST
    if os.path.isfile(os.path.join(SRC, "run.py")):
        return SRC
    return None


def stage(force=False):
    if not os.path.isdir(SRC):
        return staged_path()
    for item in ITEMS:
        s, d = os.path.join(SRC, item), os.path.join(DST, item)
        if not os.path.exists(s):
            continue
        if os.path.isdir(s):
            if force and os.path.isdir(d):
                shutil.rmtree(d)
            shutil.copytree(s, d, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        else:
            os.makedirs(os.path.dirname(d), exist_ok=True)
            if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
                shutil.copyfile(s, d)
    return DST


if __name__ == "__main__":
    print(stage(force=True))
