#!/usr/bin/env python
"""End-to-end sanity band: train the logP GraphCNN of example_configs/logp_gcnn_config.py on the reference's own
logP_labels.csv through the UNMODIFIED run.py, once with the reference's eager encoder and once with the B200 encoder (route 2,
run_b200.py), same seed, and print the validation R^2 lines of both (the tutorial reports R^2 ~ 0.90 after 101 epochs with RDKit
features, docs gcnn_tutorial.rst:158-203; features here are the RDKit-free approximation of openchem_b200/smiles.py).
    python tools/train_logp.py [epochs]        (needs a GPU and baseline/_ref)"""
import os
import re
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _dropin import run_reference_driver  # noqa: E402

epochs = sys.argv[1] if len(sys.argv) > 1 else "30"
env = {"OCB200_DATA": "logp_csv", "OCB200_EPOCHS": epochs, "OCB200_BATCH": "256", "OCB200_PRINT_EVERY": "5", "OCB200_SAVE_EVERY": "1000",
       "CUDA_VISIBLE_DEVICES": "0"}
for route, prec in (("shadow", "fp32"), ("reference", "fp32")):
    t = time.time()
    r = run_reference_driver(route, "logp_gcnn_reference_imports.py", "/tmp/ocb200_logp_%s" % route, env_extra=dict(env, OCB200_PRECISION=prec), timeout=3000)
    log = r.stdout + r.stderr
    lines = [l.split("openchem.")[-1] for l in log.splitlines() if "EVALUATION" in l or "TRAINING" in l]
    print("== %s encoder: rc %d, %.0f s wall" % ("B200" if route == "shadow" else "reference (eager CUDA)", r.returncode, time.time() - t))
    for l in lines[-6:]:
        print("   ", l)
    if r.returncode:
        print(log[-2000:])
    m = re.findall(r"EVALUATION: .*?Metrics: ([0-9.eE+-]+)", log)
    print("   final validation R^2:", m[-1] if m else None, flush=True)
