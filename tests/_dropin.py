"""Helpers for the drop-in tests: run the UNMODIFIED reference drivers (run.py / launch.py) in a subprocess."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUBS = os.path.join(ROOT, "tests", "stubs")


def reference_root():
    sys.path.insert(0, ROOT)
    from oracle import stage_reference
    return stage_reference.staged_path()


def run_reference_driver(route, config, logdir, mode="train_eval", env_extra=None, launch_nproc=0, timeout=900, extra_args=()):
    """route: 'reference' (reference alone), 'config' (route 1: this repo importable, reference modules win the namespace),
    'shadow' (route 2: through run_b200.py, this repo's openchem/ shadow wins). Returns the CompletedProcess."""
    ref = reference_root()
    assert ref is not None, "no reference checkout staged"
    env = dict(os.environ)
    env.update({"PYTHONDONTWRITEBYTECODE": "1", "OCB200_LOGDIR": logdir, "OPENCHEM_ROOT": ref,
                "MPLBACKEND": "Agg", "TQDM_DISABLE": "1"})
    env.update(env_extra or {})
    if route == "shadow":
        script = os.path.join(ROOT, "run_b200.py")
        env["PYTHONPATH"] = os.pathsep.join([ROOT, ref, STUBS])
    else:
        script = os.path.join(ref, "run.py")
        env["PYTHONPATH"] = os.pathsep.join([ref, ROOT, STUBS])
    args = ["--config_file=" + os.path.join(ROOT, "configs", config), "--mode=" + mode] + list(extra_args)
    if launch_nproc:
        cmd = [sys.executable, os.path.join(ref, "launch.py"), "--nproc_per_node=%d" % launch_nproc, "--master_addr=127.0.0.1",
               "--master_port=29731", script] + args
    else:
        cmd = [sys.executable, script] + args
    os.makedirs(logdir, exist_ok=True)
    # run.py's SummaryWriter() writes ./runs in the CWD (openchem_model.py:130): keep that inside the scratch directory
    return subprocess.run(cmd, env=env, cwd=logdir, capture_output=True, text=True, timeout=timeout)
