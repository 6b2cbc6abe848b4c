"""Trampoline for drop-in route 2: run the UNMODIFIED reference run.py with this repo's `openchem/` namespace
shadow ahead of the reference on sys.path, so `openchem.layers.gcn` / `openchem.modules.encoders.gcn_encoder` resolve to the
B200 classes and every other `openchem.*` module to the untouched reference.

    python run_b200.py --config_file=<any reference GraphCNN config> --mode=train_eval
    python <OpenChem>/launch.py --nproc_per_node=8 run_b200.py --config_file=... --mode=train_eval

`python <OpenChem>/run.py` itself would put the reference's directory FIRST on sys.path (the script directory), which is why a
trampoline is needed. The reference checkout is found through $OPENCHEM_ROOT, else baseline/_ref (staged copy), else
/root/reference. Nothing of the reference is edited: its run.py is executed with runpy as __main__.
"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_root():
    for cand in (os.environ.get("OPENCHEM_ROOT"), os.path.join(HERE, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "run.py")):
            return cand
    raise SystemExit("run_b200.py: no OpenChem checkout found (set OPENCHEM_ROOT)")


def main():
    ref = reference_root()
    # this repo first (shadow modules + openchem_b200), then the reference (everything else of the namespace package)
    sys.path[:] = [HERE, ref] + [p for p in sys.path if os.path.abspath(p or ".") not in (HERE, ref)]
    sys.argv[0] = os.path.join(ref, "run.py")
    runpy.run_path(sys.argv[0], run_name="__main__")


if __name__ == "__main__":
    main()
