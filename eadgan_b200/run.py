"""``python -m eadgan_b200.run <reference_script.py> [script args]``

Patches torch (eadgan_b200.patch) and runs the unmodified reference training script
with ``runpy`` from the current working directory, which must hold whatever dataset /
artefact files the script opens.  The script's ``utils_*.py`` helper
module is shadowed by its device-side restatement (eadgan_b200/shadow; EADGAN_NO_SHADOW=1 turns
that off).  Under ``torchrun`` every rank runs the script on its own data shard: optimisers
all-reduce their gradients and train-mode BatchNorm layers synchronise their statistics.
"""
import os
import runpy
import sys


def main():
    if len(sys.argv) < 2:
        print(__doc__)
        raise SystemExit(2)
    script = os.path.abspath(sys.argv[1])
    from . import parallel
    from .patch import patch as apply_patch   # NB: the package re-exports the FUNCTION as eadgan_b200.patch
    parallel.init_from_env()
    apply_patch()
    sys.argv = [script] + sys.argv[2:]
    sys.path.insert(0, os.path.dirname(script))
    from .shadow import shadow_dir_for
    shadow = None if os.environ.get("EADGAN_NO_SHADOW") == "1" else shadow_dir_for(script)
    if shadow is not None:
        # ``from utils_x import *`` in the script now finds the device-side restatements of the affine glue first
        sys.path.insert(0, shadow)
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
