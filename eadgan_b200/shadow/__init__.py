"""Shadow modules for the reference's per-dataset ``utils_*.py`` helper files.

The training scripts do ``from utils_rpqxy import *`` (celebA/EAD-GAN_celebA.py:24, dSprites/rp.py:26-27, ...) and call
``get_matrix*`` / ``affine_regularzier*`` 3-6 times per iteration.  The reference implementations build
``torch.eye(3)`` on the HOST and assign CUDA slices into it: 8-9 synchronising device-to-host copies per call, three
3x3 matmuls on the CPU, a copy back, plus ``torch.inverse`` (another host synchronisation) -- all inside the autograd
graph.  ``python -m eadgan_b200.run <script>`` puts the directory returned by
``shadow_dir_for(script)`` in front of the script's own directory on ``sys.path``, so the UNMODIFIED script imports
these modules instead: same function names, argument meaning and results, evaluated on the device by the fused
kernels of csrc/glue.cu (eadgan_b200.affine) without any host round trip.

One sub-directory per reference directory (two of them hold a different ``utils_pxy.py``)."""
import os

_HERE = os.path.dirname(os.path.abspath(__file__))


def shadow_dir_for(script_path):
    """the shadow directory for a reference script, chosen by the files it sits next to, or None"""
    script_dir = os.path.dirname(os.path.abspath(script_path))
    try:
        present = {f for f in os.listdir(script_dir) if f.startswith("utils_") and f.endswith(".py")}
    except OSError:
        return None
    best = None
    for name in sorted(os.listdir(_HERE)):
        d = os.path.join(_HERE, name)
        if not os.path.isdir(d) or name.startswith("_"):
            continue
        mine = {f for f in os.listdir(d) if f.startswith("utils_") and f.endswith(".py")}
        if present and present <= mine and (best is None or os.path.basename(script_dir) == name):
            best = d
    return best
