"""STOCK-path helper module of tests/scripts/dSprites/rp_mini.py, named like the reference's dSprites/utils_rp.py and
exporting its function names -- here bound to the oracle's restatements (oracle/torch_oracle.py, pinned to the
reference).  Under ``python -m eadgan_b200.run`` this file is SHADOWED by eadgan_b200/shadow/dSprites/utils_rp.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from oracle import torch_oracle as _O  # noqa: E402

get_matrix = get_matrix_D = _O.dsprites_get_matrix
affine_regularzier = _O.dsprites_affine_regularizer
