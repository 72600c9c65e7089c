"""STOCK-path helper module of tests/scripts/dSprites/rp_mini.py, named like the reference's dSprites/utils_pxy.py
(see utils_rp.py next to this file); shadowed by eadgan_b200/shadow/dSprites/utils_pxy.py under the shim."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from oracle import torch_oracle as _O  # noqa: E402

get_matrix_pxy = _O.pxy_get_matrix
get_matrix_pxy_align = _O.dsprites_align_matrix
affine_regularzier_pxy = _O.pxy_affine_regularizer
