"""STOCK-path helper module of tests/scripts/MNIST/mnist_mini.py, named like the reference's MNIST/utils_rpqmnxy.py:
like it, it loads the frozen approximator from ``rpqmnxy_approximator.pt`` in the working directory at import time.
Function names bound to the oracle's restatements; shadowed by eadgan_b200/shadow/MNIST/utils_rpqmnxy.py under the shim."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from oracle import torch_oracle as _O  # noqa: E402

BFGS_approximator = _O.MnistAffineApproximator()
BFGS_approximator.cuda()
BFGS_approximator.load_state_dict(torch.load("rpqmnxy_approximator.pt"))
BFGS_approximator.eval()
get_matrix = _O.mnist_get_matrix


def affine_regularizer(real_code, trans_code):
    return _O.mnist_affine_regularizer(real_code, trans_code, BFGS_approximator)
