"""Seeded synthetic inputs of the benchmark (the datasets are not available offline).

The same generators exist in oracle/torch_oracle.py -- test infrastructure that the product and the timed arm of
bench.py must not import; tests/test_cpu.py asserts that the two produce identical tensors."""
import torch


def celeba_images(batch, seed=0):
    """images ~ U(-1, 1) [batch, 3, 64, 64] (CelebA crops normalised to [-1, 1], celebA/EAD-GAN_celebA.py:194-203)."""
    g = torch.Generator().manual_seed(1000 + seed)
    return torch.rand(batch, 3, 64, 64, generator=g) * 2 - 1
