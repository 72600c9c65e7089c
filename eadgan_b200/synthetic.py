"""Seeded synthetic inputs of the benchmark (the datasets are not available offline).

The same generators exist in oracle/torch_oracle.py -- test infrastructure that the product and the timed arm of
bench.py must not import; tests/test_cpu.py asserts that the two produce identical tensors."""
import torch


def celeba_images(batch, seed=0):
    """images ~ U(-1, 1) [batch, 3, 64, 64] (CelebA crops normalised to [-1, 1], celebA/EAD-GAN_celebA.py:194-203)."""
    g = torch.Generator().manual_seed(1000 + seed)
    return torch.rand(batch, 3, 64, 64, generator=g) * 2 - 1


def dsprites_images(batch, seed=0):
    """binary {0,1} uint8 sprites [batch, 64, 64]: one filled axis-aligned ellipse / square per image (dSprites is
    64x64 binary shapes; dSprites/rp.py:241-246,369-370 feeds them as uint8 -> float)."""
    import numpy as np
    rs = np.random.RandomState(2000 + seed)
    yy, xx = np.mgrid[0:64, 0:64]
    out = np.zeros((batch, 64, 64), dtype=np.uint8)
    for b in range(batch):
        cx, cy = rs.uniform(20, 44, 2)
        r = rs.uniform(5, 12)
        if rs.rand() < 0.5:
            out[b] = (((xx - cx) / r) ** 2 + ((yy - cy) / (0.7 * r)) ** 2 <= 1).astype(np.uint8)
        else:
            out[b] = ((abs(xx - cx) <= r) & (abs(yy - cy) <= r)).astype(np.uint8)
    return torch.from_numpy(out)


def sample_celeba(rs, batch, latent_dim=200, code_dim=8, n_classes=10):
    """the reference's HOST draws in its order (celebA/EAD-GAN_celebA.py:308-317): z, code, labels"""
    z = rs.normal(0, 1, (batch, latent_dim))
    code = rs.uniform(-1, 1, (batch, code_dim))
    labels = rs.randint(0, n_classes, batch)
    return (torch.tensor(z, dtype=torch.float32), torch.tensor(code, dtype=torch.float32),
            torch.tensor(labels, dtype=torch.long))
