"""-m gpu: the device sampler (csrc/sample.cu, Philox4x32-10) against its numpy restatement oracle/philox_ref.py,
which tests/test_cpu.py pins to the generator's published known-answer vectors.  Uniform draws and integers are
bit-exact; normals (Box-Muller through logf / sincosf) to 2e-6.  Sharded draws equal the single-device draws."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,cols,row0", [(7, 8, 0), (64, 200, 0), (5, 3, 11), (33, 7, 2), (1024, 200, 3072)])
def test_streams_match_the_host_restatement(cuda, rows, cols, row0):
    from eadgan_b200.sampling import DeviceSampler
    from oracle import philox_ref as R
    seed = 0x1234_5678_9abc_def1
    s = DeviceSampler(seed, cuda, row0=row0)
    for step in range(3):
        u = s.uniform(1, rows, cols, -1.0, 1.0).cpu().numpy()
        assert np.array_equal(u, R.uniform(seed, step, 1, row0, rows, cols, -1.0, 1.0))
        assert u.min() >= -1.0 and u.max() < 1.0
        k = s._draw(2, 2, rows, cols, n=10).cpu().numpy()
        assert np.array_equal(k, R.randint(seed, step, 2, row0, rows, cols, 10))
        z = s.normal(0, rows, cols).cpu().numpy()
        assert np.abs(z - R.normal(seed, step, 0, row0, rows, cols)).max() <= 2e-6
        s.advance()


def test_sharded_draws_equal_the_global_draws(cuda):
    from eadgan_b200.sampling import DeviceSampler
    whole = DeviceSampler(7, cuda).celeba(64)
    parts = [DeviceSampler(7, cuda, row0=r * 16).celeba(16) for r in range(4)]
    for i in range(3):
        assert torch.equal(whole[i], torch.cat([p[i] for p in parts]))
    z, code, labels = whole
    assert abs(float(z.mean())) < 0.05 and abs(float(z.std()) - 1) < 0.05
    assert int(labels.min()) == 0 and int(labels.max()) == 9 and float(code.abs().max()) <= 1


def test_sampled_step_in_a_graph_draws_fresh_latents_every_replay(cuda):
    """the whole end-to-end iteration (sampling kernels + iteration counter + step) replays as one CUDA graph and
    equals the eager sampled step, replay after replay."""
    import os
    from eadgan_b200 import synthetic
    from eadgan_b200.graph import GraphedStep
    from eadgan_b200.sampling import DeviceSampler, SampledStep
    from eadgan_b200.steps.celeba import CelebAStep
    os.environ["EADGAN_PRECISION"] = "bf16"
    imgs = [synthetic.celeba_images(16, i).to(cuda) for i in range(3)]
    eager = SampledStep(CelebAStep(seed=0, device=cuda), DeviceSampler(5, cuda), "celeba")
    want = [{k: float(v) for k, v in eager(x).items()} for x in (imgs[0], imgs[0], imgs[1], imgs[2])]
    g = GraphedStep(SampledStep(CelebAStep(seed=0, device=cuda), DeviceSampler(5, cuda), "celeba"), [imgs[0]], warmup=1)
    got = [{k: float(v) for k, v in g(x).items()} for x in (imgs[0], imgs[1], imgs[2])]   # 1 eager warm-up + capture pass
    for a, b in zip(got, want[1:]):
        for k in a:
            assert abs(a[k] - b[k]) <= 3e-2 * max(1.0, abs(b[k])), (k, a, b)
    assert got[0] != got[1]
