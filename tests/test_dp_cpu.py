"""CPU (gloo, world_size 2): host-side logic of eadgan_b200.parallel -- contiguous sharding, per-optimizer
bucket plans, hook-driven overlapped all-reduce, reduce-at-step fallback, ownership of gradients by phase.
The kernels themselves need a GPU; here the gradients are produced by stock autograd on tiny CPU tensors."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class _Opt:
    """the two attributes DataParallel reads from an optimizer"""

    def __init__(self, params):
        self.param_groups = [{"params": list(params)}]


def _worker(rank, world, port, overlap, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from eadgan_b200 import parallel
    dp = parallel.init(rank, world, torch.device("cpu"), backend="gloo", bucket_bytes=64, overlap=overlap)
    torch.manual_seed(0)
    a = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(7))]
    b = [torch.nn.Parameter(torch.randn(4, 4)), torch.nn.Parameter(torch.randn(2))]
    opt_a, opt_b, opt_ab = _Opt(a), _Opt(b), _Opt(a + b)
    x_full = torch.arange(8 * 3, dtype=torch.float32).view(8, 3) / 10
    x = parallel.shard(x_full, rank, world)
    assert x.shape[0] == 4 and torch.equal(x, x_full[rank * 4:(rank + 1) * 4])

    def loss_fn(xs):
        return ((xs @ a[0].t()).sum() * a[1].sum() + (b[0] ** 2).sum() * xs.mean() + b[1].sum() * xs.sum())

    results = {}
    for name, opt in (("a", opt_a), ("b", opt_b), ("ab", opt_ab)):
        for p in a + b:
            p.grad = None
        dp.arm(opt)
        loss_fn(x).backward()
        red = dp.reduce(opt)
        owned = opt.param_groups[0]["params"]
        assert set(red.keys()) == set(owned)          # only the phase's own parameters are communicated
        results[name] = [red[p].clone() for p in owned]
    # reference: sum over both shards of the local gradients
    refs = {}
    for name, opt in (("a", opt_a), ("b", opt_b), ("ab", opt_ab)):
        tot = None
        for r in range(world):
            for p in a + b:
                p.grad = None
            loss_fn(parallel.shard(x_full, r, world)).backward()
            gs = [p.grad.clone() for p in opt.param_groups[0]["params"]]
            tot = gs if tot is None else [t + g for t, g in zip(tot, gs)]
        refs[name] = tot
    ok = all(torch.allclose(u, v, rtol=1e-5, atol=1e-6) for k in refs for u, v in zip(results[k], refs[k]))
    # two backward passes into the same optimiser's gradients within one phase (the all-reduce launched
    # after the first accumulation must be waited for, then the buckets are re-reduced from the final .grad values)
    for p in a + b:
        p.grad = None
    dp.arm(opt_ab)
    loss_fn(x).backward()
    loss_fn(x).backward()
    red = dp.reduce(opt_ab)
    ok = ok and all(torch.allclose(red[p], 2 * r, rtol=1e-5, atol=1e-6) for p, r in zip(a + b, refs["ab"]))
    # an Adam built while a data-parallel state exists all-reduces in step() without an explicit attach()
    # (``torchrun -m eadgan_b200.run script.py`` builds its optimisers inside the unmodified script)
    from eadgan_b200.optim import Adam
    ok = ok and Adam(a, lr=1e-3)._dp is dp
    # SyncBN helper
    t = torch.tensor([1.0 + rank, 2.0], dtype=torch.float64)
    dp.allreduce_sum_(t)
    ok = ok and torch.equal(t, torch.tensor([3.0, 4.0], dtype=torch.float64))
    q.put((rank, bool(ok), len(dp._plans[id(opt_ab)]["buckets"])))
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_bucketed_allreduce_two_ranks_gloo(overlap):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (1 if overlap else 0) + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, overlap, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in out), out
    assert all(nb >= 2 for _, _, nb in out)   # 64-byte buckets: several buckets per optimizer


def test_shard_rejects_indivisible_batch():
    from eadgan_b200 import parallel
    with pytest.raises(ValueError):
        parallel.shard(torch.zeros(7, 2), 0, 2)
