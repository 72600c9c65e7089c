"""-m gpu: write-bounds checks of the tcgen05 kernels with guard bands (compute-sanitizer is closed on the GPU pool this
was developed on, so the overrun check is built into the tests).  Every output lives in the middle of a larger
allocation filled with a sentinel bit pattern; after the launch the guard bands in front of and behind it -- and, for
halo-padded NHWC outputs, the halo itself -- must be untouched, and the interior must match a launch into a normally
allocated buffer bit for bit.  Ragged batches (n not a multiple of the images per tile) are the cases that matter: the
kernels compute whole tiles and must clip the stores."""
import pytest
import torch

pytestmark = pytest.mark.gpu
GUARD = 4096            # elements on both sides (keeps the 16-byte alignment the stores need)
SENT16 = 0x7B7B         # bf16 bit pattern of the sentinel
SENT32 = 0x7B7B7B7B     # fp32 bit pattern of the sentinel


def _guarded(shape, dtype, dev):
    n = 1
    for s in shape:
        n *= s
    big = torch.empty(n + 2 * GUARD, device=dev, dtype=dtype)
    bits = big.view(torch.int16 if dtype == torch.bfloat16 else torch.int32)
    bits.fill_(SENT16 if dtype == torch.bfloat16 else SENT32)
    return big, big[GUARD:GUARD + n].view(shape)


def _check(big, out, ref, halo):
    bits = big.view(torch.int16 if big.dtype == torch.bfloat16 else torch.int32)
    sent = SENT16 if big.dtype == torch.bfloat16 else SENT32
    assert bool((bits[:GUARD] == sent).all()), "wrote in front of the output"
    assert bool((bits[-GUARD:] == sent).all()), "wrote behind the output"
    ob = out.view(torch.int16 if big.dtype == torch.bfloat16 else torch.int32)
    if halo:   # padded NHWC [n, h+2, w+2, c]: the kernels own the interior only
        assert bool((ob[:, 0] == sent).all()) and bool((ob[:, -1] == sent).all()), "wrote into the halo rows"
        assert bool((ob[:, :, 0] == sent).all()) and bool((ob[:, :, -1] == sent).all()), "wrote into the halo columns"
        assert torch.equal(out[:, 1:-1, 1:-1], ref[:, 1:-1, 1:-1]), "interior differs from the unguarded launch"
    else:
        assert torch.equal(out, ref), "output differs from the unguarded launch"


def _bf(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize("n,c,h,k", [(5, 128, 32, 256), (3, 256, 16, 512), (9, 512, 8, 1024), (37, 128, 8, 64), (5, 64, 8, 64), (6, 32, 32, 32)])
def test_fprop_dgrad_write_bounds(cuda, monkeypatch, n, c, h, k):
    from eadgan_b200 import tc
    from eadgan_b200._lib import ACT_LRELU
    torch.manual_seed(1)
    x = _bf(torch.randn(n, c, h, h, device=cuda))
    y = _bf(torch.randn(n, k, h // 2, h // 2, device=cuda))
    w = _bf(torch.randn(k, c, 4, 4, device=cuda) * 0.05)
    b = torch.randn(k, device=cuda)
    xp, yp = tc.to_padded(x), tc.to_padded(y)
    wf, wd = tc.pack_w(w, None, "fprop"), tc.pack_w(w, None, "dgrad")
    ref = tc.fprop(xp, wf, b, k, ACT_LRELU, 0.1).clone()
    big, out = _guarded((n, h // 2 + 2, h // 2 + 2, k), torch.bfloat16, cuda)
    tc.fprop(xp, wf, b, k, ACT_LRELU, 0.1, out=out)
    _check(big, out, ref, halo=True)
    bigf, outf = _guarded((n, k, h // 2, h // 2), torch.float32, cuda)       # fp32 NCHW epilogue
    tc.fprop(xp, wf, b, k, out_f32_nchw=True, out=outf)
    _check(bigf, outf, tc.fprop(xp, wf, b, k, out_f32_nchw=True).clone(), halo=False)
    for mode in ("0", "2"):        # pixel-major and (where the geometry allows it) channel-major dgrad
        monkeypatch.setenv("EADGAN_TC_DGRADT", mode)
        stats = torch.zeros(2 * c, device=cuda, dtype=torch.float64)
        refd = tc.dgrad(yp, wd, None, c, mask=xp, mask_mode=ACT_LRELU, slope=0.1).clone()
        big, out = _guarded((n, h + 2, h + 2, c), torch.bfloat16, cuda)
        tc.dgrad(yp, wd, None, c, mask=xp, mask_mode=ACT_LRELU, slope=0.1, out=out)
        _check(big, out, refd, halo=True)
        refs = tc.dgrad(yp, wd, None, c, stats=stats).clone()
        big, out = _guarded((n, h + 2, h + 2, c), torch.bfloat16, cuda)
        tc.dgrad(yp, wd, None, c, stats=torch.zeros_like(stats), out=out)
        _check(big, out, refs, halo=True)


@pytest.mark.parametrize("n", [1, 5, 37])
def test_thin_layer_write_bounds(cuda, monkeypatch, n):
    from eadgan_b200 import tc
    from eadgan_b200._lib import ACT_LRELU, ACT_TANH
    torch.manual_seed(2)
    img = _bf(torch.rand(n, 3, 64, 64, device=cuda) * 2 - 1)
    w = torch.randn(128, 3, 4, 4, device=cuda) * 0.1
    b = torch.randn(128, device=cuda)
    r = tc.thin_expand(img)
    wf, wd = tc.thin_pack_w(w, "fprop"), tc.thin_pack_w(w, "dgrad")
    for mode in ("0", "2"):        # pixel-major and channel-major forward
        monkeypatch.setenv("EADGAN_TC_DGRADT", mode)
        ref = tc.thin_fprop(r, wf, b, 3, 128, ACT_LRELU, 0.1).clone()
        big, out = _guarded((n, 34, 34, 128), torch.bfloat16, cuda)
        tc.thin_fprop(r, wf, b, 3, 128, ACT_LRELU, 0.1, out=out)
        _check(big, out, ref, halo=True)
    y = tc.to_padded(_bf(torch.randn(n, 128, 32, 32, device=cuda)))
    b3 = torch.randn(3, device=cuda)
    ref = tc.thin_dgrad(y, wd, b3, 3, ACT_TANH).clone()
    big, out = _guarded((n, 3, 64, 64), torch.float32, cuda)
    tc.thin_dgrad(y, wd, b3, 3, ACT_TANH, out=out)
    _check(big, out, ref, halo=False)
