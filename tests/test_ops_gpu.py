"""-m gpu: every generic (fp32 SIMT / streaming) kernel against the stock torch fp32 operator
on identical seeded inputs, through the C ABI (eadgan_b200.functional -> libeadgan.so).
Bound: max|a-b|/max|b| <= 1e-5 (north_star fp32 tolerance)."""
import pytest
import torch
import torch.nn.functional as TF

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5

# (cin, cout, k, stride, pad, H) -- SURVEY.md appendix B geometries at small batch
CONV_GEOS = [
    (3, 128, 4, 2, 1, 64), (128, 256, 4, 2, 1, 32), (512, 1024, 4, 2, 1, 8), (1024, 19, 4, 1, 0, 4),
    (1, 32, 4, 2, 1, 64), (32, 64, 4, 2, 1, 16), (128, 128, 3, 1, 1, 16), (64, 1, 3, 1, 1, 32),
    (1, 16, 3, 2, 1, 32), (64, 128, 3, 2, 1, 4),
]
CONVT_GEOS = [
    (218, 1024, 4, 1, 0, 1), (1024, 512, 4, 2, 1, 4), (256, 128, 4, 2, 1, 16), (128, 3, 4, 2, 1, 32),
    (64, 64, 4, 2, 1, 8), (64, 1, 4, 2, 1, 32),
]


def _ok(ours, ref32, ref64, tol=TOL):
    """fp64 referee.  ours must be within `tol` of the fp64 result, or at
    least as close to it as stock torch fp32 is (x2): accumulation-order noise is not a defect."""
    e_ours, e_ref = rel_err(ours, ref64), rel_err(ref32, ref64)
    return e_ours <= max(tol, 2.0 * e_ref), (e_ours, e_ref)


def _grads(out, ins, seed=3):
    g = torch.Generator(device="cpu").manual_seed(seed)
    go = torch.randn(out.shape, generator=g).to(out.device)
    return torch.autograd.grad(out, ins, go), go


@pytest.mark.parametrize("geo", CONV_GEOS)
@pytest.mark.parametrize("act", [None, ("lrelu", 0.1)])
def test_conv2d(cuda, geo, act):
    import eadgan_b200.functional as Fn
    from eadgan_b200._lib import ACT_LRELU, ACT_NONE
    cin, cout, k, s, p, H = geo
    torch.manual_seed(0)
    B = 5
    x = torch.randn(B, cin, H, H, device=cuda, requires_grad=True)
    w = (torch.randn(cout, cin, k, k, device=cuda) * 0.05).requires_grad_()
    b = torch.randn(cout, device=cuda, requires_grad=True)
    x64, w64, b64 = (t.detach().double().requires_grad_() for t in (x, w, b))
    ref, ref64 = TF.conv2d(x, w, b, stride=s, padding=p), TF.conv2d(x64, w64, b64, stride=s, padding=p)
    if act:
        ref, ref64 = TF.leaky_relu(ref, act[1]), TF.leaky_relu(ref64, act[1])
    out = Fn.conv2d(x, w, b, s, p, ACT_LRELU if act else ACT_NONE, act[1] if act else 0.0)
    assert _ok(out, ref, ref64)[0], _ok(out, ref, ref64)
    (gx, gw, gb), go = _grads(out, (x, w, b))
    rs32 = torch.autograd.grad(ref, (x, w, b), go)
    rs64 = torch.autograd.grad(ref64, (x64, w64, b64), go.double())
    for o, r32, r64 in zip((gx, gw, gb), rs32, rs64):
        assert _ok(o, r32, r64)[0], _ok(o, r32, r64)


@pytest.mark.parametrize("geo", CONVT_GEOS)
@pytest.mark.parametrize("act", [None, "tanh"])
def test_conv_transpose2d(cuda, geo, act):
    import eadgan_b200.functional as Fn
    from eadgan_b200._lib import ACT_NONE, ACT_TANH
    cin, cout, k, s, p, H = geo
    torch.manual_seed(1)
    B = 3
    x = torch.randn(B, cin, H, H, device=cuda, requires_grad=True)
    w = (torch.randn(cin, cout, k, k, device=cuda) * 0.05).requires_grad_()
    b = torch.randn(cout, device=cuda, requires_grad=True)
    x64, w64, b64 = (t.detach().double().requires_grad_() for t in (x, w, b))
    ref = TF.conv_transpose2d(x, w, b, stride=s, padding=p)
    ref64 = TF.conv_transpose2d(x64, w64, b64, stride=s, padding=p)
    if act:
        ref, ref64 = torch.tanh(ref), torch.tanh(ref64)
    out = Fn.conv_transpose2d(x, w, b, s, p, ACT_TANH if act else ACT_NONE)
    assert out.shape == ref.shape
    assert _ok(out, ref, ref64)[0], _ok(out, ref, ref64)
    (gx, gw, gb), go = _grads(out, (x, w, b))
    rs32 = torch.autograd.grad(ref, (x, w, b), go)
    rs64 = torch.autograd.grad(ref64, (x64, w64, b64), go.double())
    for o, r32, r64 in zip((gx, gw, gb), rs32, rs64):
        assert _ok(o, r32, r64)[0], _ok(o, r32, r64)


@pytest.mark.parametrize("shape", [(7, 79, 8192), (64, 1024, 128), (5, 128, 1), (9, 512, 10)])
def test_linear(cuda, shape):
    import eadgan_b200.functional as Fn
    B, i, o = shape
    torch.manual_seed(2)
    x = torch.randn(B, i, device=cuda, requires_grad=True)
    w = (torch.randn(o, i, device=cuda) * 0.05).requires_grad_()
    b = torch.randn(o, device=cuda, requires_grad=True)
    ref = TF.linear(x, w, b)
    out = Fn.linear(x, w, b)
    assert rel_err(out, ref) <= TOL
    (gx, gw, gb), go = _grads(out, (x, w, b))
    rx, rw, rb = torch.autograd.grad(ref, (x, w, b), go)
    assert rel_err(gx, rx) <= TOL and rel_err(gw, rw) <= TOL and rel_err(gb, rb) <= TOL


@pytest.mark.parametrize("cfg", [(6, 512, 8, 1e-5, "relu"), (4, 128, 32, 1e-5, None), (8, 64, 16, 0.8, "lrelu"),
                                 (3, 32, 8, 0.8, None)])
@pytest.mark.parametrize("channels_last", [False, True])
def test_batchnorm_train(cuda, cfg, channels_last):
    import eadgan_b200.nn as enn
    B, C, H, eps, act = cfg
    torch.manual_seed(4)
    x = (torch.randn(B, C, H, H, device=cuda) * 1.7 + 0.3)
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
    x.requires_grad_()
    ours = enn.BatchNorm2d(C, eps).to(cuda)
    ref = torch.nn.BatchNorm2d(C, eps).to(cuda)
    with torch.no_grad():
        ours.weight.normal_(1, 0.2); ours.bias.normal_(0, 0.2)
        ref.weight.copy_(ours.weight); ref.bias.copy_(ours.bias)
    if act == "relu":
        seq_o, seq_r = enn.Sequential(ours, enn.ReLU()), torch.nn.Sequential(ref, torch.nn.ReLU())
    elif act == "lrelu":
        seq_o = enn.Sequential(ours, enn.LeakyReLU(0.2, inplace=True))
        seq_r = torch.nn.Sequential(ref, torch.nn.LeakyReLU(0.2, inplace=True))
    else:
        seq_o, seq_r = enn.Sequential(ours), torch.nn.Sequential(ref)
    xr = x.detach().clone().contiguous().requires_grad_()
    yo, yr = seq_o(x), seq_r(xr)
    assert rel_err(yo, yr) <= TOL
    assert rel_err(ours.running_mean, ref.running_mean) <= TOL
    assert rel_err(ours.running_var, ref.running_var) <= TOL
    assert int(ours.num_batches_tracked) == 1
    go = torch.randn_like(yr)
    gx, gg, gb = torch.autograd.grad(yo, (x, ours.weight, ours.bias), go)
    rx, rg, rb = torch.autograd.grad(yr, (xr, ref.weight, ref.bias), go)
    assert rel_err(gx, rx) <= 2e-5 and rel_err(gg, rg) <= 2e-5 and rel_err(gb, rb) <= 2e-5
    # eval mode uses the running statistics
    ours.eval(); ref.eval()
    assert rel_err(seq_o(x.detach()), seq_r(xr.detach())) <= TOL


@pytest.mark.parametrize("kind", ["lrelu", "relu", "tanh", "sigmoid"])
def test_activations(cuda, kind):
    import eadgan_b200.nn as enn
    torch.manual_seed(5)
    x = torch.randn(3, 7, 5, 11, device=cuda, requires_grad=True)
    mo, mr = {"lrelu": (enn.LeakyReLU(0.2), torch.nn.LeakyReLU(0.2)), "relu": (enn.ReLU(), torch.nn.ReLU()),
              "tanh": (enn.Tanh(), torch.nn.Tanh()), "sigmoid": (enn.Sigmoid(), torch.nn.Sigmoid())}[kind]
    yo, yr = mo(x), mr(x)
    assert rel_err(yo, yr) <= TOL
    go = torch.randn_like(yr)
    assert rel_err(torch.autograd.grad(yo, x, go)[0], torch.autograd.grad(yr, x, go)[0]) <= TOL


def test_inplace_leaky_relu(cuda):
    import eadgan_b200.nn as enn
    x = torch.randn(4, 8, device=cuda)
    ref = TF.leaky_relu(x, 0.1)
    y = x.clone()
    out = enn.LeakyReLU(0.1, inplace=True)(y)
    assert out.data_ptr() == y.data_ptr() and torch.equal(out, ref)


def test_softmax_upsample(cuda):
    import eadgan_b200.functional as Fn
    torch.manual_seed(6)
    x = torch.randn(33, 10, device=cuda, requires_grad=True)
    yo, yr = Fn.softmax(x), torch.softmax(x, 1)
    assert rel_err(yo, yr) <= TOL
    go = torch.randn_like(yr)
    assert rel_err(torch.autograd.grad(yo, x, go)[0], torch.autograd.grad(yr, x, go)[0]) <= TOL
    u = torch.randn(2, 5, 8, 8, device=cuda, requires_grad=True)
    uo, ur = Fn.upsample2x(u), TF.interpolate(u, scale_factor=2)
    assert torch.equal(uo, ur)
    go = torch.randn_like(ur)
    assert rel_err(torch.autograd.grad(uo, u, go)[0], torch.autograd.grad(ur, u, go)[0]) <= TOL


@pytest.mark.parametrize("shape", [(128, 3, 4, 4), (1024, 512, 4, 4), (32, 1, 4, 4), (128, 1024), (16, 1, 3, 3)])
def test_spectral_norm(cuda, shape):
    """weight, u/v evolution over 3 training forwards, sigma gradient, eval mode."""
    import eadgan_b200.nn as enn
    torch.manual_seed(7)
    if len(shape) == 4:
        mk = lambda ns: ns.Conv2d(shape[1], shape[0], shape[2], 2, 1)
        x = torch.randn(2, shape[1], 8, 8, device=cuda)
    else:
        mk = lambda ns: ns.Linear(shape[1], shape[0])
        x = torch.randn(2, shape[1], device=cuda)
    torch.manual_seed(8)
    ours = enn.spectral_norm(mk(enn)).to(cuda)
    torch.manual_seed(8)
    ref = torch.nn.utils.spectral_norm(mk(torch.nn)).to(cuda)
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    for k in ref.state_dict():
        assert torch.equal(ours.state_dict()[k], ref.state_dict()[k]), k
    for it in range(3):
        yo, yr = ours(x), ref(x)
        assert rel_err(ours.weight, ref.weight) <= TOL, it
        assert rel_err(ours.weight_u, ref.weight_u) <= TOL and rel_err(ours.weight_v, ref.weight_v) <= TOL
        assert rel_err(yo, yr) <= 2e-5
    go = torch.randn_like(yr)
    go_ = torch.autograd.grad(yo, ours.weight_orig, go)[0]
    gr_ = torch.autograd.grad(yr, ref.weight_orig, go)[0]
    assert rel_err(go_, gr_) <= 2e-5
    ours.eval(); ref.eval()
    u0 = ours.weight_u.clone()
    assert rel_err(ours(x), ref(x)) <= 2e-5 and torch.equal(u0, ours.weight_u)


def test_losses(cuda):
    import eadgan_b200.nn as enn
    import eadgan_b200.functional as Fn
    torch.manual_seed(9)
    for shape in [(37,), (64, 1), (4100,)]:
        p = torch.rand(*shape, device=cuda).clamp(1e-4, 1 - 1e-4).requires_grad_()
        t = (torch.rand(*shape, device=cuda) > 0.5).float()
        lo, lr = enn.BCELoss()(p, t), torch.nn.BCELoss()(p, t)
        assert rel_err(lo, lr) <= TOL
        assert rel_err(torch.autograd.grad(lo * 0.5, p)[0], torch.autograd.grad(lr * 0.5, p)[0]) <= TOL
    pe = torch.tensor([0.0, 1.0, 0.5], device=cuda)  # log clamp at -100
    te = torch.tensor([1.0, 0.0, 1.0], device=cuda)
    assert rel_err(enn.BCELoss()(pe, te), torch.nn.BCELoss()(pe, te)) <= TOL
    a = torch.randn(50, 8, device=cuda, requires_grad=True)
    b = torch.randn(50, 8, device=cuda)
    lo, lr = enn.MSELoss()(a, b), torch.nn.MSELoss()(a, b)
    assert rel_err(lo, lr) <= TOL
    assert rel_err(torch.autograd.grad(lo, a)[0], torch.autograd.grad(lr, a)[0]) <= TOL
    # CrossEntropyLoss applied to softmax OUTPUTS, as the reference does (celebA/EAD-GAN_celebA.py:383)
    logits = torch.randn(41, 10, device=cuda, requires_grad=True)
    lab = torch.randint(0, 10, (41,), device=cuda)
    lo = enn.CrossEntropyLoss()(Fn.softmax(logits), lab)
    lr = torch.nn.CrossEntropyLoss()(torch.softmax(logits, 1), lab)
    assert rel_err(lo, lr) <= TOL
    assert rel_err(torch.autograd.grad(lo, logits)[0], torch.autograd.grad(lr, logits)[0]) <= 2e-5
    # mutual_info_loss (dSprites/rp.py:225-232)
    q = torch.softmax(torch.randn(30, 3, device=cuda), 1).requires_grad_()
    c = torch.softmax(torch.randn(30, 3, device=cuda), 1)
    ref = torch.mean(-torch.sum(torch.log(q + 1e-8) * c, dim=1)) + torch.mean(-torch.sum(torch.log(c + 1e-8) * c, dim=1))
    lo = Fn.mutual_info_loss(q, c)
    assert rel_err(lo, ref) <= TOL
    assert rel_err(torch.autograd.grad(lo, q)[0], torch.autograd.grad(ref, q)[0]) <= TOL


@pytest.mark.parametrize("lr,betas", [(1e-3, (0.5, 0.999)), (2e-4, (0.5, 0.999)), (1e-4, (0.9, 0.99))])
def test_adam_matches_torch(cuda, lr, betas):
    """t = 1, 2, 3 on identical gradients; m, v and p compared (SURVEY.md section 7.3-1 iii)."""
    from eadgan_b200.optim import Adam
    torch.manual_seed(10)
    shapes = [(1024, 512, 4, 4), (19,), (128, 3, 4, 4), (7, 5), (1,)]
    po = [torch.randn(*s, device=cuda).requires_grad_() for s in shapes]
    pr = [p.detach().clone().requires_grad_() for p in po]
    oo, orf = Adam(po, lr=lr, betas=betas), torch.optim.Adam(pr, lr=lr, betas=betas, foreach=False)
    for t in range(3):
        for a, b in zip(po, pr):
            g = torch.randn_like(a) * (10.0 ** (t - 1))
            a.grad, b.grad = g.clone(), g.clone()
        oo.step(); orf.step()
        for a, b in zip(po, pr):
            assert rel_err(a, b) <= 2e-6
            assert rel_err(oo.state[a]["exp_avg"], orf.state[b]["exp_avg"]) <= 2e-6
            assert rel_err(oo.state[a]["exp_avg_sq"], orf.state[b]["exp_avg_sq"]) <= 2e-6


def test_cpu_tensor_is_a_hard_error(cuda):
    import eadgan_b200.nn as enn
    with pytest.raises(RuntimeError, match="no CPU"):
        enn.Conv2d(3, 4, 4, 2, 1)(torch.zeros(1, 3, 8, 8))
