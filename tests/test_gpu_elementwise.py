"""The HBM-bound passes of a residual block (BatchNorm statistics / apply / backward, residual combine, column sums)
through the C ABI against fp64 torch restatements of the reference's op sequence (networks/ResidualBlocks.py:84-97:
bn -> relu -> conv -> dropout -> bn -> relu -> conv -> dropout; out = a * BN(shortcut) + b * out), on geometries that
exercise every code path of the staged kernels: bordered inputs / outputs, chunks that do not divide the row, C = 640
(80 octets: 3 pixels per chunk), 1-D element masks, tiny 1x1 maps, fp32 and bf16 storage."""
import pytest
import torch

pytestmark = pytest.mark.gpu

EPS = 1e-5

# (B, H, W, C, nd)
GEOS = [
    (4, 8, 8, 128, 2),
    (3, 5, 7, 64, 2),        # ragged: 7 pixels x 8 octets, odd sizes
    (2, 4, 4, 640, 2),       # 80 octets per pixel: chunks of 2 pixels
    (5, 1, 1, 640, 2),       # 1x1 maps
    (2, 16, 40, 256, 2),     # 32 octets: 8 pixels per chunk, 5 chunks per row
    (3, 1, 96, 128, 1),      # 1-D text: element masks
    (2, 1, 33, 384, 1),
    (6, 64, 64, 128, 2),     # the bench's largest layer shape (small batch)
]


def _eng(dtype):
    from mopoe_mimic_b200 import _lib as L
    from mopoe_mimic_b200.engine import Engine
    return Engine('cuda', dtype, L.IMPL_AUTO), L


def _act(vals, ph, pw, dtype, fill=7.0):
    """[B,H,W,C] fp64 cuda -> Act with a border filled with a sentinel (inputs) """
    from mopoe_mimic_b200.engine import Act
    B, H, W, Cc = vals.shape
    t = torch.full((B, H + 2 * ph, W + 2 * pw, Cc), fill, dtype=dtype, device='cuda')
    t[:, ph:ph + H, pw:pw + W] = vals.to(dtype)
    return Act(t, B, H, W, Cc, ph, pw)


def _out(B, H, W, Cc, ph, pw, dtype):
    from mopoe_mimic_b200.engine import Act
    t = torch.full((B, H + 2 * ph, W + 2 * pw, Cc), 5.0, dtype=dtype, device='cuda')
    return Act(t, B, H, W, Cc, ph, pw)


def _border_is_zero(act):
    t = act.t.float().clone()
    t[:, act.ph:act.ph + act.H, act.pw:act.pw + act.W] = 0
    return bool((t == 0).all())


def _rand(shape, seed, dtype):
    g = torch.Generator(device='cuda').manual_seed(seed)
    x = torch.randn(shape, generator=g, device='cuda', dtype=torch.float32)
    return x.to(dtype).double()           # exactly representable in the storage dtype


def _mask(B, H, W, Cc, nd, seed, L):
    g = torch.Generator(device='cuda').manual_seed(seed)
    if nd == 2:
        m = (torch.rand(B, Cc, generator=g, device='cuda') > 0.5).to(torch.uint8)
        return m, L.MASK_BC, (2.0 * m.double()).view(B, 1, 1, Cc)
    m = (torch.rand(B, W, Cc, generator=g, device='cuda') > 0.5).to(torch.uint8)
    return m, L.MASK_ELEM, (2.0 * m.double()).view(B, 1, W, Cc)


def _tol(dtype):
    return (2e-5, 2e-5) if dtype == torch.float32 else (4e-3, 4e-3)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('B,H,W,Cc,nd', GEOS)
@pytest.mark.parametrize('masked', [False, True])
def test_bn_stats_apply_and_backward(dtype, B, H, W, Cc, nd, masked):
    eng, L = _eng(dtype)
    pin = 1                                   # input border (as the strided convs need it)
    ph_in, pw_in = (pin if nd == 2 else 0), pin
    x64 = _rand((B, H, W, Cc), 1, dtype)
    x = _act(x64, ph_in, pw_in, dtype)
    if masked:
        m, mode, mk = _mask(B, H, W, Cc, nd, 2, L)
    else:
        m, mode, mk = None, L.MASK_NONE, torch.ones(1, 1, 1, 1, dtype=torch.float64, device='cuda')
    gamma = (torch.rand(Cc, device='cuda') + 0.5)
    beta = torch.randn(Cc, device='cuda') * 0.3
    rmean, rvar = torch.zeros(Cc, device='cuda'), torch.ones(Cc, device='cuda')
    # ---- statistics
    st = eng.bn_stats(x, m, mode, rmean, rvar)
    v = x64 * mk
    mean = v.mean(dim=(0, 1, 2))
    var = v.var(dim=(0, 1, 2), unbiased=False)
    n = B * H * W
    assert torch.allclose(st[0].double(), mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(st[1].double(), 1.0 / torch.sqrt(var + EPS), rtol=1e-5, atol=1e-6)
    assert torch.allclose(rmean.double(), 0.1 * mean, rtol=1e-5, atol=1e-6)
    unb = var * n / (n - 1) if n > 1 else var
    assert torch.allclose(rvar.double(), 0.9 + 0.1 * unb, rtol=1e-5, atol=1e-6)
    # ---- apply (+ReLU) into a bordered output
    pho, pwo = (1 if nd == 2 else 0), 1
    a = _out(B, H, W, Cc, pho, pwo, dtype)
    eng.bn_apply(x, m, mode, st, gamma, beta, True, a)
    torch.cuda.synchronize()
    is64 = st[1].double()
    ref = torch.relu((v - st[0].double()) * is64 * gamma.double() + beta.double())
    rt, at = _tol(dtype)
    assert torch.allclose(a.interior().double(), ref, rtol=rt, atol=at)
    assert _border_is_zero(a)
    # ---- backward of y = relu(BN(x * 2mask)) given dy, + addend, into a bordered output
    dy64 = _rand((B, H, W, Cc), 3, dtype)
    add64 = _rand((B, H, W, Cc), 4, dtype)
    dy = _act(dy64, 0, 0, dtype)
    addend = _act(add64, 0, 0, dtype)
    dg, db = torch.zeros(Cc, device='cuda'), torch.zeros(Cc, device='cuda')
    dx = _out(B, H, W, Cc, ph_in, pw_in, dtype)
    eng.bn_bwd(dy, a, 1.0, x, m, mode, st, gamma, dg, db, addend, dx)
    torch.cuda.synchronize()
    gate = (a.interior().double() > 0).double()
    g = dy64 * gate
    xh = (v - st[0].double()) * is64
    sg, sgx = g.sum(dim=(0, 1, 2)), (g * xh).sum(dim=(0, 1, 2))
    ref_dx = gamma.double() * is64 * (g - sg / n - xh * sgx / n) * mk + add64
    scale = float(ref_dx.abs().max()) + 1e-30
    err = float((dx.interior().double() - ref_dx).abs().max()) / scale
    assert err < (1e-4 if dtype == torch.float32 else 1.2e-2), err
    assert _border_is_zero(dx)
    assert torch.allclose(db.double(), sg, rtol=1e-4, atol=1e-4 * float(sg.abs().max() + 1))
    assert torch.allclose(dg.double(), sgx, rtol=1e-4, atol=1e-4 * float(sgx.abs().max() + 1))
    # ---- without gate / addend, accumulate into dgamma / dbeta
    dx2 = _out(B, H, W, Cc, 0, 0, dtype)
    eng.bn_bwd(dy, None, 0.5, x, m, mode, st, gamma, dg, db, None, dx2, accumulate=True)
    torch.cuda.synchronize()
    g2 = 0.5 * dy64
    sg2, sgx2 = g2.sum(dim=(0, 1, 2)), (g2 * xh).sum(dim=(0, 1, 2))
    ref2 = gamma.double() * is64 * (g2 - sg2 / n - xh * sgx2 / n) * mk
    err = float((dx2.interior().double() - ref2).abs().max()) / (float(ref2.abs().max()) + 1e-30)
    assert err < (1e-4 if dtype == torch.float32 else 1.2e-2), err
    assert torch.allclose(db.double(), sg + sg2, rtol=1e-4, atol=1e-4 * float(sg.abs().max() + 1))


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('B,H,W,Cc,nd', GEOS)
def test_combine_and_its_backward(dtype, B, H, W, Cc, nd):
    eng, L = _eng(dtype)
    r64, c64 = _rand((B, H, W, Cc), 5, dtype), _rand((B, H, W, Cc), 6, dtype)
    r, c = _act(r64, 0, 0, dtype), _act(c64, 0, 0, dtype)
    m, mode, mk = _mask(B, H, W, Cc, nd, 7, L)
    gamma = (torch.rand(Cc, device='cuda') + 0.5)
    beta = torch.randn(Cc, device='cuda') * 0.3
    st = eng.bn_stats(r, None, L.MASK_NONE)
    pho, pwo = (1 if nd == 2 else 0), 1
    y = _out(B, H, W, Cc, pho, pwo, dtype)
    eng.combine(r, st, gamma, beta, c, m, mode, 2.0, 0.3, y)
    torch.cuda.synchronize()
    is64, mu64 = st[1].double(), st[0].double()
    ref = 2.0 * ((r64 - mu64) * is64 * gamma.double() + beta.double()) + 0.3 * (c64 * mk)
    rt, at = _tol(dtype)
    assert torch.allclose(y.interior().double(), ref, rtol=rt, atol=at * 3)
    assert _border_is_zero(y)
    # combine + the statistics of its (stored) output for the next block's bn1, in one pass
    y2 = _out(B, H, W, Cc, pho, pwo, dtype)
    rm, rv = torch.zeros(Cc, device='cuda'), torch.ones(Cc, device='cuda')
    _, st_next = eng.combine(r, st, gamma, beta, c, m, mode, 2.0, 0.3, y2, bn=(rm, rv))
    torch.cuda.synchronize()
    assert torch.equal(y2.t, y.t)
    yv = y.interior().double()
    n = B * H * W
    mean, var = yv.mean(dim=(0, 1, 2)), yv.var(dim=(0, 1, 2), unbiased=False)
    assert torch.allclose(st_next[0].double(), mean, rtol=1e-5, atol=2e-6)
    assert torch.allclose(st_next[1].double(), 1.0 / torch.sqrt(var + EPS), rtol=1e-5, atol=1e-6)
    assert torch.allclose(rm.double(), 0.1 * mean, rtol=1e-5, atol=2e-6)
    assert torch.allclose(rv.double(), 0.9 + 0.1 * (var * n / (n - 1) if n > 1 else var), rtol=1e-5, atol=1e-6)
    # backward: dr = BN-backward(a * dy), dc = b * dy * 2mask
    dy64 = _rand((B, H, W, Cc), 8, dtype)
    dy = _act(dy64, pho, pwo, dtype)
    dg, db = torch.zeros(Cc, device='cuda'), torch.zeros(Cc, device='cuda')
    dr, dc = _out(B, H, W, Cc, pho, pwo, dtype), _out(B, H, W, Cc, pho, pwo, dtype)
    eng.combine_bwd(dy, 2.0, r, st, gamma, dg, db, m, mode, 0.3, dr, dc)
    torch.cuda.synchronize()
    n = B * H * W
    g = 2.0 * dy64
    xh = (r64 - mu64) * is64
    sg, sgx = g.sum(dim=(0, 1, 2)), (g * xh).sum(dim=(0, 1, 2))
    ref_dr = gamma.double() * is64 * (g - sg / n - xh * sgx / n)
    ref_dc = 0.3 * dy64 * mk
    e1 = float((dr.interior().double() - ref_dr).abs().max()) / (float(ref_dr.abs().max()) + 1e-30)
    e2 = float((dc.interior().double() - ref_dc).abs().max()) / (float(ref_dc.abs().max()) + 1e-30)
    lim = 1e-4 if dtype == torch.float32 else 1.2e-2
    assert e1 < lim and e2 < lim, (e1, e2)
    assert _border_is_zero(dr) and _border_is_zero(dc)
    assert torch.allclose(db.double(), sg, rtol=1e-4, atol=1e-4 * float(sg.abs().max() + 1))
    assert torch.allclose(dg.double(), sgx, rtol=1e-4, atol=1e-4 * float(sgx.abs().max() + 1))
    # column sums (bias gradients)
    cs = eng.colsum(dy)
    assert torch.allclose(cs.double(), dy64.sum(dim=(0, 1, 2)), rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('B,H,W,Cc,nd', [GEOS[0], GEOS[2], GEOS[5], GEOS[7]])
@pytest.mark.parametrize('masked', [False, True])
def test_bn_backward_with_the_gate_recomputed(dtype, B, H, W, Cc, nd, masked):
    """With the BatchNorm's own (gamma, beta) both backward passes recompute the ReLU gate from x with the forward pass's
    instruction sequence instead of reading the saved activation (mopoe_bn_bwd_reduce / _apply, gate_gamma / gate_beta):
    one staged operand less, and every decision — hence every output bit — identical to the pass that reads the gate."""
    eng, L = _eng(dtype)
    x64 = _rand((B, H, W, Cc), 11, dtype)
    x = _act(x64, 1 if nd == 2 else 0, 1, dtype)
    if masked:
        m, mode, mk = _mask(B, H, W, Cc, nd, 12, L)
    else:
        m, mode, mk = None, L.MASK_NONE, None
    gamma = (torch.rand(Cc, device='cuda') + 0.5)
    beta = torch.randn(Cc, device='cuda') * 0.3
    gamma[3], beta[3] = 0.01, 0.5
    gamma[Cc - 2], beta[Cc - 2] = 0.0, 0.25          # constant activation relu(beta): gate open everywhere
    gamma[5] = -0.7                                  # negative scale
    st = eng.bn_stats(x, m, mode)
    a = _out(B, H, W, Cc, 0, 0, dtype)
    eng.bn_apply(x, m, mode, st, gamma, beta, True, a)
    dy = _act(_rand((B, H, W, Cc), 13, dtype), 0, 0, dtype)
    addend = _act(_rand((B, H, W, Cc), 14, dtype), 0, 0, dtype)
    res = {}
    for name, b_arg in (('read', None), ('recomputed', beta)):
        dg, db = torch.zeros(Cc, device='cuda'), torch.zeros(Cc, device='cuda')
        dx = _out(B, H, W, Cc, 0, 0, dtype)
        dx2 = _out(B, H, W, Cc, 1 if nd == 2 else 0, 1, dtype)
        eng.bn_bwd(dy, a, 1.0, x, m, mode, st, gamma, dg, db, None, dx, beta=b_arg)
        eng.bn_bwd(dy, a, 0.5, x, m, mode, st, gamma, dg, db, addend, dx2, accumulate=True, beta=b_arg)
        torch.cuda.synchronize()
        res[name] = (dg, db, dx.t, dx2.t)
    for got, want in zip(res['recomputed'], res['read']):
        assert torch.equal(got, want)
    # (and the gate itself: open exactly where the stored activation is positive — checked through sum g)
    gate = (a.interior().double() > 0).double()
    sg = (dy.interior().double() * gate).sum(dim=(0, 1, 2)) * 1.5
    assert torch.allclose(res['recomputed'][1].double(), sg, rtol=1e-4, atol=1e-4 * float(sg.abs().max() + 1))
