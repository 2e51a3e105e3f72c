"""Implicit-GEMM kernels through the C ABI against torch CPU convolutions (the reference's own call sites:
nn.Conv*/nn.ConvTranspose*), for both kernel families (CUDA-core fp32 path and tcgen05 bf16 path)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _eng(dtype, impl):
    from mopoe_mimic_b200 import _lib as L
    from mopoe_mimic_b200.engine import Engine
    e = Engine('cuda', dtype, {'simt': L.IMPL_SIMT, 'tc': L.IMPL_TC, 'auto': L.IMPL_AUTO}[impl])
    return e


def _act(x_nchw, pad, dtype, nd):
    """NCHW / NCL cpu tensor -> bordered channels-last Act on the GPU"""
    from mopoe_mimic_b200.engine import Act
    if nd == 1:
        B, Cc, W = x_nchw.shape
        H = 1
        cl = x_nchw.permute(0, 2, 1).reshape(B, 1, W, Cc)
        ph, pw = 0, pad
    else:
        B, Cc, H, W = x_nchw.shape
        cl = x_nchw.permute(0, 2, 3, 1)
        ph = pw = pad
    t = torch.zeros(B, H + 2 * ph, W + 2 * pw, Cc, dtype=dtype, device='cuda')
    t[:, ph:ph + H, pw:pw + W] = cl.to(dtype).cuda()
    return Act(t, B, H, W, Cc, ph, pw)


def _to_nchw(act, nd):
    t = act.interior().float().cpu()
    return t.permute(0, 3, 1, 2) if nd == 2 else t[:, 0].permute(0, 2, 1)


def _rand(shape, seed, scale=1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(shape, generator=g) * 2 - 1) * scale
    return x.to(dtype).float()      # values exactly representable in `dtype`


CASES = [
    # (nd, B, Cin, Cout, spatial)
    (2, 4, 64, 128, 16),
    (2, 2, 128, 256, 32),
    (2, 3, 64, 192, 8),
    (2, 16, 128, 160, 4),
    (1, 4, 64, 128, 64),
    (1, 2, 128, 96, 512),
    (2, 5, 128, 640, 8),
    # >= 148 output tiles: the CTA-pair kernel (tcgen05.mma.cta_group::2, 256-row MMAs over two SMs)
    (2, 32, 64, 128, 64),        # 256 m-tiles x 1 n-tile (fprop), 4 phases x 1024 m-tiles (dgrad)
    (2, 31, 64, 256, 40),        # 155 m-tiles: ODD -> the last pair's second CTA works on an out-of-range tile
    (1, 64, 64, 128, 1024),      # 1-D
    (2, 9, 128, 384, 32),        # BN = 192: 96 weight columns per CTA
    (2, 16, 128, 512, 16),       # wgrad: 4 n-tiles -> 2 n-tile pairs per window tile (CTA-pair weight-gradient kernel)
    (2, 256, 64, 640, 8),        # N = 640 at M = 4096 rows: 192-column tiles (one wave of 128 tiles), last tile overhangs N
]


@pytest.mark.parametrize('impl,dtype', [('simt', torch.float32), ('simt', torch.bfloat16), ('tc', torch.bfloat16)])
@pytest.mark.parametrize('nd,B,ci,co,sp', CASES)
def test_conv_k4s2p1_fwd_dgrad_wgrad(impl, dtype, nd, B, ci, co, sp):
    from mopoe_mimic_b200.engine import conv_form, conv_form_grad, phase_form
    eng = _eng(dtype, impl)
    shp = (B, ci, sp) if nd == 1 else (B, ci, sp, sp)
    wshape = (co, ci, 4) if nd == 1 else (co, ci, 4, 4)
    x = _rand(shp, 1, 1.0, dtype)
    w = _rand(wshape, 2, 0.05, dtype)
    bias = _rand((co,), 3, 0.1)
    conv = F.conv1d if nd == 1 else F.conv2d
    convt = F.conv_transpose1d if nd == 1 else F.conv_transpose2d
    tol = 2e-5 if dtype == torch.float32 else 1.5e-2
    # forward
    ref = conv(x, w, bias, stride=2, padding=1)
    xa = _act(x, 1, dtype, nd)
    out = eng.gemm_down(xa, conv_form(w.cuda(), dtype), bias.cuda(), 4, 2, 1, co)
    torch.cuda.synchronize()
    got = _to_nchw(out, nd)
    assert (got - ref).abs().max() <= tol * ref.abs().max()
    # dgrad = transposed conv of the output gradient
    gshape = ref.shape
    g = _rand(gshape, 4, 1.0, dtype)
    ref_dx = convt(g, w, None, stride=2, padding=1)
    ga = _act(g, 1, dtype, nd)
    dx = eng.gemm_up(ga, phase_form(w.cuda(), dtype), None, ci)
    torch.cuda.synchronize()
    assert (_to_nchw(dx, nd) - ref_dx).abs().max() <= tol * ref_dx.abs().max()
    # wgrad
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    conv(xr, wr, None, stride=2, padding=1).backward(g)
    from mopoe_mimic_b200 import _lib as L
    if impl == 'tc' and not L.load().mopoe_tc_wgrad_built():
        eng.impl = L.IMPL_AUTO
    gw = eng.wgrad_down(xa, 4, 2, 1, ga)
    torch.cuda.synchronize()
    gw = conv_form_grad(gw, wshape).cpu()
    assert (gw - wr.grad).abs().max() <= tol * wr.grad.abs().max()


@pytest.mark.parametrize('impl,dtype', [('simt', torch.float32), ('tc', torch.bfloat16)])
def test_pointwise_and_rows(impl, dtype):
    eng = _eng(dtype, impl)
    B, Cc, N, sp = 4, 128, 192, 8
    x = _rand((B, Cc, sp, sp), 5, 1.0, dtype)
    w = _rand((N, Cc), 6, 0.05, dtype)
    bias = _rand((N,), 7, 0.1)
    ref = F.conv2d(x, w.view(N, Cc, 1, 1), bias)
    tol = 2e-5 if dtype == torch.float32 else 1.5e-2
    for pad in (0, 1):
        xa = _act(x, pad, dtype, 2)
        out = eng.gemm_rows(xa, w.to(dtype).cuda().contiguous(), bias.cuda(), N)
        torch.cuda.synchronize()
        assert (_to_nchw(out, 2) - ref).abs().max() <= tol * ref.abs().max()


@pytest.mark.parametrize('impl,dtype', [('simt', torch.float32), ('tc', torch.bfloat16)])
def test_valid_4to1_and_1to4(impl, dtype):
    """k4/s2/p0 conv on 4x4 -> 1x1 (encoder tail) and k4/s1/p0 deconv 1x1 -> 4x4 (decoder head)"""
    from mopoe_mimic_b200.engine import conv_form, full_form
    eng = _eng(dtype, impl)
    B, ci, co = 128, 128, 64
    tol = 2e-5 if dtype == torch.float32 else 1.5e-2
    x = _rand((B, ci, 4, 4), 8, 1.0, dtype)
    w = _rand((co, ci, 4, 4), 9, 0.05, dtype)
    ref = F.conv2d(x, w, None, stride=2, padding=0)
    xa = _act(x, 1, dtype, 2)
    out = eng.gemm_down(xa, conv_form(w.cuda(), dtype), None, 4, 2, 0, co)
    torch.cuda.synchronize()
    assert (_to_nchw(out, 2) - ref).abs().max() <= tol * ref.abs().max()
    z = _rand((B, ci, 1, 1), 10, 1.0, dtype)
    wt = _rand((ci, co, 4, 4), 11, 0.05, dtype)
    ref2 = F.conv_transpose2d(z, wt, None, stride=1, padding=0)
    za = _act(z, 0, dtype, 2)
    out2 = eng.gemm_rows(za, full_form(wt.cuda(), dtype), None, 16 * co, out_shape=(B, 4, 4, co))
    torch.cuda.synchronize()
    assert (_to_nchw(out2, 2) - ref2).abs().max() <= tol * ref2.abs().max()


@pytest.mark.parametrize('B,ci,co,bias_on,bn_on', [(256, 512, 640, False, False), (200, 640, 640, True, True), (128, 128, 64, True, False)])
def test_split_k_weight_bound_gemm(B, ci, co, bias_on, bn_on):
    """the 4x4 -> 1x1 conv at the bottom of the image encoder (M = batch rows, K = 16 * Cin = 8192-10240): few output tiles,
    long reduction -> split-K work items + fixed-order finish kernel (mopoe_conv_gemm_splitk); with and without the
    BatchNorm statistics of the output; the library must also report the path as taken for these shapes."""
    import ctypes as C
    import os
    from mopoe_mimic_b200 import _lib as L
    from mopoe_mimic_b200.engine import conv_form
    if os.environ.get('MOPOE_GEMM_SPLITK', '0') != '1':
        pytest.skip('split-K is opt-in (MOPOE_GEMM_SPLITK=1): no step-time gain measured; run this test with the switch set')
    dtype = torch.bfloat16
    eng = _eng(dtype, 'tc')
    x = _rand((B, ci, 4, 4), 8, 1.0, dtype)
    w = _rand((co, ci, 4, 4), 9, 0.02, dtype)
    bias = _rand((co,), 10, 0.5) if bias_on else None
    ref = F.conv2d(x, w, bias, stride=2, padding=0)
    xa = _act(x, 1, dtype, 2)
    win, OH, OW = eng.win_down(xa, 4, 2, 0)
    from mopoe_mimic_b200.engine import Act
    probe = Act.empty(B, OH, OW, co, 0, 0, dtype, eng.device)
    assert L.load().mopoe_conv_gemm_splitk_ws(C.byref(win), C.byref(eng.rows_of(probe)), eng.impl) > 0
    wc = conv_form(w.cuda(), dtype)
    bn = None
    if bn_on:
        rm, rv = torch.zeros(co, device='cuda'), torch.ones(co, device='cuda')
        bn = (None, L.MASK_NONE, rm, rv)
    res = eng.gemm_down(xa, wc, bias.cuda() if bias_on else None, 4, 2, 0, co, bn=bn)
    out, st = res if bn_on else (res, None)
    torch.cuda.synchronize()
    got = _to_nchw(out, 2)
    assert (got - ref).abs().max() <= 1.5e-2 * ref.abs().max()
    if bn_on:
        v = out.interior().double()
        assert torch.allclose(st[0].double(), v.mean(dim=(0, 1, 2)), rtol=1e-5, atol=1e-5)
        assert torch.allclose(st[1].double(), 1 / torch.sqrt(v.var(dim=(0, 1, 2), unbiased=False) + 1e-5), rtol=1e-5, atol=1e-5)
    # deterministic: the finish kernel sums the slices in a fixed order
    out2 = eng.gemm_down(xa, wc, bias.cuda() if bias_on else None, 4, 2, 0, co)
    torch.cuda.synchronize()
    assert torch.equal(out2.t, out.t)


BN_CASES = [
    # (kind, nd, B, Cin, Cout, spatial, mask)
    ('rows', 2, 4, 128, 128, 16, 'bc'),          # conv1 of a 2-D block: Dropout2d mask [B, C]
    ('rows', 2, 3, 64, 256, 7, 'bc'),            # ragged: 147 rows, not a multiple of the 128-row tile
    ('rows', 1, 4, 128, 384, 100, 'elem'),       # conv1 of a 1-D block: elementwise mask [B, L, C]; BN = 192
    ('rows', 2, 2, 64, 640, 4, 'bc'),            # wide layer: 5 n-tiles of 128
    ('rows', 2, 2, 64, 96, 8, 'bc'),             # BN = 96: register epilogue + separate statistics pass inside the library
    ('down', 2, 4, 64, 128, 16, None),           # shortcut conv k4 s2 p1
    ('down', 1, 3, 64, 512, 64, None),
    ('up', 2, 4, 128, 64, 8, None),              # shortcut deconv: 4 sub-pixel phases in one launch
    ('up', 1, 5, 64, 256, 32, None),             # 2 phases
    # CTA-pair kernel (>= 148 tiles)
    ('rows', 2, 20, 128, 128, 32, 'bc'),         # 160 m-tiles, Dropout2d mask
    ('rows', 1, 21, 128, 256, 1000, 'elem'),     # 165 m-tiles (odd pairs), elementwise mask
    ('down', 2, 19, 64, 256, 64, None),          # 152 m-tiles
    ('up', 2, 10, 128, 64, 16, None),            # 4 phases x 20 m-tiles x 1 n-tile = 80 ... below the threshold: single-CTA path
    ('up', 2, 40, 128, 128, 16, None),           # 4 phases x 80 m-tiles = 320 tiles
]


@pytest.mark.parametrize('impl,dtype', [('tc', torch.bfloat16), ('simt', torch.float32)])
@pytest.mark.parametrize('kind,nd,B,ci,co,sp,mask', BN_CASES)
def test_gemm_with_fused_batchnorm_statistics(impl, dtype, kind, nd, B, ci, co, sp, mask):
    """mopoe_conv_gemm_bn: the GEMM output must equal the plain launch bit for bit, and the statistics must equal the
    separate reduction pass over the stored output (the path it replaces) and torch's own batch statistics."""
    from mopoe_mimic_b200 import _lib as L
    from mopoe_mimic_b200.engine import conv_form, phase_form
    eng = _eng(dtype, impl)
    shp = (B, ci, sp) if nd == 1 else (B, ci, sp, sp)
    x = _rand(shp, 21, 1.0, dtype)
    bias = _rand((co,), 23, 0.5).cuda()
    g = torch.Generator().manual_seed(24)
    rm0, rv0 = torch.randn(co, generator=g), torch.rand(co, generator=g) + 0.5

    def run(bn):
        if kind == 'rows':
            w = _rand((co, ci), 22, 0.05, dtype).to(dtype).cuda().contiguous()
            return eng.gemm_rows(_act(x, 0, dtype, nd), w, bias, co, bn=bn)
        if kind == 'down':
            w = _rand((co, ci, 4) if nd == 1 else (co, ci, 4, 4), 22, 0.05, dtype)
            return eng.gemm_down(_act(x, 1, dtype, nd), conv_form(w.cuda(), dtype), bias, 4, 2, 1, co, bn=bn)
        w = _rand((ci, co, 4) if nd == 1 else (ci, co, 4, 4), 22, 0.05, dtype)
        return eng.gemm_up(_act(x, 1, dtype, nd), phase_form(w.cuda(), dtype), bias, co, bn=bn)
    plain = run(None)
    rows = plain.B * plain.H * plain.W
    mk, mode = None, L.MASK_NONE
    if mask == 'bc':
        mk, mode = (torch.rand(B * co, generator=g) < 0.5).to(torch.uint8).cuda(), L.MASK_BC
    elif mask == 'elem':
        mk, mode = (torch.rand(rows * co, generator=g) < 0.5).to(torch.uint8).cuda(), L.MASK_ELEM
    rm, rv = rm0.clone().cuda(), rv0.clone().cuda()
    out, st = run((mk, mode, rm, rv))
    torch.cuda.synchronize()
    assert torch.equal(out.t, plain.t)
    # the path it replaces: the separate statistics pass over the stored output
    rm2, rv2 = rm0.clone().cuda(), rv0.clone().cuda()
    st_ref = eng.bn_stats(plain, mk, mode, rm2, rv2)
    torch.cuda.synchronize()
    torch.testing.assert_close(st, st_ref, rtol=2e-5, atol=2e-6)
    torch.testing.assert_close(rm, rm2, rtol=2e-5, atol=2e-6)
    torch.testing.assert_close(rv, rv2, rtol=2e-5, atol=2e-6)
    # and torch's definition (biased variance for the normalisation, eps 1e-5)
    v = plain.interior().double().reshape(rows, co)
    if mask == 'bc':
        v = (v.view(B, -1, co) * (2.0 * mk.view(B, 1, co).double())).reshape(rows, co)
    elif mask == 'elem':
        v = v * (2.0 * mk.view(rows, co).double())
    mean, var = v.mean(0), v.var(0, unbiased=False)
    torch.testing.assert_close(st[0].double(), mean, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(st[1].double(), 1.0 / torch.sqrt(var + 1e-5), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('B,ci,co,sp', [(3, 64, 128, 10), (2, 64, 256, 16), (5, 128, 640, 6)])
def test_tma_store_epilogue_writes_nothing_outside_its_rows(B, ci, co, sp):
    """Guard-band check of the TMA-store epilogue (compute-sanitizer is closed on this pool): the GEMM writes the interior of
    a BORDERED output carved out of a larger sentinel-filled buffer; ragged tiles (B * OH * OW not a multiple of 128), the
    border pixels and the guard bands before / after the tensor must keep the sentinel."""
    from mopoe_mimic_b200.engine import Act, conv_form
    eng = _eng(torch.bfloat16, 'tc')
    x = _rand((B, ci, sp, sp), 31, 1.0, torch.bfloat16)
    w = _rand((co, ci, 4, 4), 32, 0.05, torch.bfloat16)
    ref = F.conv2d(x, w, None, stride=2, padding=1)
    OH = sp // 2
    guard = 4096
    n = B * (OH + 2) * (OH + 2) * co
    buf = torch.full((guard + n + guard,), 7.0, dtype=torch.bfloat16, device='cuda')
    out = Act(buf[guard:guard + n].view(B, OH + 2, OH + 2, co), B, OH, OH, co, 1, 1)
    eng.gemm_down(_act(x, 1, torch.bfloat16, 2), conv_form(w.cuda(), torch.bfloat16), None, 4, 2, 1, co, out=out)
    torch.cuda.synchronize()
    assert (_to_nchw(out, 2) - ref).abs().max() <= 1.5e-2 * ref.abs().max()
    assert bool((buf[:guard] == 7.0).all()) and bool((buf[guard + n:] == 7.0).all())
    full = out.t.clone()
    full[:, 1:1 + OH, 1:1 + OH, :] = 7.0
    assert bool((full == 7.0).all())          # the zero border of the output tensor was not touched


RES_CASES = [
    # (kind, nd, B, ci, co, sp, mask, bias, next_bn, out_pad)
    ('down', 2, 4, 64, 128, 16, 'bc', False, True, 1),      # conv2 of a 2-D down block, Dropout2d mask, bordered output
    ('down', 2, 3, 64, 192, 20, 'bc', True, True, 0),       # ragged tiles (300 rows), conv bias, 192-column tile
    ('down', 2, 5, 128, 640, 8, 'bc', False, False, 1),     # wide layer, no statistics
    ('down', 1, 3, 64, 256, 200, 'elem', True, True, 1),    # 1-D block: elementwise mask laid out like r
    ('up', 2, 4, 128, 64, 8, 'bc', False, True, 1),         # deconv block: 4 sub-pixel phases in one launch
    ('up', 1, 5, 64, 128, 50, 'elem', True, True, 1),       # 1-D deconv: 2 phases, elementwise mask
    ('up', 2, 3, 64, 128, 6, None, False, True, 0),         # no dropout mask (factor 1)
    # CTA-pair kernel
    ('down', 2, 19, 64, 256, 64, 'bc', False, True, 1),     # 152 m-tiles
    ('down', 2, 31, 64, 256, 40, 'bc', True, True, 1),      # 155 m-tiles: odd -> phantom tile in the last pair
    ('up', 2, 40, 256, 256, 16, 'bc', False, True, 1),      # 4 phases x 80 m-tiles, 16 k-steps
    ('down', 1, 64, 64, 256, 1024, 'elem', False, True, 1),
]


@pytest.mark.parametrize('kind,nd,B,ci,co,sp,mask,bias_on,next_bn,out_pad', RES_CASES)
def test_gemm_with_the_residual_combine_in_its_epilogue(kind, nd, B, ci, co, sp, mask, bias_on, next_bn, out_pad):
    """mopoe_conv_gemm_res against the two launches it replaces (mopoe_conv_gemm_batched -> mopoe_combine_bn) and against
    the definition `a * BN(residual) + b * dropout(conv2)` (ResidualBlocks.py:92-96) in fp64 on the same bf16 operands.
    The fused epilogue combines the fp32 accumulator, the unfused path the bf16-rounded conv result: they agree to bf16
    rounding, the fused one being the closer to the definition."""
    from mopoe_mimic_b200 import _lib as L
    from mopoe_mimic_b200.engine import Act, conv_form, phase_form
    dtype = torch.bfloat16
    eng = _eng(dtype, 'tc')
    shp = (B, ci, sp) if nd == 1 else (B, ci, sp, sp)
    x = _rand(shp, 41, 1.0, dtype)
    bias = _rand((co,), 43, 0.5).cuda() if bias_on else None
    g = torch.Generator().manual_seed(44)
    if kind == 'down':
        w = _rand((co, ci, 4) if nd == 1 else (co, ci, 4, 4), 42, 0.05, dtype)
        wp = conv_form(w.cuda(), dtype)
        conv = lambda **kw: eng.gemm_down(_act(x, 1, dtype, nd), wp, bias, 4, 2, 1, co, **kw)
        ref_c = (F.conv1d if nd == 1 else F.conv2d)(x.double(), w.double(), None, stride=2, padding=1)
    else:
        w = _rand((ci, co, 4) if nd == 1 else (ci, co, 4, 4), 42, 0.05, dtype)
        wp = phase_form(w.cuda(), dtype)
        conv = lambda **kw: eng.gemm_up(_act(x, 1, dtype, nd), wp, bias, co, **kw)
        ref_c = (F.conv_transpose1d if nd == 1 else F.conv_transpose2d)(x.double(), w.double(), None, stride=2, padding=1)
    c = conv()                                                   # the unfused conv2 result (bf16)
    OH, OW = c.H, c.W
    rows = B * OH * OW
    r = Act((torch.randn(B, OH, OW, co, generator=g) * 1.5).to(dtype).cuda(), B, OH, OW, co, 0, 0)
    st3 = torch.stack((torch.randn(co, generator=g) * 0.2, torch.rand(co, generator=g) + 0.5)).cuda()
    gamma, beta = (torch.rand(co, generator=g) + 0.5).cuda(), (torch.randn(co, generator=g) * 0.3).cuda()
    a, b = 2.0, 0.3
    mk, mode = None, L.MASK_NONE
    if mask == 'bc':
        mk, mode = (torch.rand(B * co, generator=g) < 0.5).to(torch.uint8).cuda(), L.MASK_BC
    elif mask == 'elem':
        mk, mode = (torch.rand(rows * co, generator=g) < 0.5).to(torch.uint8).cuda(), L.MASK_ELEM
    rm0, rv0 = torch.randn(co, generator=g), torch.rand(co, generator=g) + 0.5
    ph, pw = (0, out_pad) if nd == 1 else (out_pad, out_pad)

    def sentinel_out():
        guard = 4096
        n = B * (OH + 2 * ph) * (OW + 2 * pw) * co
        buf = torch.full((guard + n + guard,), 7.0, dtype=dtype, device='cuda')
        return buf, guard, n, Act(buf[guard:guard + n].view(B, OH + 2 * ph, OW + 2 * pw, co), B, OH, OW, co, ph, pw)

    # the path it replaces
    y_ref = Act.empty(B, OH, OW, co, ph, pw, dtype, 'cuda')
    rm2, rv2 = rm0.clone().cuda(), rv0.clone().cuda()
    if next_bn:
        y_ref, st_ref = eng.combine(r, st3, gamma, beta, c, mk, mode, a, b, y_ref, bn=(rm2, rv2))
    else:
        eng.combine(r, st3, gamma, beta, c, mk, mode, a, b, y_ref)
    # fused
    buf, guard, n, y = sentinel_out()
    rm, rv = rm0.clone().cuda(), rv0.clone().cuda()
    res = dict(r=r, stats=st3, gamma=gamma, beta=beta, a=a, b=b, mask=mk, mode=mode, next_bn=(rm, rv) if next_bn else None)
    got = conv(out=y, res=res)
    torch.cuda.synchronize()
    assert got is not None, 'the fused residual epilogue must apply to this problem'
    # nothing outside the activation was written (guard bands), and its border is zero (written by the same launch)
    assert bool((buf[:guard] == 7.0).all()) and bool((buf[guard + n:] == 7.0).all())
    full = y.t.clone()
    full[:, ph:ph + OH, pw:pw + OW, :] = 0.0
    assert bool((full == 0.0).all())
    # the definition, fp64 on the same operands
    cc = ref_c.permute(0, 2, 3, 1) if nd == 2 else ref_c.permute(0, 2, 1).reshape(B, 1, OW, co)
    if bias is not None:
        cc = cc + bias.double().cpu()
    if mask == 'bc':
        cc = cc * (2.0 * mk.cpu().view(B, 1, 1, co).double())
    elif mask == 'elem':
        cc = cc * (2.0 * mk.cpu().view(B, OH, OW, co).double())
    mean, invstd = st3[0].double().cpu(), st3[1].double().cpu()
    bn_r = (r.t.double().cpu() - mean) * invstd * gamma.double().cpu() + beta.double().cpu()
    want = a * bn_r + b * cc
    yf, yr = y.interior().double().cpu(), y_ref.interior().double().cpu()
    scale = want.abs().max()
    err_fused, err_unfused = (yf - want).abs().max(), (yr - want).abs().max()
    assert err_fused <= 2.0 ** -8 * scale, (float(err_fused), float(scale))       # one bf16 rounding of the result
    assert (yf - want).abs().mean() <= 1.02 * (yr - want).abs().mean()
    assert err_unfused <= 2.0 ** -7 * scale
    assert (yf - yr).abs().max() <= 2.0 ** -7 * scale
    if next_bn:
        # statistics of the STORED output, as the separate pass computes them
        v = y.interior().double().reshape(rows, co)
        m_, var = v.mean(0), v.var(0, unbiased=False)
        torch.testing.assert_close(got[0].double(), m_, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(got[1].double(), 1.0 / torch.sqrt(var + 1e-5), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(got, st_ref, rtol=2e-2, atol=2e-3)             # (of a slightly different tensor)
        cnt = float(rows)
        torch.testing.assert_close(rm.double(), 0.9 * rm0.double().cuda() + 0.1 * m_, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(rv.double(), 0.9 * rv0.double().cuda() + 0.1 * var * cnt / (cnt - 1), rtol=1e-4, atol=1e-5)


BNB_CASES = [
    # (kind, nd, B, ci, co, sp, mask, accumulate): GEMM = input gradient of a k4 s2 p1 (de)conv, ci gradient channels in, co out
    ('up', 2, 4, 128, 64, 8, 'bc', False),         # dgrad of a 2-D down block's conv2: 4 phases, Dropout2d mask by sample
    ('up', 2, 3, 64, 128, 10, 'bc', True),         # ragged tiles, accumulate into dgamma / dbeta
    ('up', 2, 2, 64, 192, 4, 'bc', False),         # 16-pixel images: a warp's 32 rows span two samples (per-row masks)
    ('down', 2, 4, 64, 128, 16, 'bc', False),      # dgrad of a deconv block's conv2
    ('down', 2, 5, 128, 640, 8, None, True),       # no dropout, wide layer
    ('down', 1, 3, 64, 256, 200, 'elem', False),   # 1-D deconv block: elementwise mask, single problem
    ('up', 1, 5, 64, 128, 50, 'elem', False),      # elementwise mask over 2 phase problems: falls back to the reduction pass
    # CTA-pair kernel
    ('down', 2, 19, 64, 256, 64, 'bc', False),
    ('down', 2, 31, 64, 256, 40, 'bc', True),      # odd m-tile count: phantom tile
    ('up', 2, 40, 256, 256, 16, 'bc', False),
]


@pytest.mark.parametrize('kind,nd,B,ci,co,sp,mask,accumulate', BNB_CASES)
def test_input_gradient_gemm_with_the_batchnorm_backward_sums_in_its_epilogue(kind, nd, B, ci, co, sp, mask, accumulate):
    """mopoe_conv_gemm_bnbwd: same stored output as the plain launch (bit for bit), and sums / dgamma / dbeta equal to the
    mopoe_bn_bwd_reduce pass over that output (the pass it replaces) and to the fp64 definition."""
    from mopoe_mimic_b200 import _lib as L
    from mopoe_mimic_b200.engine import Act, conv_form, phase_form
    dtype = torch.bfloat16
    eng = _eng(dtype, 'tc')
    shp = (B, ci, sp) if nd == 1 else (B, ci, sp, sp)
    dyin = _rand(shp, 51, 1.0, dtype)
    g = torch.Generator().manual_seed(54)
    if kind == 'down':
        w = _rand((co, ci, 4) if nd == 1 else (co, ci, 4, 4), 52, 0.05, dtype)
        wp = conv_form(w.cuda(), dtype)
        gemm = lambda **kw: eng.gemm_down(_act(dyin, 1, dtype, nd), wp, None, 4, 2, 1, co, **kw)
    else:
        w = _rand((ci, co, 4) if nd == 1 else (ci, co, 4, 4), 52, 0.05, dtype)
        wp = phase_form(w.cuda(), dtype)
        gemm = lambda **kw: eng.gemm_up(_act(dyin, 1, dtype, nd), wp, None, co, **kw)
    plain = gemm()
    OH, OW = plain.H, plain.W
    rows = B * OH * OW
    x = Act((torch.randn(B, OH, OW, co, generator=g) * 1.5).to(dtype).cuda(), B, OH, OW, co, 0, 0)
    mk, mode = None, L.MASK_NONE
    if mask == 'bc':
        mk, mode = (torch.rand(B * co, generator=g) < 0.5).to(torch.uint8).cuda(), L.MASK_BC
    elif mask == 'elem':
        mk, mode = (torch.rand(rows * co, generator=g) < 0.5).to(torch.uint8).cuda(), L.MASK_ELEM
    gamma, beta = (torch.rand(co, generator=g) + 0.5).cuda(), (torch.randn(co, generator=g) * 0.3).cuda()
    st = eng.bn_stats(x, mk, mode)
    a2 = eng.bn_apply(x, mk, mode, st, gamma, beta, True, Act.empty(B, OH, OW, co, 0, 0, dtype, 'cuda'))
    dg0, db0 = torch.randn(co, generator=g).cuda(), torch.randn(co, generator=g).cuda()
    # the pass it replaces
    dg_ref, db_ref = dg0.clone(), db0.clone()
    dh_ref = eng.bn_bwd(plain, a2, 1.0, x, mk, mode, st, gamma, dg_ref, db_ref, None, Act.empty(B, OH, OW, co, 0, 0, dtype, 'cuda'),
                        accumulate=accumulate, beta=beta)
    # fused
    dg, db = dg0.clone(), db0.clone()
    out, sums = gemm(bnb=dict(x=x, mask=mk, mode=mode, stats=st, gamma=gamma, beta=beta, dgamma=dg, dbeta=db,
                              accumulate=accumulate))
    torch.cuda.synchronize()
    assert torch.equal(out.t, plain.t)
    if mask == 'elem' and kind == 'up':
        assert sums is None                      # not eligible: the GEMM ran plain, the caller runs the reduction pass
        return
    assert sums is not None, 'the fused epilogue must apply to this problem'
    dh = eng.bn_bwd(out, a2, 1.0, x, mk, mode, st, gamma, dg, db, None, Act.empty(B, OH, OW, co, 0, 0, dtype, 'cuda'),
                    accumulate=accumulate, beta=beta, sums=sums)
    torch.cuda.synchronize()
    # fp64 definition on the stored values
    d = plain.t.double().reshape(rows, co)
    xv = x.t.double().reshape(rows, co)
    if mask == 'bc':
        xv = (xv.view(B, -1, co) * (2.0 * mk.view(B, 1, co).double())).reshape(rows, co)
    elif mask == 'elem':
        xv = xv * (2.0 * mk.view(rows, co).double())
    gate = (a2.t.reshape(rows, co) > 0).double()
    gg = d * gate
    xh = (xv - st[0].double()) * st[1].double()
    s_def = torch.stack((gg.sum(0), (gg * xh).sum(0)))
    scale = torch.stack((gg.abs().sum(0), (gg * xh).abs().sum(0))) + 1e-6
    assert float(((sums.double() - s_def).abs() / scale).max()) <= 1e-5
    base_g = dg0.double() if accumulate else torch.zeros(co, dtype=torch.float64, device='cuda')
    base_b = db0.double() if accumulate else torch.zeros(co, dtype=torch.float64, device='cuda')
    assert float(((dg.double() - base_g - s_def[1]).abs() / (scale[1] + base_g.abs())).max()) <= 1e-5
    assert float(((db.double() - base_b - s_def[0]).abs() / (scale[0] + base_b.abs())).max()) <= 1e-5
    # and against the reduction pass
    assert float(((dg - dg_ref).abs() / (scale[1].float() + dg_ref.abs())).max()) <= 1e-5
    assert float(((db - db_ref).abs() / (scale[0].float() + db_ref.abs())).max()) <= 1e-5
    diff = (dh.t.float() - dh_ref.t.float()).abs().max()
    assert float(diff) <= 2.0 ** -7 * float(dh_ref.t.float().abs().max())
