"""autograd.Function wrappers: each Function runs a whole sub-graph (a residual block, a stem, a head,
a likelihood reduction) through the CUDA library in forward and in backward.  PyTorch's autograd only
chains these coarse nodes; no torch math op is on the hot path.

Reference graph of one block (networks/ResidualBlocks.py:20-33 / 51-65 / 84-97 / 118-131):
    out = conv2(relu(bn2(dropout1(conv1(relu(bn1(x))))))); out = dropout2(out)
    res = BN(shortcut_conv(x));  y = a * res + b * out
"""
import ctypes as C
import math
import os

import torch

from . import _lib as L
from .engine import Act, conv_form, conv_form_grad, full_form, phase_form

BN_EPS, BN_MOMENTUM = 1e-5, 0.1


class BlockSpec:
    """Static description of one residual block."""

    def __init__(self, name, nd, cin, cout, k, stride, pad, transposed, a=2.0, b=0.3):
        self.name, self.nd, self.cin, self.cout = name, nd, cin, cout
        self.k, self.stride, self.pad, self.transposed = k, stride, pad, transposed
        self.a, self.b = a, b
        self.inner_bias = nd == 1          # 1-D convs carry biases, 2-D inner convs do not
        if k == 4 and stride == 2 and pad == 1:
            self.kind = 'S'
        elif not transposed and k == 4 and pad == 0:
            self.kind = 'Z'                # 4 -> 1 valid conv
        elif transposed and k == 4 and stride == 1 and pad == 0:
            self.kind = 'U'                # 1 -> 4 transposed conv
        elif not transposed and k == 4 and stride == 4 and pad == 1:
            self.kind = 'Q'
        else:
            raise NotImplementedError('block geometry k=%d s=%d p=%d T=%s' % (k, stride, pad, transposed))
        self.needs_pad = 1 if self.kind in ('S', 'Q') else 0   # border the block INPUT must carry

    def out_hw(self, H, W):
        k, s, p = self.k, self.stride, self.pad
        f = (lambda n: (n - 1) * s - 2 * p + k) if self.transposed else (lambda n: (n + 2 * p - k) // s + 1)
        return (1 if self.nd == 1 else f(H)), f(W)

    def param_names(self):
        short = 'upsample' if self.transposed else 'downsample'
        n = ['bn1.weight', 'bn1.bias', 'conv1.weight']
        if self.inner_bias:
            n.append('conv1.bias')
        n += ['bn2.weight', 'bn2.bias', 'conv2.weight']
        if self.inner_bias:
            n.append('conv2.bias')
        n += [short + '.0.weight', short + '.0.bias', short + '.1.weight', short + '.1.bias']
        return n


class BlockRun:
    """Per-call context handed to ResBlockFn (not a tensor): engine, geometry, BN buffers, masks."""

    def __init__(self, eng, spec, B, H, W, in_pad, out_pad, train, buffers, masks):
        self.eng, self.spec, self.B, self.H, self.W = eng, spec, B, H, W
        self.in_pad, self.out_pad, self.train = in_pad, out_pad, train
        self.buffers = buffers          # dict: 'bn1'/'bn2'/'short' -> (running_mean, running_var)
        self.masks = masks              # (mask1, mask2) uint8 tensors or None
        # hand-off between consecutive blocks of a chain (training): this block's bn1 statistics as produced by the
        # previous block's combine pass / the buffers of the next block's bn1, whose statistics this block then produces
        self.in_stats = None
        self.next_bn = None
        self.out_stats = None


def _pads(nd, p):
    return (0 if nd == 1 else p), p


def _eval_stats(rm, rv):
    return torch.stack((rm, torch.rsqrt(rv + BN_EPS)))


def _main_fwd(eng, spec, x, Wg, bias, dtype, bn=None, res=None, out=None):
    """conv2 / shortcut forward by geometry kind -> plain Act; bn = (mask, mode, running_mean, running_var): also the
    training-mode BatchNorm statistics of the result -> (Act, stats).  res (with out = the block's output activation): the
    block's residual combine in the GEMM epilogue, see Engine._gemm_res -> None when that does not apply."""
    if res is not None and spec.kind == 'U':
        return None                                                # GEMM columns are (tap, channel): separate combine
    if spec.kind == 'S' and not spec.transposed:
        return eng.gemm_down(x, eng.packed(Wg, "conv"), bias, 4, 2, 1, spec.cout, bn=bn, res=res, out=out)
    if spec.kind == 'S':
        return eng.gemm_up(x, eng.packed(Wg, "phase"), bias, spec.cout, bn=bn, res=res, out=out)
    if spec.kind == 'Z':
        return eng.gemm_down(x, eng.packed(Wg, "conv"), bias, 4, 2, 0, spec.cout, bn=bn, res=res, out=out)
    if spec.kind == 'Q':
        return eng.gemm_down(x, eng.packed(Wg, "conv"), bias, 4, 4, 1, spec.cout, bn=bn, res=res, out=out)
    taps = 4 ** spec.nd if spec.nd == 2 else 4                     # 'U'
    bb = bias.repeat(taps) if bias is not None else None
    oh, ow = spec.out_hw(x.H, x.W)
    out = eng.gemm_rows(x, eng.packed(Wg, "full"), bb, taps * spec.cout, out_shape=(x.B, oh, ow, spec.cout))
    if bn is None:
        return out
    return out, eng.bn_stats(out, bn[0], bn[1], bn[2], bn[3])       # GEMM columns are (tap, channel) here: separate pass


def _main_dgrad(eng, spec, dout, Wg, dtype, H, W, bnb=None):
    """d/d(input) of conv2 / shortcut: dout is a bordered Act -> plain Act [B,H,W,cin].  bnb (Engine._gemm_bnbwd): the
    result feeds a BatchNorm backward whose sums the GEMM epilogue can produce -> (Act, sums or None)"""
    if spec.kind == 'S' and not spec.transposed:
        return eng.gemm_up(dout, eng.packed(Wg, "phase"), None, spec.cin, bnb=bnb)
    if spec.kind == 'S':
        return eng.gemm_down(dout, eng.packed(Wg, "conv"), None, 4, 2, 1, spec.cin, bnb=bnb)
    if spec.kind == 'U':
        return eng.gemm_down(dout, eng.packed(Wg, "conv"), None, 4, 1, 0, spec.cin, bnb=bnb)
    if spec.kind == 'Z':
        taps = 16 if spec.nd == 2 else 4
        out = eng.gemm_rows(dout, eng.packed(Wg, "full"), None, taps * spec.cin, out_shape=(dout.B, H, W, spec.cin))
    elif spec.kind == 'Q':
        out = eng.gemm_unfold(dout, eng.packed(Wg, "full"), spec.cin, H, W, 4, 1)
    else:
        raise NotImplementedError('dgrad for block kind %r' % spec.kind)
    return out if bnb is None else (out, None)


def _main_wgrad(eng, spec, xin, dout, param):
    """weight gradient of conv2 / shortcut: accumulated straight into param.grad when the model's gradients live in
    the flat buffer (returns None for autograd), else returned in the parameter's layout"""
    if spec.transposed:
        s = 1 if spec.kind == 'U' else 2
        args = (dout, 4, s, spec.pad if spec.kind == 'S' else 0, xin)
    else:
        args = (xin, 4, spec.stride, spec.pad, dout)
    if eng.wgrad_down_param(*args, param):
        return None
    return conv_form_grad(eng.wgrad_down(*args), param.shape)


def grad_slot(param):
    """the parameter's view into the flat gradient buffer, when kernels may accumulate into it directly (the autograd
    node then returns None for it: no AccumulateGrad add kernel, no temporary)"""
    g = param.grad if param is not None else None
    return g if (g is not None and g.is_contiguous() and g.dtype == torch.float32) else None


class ResBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_t, run, *params):
        eng, sp = run.eng, run.spec
        dt = x_t.dtype
        P = dict(zip(sp.param_names(), params))
        short = 'upsample' if sp.transposed else 'downsample'
        B, H, W = run.B, run.H, run.W
        iph, ipw = _pads(sp.nd, run.in_pad)
        x = Act.like(x_t, B, H, W, sp.cin, iph, ipw)
        mode = L.MASK_NONE if not run.train else (L.MASK_BC if sp.nd == 2 else L.MASK_ELEM)
        m1, m2 = run.masks if run.train else (None, None)
        bufs = run.buffers
        # bn1 -> relu
        if not run.train:
            st1 = _eval_stats(*bufs['bn1'])
        elif run.in_stats is not None:
            st1 = run.in_stats           # (the previous block's combine pass produced them, running statistics included)
        else:
            st1 = eng.bn_stats(x, None, L.MASK_NONE, *bufs['bn1'])
        a1 = eng.bn_apply(x, None, L.MASK_NONE, st1, P['bn1.weight'], P['bn1.bias'], True,
                          Act.empty(B, H, W, sp.cin, 0, 0, dt, eng.device))
        # conv1 (1x1)
        w1f = eng.packed(P['conv1.weight'], 'matT' if sp.transposed else 'mat')
        # (training: the statistics of dropout1(conv1(.)) for bn2 come out of the GEMM's epilogue)
        if run.train:
            hh, st2 = eng.gemm_rows(a1, w1f, P.get('conv1.bias'), sp.cin, bn=(m1, mode) + tuple(bufs['bn2']))
        else:
            hh, st2 = eng.gemm_rows(a1, w1f, P.get('conv1.bias'), sp.cin), _eval_stats(*bufs['bn2'])
        # dropout1 -> bn2 -> relu  (written with the border conv2 needs)
        pph, ppw = _pads(sp.nd, sp.needs_pad)
        a2 = eng.bn_apply(hh, m1, mode, st2, P['bn2.weight'], P['bn2.bias'], True,
                          Act.empty(B, H, W, sp.cin, pph, ppw, dt, eng.device))
        # shortcut conv, then conv2 with the combine y = a * BN3(r) + b * dropout2(conv2) in its epilogue where the
        # library can (bf16 tcgen05 path), else conv2 -> combine pass
        if run.train:
            r, st3 = _main_fwd(eng, sp, x, P[short + '.0.weight'], P[short + '.0.bias'], dt,
                               bn=(None, L.MASK_NONE) + tuple(bufs['short']))
        else:
            r, st3 = _main_fwd(eng, sp, x, P[short + '.0.weight'], P[short + '.0.bias'], dt), _eval_stats(*bufs['short'])
        oph, opw = _pads(sp.nd, run.out_pad)
        y = Act.empty(B, r.H, r.W, sp.cout, oph, opw, dt, eng.device)
        next_bn = run.next_bn if (run.train and run.next_bn is not None and eng.fuse_next_stats) else None
        fused = None
        if run.train and eng.fuse_res and dt == torch.bfloat16:
            fused = _main_fwd(eng, sp, a2, P['conv2.weight'], P.get('conv2.bias'), dt, out=y,
                              res=dict(r=r, stats=st3, gamma=P[short + '.1.weight'], beta=P[short + '.1.bias'], a=sp.a, b=sp.b,
                                       mask=m2, mode=mode, next_bn=next_bn))
        if fused is not None:
            if next_bn is not None:
                run.out_stats = fused
        else:
            c = _main_fwd(eng, sp, a2, P['conv2.weight'], P.get('conv2.bias'), dt)
            if next_bn is not None:
                y, run.out_stats = eng.combine(r, st3, P[short + '.1.weight'], P[short + '.1.bias'], c, m2, mode, sp.a, sp.b, y,
                                               bn=next_bn)
            else:
                eng.combine(r, st3, P[short + '.1.weight'], P[short + '.1.bias'], c, m2, mode, sp.a, sp.b, y)
        ctx.run = run
        ctx.geo = (r.H, r.W)
        ctx.save_for_backward(x_t, a1.t, hh.t, a2.t, r.t, st1, st2, st3, *params)
        ctx.mask_refs = (m1, m2, mode)
        return y.t

    @staticmethod
    def backward(ctx, dy_t):
        run = ctx.run
        eng, sp = run.eng, run.spec
        if not run.train:
            raise RuntimeError('backward through an eval-mode block is not supported')
        x_t, a1_t, hh_t, a2_t, r_t, st1, st2, st3, *params = ctx.saved_tensors
        P = dict(zip(sp.param_names(), params))
        short = 'upsample' if sp.transposed else 'downsample'
        dt = x_t.dtype
        B, H, W = run.B, run.H, run.W
        OH, OW = ctx.geo
        m1, m2, mode = ctx.mask_refs
        iph, ipw = _pads(sp.nd, run.in_pad)
        oph, opw = _pads(sp.nd, run.out_pad)
        pph, ppw = _pads(sp.nd, sp.needs_pad)
        bph, bpw = _pads(sp.nd, 1)
        x = Act.like(x_t, B, H, W, sp.cin, iph, ipw)
        a1 = Act.like(a1_t, B, H, W, sp.cin)
        hh = Act.like(hh_t, B, H, W, sp.cin)
        a2 = Act.like(a2_t, B, H, W, sp.cin, pph, ppw)
        r = Act.like(r_t, B, OH, OW, sp.cout)
        dy = Act.like(dy_t.contiguous(), B, OH, OW, sp.cout, oph, opw)
        G = {}
        po = run.param_objs

        def bn_slots(prefix):
            """(dgamma, dbeta, accumulate): the flat-gradient views when present, else fresh tensors returned to autograd"""
            pw, pb = po[prefix + '.weight'], po[prefix + '.bias']
            gw, gb = grad_slot(pw), grad_slot(pb)
            if gw is not None and gb is not None:
                G[prefix + '.weight'] = G[prefix + '.bias'] = None
                return gw, gb, True
            G[prefix + '.weight'], G[prefix + '.bias'] = eng.f32(pw.numel()), eng.f32(pb.numel())
            return G[prefix + '.weight'], G[prefix + '.bias'], False

        def bias_grad(name, act):
            slot = grad_slot(po[name])
            if slot is not None:
                eng.colsum(act, slot, accumulate=True)
                return None
            return eng.colsum(act)

        # weight gradients (and the bias column sums) that accumulate straight into the flat gradient run on the side
        # stream of this branch (Engine.wgrad_side): nothing in this backward chain reads them
        flat = all(grad_slot(po[n]) is not None for n in sp.param_names())

        def deferred(fn, *operands):
            side = eng.wgrad_side() if flat else None
            if side is None:
                return fn()
            eng.wgrad_keep(*[o.t for o in operands])
            with torch.cuda.stream(side):
                return fn()

        # y = a*BN3(r) + b*(c*2m2)
        dg, db, acc = bn_slots(short + '.1')
        dr, dc = eng.combine_bwd(dy, sp.a, r, st3, P[short + '.1.weight'], dg, db,
                                 m2, mode, sp.b, Act.empty(B, OH, OW, sp.cout, bph, bpw, dt, eng.device),
                                 Act.empty(B, OH, OW, sp.cout, bph, bpw, dt, eng.device), accumulate=acc)
        # shortcut conv
        Ws_ = P[short + '.0.weight']
        G[short + '.0.weight'] = deferred(lambda: _main_wgrad(eng, sp, x, dr, po[short + '.0.weight']), x, dr)
        # the shortcut bias feeds a train-mode BatchNorm: its gradient sum(dr) is analytically zero (BN backward
        # output sums to zero per channel); the reference only accumulates rounding noise there
        G[short + '.0.bias'] = None if grad_slot(po[short + '.0.bias']) is not None else \
            torch.zeros(sp.cout, dtype=torch.float32, device=eng.device)
        dxs = _main_dgrad(eng, sp, dr, Ws_, dt, H, W)
        # conv2
        W2 = P['conv2.weight']
        G['conv2.weight'] = deferred(lambda: _main_wgrad(eng, sp, a2, dc, po['conv2.weight']), a2, dc)
        if sp.inner_bias:
            G['conv2.bias'] = deferred(lambda: bias_grad('conv2.bias', dc), dc)
        # relu, bn2, dropout1: the two per-channel sums of the bn2 backward come out of the conv2 input-gradient GEMM's epilogue
        # where the library can (bf16 tcgen05 path), else out of a reduction pass over (da2, hh)
        dg, db, acc = bn_slots('bn2')
        da2, sums2 = _main_dgrad(eng, sp, dc, W2, dt, H, W,
                                 bnb=dict(x=hh, mask=m1, mode=mode, stats=st2, gamma=P['bn2.weight'], beta=P['bn2.bias'],
                                          dgamma=dg, dbeta=db, accumulate=acc))
        dh = eng.bn_bwd(da2, a2, 1.0, hh, m1, mode, st2, P['bn2.weight'], dg, db, None,
                        Act.empty(B, H, W, sp.cin, 0, 0, dt, eng.device), accumulate=acc, beta=P['bn2.bias'], sums=sums2)
        # conv1 (1x1): weight [n_out, c_in, 1..] (conv) or [c_in, n_out, 1..] (transposed conv)
        w1p = po['conv1.weight']
        done = deferred(lambda: eng.wgrad_rows_param(dh, a1, w1p) if sp.transposed else eng.wgrad_rows_param(a1, dh, w1p),
                        a1, dh)
        if done:
            G['conv1.weight'] = None
        else:
            g1 = eng.wgrad_rows(a1, dh)                               # [n_out, c_in]
            G['conv1.weight'] = (g1.t() if sp.transposed else g1).reshape(P['conv1.weight'].shape)
        if sp.inner_bias:
            G['conv1.bias'] = deferred(lambda: bias_grad('conv1.bias', dh), dh)
        w1b = eng.packed(P['conv1.weight'], 'mat' if sp.transposed else 'matT')   # [c_in, n_out]
        da1 = eng.gemm_rows(dh, w1b, None, sp.cin)
        # relu, bn1 (+ the shortcut's input gradient)
        dg, db, acc = bn_slots('bn1')
        dx = eng.bn_bwd(da1, a1, 1.0, x, None, L.MASK_NONE, st1, P['bn1.weight'], dg, db, dxs,
                        Act.empty(B, H, W, sp.cin, iph, ipw, dt, eng.device), accumulate=acc, beta=P['bn1.bias'])
        grads = [G[n] for n in sp.param_names()]
        return (dx.t.view_as(x_t), None, *grads)


# ---- stems and heads ------------------------------------------------------------------------------------------
def _tc_taps(eng, act):
    """the tap gradients of the single-channel layers run as tcgen05 weight-gradient GEMMs in bf16 mode"""
    import os
    return (act.dtype == torch.bfloat16 and eng.impl != L.IMPL_SIMT and act.C % 64 == 0
            and os.environ.get('MOPOE_TC_TAPS', '1') != '0')


PATCH_COLS = 64        # 9 taps zero-padded to one 64-element k-block of the tcgen05 GEMMs


def _patches(eng, img, B, SH, SW):
    """3x3 / stride-2 / pad-1 patches of the fp32 image [B, SH, SW] as a bf16 activation [B, SH/2, SW/2, 64] (columns 9.. zero)"""
    pt = torch.empty(B * (SH // 2) * (SW // 2), PATCH_COLS, dtype=torch.bfloat16, device=eng.device)
    L.call('mopoe_im2col3x3s2', L.ptr(img), B, SH, SW, PATCH_COLS, L.ptr(pt), L.stream_ptr())
    return Act(pt, B, SH // 2, SW // 2, PATCH_COLS, 0, 0)


def _tap_grad_tc(eng, patches, act):
    """dW[c, t] = sum_m act[m, c] * patches[m, t]: ONE weight-gradient GEMM with the 128-channel activation as the window
    operand and the first 16 patch columns as the row operand.  Returns [C, 9] fp32."""
    g = eng.wgrad_rows(act, patches, n=16)          # [16, C]
    return g[:9].t().contiguous()


class ImgStemFn(torch.autograd.Function):
    """nn.Conv2d(1, C, 3, 2, 1, bias=False) on the NCHW fp32 image (FeatureExtractorImg.py:29-34, :72).  bf16 mode: the
    3x3 patches become a [M, 64] bf16 operand and the layer — forward and weight gradient — runs on the tensor cores."""

    @staticmethod
    def forward(ctx, x, w, eng, out_pad):
        B, _, H, W = x.shape
        Cc = w.shape[0]
        x = x.contiguous().float()
        y = Act.empty(B, H // 2, W // 2, Cc, out_pad, out_pad, eng.dtype, eng.device)
        if _tc_taps(eng, y):
            pt = _patches(eng, x, B, H, W)
            eng.zero_border(y)
            eng.gemm_rows(pt, eng.packed(w.view(Cc, 9), 'mat', bpad=PATCH_COLS), None, Cc, out=y)
            ctx.save_for_backward(pt.t, w)
            ctx.tc = True
        else:
            L.call('mopoe_conv3x3s2_c1_fwd', L.ptr(x), L.ptr(w), B, H, W, C.byref(y.view()), L.stream_ptr())
            ctx.save_for_backward(x, w)
            ctx.tc = False
        ctx.eng, ctx.out_pad, ctx.geo = eng, out_pad, (B, H, W)
        return y.t

    @staticmethod
    def backward(ctx, dy_t):
        x, w = ctx.saved_tensors
        eng = ctx.eng
        B, H, W = ctx.geo
        Cc = w.shape[0]
        dy = Act.like(dy_t.contiguous(), B, H // 2, W // 2, Cc, ctx.out_pad, ctx.out_pad)
        if ctx.tc:
            pt = Act(x, B, H // 2, W // 2, PATCH_COLS, 0, 0)
            return None, _tap_grad_tc(eng, pt, dy).view_as(w), None, None
        nc = eng.nchunk(B * (H // 2) * (W // 2), Cc)
        ws = eng.ws64(nc * 9 * Cc)
        dw = eng.f32(*w.shape)
        L.call('mopoe_conv3x3s2_c1_wgrad', L.ptr(x), C.byref(dy.view()), B, H, W, L.ptr(dw), 0, L.ptr(ws), nc,
               L.stream_ptr())
        return None, dw, None, None


class ImgLastFn(torch.autograd.Function):
    """nn.ConvTranspose2d(C, 1, 3, 2, 1, output_padding=1) -> fp32 NCHW loc (DataGeneratorImg.py:84-90).  bf16 mode: the
    9 tap products per input pixel are a [M, C] x [C, 16] GEMM (then the 2x2 quads are assembled); the input gradient is
    patches(dout) [M, 64] x [64, C] and the weight gradient a 16-wide wgrad — all three on the tensor cores."""

    @staticmethod
    def forward(ctx, x_t, w, bias, eng, B, H, W):
        Cc = w.shape[0]
        x = Act.like(x_t, B, H, W, Cc)
        out = eng.f32(B, 1, 2 * H, 2 * W)
        if _tc_taps(eng, x):
            taps = eng.gemm_rows(x, eng.packed(w.view(Cc, 9), 'matT', bpad=16), None, 16, out_dtype=torch.float32)
            L.call('mopoe_deconv3x3s2_c1_assemble', L.ptr(taps.t), 16, L.ptr(bias), L.ptr(out), B, H, W, L.stream_ptr())
        else:
            nbytes = L.load().mopoe_deconv3x3s2_c1_fwd_ws(C.byref(x.view()))
            ws = eng.wsf(nbytes)
            L.call('mopoe_deconv3x3s2_c1_fwd', C.byref(x.view()), L.ptr(w), L.ptr(bias), L.ptr(out), L.ptr(ws), nbytes,
                   L.stream_ptr())
        ctx.save_for_backward(x_t, w)
        ctx.eng, ctx.geo = eng, (B, H, W)
        return out

    @staticmethod
    def backward(ctx, dout):
        x_t, w = ctx.saved_tensors
        eng = ctx.eng
        B, H, W = ctx.geo
        Cc = w.shape[0]
        x = Act.like(x_t, B, H, W, Cc)
        nc = eng.nchunk(B * H * W, Cc)
        ws = eng.ws64(nc * (9 * Cc + 1))
        db = eng.f32(1)
        dout = dout.contiguous()
        if _tc_taps(eng, x):
            pt = _patches(eng, dout, B, 2 * H, 2 * W)
            dx = eng.gemm_rows(pt, eng.packed(w.view(Cc, 9), 'mat', bpad=PATCH_COLS), None, Cc)
            dw = _tap_grad_tc(eng, pt, x).view_as(w)
            L.call('mopoe_deconv3x3s2_c1_bwd', C.byref(x.view()), L.ptr(w), L.ptr(dout), None, None, L.ptr(db), 0,
                   L.ptr(ws), nc, L.stream_ptr())                    # (only the bias gradient is left to this call)
        else:
            dx = Act.empty(B, H, W, Cc, 0, 0, x_t.dtype, eng.device)
            dw = eng.f32(*w.shape)
            L.call('mopoe_deconv3x3s2_c1_bwd', C.byref(x.view()), L.ptr(w), L.ptr(dout), C.byref(dx.view()),
                   L.ptr(dw), L.ptr(db), 0, L.ptr(ws), nc, L.stream_ptr())
        return dx.t.view_as(x_t), dw, db, None, None, None, None


def attach_token_indices(onehot, idx):
    """tell the text encoder that the one-hot rows `onehot` [B, L, V] were expanded from the byte indices `idx` [B, L]
    (uint8, same device): its first layer then runs as a gather.  The caller guarantees that they agree."""
    assert idx.dtype == torch.uint8 and idx.is_cuda and tuple(idx.shape) == tuple(onehot.shape[:2])
    onehot._mopoe_token_idx = idx.contiguous()
    return onehot


def token_indices_of(x):
    idx = getattr(x, '_mopoe_token_idx', None)
    if idx is None or not x.is_cuda or idx.device != x.device or tuple(idx.shape) != tuple(x.shape[:2]):
        return None
    return idx


class TextStemFn(torch.autograd.Function):
    """x.transpose(-2,-1) -> nn.Conv1d(71, C, 4, 2, 1) (char_encoding/FeatureExtractorText.py:30-31, :71-72):
    the [B, L, F] input already IS channels-last; it is copied once into a bordered, channel-padded buffer."""

    @staticmethod
    def forward(ctx, x, w, bias, eng, out_pad, idx=None):
        B, Lq, Fq = x.shape
        Cc = w.shape[0]
        Fp = (Fq + 15) // 16 * 16
        # idx: the one-hot rows came from byte indices that are still on the device (the 1-byte-per-token wire format,
        # train.GraphedTrainStep(token_indices=True) / attach_token_indices): the layer is a gather of 4 weight columns
        ctx.gather = (idx is not None and not ctx.needs_input_grad[0] and Lq % 2 == 0 and Cc % 8 == 0 and Fq <= 255
                      and os.environ.get('MOPOE_TEXT_GATHER', '1') != '0')
        if ctx.gather:
            y = Act(torch.zeros((B, 1, Lq // 2 + 2 * out_pad, Cc), dtype=eng.dtype, device=eng.device), B, 1, Lq // 2, Cc,
                    0, out_pad)
            table = eng.packed(w, 'full')                          # [(t * V + v), c] in the activation dtype
            L.call('mopoe_text_stem_gather_fwd', L.ptr(idx), B, Lq, Fq, L.ptr(table), L.dtype_code(eng.dtype), L.ptr(bias),
                   C.byref(y.view()), L.stream_ptr())
            ctx.save_for_backward(idx, w)
            ctx.eng, ctx.geo = eng, (B, Lq, Fq, Fp, out_pad)
            return y.t
        x = x.contiguous().float()
        xin = Act.empty(B, 1, Lq, Fp, 0, 1, eng.dtype, eng.device)
        src = L.View(x.data_ptr(), L.F32, B, 1, Lq, Fq, 0, 0, 0, Lq * Fq, Lq * Fq, Fq)
        eng.convert(src, False, xin)
        y = Act(torch.zeros((B, 1, Lq // 2 + 2 * out_pad, Cc), dtype=eng.dtype, device=eng.device), B, 1, Lq // 2, Cc,
                0, out_pad)
        eng.gemm_down(xin, eng.packed(w, 'conv', bpad=Fp), bias, 4, 2, 1, Cc, out=y)
        ctx.save_for_backward(xin.t, w)
        ctx.eng, ctx.geo = eng, (B, Lq, Fq, Fp, out_pad)
        return y.t

    @staticmethod
    def backward(ctx, dy_t):
        xin_t, w = ctx.saved_tensors
        eng = ctx.eng
        B, Lq, Fq, Fp, out_pad = ctx.geo
        Cc = w.shape[0]
        dy = Act.like(dy_t.contiguous(), B, 1, Lq // 2, Cc, 0, out_pad)
        if ctx.gather:              # the one-hot rows the weight-gradient GEMM reads: built from the byte indices, here
            xin = Act.empty(B, 1, Lq, Fp, 0, 1, eng.dtype, eng.device)
            L.call('mopoe_text_onehot_act', L.ptr(xin_t), B, Lq, Fq, C.byref(xin.view()), L.stream_ptr())
        else:
            xin = Act.like(xin_t, B, 1, Lq, Fp, 0, 1)
        g = eng.wgrad_down(xin, 4, 2, 1, dy)                   # [Cc, 4*Fp]
        dw = g.view(Cc, 4, Fp)[:, :, :Fq].permute(0, 2, 1).contiguous()
        db = eng.colsum(dy)
        dx = None
        if ctx.needs_input_grad[0]:            # word encoding: x is the embedded token sequence
            if Fp != Fq or out_pad < 1:
                raise NotImplementedError('input gradient of the text stem needs a feature count that is a multiple of 16')
            dxa = eng.gemm_up(dy, eng.packed(w, 'phase'), None, Fq, out_dtype=torch.float32)     # [B, 1, Lq, Fq] fp32
            dx = dxa.t.view(B, Lq, Fq)
        return dx, dw, db, None, None, None


class EmbeddingFn(torch.autograd.Function):
    """nn.Embedding(vocab, D, padding_idx=0) on token indices [B, L] (shipped as floats, as the reference does) -> fp32
    [B, L, D]  (word_encoding/mmvae_text_enc.py:27-28,69).  The look-up is the GEMM onehot[B*L, V] x E[V, D] and its
    gradient the matching weight-gradient GEMM — the same tcgen05 / SIMT kernels as every other layer; the padding row
    receives no gradient."""

    @staticmethod
    def forward(ctx, idx_f, E, eng):
        B, Lq = idx_f.shape
        V, D = E.shape
        Vp = (V + 63) // 64 * 64
        rows = B * Lq
        oh = torch.empty(rows, Vp, dtype=eng.dtype, device=eng.device)
        L.call('mopoe_onehot', L.ptr(idx_f.contiguous().float()), rows, V, Vp, L.ptr(oh), L.dtype_code(eng.dtype), None,
               L.stream_ptr())
        xa = Act(oh, rows, 1, 1, Vp, 0, 0)
        # GEMM operand [n = D, k = Vp] = E^T (zero columns for the vocabulary padding); V x D is tiny, torch re-lays it out
        w_op = torch.zeros(D, Vp, dtype=eng.dtype, device=eng.device)
        w_op[:, :V] = E.detach().t().to(eng.dtype)
        impl, eng.impl = eng.impl, L.IMPL_SIMT            # (shapes like N = 16 are outside what the tcgen05 path is tested on)
        try:
            out = eng.gemm_rows(xa, w_op, None, D, out_dtype=torch.float32)
        finally:
            eng.impl = impl
        ctx.save_for_backward(oh, E)
        ctx.eng, ctx.geo = eng, (B, Lq, V, Vp, D)
        return out.t.view(B, Lq, D)

    @staticmethod
    def backward(ctx, d_out):
        oh, E = ctx.saved_tensors
        eng = ctx.eng
        B, Lq, V, Vp, D = ctx.geo
        rows = B * Lq
        dy = Act(d_out.contiguous().to(eng.dtype).view(rows, 1, 1, D), rows, 1, 1, D, 0, 0)
        xa = Act(oh, rows, 1, 1, Vp, 0, 0)
        # dE[v, d] = sum_rows onehot[row, v] * d_out[row, d]:  wgrad with the one-hot rows as "dY" and d_out as the window
        impl, eng.impl = eng.impl, L.IMPL_SIMT
        try:
            g = eng.wgrad_rows(dy, xa)                          # [Vp, D] fp32
        finally:
            eng.impl = impl
        dE = g[:V].clone()
        dE[0].zero_()                                           # padding_idx = 0
        return None, dE, None


class LinearFn(torch.autograd.Function):
    """nn.Linear on [B, K] rows -> [B, N] (FeatureCompressor.py:13-19, ConvNetworks*Mimic.py feature_generator).
    in_act: x_t is an activation (act dtype, possibly bordered 1x1); else x_t is an fp32 [B, K] tensor."""

    @staticmethod
    def forward(ctx, x_t, w, bias, eng, B, in_pad, nd, out_act):
        N, K = w.shape
        if in_pad is None:                 # fp32 latent rows
            xa = Act.like(x_t.contiguous().to(eng.dtype), B, 1, 1, K)
        else:
            xa = Act.like(x_t, B, 1, 1, K, *_pads(nd, in_pad))
        out = eng.gemm_rows(xa, eng.packed(w, 'mat'), bias, N, out_dtype=eng.dtype if out_act else torch.float32)
        ctx.save_for_backward(xa.t, w)
        ctx.eng, ctx.geo = eng, (B, in_pad, nd, out_act, x_t.dtype, tuple(x_t.shape))
        return out.t.view(B, N) if not out_act else out.t

    @staticmethod
    def backward(ctx, dy_t):
        xa_t, w = ctx.saved_tensors
        eng = ctx.eng
        B, in_pad, nd, out_act, x_dtype, x_shape = ctx.geo
        N, K = w.shape
        ph, pw = (0, 0) if in_pad is None else _pads(nd, in_pad)
        xa = Act.like(xa_t, B, 1, 1, K, ph, pw)
        dy = Act.like(dy_t.contiguous().to(eng.dtype), B, 1, 1, N)
        dw = eng.wgrad_rows(xa, dy)
        db = eng.colsum(dy)
        dxp = eng.gemm_rows(dy, eng.packed(w, 'matT'), None, K,
                            out_dtype=torch.float32 if in_pad is None else eng.dtype)
        if in_pad is None:
            dx = dxp.t.view(x_shape).to(x_dtype)
        elif ph == 0 and pw == 0:
            dx = dxp.t.view(x_shape)
        else:
            full = Act(torch.zeros(x_shape, dtype=eng.dtype, device=eng.device), B, 1, 1, K, ph, pw)
            full.interior().copy_(dxp.t.view(B, 1, 1, K))
            dx = full.t
        return dx, dw, db, None, None, None, None, None


class TextLastFn(torch.autograd.Function):
    """nn.ConvTranspose1d(C, 71, 4, 2, 1) -> fp32 pre-softmax scores [B, L, 71]
    (char_encoding/DataGeneratorText.py:44-49, :73-74); the LogSoftmax lives in the likelihood kernel."""

    @staticmethod
    def forward(ctx, x_t, w, bias, eng, B, Lq, in_pad):
        Cc, Fq = w.shape[0], w.shape[1]
        x = Act.like(x_t, B, 1, Lq, Cc, 0, in_pad)
        out = eng.gemm_up(x, eng.packed(w, 'phase'), bias, Fq, out_dtype=torch.float32)
        ctx.save_for_backward(x_t, w)
        ctx.eng, ctx.geo = eng, (B, Lq, in_pad)
        return out.t.view(B, 2 * Lq, Fq)

    @staticmethod
    def backward(ctx, dy_t):
        x_t, w = ctx.saved_tensors
        eng = ctx.eng
        B, Lq, in_pad = ctx.geo
        Cc, Fq = w.shape[0], w.shape[1]
        Fp = (Fq + 15) // 16 * 16
        x = Act.like(x_t, B, 1, Lq, Cc, 0, in_pad)
        # bordered, channel-padded copy of the fp32 gradient in the activation dtype
        dyp = Act.empty(B, 1, 2 * Lq, Fp, 0, 1, eng.dtype, eng.device)
        dyc = dy_t.contiguous()
        src = L.View(dyc.data_ptr(), L.F32, B, 1, 2 * Lq, Fq, 0, 0, 0, 2 * Lq * Fq, 2 * Lq * Fq, Fq)
        eng.convert(src, False, dyp)
        db = eng.colsum(dyp)[:Fq]
        # dgrad of a stride-2 deconv = stride-2 conv over the bordered gradient, conv-form weights [ci,(kx,co)]
        dx = eng.gemm_down(dyp, eng.packed(w, 'conv', bpad=Fp), None, 4, 2, 1, Cc)
        g = eng.wgrad_down(dyp, 4, 2, 1, x)                     # [Cc, 4*Fp]
        dw = g.view(Cc, 4, Fp)[:, :, :Fq].permute(0, 2, 1).contiguous()
        if in_pad:
            full = Act(torch.zeros_like(x_t), B, 1, Lq, Cc, 0, in_pad)
            full.interior().copy_(dx.t)
            dxt = full.t
        else:
            dxt = dx.t.view_as(x_t)
        return dxt, dw, db, None, None, None, None


# ---- likelihood reductions ------------------------------------------------------------------------------------
class LaplaceLogProbSumFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, loc, target, scale, eng):
        loc = loc.contiguous()
        target = target.contiguous().float()
        n = loc.numel()
        nc = int(min(148 * 8, max(1, n // 4096)))
        out = eng.f32(1)
        L.annotate(kind='laplace_nll', bytes=8 * n)            # loc + target, fp32
        L.call('mopoe_laplace_logprob_sum', L.ptr(loc), L.ptr(target), n, float(scale), L.ptr(out),
               L.ptr(eng.ws64(nc)), nc, L.stream_ptr())
        ctx.save_for_backward(loc, target)
        ctx.eng, ctx.scale = eng, float(scale)
        return out.view(())

    @staticmethod
    def backward(ctx, g):
        loc, target = ctx.saved_tensors
        dloc = torch.empty_like(loc)
        # SURVEY.md §8d counts forward + gradient as 3 passes (loc, target, dloc); the backward re-reads loc and target
        L.annotate(kind='laplace_nll', bytes=4 * loc.numel())
        L.call('mopoe_laplace_logprob_bwd', L.ptr(loc), L.ptr(target), loc.numel(), ctx.scale,
               L.ptr(g.contiguous().float()), L.ptr(dloc), L.stream_ptr())
        return dloc, None, None, None


class CategoricalLogProbSumFn(torch.autograd.Function):
    """scores: pre-softmax [B, L, V]; target: one-hot [B, L, V], or token indices [B, L] (word encoding).

    Precondition (as in the reference's data): target rows are strictly one-hot.  The kernel gathers at argmax(target);
    OneHotCategorical.log_prob computes sum(target * logits), which differs for all-zero or soft rows."""

    @staticmethod
    def forward(ctx, scores, target, eng):
        scores = scores.contiguous()
        V = scores.shape[-1]
        rows = scores.numel() // V
        nc = int(min(148 * 8, max(1, rows // 64)))
        out = eng.f32(1)
        lib = L.load()
        if target.dim() == scores.dim() - 1:            # indices (MimicText.calc_log_prob one-hot encodes them, :37-40)
            idx = target.contiguous().reshape(-1).to(torch.int32)
            tgt, idx_in, idx_out = None, idx, None
        else:
            tgt = target.contiguous().float()
            idx = torch.empty(rows, dtype=torch.int32, device=scores.device)
            idx_in, idx_out = None, idx
        lse = eng.f32(rows) if lib.mopoe_categorical_has_lse(L.ptr(scores), L.ptr(tgt), V) else None
        L.annotate(kind='categorical_nll', bytes=4 * rows * (V * (2 if tgt is not None else 1) + (1 if tgt is None else 0)))
        L.call('mopoe_categorical_logprob_sum', L.ptr(scores), L.ptr(tgt), L.ptr(idx_in), rows, V, None, L.ptr(idx_out),
               L.ptr(lse), L.ptr(out), L.ptr(eng.ws64(nc)), nc, L.stream_ptr())
        ctx.has_lse = lse is not None
        ctx.save_for_backward(scores, idx, *([lse] if lse is not None else []))
        return out.view(())

    @staticmethod
    def backward(ctx, g):
        scores, idx = ctx.saved_tensors[:2]
        lse = ctx.saved_tensors[2] if ctx.has_lse else None
        V = scores.shape[-1]
        d = torch.empty_like(scores)
        L.annotate(kind='categorical_nll', bytes=4 * scores.numel())      # the gradient write (3 passes in all)
        L.call('mopoe_categorical_logprob_bwd', L.ptr(scores), L.ptr(idx), L.ptr(lse), scores.numel() // V, V,
               L.ptr(g.contiguous().float()), L.ptr(d), L.stream_ptr())
        return d, None, None


def laplace_log_prob(loc, value, scale, eng):
    """elementwise Laplace(loc, scale).log_prob(value) (no grad; the evaluation callers' `.log_prob`)"""
    loc = loc.detach().contiguous().float()
    value = value.detach().float().expand_as(loc).contiguous()
    L.require_cuda(loc, value)
    out = torch.empty_like(loc)
    L.call('mopoe_laplace_logprob_elem', L.ptr(loc), L.ptr(value), loc.numel(), float(scale), L.ptr(out), L.stream_ptr())
    return out


def log_softmax_rows(scores, eng):
    """log_softmax over the last dim through the categorical kernel (no grad; API/`.logits` use)."""
    scores = scores.detach().contiguous()
    V = scores.shape[-1]
    rows = scores.numel() // V
    out = torch.empty_like(scores)
    idx = torch.zeros(rows, dtype=torch.int32, device=scores.device)
    dummy = eng.f32(1)
    nc = int(min(148 * 8, max(1, rows // 64)))
    L.call('mopoe_categorical_logprob_sum', L.ptr(scores), None, L.ptr(idx), rows, V, L.ptr(out), None, None, L.ptr(dummy),
           L.ptr(eng.ws64(nc)), nc, L.stream_ptr())
    return out
