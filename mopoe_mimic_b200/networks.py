"""Encoders / decoders with the reference's constructor signatures, attribute names and state_dict
layout (networks/ConvNetworksImgMimic.py:20-54, ConvNetworksTextMimic.py:11-68 and the modules they
build), but whose forward/backward run entirely in libmopoe_b200.so.

The nn.Module tree below exists to own parameters under the reference's names (so checkpoints
interchange); the container modules are never "called" layer by layer.
"""
import math
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _lib as L
from .blocks import (BN_EPS, BlockRun, BlockSpec, EmbeddingFn, ImgLastFn, ImgStemFn, LinearFn, ResBlockFn, TextLastFn, token_indices_of,
                     TextStemFn, _pads)
from .engine import Engine

RES_A, RES_B = 2.0, 0.3      # FeatureExtractorImg.py:24, DataGeneratorImg.py:30, char_encoding/*.py:6


def compute_dtype(flags):
    cd = getattr(flags, 'compute_dtype', 'bf16')
    if cd in ('bf16', torch.bfloat16):
        return torch.bfloat16
    if cd in ('fp32', torch.float32):
        return torch.float32
    raise ValueError('flags.compute_dtype must be "bf16" or "fp32", got %r' % (cd,))


class Runtime:
    """Shared per-model execution state: engine (scratch, dtype), injected noise for parity runs, RNG seed."""

    def __init__(self, flags):
        self.flags = flags
        self.engine = None
        self.injected_masks = None    # dict name -> uint8 keep-mask in our layout ([B,C] 2-D, [B,L,C] 1-D)
        self.injected_eps = None      # [B, class_dim] fp32
        self.injected_eps_style = None   # modality -> [B, style_dim] fp32 (factorized representation)
        self.schedule = None          # optional list of (masks, eps[, eps_style]), one per model forward call (poe passes)
        self.seed = None
        self.on_decoders_done = None  # callback fired in backward once every decoder gradient is final (bucketed DP exchange)

    def eng(self, device):
        if self.engine is None:
            self.engine = Engine(device, compute_dtype(self.flags))
        return self.engine

    def mask(self, eng, name, n):
        if self.injected_masks is not None:
            m = self.injected_masks[name]
            assert m.dtype == torch.uint8 and m.numel() == n, (name, m.shape, n)
            return m.contiguous()
        if self.seed is None:
            self.seed = torch.initial_seed()
        return eng.dropout_mask(n, self.seed)


class _Params(nn.Module):
    """Parameter holder standing in for nn.Conv*/nn.ConvTranspose*/nn.Linear (same names, shapes, default init)."""

    def __init__(self, wshape, bias_n=None):
        super().__init__()
        fan_in = wshape[1] * int(math.prod(wshape[2:]))
        bound = 1.0 / math.sqrt(fan_in)
        self.weight = nn.Parameter(torch.empty(wshape).uniform_(-bound, bound))
        if bias_n:
            self.bias = nn.Parameter(torch.empty(bias_n).uniform_(-bound, bound))
        else:
            self.register_parameter('bias', None)


class _BN(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer('running_mean', torch.zeros(c))
        self.register_buffer('running_var', torch.ones(c))
        self.register_buffer('num_batches_tracked', torch.tensor(0, dtype=torch.long))


class ResBlock(nn.Module):
    """Parameter layout of ResidualBlock{1d,2d}{Conv,TransposeConv} (networks/ResidualBlocks.py)."""

    def __init__(self, spec):
        super().__init__()
        self.spec = spec
        nd, cin, cout, k = spec.nd, spec.cin, spec.cout, spec.k
        one, ks = (1,) * nd, (k,) * nd
        w2 = ((cin, cout) if spec.transposed else (cout, cin)) + ks
        conv1 = _Params((cin, cin) + one, cin if spec.inner_bias else None)
        conv2 = _Params(w2, cout if spec.inner_bias else None)
        if nd == 1:      # registration order = reference state_dict order
            self.bn1, self.conv1, self.bn2, self.conv2 = _BN(cin), conv1, _BN(cin), conv2
        else:
            self.conv1, self.bn1, self.bn2, self.conv2 = conv1, _BN(cin), _BN(cin), conv2
        short = nn.Sequential(_Params(w2, cout), _BN(cout))
        setattr(self, 'upsample' if spec.transposed else 'downsample', short)

    def short(self):
        return getattr(self, 'upsample' if self.spec.transposed else 'downsample')

    def bns(self):
        return [self.bn1, self.bn2, self.short()[1]]

    def run(self, rt, eng, x_t, B, H, W, in_pad, out_pad, train, prefix, chain=None, next_blk=None):
        """chain: a dict shared by the consecutive blocks of one network call.  In training the block's combine pass also
        produces the batch statistics of next_blk.bn1 (which reads this block's output) and hands them over through it."""
        sp = self.spec
        masks = None
        if train:
            oh, ow = sp.out_hw(H, W)
            if sp.nd == 2:
                n1, n2 = B * sp.cin, B * sp.cout
            else:
                n1, n2 = B * W * sp.cin, B * ow * sp.cout
            masks = (rt.mask(eng, prefix + '.dropout1', n1), rt.mask(eng, prefix + '.dropout2', n2))
        bufs = {'bn1': (self.bn1.running_mean, self.bn1.running_var),
                'bn2': (self.bn2.running_mean, self.bn2.running_var),
                'short': (self.short()[1].running_mean, self.short()[1].running_var)}
        run = BlockRun(eng, sp, B, H, W, in_pad, out_pad, train, bufs, masks)
        named = dict(self.named_parameters())
        params = [named[n] for n in sp.param_names()]
        run.param_objs = dict(zip(sp.param_names(), params))     # the Parameter objects (their .grad is the flat view)
        if train and chain is not None:
            prev = chain.pop('stats', None)
            if prev is not None and prev[0] == x_t.data_ptr():
                run.in_stats = prev[1]
            if next_blk is not None:
                run.next_bn = (next_blk.bn1.running_mean, next_blk.bn1.running_var)
        y = ResBlockFn.apply(x_t, run, *params)
        if run.out_stats is not None:
            chain['stats'] = (y.data_ptr(), run.out_stats)
        return y


def _chain_pads(specs, last_pad):
    """in_pad of every block (what its own geometry needs) and out_pad (= next block's in_pad)"""
    ins = [s.needs_pad for s in specs]
    outs = ins[1:] + [last_pad]
    return ins, outs


class _Net(nn.Module):
    """Common plumbing: runtime attachment, BN counters."""

    def __init__(self, flags):
        super().__init__()
        self.flags = flags
        object.__setattr__(self, 'rt', Runtime(flags))
        object.__setattr__(self, 'prefix', type(self).__name__)

    def _all_bns(self):
        return [m for m in self.modules() if isinstance(m, _BN)]

    def _bump(self):
        torch._foreach_add_([b.num_batches_tracked for b in self._all_bns()], 1)


# ---- image encoder ------------------------------------------------------------------------------------------------
class FeatureExtractorImg(nn.Module):
    def __init__(self, flags):
        super().__init__()
        d = flags.DIM_img
        self.conv1 = _Params((d, flags.image_channels, 3, 3))
        if flags.image_channels != 1:
            raise NotImplementedError('image_channels must be 1 (MIMIC-CXR); got %d' % flags.image_channels)
        cfg = [(d, 2 * d, 2, 1), (2 * d, 3 * d, 2, 1), (3 * d, 4 * d, 2, 1)]
        if flags.img_size == 64:
            cfg += [(4 * d, 5 * d, 2, 0)]
        elif flags.img_size == 128:
            cfg += [(4 * d, 5 * d, 2, 1), (5 * d, 5 * d, 2, 0)]
        elif flags.img_size == 256:
            cfg += [(4 * d, 5 * d, 4, 1), (5 * d, 5 * d, 2, 0)]
        else:
            raise NotImplementedError('img_size %r' % flags.img_size)
        self.specs = []
        for i, (ci, co, s, p) in enumerate(cfg):
            sp = BlockSpec('resblock_%d' % (i + 1), 2, ci, co, 4, s, p, False, RES_A, RES_B)
            self.specs.append(sp)
            setattr(self, 'resblock_%d' % (i + 1), nn.Sequential(ResBlock(sp)))


class LinearFeatureCompressor(nn.Module):
    def __init__(self, in_channels, out_channels_style, out_channels_content):
        super().__init__()
        if out_channels_style:                 # registered FIRST, as in the reference (state_dict order)
            self.style_mu = _Params((out_channels_style, in_channels), out_channels_style)
            self.style_logvar = _Params((out_channels_style, in_channels), out_channels_style)
        else:
            self.style_mu = None
            self.style_logvar = None
        self.content_mu = _Params((out_channels_content, in_channels), out_channels_content)
        self.content_logvar = _Params((out_channels_content, in_channels), out_channels_content)


class EncoderImg(_Net):
    """EncoderImg(flags, style_dim)(x_img [B,1,px,px]) -> (mu_content, logvar_content)  (ConvNetworksImgMimic.py:20-36)"""

    def __init__(self, flags, style_dim=0):
        super().__init__(flags)
        if getattr(flags, 'feature_extractor_img', 'resnet') != 'resnet':
            raise NotImplementedError('only the resnet feature extractor is built (densenet needs downloaded weights)')
        self.feature_extractor = FeatureExtractorImg(flags)
        self.feature_compressor = LinearFeatureCompressor(5 * flags.DIM_img, style_dim, flags.class_dim)

    def forward(self, x_img):
        L.require_cuda(x_img)
        rt, fe = self.rt, self.feature_extractor
        eng = rt.eng(x_img.device)
        train = self.training
        B, _, H, W = x_img.shape
        specs = fe.specs
        ins, outs = _chain_pads(specs, 0)
        h = ImgStemFn.apply(x_img, fe.conv1.weight, eng, ins[0])
        H, W = H // 2, W // 2
        blks, chain = [getattr(fe, sp.name)[0] for sp in specs], {}
        for i, sp in enumerate(specs):
            h = blks[i].run(rt, eng, h, B, H, W, ins[i], outs[i], train, '%s.feature_extractor.%s.0' % (self.prefix, sp.name),
                            chain, blks[i + 1] if i + 1 < len(blks) else None)
            H, W = sp.out_hw(H, W)
        assert H == 1 and W == 1, 'feature extractor must end at 1x1 (got %dx%d)' % (H, W)
        fc = self.feature_compressor
        mu = LinearFn.apply(h, fc.content_mu.weight, fc.content_mu.bias, eng, B, 0, 2, False)
        lv = LinearFn.apply(h, fc.content_logvar.weight, fc.content_logvar.bias, eng, B, 0, 2, False)
        if train:
            self._bump()
        if fc.style_mu is not None:            # content first, style after (ConvNetworksImgMimic.py:31-33)
            smu = LinearFn.apply(h, fc.style_mu.weight, fc.style_mu.bias, eng, B, 0, 2, False)
            slv = LinearFn.apply(h, fc.style_logvar.weight, fc.style_logvar.bias, eng, B, 0, 2, False)
            return mu, lv, smu, slv
        return mu, lv


# ---- image decoder ------------------------------------------------------------------------------------------------
class DataGeneratorImg(nn.Module):
    def __init__(self, flags):
        super().__init__()
        d = flags.DIM_img
        cfg = [(5 * d, 4 * d, 1, 0), (4 * d, 3 * d, 2, 1), (3 * d, 2 * d, 2, 1), (2 * d, d, 2, 1)]
        if flags.img_size == 128:
            cfg += [(d, d, 2, 1)]
        if flags.img_size == 256:
            cfg += [(d, d, 2, 1), (d, d, 2, 1)]
        self.specs = []
        mods = []
        for i, (ci, co, s, p) in enumerate(cfg):
            sp = BlockSpec(str(i), 2, ci, co, 4, s, p, True, RES_A, RES_B)
            self.specs.append(sp)
            mods.append(nn.Sequential(ResBlock(sp)))
        mods.append(_Params((d, flags.image_channels, 3, 3), flags.image_channels))
        self.generator = nn.Sequential(*mods)


class DecoderImg(_Net):
    """DecoderImg(flags, style_dim)(z_style, z_content) -> (img_hat [B,1,px,px], scale=0.75)  (ConvNetworksImgMimic.py:39-54)"""

    def __init__(self, flags, style_dim=0):
        super().__init__(flags)
        self.feature_generator = _Params((5 * flags.DIM_img, style_dim + flags.class_dim), 5 * flags.DIM_img)
        self.img_generator = DataGeneratorImg(flags)
        self.register_buffer('_scale', torch.tensor(0.75), persistent=False)

    def forward(self, z_style, z_content):
        # factorized representation: z = cat(style, content) (ConvNetworksImgMimic.py:47-48, ConvNetworksTextMimic.py:52-55)
        z = z_content if z_style is None else torch.cat((z_style, z_content), dim=1)
        L.require_cuda(z)
        rt, gen = self.rt, self.img_generator
        eng = rt.eng(z.device)
        train = self.training
        B = z.shape[0]
        fg = self.feature_generator
        h = LinearFn.apply(z, fg.weight, fg.bias, eng, B, None, 2, True)
        specs = gen.specs
        ins, outs = _chain_pads(specs, 0)
        if ins[0]:
            raise AssertionError('first decoder block takes an unbordered 1x1 input')
        H = W = 1
        blks, chain = [gen.generator[i][0] for i in range(len(specs))], {}
        for i, sp in enumerate(specs):
            h = blks[i].run(rt, eng, h, B, H, W, ins[i], outs[i], train,
                            '%s.img_generator.generator.%d.0' % (self.prefix, i), chain, blks[i + 1] if i + 1 < len(blks) else None)
            H, W = sp.out_hw(H, W)
        last = gen.generator[len(specs)]
        img = ImgLastFn.apply(h, last.weight, last.bias, eng, B, H, W)
        if train:
            self._bump()
        return img, self._scale.to(z.device)


# ---- text encoder / decoder (char encoding) ---------------------------------------------------------------------------
class FeatureExtractorText(nn.Module):
    def __init__(self, flags):
        super().__init__()
        d = flags.DIM_text
        self.conv1 = _Params((d, flags.num_features, 4), d)
        chans = [(d, 2 * d), (2 * d, 3 * d), (3 * d, 4 * d), (4 * d, 4 * d), (4 * d, 4 * d), (4 * d, 5 * d),
                 (5 * d, 5 * d), (5 * d, 5 * d)]
        self.specs = []
        for i, (ci, co) in enumerate(chans):
            sp = BlockSpec('resblock_%d' % (i + 1), 1, ci, co, 4, 2, 1 if i < 7 else 0, False, RES_A, RES_B)
            self.specs.append(sp)
            setattr(self, 'resblock_%d' % (i + 1), nn.Sequential(ResBlock(sp)))


class FeatureExtractorTextWord(nn.Module):
    """word_encoding/mmvae_text_enc.py:22-85: Embedding(vocab, DIM, padding_idx=0) -> Conv1d(DIM, DIM, 4, 2, 1) -> the same
    8 residual blocks as the char extractor, of which only the first 6 RUN when len_sequence <= 500."""

    def __init__(self, flags):
        super().__init__()
        d = flags.DIM_text
        self.embedding = _Params((flags.vocab_size, d))
        with torch.no_grad():                                   # nn.Embedding init: N(0, 1), padding row zero
            self.embedding.weight.normal_(0.0, 1.0)
            self.embedding.weight[0].zero_()
        self.conv1 = _Params((d, d, 4), d)
        chans = [(d, 2 * d), (2 * d, 3 * d), (3 * d, 4 * d), (4 * d, 4 * d), (4 * d, 4 * d), (4 * d, 5 * d),
                 (5 * d, 5 * d), (5 * d, 5 * d)]
        self.specs = []
        for i, (ci, co) in enumerate(chans):
            sp = BlockSpec('resblock_%d' % (i + 1), 1, ci, co, 4, 2, 1 if i < 7 else 0, False, RES_A, RES_B)
            self.specs.append(sp)
            setattr(self, 'resblock_%d' % (i + 1), nn.Sequential(ResBlock(sp)))
        self.used = 8 if flags.len_sequence > 500 else 6


class EncoderText(_Net):
    """EncoderText(flags, style_dim)(x_text [B, L, num_features]) -> (mu, logvar)  (ConvNetworksTextMimic.py:11-36)"""

    def __init__(self, flags, style_dim=0):
        super().__init__(flags)
        self.word = flags.text_encoding == 'word'
        self.feature_extractor = FeatureExtractorTextWord(flags) if self.word else FeatureExtractorText(flags)
        self.feature_compressor = LinearFeatureCompressor(5 * flags.DIM_text, style_dim, flags.class_dim)

    def forward(self, x_text):
        L.require_cuda(x_text)
        rt, fe = self.rt, self.feature_extractor
        eng = rt.eng(x_text.device)
        train = self.training
        if self.word:               # token indices [B, L] -> embedded sequence [B, L, DIM] (fp32, differentiable)
            x_text = EmbeddingFn.apply(x_text, fe.embedding.weight, eng)
        B, Lq, _ = x_text.shape
        specs = fe.specs[:fe.used] if self.word else fe.specs
        ins, outs = _chain_pads(specs, 0)
        h = TextStemFn.apply(x_text, fe.conv1.weight, fe.conv1.bias, eng, ins[0], None if self.word else token_indices_of(x_text))
        H, W = 1, Lq // 2
        blks, chain = [getattr(fe, sp.name)[0] for sp in specs], {}
        for i, sp in enumerate(specs):
            h = blks[i].run(rt, eng, h, B, H, W, ins[i], outs[i], train, '%s.feature_extractor.%s.0' % (self.prefix, sp.name),
                            chain, blks[i + 1] if i + 1 < len(blks) else None)
            H, W = sp.out_hw(H, W)
        assert W == 1, 'text feature extractor must end at length 1 (got %d)' % W
        fc = self.feature_compressor
        mu = LinearFn.apply(h, fc.content_mu.weight, fc.content_mu.bias, eng, B, 0, 1, False)
        lv = LinearFn.apply(h, fc.content_logvar.weight, fc.content_logvar.bias, eng, B, 0, 1, False)
        if train:
            self._bump()
        if fc.style_mu is not None:
            smu = LinearFn.apply(h, fc.style_mu.weight, fc.style_mu.bias, eng, B, 0, 1, False)
            slv = LinearFn.apply(h, fc.style_logvar.weight, fc.style_logvar.bias, eng, B, 0, 1, False)
            return mu, lv, smu, slv
        return mu, lv


class DataGeneratorText(nn.Module):
    def __init__(self, flags):
        super().__init__()
        d = flags.DIM_text
        chans = [(5 * d, 5 * d), (5 * d, 5 * d), (5 * d, 5 * d), (5 * d, 4 * d), (4 * d, 4 * d), (4 * d, 3 * d),
                 (3 * d, 2 * d), (2 * d, d)]
        self.specs = []
        for i, (ci, co) in enumerate(chans):
            sp = BlockSpec('resblock_%d' % (i + 1), 1, ci, co, 4, 1 if i == 0 else 2, 0 if i == 0 else 1, True,
                           RES_A, RES_B)
            self.specs.append(sp)
            setattr(self, 'resblock_%d' % (i + 1), nn.Sequential(ResBlock(sp)))
        self.conv2 = _Params((d, flags.num_features, 4), flags.num_features)


class DataGeneratorTextWord(nn.Module):
    """word_encoding/DataGeneratorText.py:29-98: one nn.Sequential `generator` of transposed residual blocks, then
    ConvTranspose1d(DIM, vocab, 4, 2, 1) (len_sequence >= 512) or Conv1d(DIM, vocab, 1) (len_sequence == 128); the
    LogSoftmax module (no parameters) lives in the likelihood kernel."""

    def __init__(self, flags):
        super().__init__()
        d = flags.DIM_text
        chans = [(5 * d, 5 * d), (5 * d, 5 * d), (5 * d, 5 * d), (5 * d, 4 * d), (4 * d, 4 * d)]
        if flags.len_sequence >= 512:
            chans += [(4 * d, 3 * d), (3 * d, 2 * d), (2 * d, d)]
        elif flags.len_sequence == 128:
            chans += [(4 * d, d)]
        else:
            raise NotImplementedError('The output shapes of this network will not work for len_sequence: %d' % flags.len_sequence)
        self.specs, mods = [], []
        for i, (ci, co) in enumerate(chans):
            sp = BlockSpec(str(i), 1, ci, co, 4, 1 if i == 0 else 2, 0 if i == 0 else 1, True, RES_A, RES_B)
            self.specs.append(sp)
            mods.append(nn.Sequential(ResBlock(sp)))
        self.pointwise_last = flags.len_sequence == 128
        if self.pointwise_last and flags.vocab_size % 8 != 0:
            raise NotImplementedError('the pointwise vocabulary head needs vocab_size %% 8 == 0 (got %d): pad the vocabulary'
                                      % flags.vocab_size)
        if self.pointwise_last:
            mods.append(_Params((flags.vocab_size, d, 1), flags.vocab_size))        # nn.Conv1d(DIM, vocab, 1)
        else:
            mods.append(_Params((d, flags.vocab_size, 4), flags.vocab_size))        # nn.ConvTranspose1d(DIM, vocab, 4, 2, 1)
        self.generator = nn.Sequential(*mods)


class DecoderText(_Net):
    """DecoderText(flags, style_dim)(z_style, z_content) -> [text_hat [B, L, num_features]]  (ConvNetworksTextMimic.py:39-68).

    The reference returns log-softmaxed scores; here text_hat holds the PRE-softmax scores wrapped so that the
    likelihood kernel can fuse LogSoftmax with the categorical log-prob (see modalities.CategoricalLikelihood,
    which exposes the normalised `.logits`)."""

    def __init__(self, flags, style_dim=0):
        super().__init__(flags)
        self.word = flags.text_encoding == 'word'
        self.feature_generator = _Params((5 * flags.DIM_text, style_dim + flags.class_dim), 5 * flags.DIM_text)
        self.text_generator = DataGeneratorTextWord(flags) if self.word else DataGeneratorText(flags)

    def forward(self, z_style, z_content):
        # factorized representation: z = cat(style, content) (ConvNetworksImgMimic.py:47-48, ConvNetworksTextMimic.py:52-55)
        z = z_content if z_style is None else torch.cat((z_style, z_content), dim=1)
        L.require_cuda(z)
        # "predict in batches to spare GPU memory" (ConvNetworksTextMimic.py:59-64): more rows than flags.batch_size (the
        # importance-sampled likelihood decodes B*K rows) are decoded in batch_size chunks and concatenated; in train
        # mode the BatchNorm statistics are then per chunk, exactly as in the reference
        nb = int(self.flags.batch_size)
        if z.shape[0] > nb:
            return [torch.cat([self._decode_rows(z[i:i + nb]) for i in range(0, z.shape[0], nb)])]
        return [self._decode_rows(z)]

    def _decode_rows(self, z):
        rt, gen = self.rt, self.text_generator
        eng = rt.eng(z.device)
        train = self.training
        B = z.shape[0]
        fg = self.feature_generator
        h = LinearFn.apply(z, fg.weight, fg.bias, eng, B, None, 1, True)
        specs = gen.specs
        last_pad = 0 if (self.word and gen.pointwise_last) else 1        # a 1x1 conv needs no border on its input
        ins, outs = _chain_pads(specs, last_pad)
        H, W = 1, 1
        if self.word:
            blks = [gen.generator[i][0] for i in range(len(specs))]
            names = ['%s.text_generator.generator.%d.0' % (self.prefix, i) for i in range(len(specs))]
        else:
            blks = [getattr(gen, sp.name)[0] for sp in specs]
            names = ['%s.text_generator.%s.0' % (self.prefix, sp.name) for sp in specs]
        chain = {}
        for i, sp in enumerate(specs):
            h = blks[i].run(rt, eng, h, B, H, W, ins[i], outs[i], train, names[i], chain, blks[i + 1] if i + 1 < len(blks) else None)
            H, W = sp.out_hw(H, W)
        if self.word and gen.pointwise_last:
            last = gen.generator[len(specs)]
            V = last.weight.shape[0]
            scores = LinearFn.apply(h, last.weight.view(V, -1), last.bias, eng, B * W, 0, 1, False).view(B, W, V)
        elif self.word:
            last = gen.generator[len(specs)]
            scores = TextLastFn.apply(h, last.weight, last.bias, eng, B, W, 1)
        else:
            scores = TextLastFn.apply(h, gen.conv2.weight, gen.conv2.bias, eng, B, W, 1)
        if train:
            self._bump()
        return scores
