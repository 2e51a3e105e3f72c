"""CPU oracle for the MoPoE-MIMIC training step.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under
``mopoe_mimic_b200/`` imports it, and the product raises if its CUDA library is missing.

It is a *functional* restatement (a flat ``state`` dict of tensors + plain functions on
``torch.nn.functional``) of the reference's module tree, so that weights, dropout masks and
the reparameterisation noise can be injected explicitly.  All numerics live in PyTorch CPU ops
(the reference's own third-party dependency, ``torch~=1.6`` in requirements.txt:3; torch 2.11 here).

Parity pin: the reference's own tests hold NO golden vectors for this path.  The
oracle is pinned instead against the reference itself, run live in the build container by
``oracle/gen_golden.py`` (imports /root/reference from a writable copy, loads the same state,
patches dropout / reparameterize to the injected masks / eps) — outputs are committed under
``tests/golden/`` and checked by ``tests/test_oracle_golden.py``.

Reference citations are relative to /root/reference/mimic/.
"""
from __future__ import annotations

import hashlib
import math
from collections import OrderedDict
from itertools import chain, combinations
from types import SimpleNamespace

import torch
import torch.nn.functional as F

BN_EPS = 1e-5          # torch BatchNorm default (no override in networks/ResidualBlocks.py)
BN_MOMENTUM = 0.1
RES_A, RES_B = 2.0, 0.3  # FeatureExtractorImg.py:24, DataGeneratorImg.py:30, char_encoding/*:6
LAPLACE_SCALE = 0.75   # networks/ConvNetworksImgMimic.py:54
POE_EPS = 1e-8         # evaluation/divergence_measures/mm_div.py:10


def default_flags(**kw):
    """Flag fields the hot path reads; defaults = utils/flags.py, BaseFlags.py."""
    f = dict(batch_size=16, class_dim=128, img_size=128, image_channels=1, DIM_img=128, DIM_text=128,
             text_encoding='char', len_sequence=1024, num_features=71, vocab_size=0,
             method='joint_elbo', mods=('PA', 'Lateral', 'text'),
             beta=5.0, beta_style=1.0, beta_content=1.0,
             rec_weights={'PA': 0.33, 'Lateral': 0.33, 'text': 0.33},
             alpha_modalities=[0.25, 0.25, 0.25, 0.25],
             # factorized representation (flags.py: factorized_representation, style_{pa,lat,text}_dim): modality-specific
             # style latents next to the shared content latent; style_weights = experiment.py's beta_m*_style (1.0)
             style_dims={'PA': 0, 'Lateral': 0, 'text': 0}, style_weights={'PA': 1.0, 'Lateral': 1.0, 'text': 1.0})
    f.update(kw)
    return SimpleNamespace(**f)


def style_dim(flags, m):
    return int(getattr(flags, 'style_dims', {}).get(m, 0))


def factorized(flags):
    return any(style_dim(flags, m) > 0 for m in flags.mods)


# --------------------------------------------------------------------------------------
# architecture tables (what the reference's constructors build)
# --------------------------------------------------------------------------------------
def img_encoder_blocks(flags):
    """FeatureExtractorImg.py:29-59 -> list of (cin, cout, k, stride, pad)."""
    d = flags.DIM_img
    blocks = [(d, 2 * d, 4, 2, 1), (2 * d, 3 * d, 4, 2, 1), (3 * d, 4 * d, 4, 2, 1)]
    if flags.img_size == 64:
        blocks += [(4 * d, 5 * d, 4, 2, 0)]
    elif flags.img_size == 128:
        blocks += [(4 * d, 5 * d, 4, 2, 1), (5 * d, 5 * d, 4, 2, 0)]
    elif flags.img_size == 256:
        blocks += [(4 * d, 5 * d, 4, 4, 1), (5 * d, 5 * d, 4, 2, 0)]
    else:
        raise NotImplementedError(flags.img_size)
    return blocks


def img_decoder_blocks(flags):
    """DataGeneratorImg.py:33-79 -> list of (cin, cout, k, stride, pad)."""
    d = flags.DIM_img
    blocks = [(5 * d, 4 * d, 4, 1, 0), (4 * d, 3 * d, 4, 2, 1), (3 * d, 2 * d, 4, 2, 1), (2 * d, d, 4, 2, 1)]
    if flags.img_size == 128:
        blocks += [(d, d, 4, 2, 1)]
    if flags.img_size == 256:
        blocks += [(d, d, 4, 2, 1), (d, d, 4, 2, 1)]
    return blocks


def text_encoder_blocks(flags):
    """char_encoding/FeatureExtractorText.py:32-56."""
    d = flags.DIM_text
    chans = [(d, 2 * d), (2 * d, 3 * d), (3 * d, 4 * d), (4 * d, 4 * d), (4 * d, 4 * d), (4 * d, 5 * d),
             (5 * d, 5 * d), (5 * d, 5 * d)]
    return [(ci, co, 4, 2, 1 if i < 7 else 0) for i, (ci, co) in enumerate(chans)]


def is_word(flags):
    return getattr(flags, 'text_encoding', 'char') == 'word'


def text_encoder_blocks_used(flags):
    """word encoding (word_encoding/mmvae_text_enc.py:76-79): resblock_7/8 are constructed but only RUN when
    len_sequence > 500; the char encoder always runs all 8"""
    blocks = text_encoder_blocks(flags)
    return blocks[:6] if (is_word(flags) and flags.len_sequence <= 500) else blocks


def text_decoder_blocks(flags):
    """char_encoding/DataGeneratorText.py:28-43."""
    d = flags.DIM_text
    chans = [(5 * d, 5 * d), (5 * d, 5 * d), (5 * d, 5 * d), (5 * d, 4 * d), (4 * d, 4 * d), (4 * d, 3 * d),
             (3 * d, 2 * d), (2 * d, d)]
    if is_word(flags) and flags.len_sequence == 128:     # word_encoding/DataGeneratorText.py:58-65
        chans = chans[:5] + [(4 * d, d)]
    elif is_word(flags) and flags.len_sequence < 512:
        raise NotImplementedError('word decoder: len_sequence %d' % flags.len_sequence)
    return [(ci, co, 4, 1 if i == 0 else 2, 0 if i == 0 else 1) for i, (ci, co) in enumerate(chans)]


ENC_NAME = {'PA': 'encoder_pa', 'Lateral': 'encoder_lat', 'text': 'encoder_text'}
DEC_NAME = {'PA': 'decoder_pa', 'Lateral': 'decoder_lat', 'text': 'decoder_text'}


def text_dec_block_name(flags, i):
    return ('.text_generator.generator.%d.0' % i) if is_word(flags) else ('.text_generator.resblock_%d.0' % (i + 1))


def _bn_spec(spec, p, c):
    spec[p + '.weight'] = (c,)
    spec[p + '.bias'] = (c,)
    spec[p + '.running_mean'] = (c,)
    spec[p + '.running_var'] = (c,)
    spec[p + '.num_batches_tracked'] = ()


def _block_spec(spec, p, cin, cout, k, nd, transposed, inner_bias, short):
    """ResidualBlocks.py:5-131 + the make_res_block_* helpers: registration order of the modules."""
    ks = (k,) * nd
    one = (1,) * nd
    w2 = (cin, cout) + ks if transposed else (cout, cin) + ks
    if nd == 1:   # 1-D blocks register bn1, conv1, bn2, conv2, shortcut (ResidualBlocks.py:8-16)
        _bn_spec(spec, p + '.bn1', cin)
        spec[p + '.conv1.weight'] = (cin, cin) + one
        spec[p + '.conv1.bias'] = (cin,)
        _bn_spec(spec, p + '.bn2', cin)
        spec[p + '.conv2.weight'] = w2
        spec[p + '.conv2.bias'] = (cout,)
    else:         # 2-D blocks register conv1, bn1, bn2, conv2 (ResidualBlocks.py:71-80), no bias
        spec[p + '.conv1.weight'] = (cin, cin) + one
        _bn_spec(spec, p + '.bn1', cin)
        _bn_spec(spec, p + '.bn2', cin)
        spec[p + '.conv2.weight'] = w2
    spec[p + '.%s.0.weight' % short] = w2
    spec[p + '.%s.0.bias' % short] = (cout,)
    _bn_spec(spec, p + '.%s.1' % short, cout)


def param_spec(flags):
    """Ordered name -> shape for every state_dict entry (params + BN buffers) of the model."""
    spec = OrderedDict()
    D = flags.class_dim
    for m in flags.mods:
        e = ENC_NAME[m]
        if m == 'text':
            d = flags.DIM_text
            if is_word(flags):       # word_encoding/mmvae_text_enc.py:27-30: Embedding(vocab, DIM, padding_idx=0), Conv1d(DIM, DIM)
                spec[e + '.feature_extractor.embedding.weight'] = (flags.vocab_size, d)
                spec[e + '.feature_extractor.conv1.weight'] = (d, d, 4)
            else:
                spec[e + '.feature_extractor.conv1.weight'] = (d, flags.num_features, 4)
            spec[e + '.feature_extractor.conv1.bias'] = (d,)
            for i, (ci, co, k, s, p) in enumerate(text_encoder_blocks(flags)):
                _block_spec(spec, e + '.feature_extractor.resblock_%d.0' % (i + 1), ci, co, k, 1, False, True,
                            'downsample')
        else:
            d = flags.DIM_img
            spec[e + '.feature_extractor.conv1.weight'] = (d, flags.image_channels, 3, 3)
            for i, (ci, co, k, s, p) in enumerate(img_encoder_blocks(flags)):
                _block_spec(spec, e + '.feature_extractor.resblock_%d.0' % (i + 1), ci, co, k, 2, False, False,
                            'downsample')
        sd = style_dim(flags, m)
        if sd:                                    # FeatureCompressor.py:11-16: style heads are registered first
            for head in ('style_mu', 'style_logvar'):
                spec[e + '.feature_compressor.%s.weight' % head] = (sd, 5 * d)
                spec[e + '.feature_compressor.%s.bias' % head] = (sd,)
        for head in ('content_mu', 'content_logvar'):
            spec[e + '.feature_compressor.%s.weight' % head] = (D, 5 * d)
            spec[e + '.feature_compressor.%s.bias' % head] = (D,)
    for m in flags.mods:
        dn = DEC_NAME[m]
        if m == 'text':
            d = flags.DIM_text
            spec[dn + '.feature_generator.weight'] = (5 * d, style_dim(flags, m) + D)
            spec[dn + '.feature_generator.bias'] = (5 * d,)
            blocks = text_decoder_blocks(flags)
            for i, (ci, co, k, s, p) in enumerate(blocks):
                _block_spec(spec, dn + text_dec_block_name(flags, i), ci, co, k, 1, True, True, 'upsample')
            if is_word(flags):       # word_encoding/DataGeneratorText.py:49-65: the modules live in one nn.Sequential
                last = dn + '.text_generator.generator.%d' % len(blocks)
                if flags.len_sequence == 128:
                    spec[last + '.weight'] = (flags.vocab_size, d, 1)            # nn.Conv1d(DIM, vocab, 1)
                else:
                    spec[last + '.weight'] = (d, flags.vocab_size, 4)            # nn.ConvTranspose1d(DIM, vocab, 4, 2, 1)
                spec[last + '.bias'] = (flags.vocab_size,)
            else:
                spec[dn + '.text_generator.conv2.weight'] = (d, flags.num_features, 4)
                spec[dn + '.text_generator.conv2.bias'] = (flags.num_features,)
        else:
            d = flags.DIM_img
            spec[dn + '.feature_generator.weight'] = (5 * d, style_dim(flags, m) + D)
            spec[dn + '.feature_generator.bias'] = (5 * d,)
            blocks = img_decoder_blocks(flags)
            for i, (ci, co, k, s, p) in enumerate(blocks):
                _block_spec(spec, dn + '.img_generator.generator.%d.0' % i, ci, co, k, 2, True, False, 'upsample')
            spec[dn + '.img_generator.generator.%d.weight' % len(blocks)] = (d, flags.image_channels, 3, 3)
            spec[dn + '.img_generator.generator.%d.bias' % len(blocks)] = (flags.image_channels,)
    return spec


def _seed_for(name, seed):
    h = hashlib.sha256(('%d:%s' % (seed, name)).encode()).digest()
    return int.from_bytes(h[:6], 'little')


def seeded_uniform(name, seed, shape, lo=0.0, hi=1.0, dtype=torch.float64):
    """Deterministic per-name tensor, independent of creation order (CPU mt19937, fp64 draw)."""
    g = torch.Generator(device='cpu')
    g.manual_seed(_seed_for(name, seed))
    return (torch.rand(tuple(shape), generator=g, dtype=torch.float64) * (hi - lo) + lo).to(dtype)


def make_state(flags, seed=0, dtype=torch.float32):
    """Deterministic synthetic weights with torch's default-init *scales* (U(+-1/sqrt(fan_in)));
    BN affine is perturbed away from (1, 0) so gamma/beta gradients are exercised."""
    st = OrderedDict()
    for name, shape in param_spec(flags).items():
        if name.endswith('num_batches_tracked'):
            st[name] = torch.zeros((), dtype=torch.int64)
        elif name.endswith('running_mean'):
            st[name] = torch.zeros(shape, dtype=dtype)
        elif name.endswith('running_var'):
            st[name] = torch.ones(shape, dtype=dtype)
        elif name.endswith('embedding.weight'):
            # nn.Embedding default init is N(0, 1) with the padding_idx row zeroed; a seeded uniform of the same scale
            w = seeded_uniform(name, seed, shape, -1.7, 1.7, dtype)
            w[0].zero_()
            st[name] = w
        elif len(shape) == 1 and ('.bn' in name or 'sample.1.' in name):
            if name.endswith('weight'):
                st[name] = seeded_uniform(name, seed, shape, 0.8, 1.2, dtype)
            else:
                st[name] = seeded_uniform(name, seed, shape, -0.1, 0.1, dtype)
        else:
            # torch default init: U(+-1/sqrt(fan_in)), fan_in = weight.size(1) * receptive field
            # (the same rule for Conv, ConvTranspose and Linear); a bias uses its sibling weight's fan_in
            wshape = param_spec_cached(flags)[name[:-4] + 'weight'] if len(shape) == 1 else shape
            fan_in = wshape[1] * int(math.prod(wshape[2:]))
            b = 1.0 / math.sqrt(fan_in)
            st[name] = seeded_uniform(name, seed, shape, -b, b, dtype)
    return st


_spec_cache = {}


def param_spec_cached(flags):
    key = (flags.class_dim, flags.img_size, flags.DIM_img, flags.DIM_text, tuple(flags.mods), flags.num_features,
           tuple(style_dim(flags, m) for m in flags.mods), getattr(flags, 'text_encoding', 'char'),
           getattr(flags, 'vocab_size', 0), flags.len_sequence)
    if key not in _spec_cache:
        _spec_cache[key] = param_spec(flags)
    return _spec_cache[key]


def make_batch(flags, seed=1, dtype=torch.float32, batch=None):
    """Synthetic inputs of the reference's own shapes (dataio/MimicDataset.py:414-428), with TRUE
    one-hot text (torch>=2 validates OneHotCategorical targets)."""
    B = batch or flags.batch_size
    out = OrderedDict()
    for m in flags.mods:
        if m == 'text':
            nf = flags.vocab_size if is_word(flags) else flags.num_features
            idx = (seeded_uniform('text', seed, (B, flags.len_sequence)) * nf).long()
            idx.clamp_(0, nf - 1)
            # word encoding ships token indices [B, L] (0 = padding), char encoding one-hot rows [B, L, 71]
            out[m] = idx.to(dtype) if is_word(flags) else F.one_hot(idx, flags.num_features).to(dtype)
        else:
            out[m] = seeded_uniform(m, seed, (B, flags.image_channels, flags.img_size, flags.img_size), dtype=dtype)
    return out


def dropout_sites(flags, batch=None):
    """name -> mask shape for every Dropout/Dropout2d call of one train-mode forward, in the order
    the reference consumes RNG.  2-D: [B,C,1,1] (Dropout2d), 1-D: [B,C,L]."""
    B = batch or flags.batch_size
    sites = OrderedDict()

    def enc(m):
        e = ENC_NAME[m] + '.feature_extractor.resblock_%d.0'
        if m == 'text':
            L = flags.len_sequence // 2
            for i, (ci, co, k, s, p) in enumerate(text_encoder_blocks_used(flags)):
                sites[(e % (i + 1)) + '.dropout1'] = (B, ci, L)
                L = (L + 2 * p - k) // s + 1
                sites[(e % (i + 1)) + '.dropout2'] = (B, co, L)
        else:
            for i, (ci, co, k, s, p) in enumerate(img_encoder_blocks(flags)):
                sites[(e % (i + 1)) + '.dropout1'] = (B, ci, 1, 1)
                sites[(e % (i + 1)) + '.dropout2'] = (B, co, 1, 1)

    def dec(m):
        if m == 'text':
            L = 1
            for i, (ci, co, k, s, p) in enumerate(text_decoder_blocks(flags)):
                dn = DEC_NAME[m] + text_dec_block_name(flags, i)
                sites[dn + '.dropout1'] = (B, ci, L)
                L = (L - 1) * s - 2 * p + k
                sites[dn + '.dropout2'] = (B, co, L)
        else:
            dn = DEC_NAME[m] + '.img_generator.generator.%d.0'
            for i, (ci, co, k, s, p) in enumerate(img_decoder_blocks(flags)):
                sites[(dn % i) + '.dropout1'] = (B, ci, 1, 1)
                sites[(dn % i) + '.dropout2'] = (B, co, 1, 1)

    for m in flags.mods:
        enc(m)
    for m in flags.mods:
        dec(m)
    return sites


def make_noise(flags, seed=2, dtype=torch.float32, batch=None):
    """Injected randomness: Bernoulli(0.5) keep-masks (values {0,1}) for every dropout site and the
    reparameterisation eps ~ N(0,1) [B, class_dim] (Box-Muller on the seeded uniform stream)."""
    B = batch or flags.batch_size
    masks = OrderedDict()
    for name, shape in dropout_sites(flags, B).items():
        masks[name] = (seeded_uniform(name, seed, shape) < 0.5).to(dtype)
    u1 = seeded_uniform('eps.u1', seed, (B, flags.class_dim)).clamp_min(1e-12)
    u2 = seeded_uniform('eps.u2', seed, (B, flags.class_dim))
    eps = (torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2.0 * math.pi * u2)).to(dtype)
    return masks, eps


def make_style_noise(flags, seed=2, dtype=torch.float32, batch=None):
    """reparameterisation eps of the style latents: modality -> [B, style_dim] (None when not factorized)"""
    if not factorized(flags):
        return None
    B = batch or flags.batch_size
    out = OrderedDict()
    for m in flags.mods:
        sd = style_dim(flags, m)
        u1 = seeded_uniform('eps_style.u1.' + m, seed, (B, sd)).clamp_min(1e-12)
        u2 = seeded_uniform('eps_style.u2.' + m, seed, (B, sd))
        out[m] = (torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2.0 * math.pi * u2)).to(dtype)
    return out


# --------------------------------------------------------------------------------------
# networks
# --------------------------------------------------------------------------------------
class _Ctx:
    """Carries state, masks, train flag and collects the BN running-stat updates."""

    def __init__(self, state, masks, train):
        self.s, self.masks, self.train = state, masks, train
        self.bn_updates = OrderedDict()


def _bn(ctx, p, x):
    """nn.BatchNorm{1,2}d defaults: train -> biased batch var for normalisation, running stats
    updated with momentum 0.1 and the UNBIASED var; eval -> running stats."""
    s = ctx.s
    if ctx.train:
        dims = [0] + list(range(2, x.dim()))
        n = x.numel() // x.shape[1]
        with torch.no_grad():
            mean = x.mean(dims)
            var_b = x.var(dims, unbiased=False)
            ctx.bn_updates[p + '.running_mean'] = (1 - BN_MOMENTUM) * s[p + '.running_mean'] + BN_MOMENTUM * mean
            ctx.bn_updates[p + '.running_var'] = ((1 - BN_MOMENTUM) * s[p + '.running_var']
                                                  + BN_MOMENTUM * var_b * (n / max(n - 1, 1)))
        return F.batch_norm(x, None, None, s[p + '.weight'], s[p + '.bias'], True, BN_MOMENTUM, BN_EPS)
    return F.batch_norm(x, s[p + '.running_mean'], s[p + '.running_var'], s[p + '.weight'], s[p + '.bias'],
                        False, BN_MOMENTUM, BN_EPS)


def _dropout(ctx, name, x):
    """nn.Dropout(p=.5) / nn.Dropout2d(p=.5): survivors scaled by 1/(1-p) = 2 (ResidualBlocks.py:10,73)."""
    if not ctx.train:
        return x
    return x * ctx.masks[name] * 2.0


def _res_block(ctx, p, x, k, stride, pad, nd, transposed):
    """ResidualBlocks.py forward():20-33 / 51-65 / 84-97 / 118-131 (all four are the same graph)."""
    s = ctx.s
    if nd == 1:
        conv, convt = F.conv1d, F.conv_transpose1d
    else:
        conv, convt = F.conv2d, F.conv_transpose2d
    short = 'upsample' if transposed else 'downsample'
    out = F.relu(_bn(ctx, p + '.bn1', x))
    b1 = s.get(p + '.conv1.bias')
    out = (convt if transposed else conv)(out, s[p + '.conv1.weight'], b1)
    out = _dropout(ctx, p + '.dropout1', out)
    out = F.relu(_bn(ctx, p + '.bn2', out))
    b2 = s.get(p + '.conv2.bias')
    if transposed:
        out = convt(out, s[p + '.conv2.weight'], b2, stride=stride, padding=pad)
        res = convt(x, s[p + '.%s.0.weight' % short], s[p + '.%s.0.bias' % short], stride=stride, padding=pad)
    else:
        out = conv(out, s[p + '.conv2.weight'], b2, stride=stride, padding=pad)
        res = conv(x, s[p + '.%s.0.weight' % short], s[p + '.%s.0.bias' % short], stride=stride, padding=pad)
    out = _dropout(ctx, p + '.dropout2', out)
    res = _bn(ctx, p + '.%s.1' % short, res)
    return RES_A * res + RES_B * out


def _compress(s, e, f):
    """LinearFeatureCompressor.forward (FeatureCompressor.py:21-28); the encoders return content first, style after
    (ConvNetworksImgMimic.py:29-36)"""
    mu = F.linear(f, s[e + '.feature_compressor.content_mu.weight'], s[e + '.feature_compressor.content_mu.bias'])
    lv = F.linear(f, s[e + '.feature_compressor.content_logvar.weight'],
                  s[e + '.feature_compressor.content_logvar.bias'])
    if e + '.feature_compressor.style_mu.weight' in s:
        smu = F.linear(f, s[e + '.feature_compressor.style_mu.weight'], s[e + '.feature_compressor.style_mu.bias'])
        slv = F.linear(f, s[e + '.feature_compressor.style_logvar.weight'],
                       s[e + '.feature_compressor.style_logvar.bias'])
        return mu, lv, smu, slv
    return mu, lv


def encoder_img(ctx, flags, e, x):
    """EncoderImg.forward (ConvNetworksImgMimic.py:29-36) -> FeatureExtractorImg.forward
    (FeatureExtractorImg.py:61-81) -> LinearFeatureCompressor.forward (FeatureCompressor.py:21-28)."""
    s = ctx.s
    h = F.conv2d(x, s[e + '.feature_extractor.conv1.weight'], None, stride=2, padding=1)
    for i, (ci, co, k, st, p) in enumerate(img_encoder_blocks(flags)):
        h = _res_block(ctx, e + '.feature_extractor.resblock_%d.0' % (i + 1), h, k, st, p, 2, False)
    return _compress(s, e, h.reshape(h.shape[0], -1))


def encoder_text(ctx, flags, e, x):
    """EncoderText.forward (ConvNetworksTextMimic.py:23-36) -> FeatureExtractorText.forward
    (char_encoding/FeatureExtractorText.py:58-81).  x: [B, L, num_features]."""
    s = ctx.s
    if is_word(flags):      # word_encoding/mmvae_text_enc.py:69-71: embedding(x.long()) (padding_idx only affects the gradient)
        x = F.embedding(x.long(), s[e + '.feature_extractor.embedding.weight'], padding_idx=0)
    h = F.conv1d(x.transpose(-2, -1), s[e + '.feature_extractor.conv1.weight'],
                 s[e + '.feature_extractor.conv1.bias'], stride=2, padding=1)
    for i, (ci, co, k, st, p) in enumerate(text_encoder_blocks_used(flags)):
        h = _res_block(ctx, e + '.feature_extractor.resblock_%d.0' % (i + 1), h, k, st, p, 1, False)
    return _compress(s, e, h.reshape(h.shape[0], -1))


def decoder_img(ctx, flags, d, z):
    """DecoderImg.forward (ConvNetworksImgMimic.py:46-54) -> DataGeneratorImg (DataGeneratorImg.py:29-98).
    Returns loc of the Laplace likelihood [B,1,px,px] (scale is the constant 0.75)."""
    s = ctx.s
    h = F.linear(z, s[d + '.feature_generator.weight'], s[d + '.feature_generator.bias'])
    h = h.view(h.shape[0], h.shape[1], 1, 1)
    blocks = img_decoder_blocks(flags)
    for i, (ci, co, k, st, p) in enumerate(blocks):
        h = _res_block(ctx, d + '.img_generator.generator.%d.0' % i, h, k, st, p, 2, True)
    n = len(blocks)
    return F.conv_transpose2d(h, s[d + '.img_generator.generator.%d.weight' % n],
                              s[d + '.img_generator.generator.%d.bias' % n], stride=2, padding=1, output_padding=1)


def decoder_text(ctx, flags, d, z):
    """DecoderText.forward (ConvNetworksTextMimic.py:51-68) -> DataGeneratorText.forward
    (char_encoding/DataGeneratorText.py:53-76).  Returns log-softmaxed logits [B, L, num_features]."""
    s = ctx.s
    h = F.linear(z, s[d + '.feature_generator.weight'], s[d + '.feature_generator.bias']).unsqueeze(-1)
    blocks = text_decoder_blocks(flags)
    for i, (ci, co, k, st, p) in enumerate(blocks):
        h = _res_block(ctx, d + text_dec_block_name(flags, i), h, k, st, p, 1, True)
    if is_word(flags):
        last = d + '.text_generator.generator.%d' % len(blocks)
        if flags.len_sequence == 128:
            h = F.conv1d(h, s[last + '.weight'], s[last + '.bias'])
        else:
            h = F.conv_transpose1d(h, s[last + '.weight'], s[last + '.bias'], stride=2, padding=1)
    else:
        h = F.conv_transpose1d(h, s[d + '.text_generator.conv2.weight'], s[d + '.text_generator.conv2.bias'],
                               stride=2, padding=1)
    return F.log_softmax(h, dim=1).transpose(-2, -1)


# --------------------------------------------------------------------------------------
# MoPoE fusion / ELBO
# --------------------------------------------------------------------------------------
def subset_keys(mod_names):
    """BaseExperiment.set_subsets (utils/BaseExperiment.py:66-82): powerset in itertools order,
    key = '_'.join(sorted(names)), members sorted by name.  Returns OrderedDict key -> [names]."""
    xs = list(mod_names)
    out = OrderedDict()
    for names in chain.from_iterable(combinations(xs, n) for n in range(len(xs) + 1)):
        out['_'.join(sorted(names))] = sorted(names)
    return out


def poe(mu, logvar):
    """mm_div.poe (evaluation/divergence_measures/mm_div.py:10-17).  mu, logvar: [m, B, D]."""
    var = torch.exp(logvar) + POE_EPS
    T = 1.0 / var
    pd_mu = torch.sum(mu * T, dim=0) / torch.sum(T, dim=0)
    pd_var = 1.0 / torch.sum(T, dim=0)
    return pd_mu, torch.log(pd_var)


def selection_bounds(num_samples, weights):
    """utils.mixture_component_selection index math (utils/utils.py:62-73) with FP32 weights:
    end_k = start_k + int(floor(B * w_k)); the last component takes the remainder."""
    w = torch.as_tensor(weights, dtype=torch.float32)
    w = w / w.sum()                       # reweight_weights (utils/utils.py:51-52) in moe_fusion
    starts, ends = [], []
    for k in range(w.shape[0]):
        i_start = 0 if k == 0 else ends[k - 1]
        if k == w.shape[0] - 1:
            i_end = num_samples
        else:
            i_end = i_start + int(torch.floor(num_samples * w[k]))
        starts.append(i_start)
        ends.append(i_end)
    ends[-1] = num_samples
    return starts, ends


def mixture_component_selection(mus, logvars, weights):
    """utils/utils.py:55-77.  mus, logvars [S,B,D] -> [B,D] by contiguous batch ranges."""
    starts, ends = selection_bounds(mus.shape[1], weights)
    S = len(starts)
    mu = torch.cat([mus[k, starts[k]:ends[k], :] for k in range(S)])
    lv = torch.cat([logvars[k, starts[k]:ends[k], :] for k in range(S)])
    return mu, lv


def kl_to_standard_normal(mu, logvar, norm_value):
    """kl_div.calc_kl_divergence, prior branch (evaluation/divergence_measures/kl_div.py:8-16)."""
    return -0.5 * torch.sum(1 - logvar.exp() - mu.pow(2) + logvar) / float(norm_value)


def inference(enc_mods, flags, present):
    """BaseMMVae.inference (utils/BaseMMVae.py:139-196) for methods moe / jsd / poe / joint_elbo.
    enc_mods: name -> (mu, logvar); present: modality names in the input batch (dict order)."""
    method = flags.method
    subsets = subset_keys(flags.mods)
    mus, logvars, distr = [], [], OrderedDict()
    for key, members in subsets.items():
        if key == '' or not all(m in present for m in members):
            continue
        mu_s = torch.stack([enc_mods[m][0] for m in members])
        lv_s = torch.stack([enc_mods[m][1] for m in members])
        if method in ('poe', 'joint_elbo'):
            if method == 'poe':           # prior expert appended to EVERY subset (BaseMMVae.py:117-124)
                z = torch.zeros_like(mu_s[:1])
                mu_s, lv_s = torch.cat((mu_s, z)), torch.cat((lv_s, z))
            s_mu, s_lv = poe(mu_s, lv_s)
        else:                              # moe: selection among the members, uniform weights
            w = torch.full((mu_s.shape[0],), 1.0 / mu_s.shape[0])
            s_mu, s_lv = mixture_component_selection(mu_s, lv_s, w)
        distr[key] = (s_mu, s_lv)
        if method in ('moe', 'jsd'):
            cond = len(members) == 1                  # fusion_condition_moe  :130-131 (jsd uses it too, :58-61)
        elif method == 'poe':
            cond = len(members) == len(present)       # fusion_condition_poe  :133-134
        else:
            cond = True                               # fusion_condition_joint:136-137
        if cond:
            mus.append(s_mu)
            logvars.append(s_lv)
    mus, logvars = torch.stack(mus), torch.stack(logvars)
    if method == 'jsd':                    # the prior N(0, I) joins the mixture as one more component (:180-186)
        zrow = torch.zeros_like(mus[:1])
        mus, logvars = torch.cat((mus, zrow)), torch.cat((logvars, zrow))
    S = mus.shape[0]
    weights = torch.full((S,), 1.0 / S, dtype=torch.float32)
    j_mu, j_lv = mixture_component_selection(mus, logvars, weights)
    return dict(modalities=enc_mods, mus=mus, logvars=logvars, weights=weights, joint=(j_mu, j_lv), subsets=distr)


def alpha_poe(alpha, mu, logvar):
    """mm_div.alpha_poe (mm_div.py:20-32): weighted product of experts = the jsd dynamic prior.  mu, logvar [K,B,D]."""
    var = torch.exp(logvar) + POE_EPS
    a = alpha.to(mu.dtype).unsqueeze(-1).unsqueeze(-1)
    T = 1 / var
    pd_var = 1.0 / torch.sum(a * T, dim=0)
    pd_mu = pd_var * torch.sum(a * mu * T, dim=0)
    return pd_mu, torch.log(pd_var)


def kl_between(mu0, logvar0, mu1, logvar1, norm_value):
    """kl_div.calc_kl_divergence, two-Gaussian branch (kl_div.py:11-13)."""
    return -0.5 * torch.sum(1 - logvar0.exp() / logvar1.exp() - (mu0 - mu1).pow(2) / logvar1.exp()
                            + logvar0 - logvar1) / float(norm_value)


def laplace_log_prob_sum(loc, target):
    """dist.Laplace(loc, 0.75).log_prob(target).sum()  (modalities/Modality.py:25-30)."""
    b = LAPLACE_SCALE
    # the reference's scale is an fp32 tensor (torch.tensor(0.75)), so log(2b) is evaluated in fp32
    log2b = float(torch.log(torch.tensor(2 * b, dtype=torch.float32)))
    return (-log2b - (target - loc).abs() / b).sum()


def categorical_log_prob_sum(logits, target):
    """dist.OneHotCategorical(logits=logits).log_prob(target).sum(): logits are re-normalised
    (idempotent after the decoder's LogSoftmax) and indexed by argmax(target)."""
    ln = logits - logits.logsumexp(dim=-1, keepdim=True)
    # word encoding: the target holds token indices and MimicText.calc_log_prob one-hot encodes it (MimicText.py:37-40)
    idx = target.long() if target.dim() == logits.dim() - 1 else target.max(-1)[1]
    return ln.gather(-1, idx.unsqueeze(-1)).sum()


def forward(state, batch, flags, masks=None, eps=None, train=True, present=None, eps_style=None):
    """VAEtrimodalMimic.forward (networks/VAEtrimodalMimic.py:31-62), tolerant of missing modalities
    in the decode loop (the intended behaviour for calc_poe_loss)."""
    ctx = _Ctx(state, masks or {}, train)
    present = list(present or [m for m in flags.mods if m in batch])
    enc = OrderedDict()
    for m in present:
        if m == 'text':
            enc[m] = encoder_text(ctx, flags, ENC_NAME[m], batch[m])
        else:
            enc[m] = encoder_img(ctx, flags, ENC_NAME[m], batch[m])
    styles = OrderedDict((m, tuple(v[2:])) for m, v in enc.items() if len(v) == 4)     # VAEtrimodalMimic.encode:64-93
    enc = OrderedDict((m, tuple(v[:2])) for m, v in enc.items())
    lat = inference(enc, flags, present)
    lat['styles'] = styles
    # calc_group_divergence_moe (mm_div.py:90-110): the reference collects the per-subset KLs in
    # `torch.zeros(num_mods)` -- an FP32 tensor whatever the model dtype -- and the weights are FP32 too
    # (BaseMMVae.py:187), so joint_divergence is an fp32 quantity even in an fp64 run.
    dyn_prior = None
    if flags.method == 'jsd':
        # divergence_dynamic_prior (BaseMMVae.py:87-99) -> calc_alphaJSD_modalities (mm_div.py:67-87) with the weights
        # forward() passes: latents['weights'] (uniform 1/(M+1), NOT re-normalised, alpha_modalities is not used here)
        w = lat['weights']
        a_mu, a_lv = alpha_poe(w, lat['mus'], lat['logvars'])
        klds_ind = torch.stack([kl_between(lat['mus'][k], lat['logvars'][k], a_mu, a_lv, flags.batch_size)
                                for k in range(lat['mus'].shape[0])]).to(torch.float32)
        dyn_prior = (a_mu, a_lv)
    else:
        w = lat['weights'] / lat['weights'].sum()
        klds_ind = torch.stack([kl_to_standard_normal(lat['mus'][k], lat['logvars'][k], flags.batch_size)
                                for k in range(lat['mus'].shape[0])]).to(torch.float32)
    joint_div = (w * klds_ind).sum()
    j_mu, j_lv = lat['joint']
    if eps is None:
        eps = torch.zeros_like(j_mu)
    z = eps * torch.exp(0.5 * j_lv) + j_mu   # utils.reparameterize (utils/utils.py:45-48)
    rec = OrderedDict()
    for m in present:
        zm = z
        if m in styles:                   # VAEtrimodalMimic.forward:49-51 + DecoderImg.forward:47-48: cat(style, content)
            s_mu, s_lv = styles[m]
            e_s = eps_style[m] if eps_style is not None else torch.zeros_like(s_mu)
            zm = torch.cat((e_s * torch.exp(0.5 * s_lv) + s_mu, z), dim=1)
        if m == 'text':
            rec[m] = decoder_text(ctx, flags, DEC_NAME[m], zm)
        else:
            rec[m] = decoder_img(ctx, flags, DEC_NAME[m], zm)
    return dict(latents=lat, joint_divergence=joint_div, individual_divs=klds_ind, dyn_prior=dyn_prior, z=z, rec=rec,
                bn_updates=ctx.bn_updates)


def step_losses(state, batch, flags, masks=None, eps=None, train=True, uni_masks=None, eps_style=None, uni_eps_style=None):
    """run_epochs.basic_routine_epoch (run_epochs.py:52-96): forward, calc_log_probs (losses.py:6-21),
    calc_klds (:24-31), calc_joint_elbo_loss (:80-89) or calc_poe_loss (:54-77, intended semantics)."""
    res = forward(state, batch, flags, masks, eps, train, eps_style=eps_style)
    Bn = float(flags.batch_size)
    log_probs, weighted = OrderedDict(), 0.0
    for m in flags.mods:
        if m == 'text':
            lp = categorical_log_prob_sum(res['rec'][m], batch[m])
        else:
            lp = laplace_log_prob_sum(res['rec'][m], batch[m])
        log_probs[m] = -lp / Bn
        weighted = weighted + flags.rec_weights[m] * log_probs[m]
    klds = OrderedDict((k, kl_to_standard_normal(mu, lv, Bn)) for k, (mu, lv) in res['latents']['subsets'].items())
    # calc_klds_style (losses.py:34-42) + calc_style_kld (:45-51)
    klds_style = OrderedDict((m + '_style', kl_to_standard_normal(mu, lv, Bn)) for m, (mu, lv) in res['latents']['styles'].items())
    kld_style = sum(flags.style_weights[m] * klds_style[m + '_style'] for m in flags.mods) if klds_style else 0.0
    if flags.method in ('moe', 'jsd', 'joint_elbo'):
        total = weighted + flags.beta * (flags.beta_style * kld_style + flags.beta_content * res['joint_divergence'])
    elif flags.method == 'poe':
        # calc_poe_loss (losses.py:54-77) + utils.calc_elbo (utils/utils.py:105-127): the style KLs of the JOINT pass enter
        # the joint ELBO (weighted sum) and every unimodal ELBO (the modality's own)
        total = weighted + flags.beta * (flags.beta_content * res['joint_divergence'] + flags.beta_style * kld_style)
        state2 = dict(state)              # the unimodal passes see the BN buffers the joint pass updated
        state2.update(res['bn_updates'])
        for i, m in enumerate(flags.mods):
            um, ue = (uni_masks or {}).get(m, (masks, eps))
            es_m = (uni_eps_style or {}).get(m, eps_style)
            r_m = forward(state2, {m: batch[m]}, flags, um, ue, train, present=[m], eps_style=es_m)
            lp = (categorical_log_prob_sum if m == 'text' else laplace_log_prob_sum)(r_m['rec'][m], batch[m])
            ks_m = flags.style_weights[m] * klds_style[m + '_style'] if klds_style else 0.0
            total = total + (-lp / Bn) + flags.beta * (flags.beta_content * klds[m] + flags.beta_style * ks_m)   # calc_elbo modality
            for k, v in r_m['bn_updates'].items():
                res['bn_updates'][k] = v
    else:
        raise NotImplementedError(flags.method)
    return dict(results=res, log_probs=log_probs, klds=klds, klds_style=klds_style, total_loss=total,
                weighted_log_prob=weighted)


def step_with_grads(state, batch, flags, masks=None, eps=None, uni_masks=None, eps_style=None, uni_eps_style=None):
    """One train-mode step: losses + d(total_loss)/d(param) for every float parameter."""
    params = {k: v for k, v in state.items() if v.is_floating_point() and 'running_' not in k}
    for v in params.values():
        v.requires_grad_(True)
        v.grad = None
    out = step_losses(state, batch, flags, masks, eps, True, uni_masks=uni_masks, eps_style=eps_style, uni_eps_style=uni_eps_style)
    out['total_loss'].backward()
    # (parameters the step never touches — the word encoder's resblock_7/8 at len_sequence <= 500 — have no gradient)
    grads = OrderedDict((k, v.grad.detach().clone()) for k, v in params.items() if v.grad is not None)
    for v in params.values():
        v.requires_grad_(False)
        v.grad = None
    out['grads'] = grads
    return out


step_with_grads_full = step_with_grads


def adam_step(params, grads, m, v, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam defaults as used by MimicExperiment.set_optimizer (utils/experiment.py:171-178):
    no weight decay, no amsgrad.  In-place on params/m/v; `step` is the 1-based step count."""
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    for k in params:
        if k not in grads:            # never used in the step (no gradient): torch.optim.Adam skips it too
            continue
        g = grads[k]
        m[k].mul_(b1).add_(g, alpha=1 - b1)
        v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v[k].sqrt() / math.sqrt(bc2)).add_(eps)
        params[k].addcdiv_(m[k], denom, value=-lr / bc1)


# --------------------------------------------------------------------------------------
# evaluation path: decode from latents, per-sample log-probs, importance-sampled likelihoods
# --------------------------------------------------------------------------------------
def decode(state, flags, z, z_style=None, train=False):
    """VAEtrimodalMimic.generate_sufficient_statistics_from_latents (networks/VAEtrimodalMimic.py:143-152): every decoder on
    the same content rows (cat(style, content) when factorized).  Returns modality -> Laplace loc / log-softmaxed logits.
    (The reference decodes text in flags.batch_size chunks, ConvNetworksTextMimic.py:59-64 — a no-op for the values in
    eval mode, where BatchNorm reads its running statistics.)"""
    ctx = _Ctx(state, {}, train)
    out = OrderedDict()
    for m in flags.mods:
        zm = z if not (z_style and z_style.get(m) is not None) else torch.cat((z_style[m], z), dim=1)
        out[m] = decoder_text(ctx, flags, DEC_NAME[m], zm) if m == 'text' else decoder_img(ctx, flags, DEC_NAME[m], zm)
    return out


def likelihood_mean(m, dec_out):
    """`.mean` of the likelihood object (BaseMMVae.generate_from_latents:210-216): Laplace loc, categorical probabilities"""
    return dec_out.exp() if m == 'text' else dec_out


def log_prob_rows(m, dec_out, x):
    """likelihood.log_prob(x).view(rows, -1).sum(dim=1)  (utils/likelihood.py:120-121, :187)"""
    rows = x.shape[0]
    if m == 'text':
        ln = dec_out - dec_out.logsumexp(dim=-1, keepdim=True)
        lp = ln.gather(-1, x.max(-1)[1].unsqueeze(-1)).squeeze(-1)
    else:
        b = LAPLACE_SCALE
        log2b = float(torch.log(torch.tensor(2 * b, dtype=torch.float32)))
        lp = -log2b - (x - dec_out).abs() / b
    return lp.reshape(rows, -1).sum(dim=1)


def _gaussian_log_pdf(x, mu, logvar):
    """utils/likelihood.py:54-66"""
    log2pi = float(math.log(2.0 * math.pi))
    return torch.sum(-0.5 * log2pi - logvar / 2. - torch.pow(x - mu, 2) / (2. * torch.exp(logvar)), dim=1)


def _log_mean_exp(x, dim=1):
    """utils/likelihood.py:39-51"""
    m = torch.max(x, dim=dim, keepdim=True)[0]
    return m + torch.log(torch.mean(torch.exp(x - m), dim=dim, keepdim=True))


def importance_likelihoods(flags, K, dec, batch, z, mu, lv):
    """calc_log_likelihood_batch without style latents (evaluation/eval_metrics/likelihood.py:17-93): per-modality
    log_marginal_estimate (utils/likelihood.py:82-141) and log_joint_estimate (:144-220) with a static N(0, I) prior.
    dec: decode() output on the K*B rows z (sample-major: row k*B + b); mu, lv: the conditioning subset's posterior [B, D]."""
    B = flags.batch_size
    mu_r = mu.unsqueeze(0).repeat(K, 1, 1).view(K * B, -1)
    lv_r = lv.unsqueeze(0).repeat(K, 1, 1).view(K * B, -1)
    log_q = _gaussian_log_pdf(z, mu_r, lv_r)
    log_p = _gaussian_log_pdf(z, torch.zeros_like(z), torch.zeros_like(z))      # unit_gaussian_log_pdf (:69-79)
    out, rows = OrderedDict(), []
    for m in flags.mods:
        x = batch[m]
        xr = x.unsqueeze(0).repeat(K, *([1] * x.dim())).view(K * B, *x.shape[1:])
        lp = log_prob_rows(m, dec[m], xr)
        rows.append(lp)
        lw = (lp + log_p - log_q).view(B, K)          # the reference views the sample-major rows as (batch, samples)
        out[m] = torch.mean(_log_mean_exp(lw, dim=1))
    # log_joint_estimate collects the per-modality rows in `torch.zeros(num_mods, B*K)` — an FP32 tensor whatever the model
    # dtype (utils/likelihood.py:183-188) — and sums over modalities there: an fp32 quantity even in an fp64 run
    lw = (torch.stack(rows).to(torch.float32).sum(0) + log_p - log_q).view(B, K)
    out['joint'] = torch.mean(_log_mean_exp(lw, dim=1))
    return out
