"""Modality descriptors and likelihood objects (reference: modalities/Modality.py:15-47, MimicPA.py:7-17,
MimicLateral.py:7-17, MimicText.py:12-40, modalities/utils.py:4-15).

The likelihood objects behave like the torch.distributions the reference builds (`.log_prob(target)`,
`.mean`, `.loc` / `.logits`) for the evaluation callers and add the fused
`log_prob_sum(target)` path that Modality.calc_log_prob uses on the hot path.
"""
import math

import torch

from .blocks import CategoricalLogProbSumFn, LaplaceLogProbSumFn, laplace_log_prob, log_softmax_rows


def _need_engine(eng, what):
    if eng is None:
        raise RuntimeError('%s needs the CUDA engine (likelihood objects are built by the model\'s decoders); there is no '
                           'torch / CPU fallback' % what)
    return eng


class LaplaceLikelihood:
    """dist.Laplace(loc, scale) stand-in.  scale is the constant 0.75 tensor the decoder returns."""

    def __init__(self, loc, scale, eng=None, scale_value=None):
        self.loc, self.scale, self._eng = loc, scale, eng
        # scale_value: the host copy of the (constant) scale — avoids a device->host sync per forward
        if scale_value is not None:
            self._scale_f = float(scale_value)
        elif not torch.is_tensor(scale):
            self._scale_f = float(scale)
        else:
            self._scale_f = float(scale)          # one device->host read (foreign callers only; the model passes scale_value)

    @property
    def mean(self):
        return self.loc

    def log_prob(self, value):
        """elementwise -log(2b) - |x - loc| / b, shape of loc (the evaluation callers reduce it per sample:
        utils/likelihood.py:120-121); `value` broadcasts over leading dims of loc as in torch.distributions"""
        return laplace_log_prob(self.loc, value, self._scale_f, _need_engine(self._eng, 'LaplaceLikelihood.log_prob'))

    def log_prob_sum(self, value):
        return LaplaceLogProbSumFn.apply(self.loc, value, self._scale_f, _need_engine(self._eng, 'LaplaceLikelihood.log_prob_sum'))


class CategoricalLikelihood:
    """dist.OneHotCategorical(logits=...) stand-in built from the decoder's PRE-softmax scores [B, L, V].

    Targets must be strictly one-hot rows (or token indices): see blocks.CategoricalLogProbSumFn."""

    def __init__(self, logits=None, scores=None, eng=None):
        self._scores = scores if scores is not None else logits
        self._eng = eng
        self._logits = None

    @property
    def logits(self):
        if self._logits is None:
            self._logits = log_softmax_rows(self._scores, _need_engine(self._eng, 'CategoricalLikelihood.logits'))
        return self._logits

    @property
    def probs(self):
        return self.logits.exp()

    @property
    def mean(self):
        return self.probs

    def log_prob(self, value):
        # one-hot rows [.., V], or token indices [..] (word encoding: MimicText.calc_log_prob one-hot encodes them, :37-40)
        idx = value.long() if value.dim() == self.logits.dim() - 1 else value.max(-1)[1]
        return self.logits.gather(-1, idx.unsqueeze(-1)).squeeze(-1)

    def log_prob_sum(self, value):
        return CategoricalLogProbSumFn.apply(self._scores, value, _need_engine(self._eng, 'CategoricalLikelihood.log_prob_sum'))


def get_likelihood(name):
    """modalities/utils.py:4-15 (laplace / categorical are the ones the MIMIC modalities use)."""
    if name == 'laplace':
        return LaplaceLikelihood
    if name == 'categorical':
        return CategoricalLikelihood
    raise NotImplementedError('likelihood %r is not used by the MIMIC modalities' % name)


class Modality:
    def calc_log_prob(self, out_dist, target, norm_value):
        """log P(target | out_dist) / norm_value  (Modality.py:25-30); fused reduction when available."""
        if hasattr(out_dist, 'log_prob_sum'):
            log_prob = out_dist.log_prob_sum(target)
        else:
            log_prob = out_dist.log_prob(target).sum()
        return log_prob / norm_value


class _MimicImg(Modality):
    def __init__(self, name, enc, dec, args):
        self.name = name
        self.likelihood_name = 'laplace'
        self.data_size = torch.Size((1, args.img_size, args.img_size))
        self.gen_quality_eval = True
        self.file_suffix = '.png'
        self.encoder = enc
        self.decoder = dec
        self.likelihood = get_likelihood(self.likelihood_name)


class MimicPA(_MimicImg):
    def __init__(self, enc, dec, args):
        super().__init__('PA', enc, dec, args)


class MimicLateral(_MimicImg):
    def __init__(self, enc, dec, args):
        super().__init__('Lateral', enc, dec, args)


class MimicText(Modality):
    def __init__(self, enc, dec, len_sequence, plotImgSize=None, font=None, args=None):
        self.name = 'text'
        self.args = args
        self.likelihood_name = 'categorical'
        self.len_sequence = len_sequence
        if args.text_encoding == 'char':
            self.alphabet = getattr(args, 'alphabet', None)
            self.data_size = torch.Size((args.num_features, len_sequence))
        elif args.text_encoding == 'word':
            self.data_size = torch.Size((args.vocab_size, len_sequence))
        else:
            raise NotImplementedError('text_encoding %r' % args.text_encoding)
        self.plot_img_size = plotImgSize
        self.font = font
        self.gen_quality_eval = False
        self.file_suffix = '.txt'
        self.encoder = enc
        self.decoder = dec
        self.likelihood = get_likelihood(self.likelihood_name)
