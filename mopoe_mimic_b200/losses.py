"""ELBO assembly with the reference's free-function contract (evaluation/losses.py:6-89,
utils/utils.py:105-127): same names, same argument meaning, same return types.  The heavy parts
(log-likelihood reductions, per-subset KLs) come from fused CUDA kernels; what is left here is scalar
bookkeeping on 0-dim device tensors.
"""
from collections import OrderedDict

import torch

from .fusion import FusionFn, FusionPlan


def calc_kl_divergence(mu0, logvar0, mu1=None, logvar1=None, norm_value=None):
    """kl_div.calc_kl_divergence (evaluation/divergence_measures/kl_div.py:8-16), prior branch, through the
    fusion kernel's KL reduction.  (The two-Gaussian branch is only used inside calc_alphaJSD_modalities, which the jsd path
    runs fused — mopoe_jsd_divergence_fwd / BaseMMVae.divergence_dynamic_prior — so it is not exposed separately.)"""
    if mu1 is not None or logvar1 is not None:
        raise NotImplementedError('KL between two Gaussians: use mm_vae.divergence_dynamic_prior (the fused jsd divergence)')
    from .engine import Engine
    B, D = mu0.shape
    plan = FusionPlan(['e0'], ['e0'], ['e0'], [['e0']], 'moe', B, D, float(norm_value) if norm_value else 1.0)
    eng = _engine_for(mu0.device)
    out = FusionFn.apply(plan, eng, torch.zeros(B, D, device=mu0.device), mu0, logvar0)
    return out[5][0]


_ENGINES = {}


def _engine_for(device):
    from .engine import Engine
    key = str(device)
    if key not in _ENGINES:
        _ENGINES[key] = Engine(device, torch.float32)
    return _ENGINES[key]


def calc_log_probs(exp, result, batch):
    """losses.py:6-21: log_probs[m] = -log p(x_m | z) / batch_size ; weighted sum with exp.rec_weights."""
    mods = exp.modalities
    log_probs = {}
    weighted_log_prob = 0.0
    for m_key in mods:
        mod = mods[m_key]
        ba = batch[0][mod.name]
        log_probs[mod.name] = -mod.calc_log_prob(out_dist=result['rec'][mod.name], target=ba,
                                                 norm_value=exp.flags.batch_size)
        weighted_log_prob += exp.rec_weights[mod.name] * log_probs[mod.name]
    return log_probs, weighted_log_prob


def calc_klds(exp, result):
    """losses.py:24-31: KL(subset || N(0,I)) / batch_size for every entry of latents['subsets'].  The fused
    inference kernel has already reduced them; a foreign result dict falls back to the standalone reduction."""
    latents = result['latents']
    fused = latents.get('_klds')
    if fused is not None and list(fused.keys()) == list(latents['subsets'].keys()):
        return OrderedDict(fused)
    klds = {}
    for key in latents['subsets']:
        mu, logvar = latents['subsets'][key]
        klds[key] = calc_kl_divergence(mu, logvar, norm_value=exp.flags.batch_size)
    return klds


def calc_klds_style(exp, result):
    """losses.py:34-42: KL(style posterior || N(0,I)) / batch_size for every '<modality>_style' entry.  The fused
    style-reparameterisation launch of forward() already produced them (latents['_klds_style'])."""
    latents = result['latents']
    fused = latents.get('_klds_style')
    klds = {}
    for key, val in latents['modalities'].items():
        if key.endswith('style') and val[0] is not None:
            klds[key] = fused[key] if fused is not None and key in fused else \
                calc_kl_divergence(val[0], val[1], norm_value=exp.flags.batch_size)
    return klds


def calc_style_kld(exp, klds):
    """losses.py:45-51"""
    weighted_klds = 0.0
    for m_key in exp.modalities.keys():
        weighted_klds = weighted_klds + exp.style_weights[m_key] * klds[m_key + '_style']
    return weighted_klds


def calc_elbo(exp, modality, recs, klds):
    """utils.calc_elbo (utils/utils.py:105-127): klds = {'content': ..., 'style': {m: ...}}"""
    flags = exp.flags
    s_weights = exp.style_weights
    kld_content = klds['content']
    if modality == 'joint':
        w_style_kld = 0.0
        rec_error = 0.0
        for m_key in exp.modalities:
            w_style_kld = w_style_kld + s_weights[m_key] * klds['style'][m_key]
            rec_error = rec_error + exp.rec_weights[m_key] * recs[m_key]
        kld_style = w_style_kld
    else:
        kld_style = s_weights[modality] * klds['style'][modality]
        rec_error = 1.0 * recs[modality]
    div = flags.beta_content * kld_content + flags.beta_style * kld_style
    return rec_error + flags.beta * div


def calc_poe_loss(exp, mods, group_divergence, klds, klds_style, batch_d, mm_vae, log_probs):
    """losses.py:54-77: one unimodal forward pass per modality + the joint ELBO.  With a factorized representation the
    style KLs of the JOINT pass enter every ELBO (the unimodal passes only contribute their reconstruction terms)."""
    klds_joint = {'content': group_divergence, 'style': {}}
    factorized = bool(getattr(exp.flags, 'factorized_representation', False))
    elbos = {}
    for m_key in mods.keys():
        mod = mods[m_key]
        kld_style_m = klds_style[m_key + '_style'] if factorized else 0.0
        klds_joint['style'][m_key] = kld_style_m
        r_mod = mm_vae({m_key: batch_d[m_key]})
        log_prob_mod = -mod.calc_log_prob(r_mod['rec'][m_key], batch_d[m_key], exp.flags.batch_size)
        klds_mod = {'content': klds[m_key], 'style': {m_key: kld_style_m}}
        elbos[m_key] = calc_elbo(exp, m_key, {m_key: log_prob_mod}, klds_mod)
    elbos['joint'] = calc_elbo(exp, 'joint', log_probs, klds_joint)
    return sum(elbos.values())


def calc_joint_elbo_loss(exp, klds_style, group_divergence, beta_style, beta_content, weighted_log_prob, beta):
    """losses.py:80-89."""
    kld_style = calc_style_kld(exp, klds_style) if getattr(exp.flags, 'factorized_representation', False) else 0.0
    kld_weighted = beta_style * kld_style + beta_content * group_divergence
    return 1.0 * weighted_log_prob + beta * kld_weighted
