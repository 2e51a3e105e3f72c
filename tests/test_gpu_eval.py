"""Evaluation path on the GPU against the reference's own outputs (tests/golden/eval_tri.pt, generated
from the unmodified reference by oracle/gen_golden_eval.py): eval-mode inference on full and partial-modality batch-1
inputs, generate / cond_generation, and the importance-sampled likelihood — decoders at B*K rows with the reference's
chunked text decode, `.log_prob` / `.mean` of the likelihood objects.  The callers' arithmetic (log_marginal_estimate /
log_joint_estimate, utils/likelihood.py:82-220) is restated here exactly as the evaluation code would run it on the
product's objects.  Tolerances: fp32 validation mode 2e-5 relative; bf16 mode 3e-2 on latents / reconstructions."""
import math
import os
from collections import OrderedDict

import pytest
import torch

from oracle import gen_golden_eval as GE
from oracle import mopoe_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _fixture(golden_dir):
    return torch.load(os.path.join(golden_dir, 'eval_tri.pt'), weights_only=False)


def _model(fx, compute_dtype):
    import mopoe_mimic_b200 as P
    ofl = H.oracle_flags(**fx['flags'])
    state = GE.eval_state(ofl, torch.float32)
    exp = P.Experiment(H.product_flags(ofl, compute_dtype))
    exp.mm_vae.load_state_dict(state)
    exp.mm_vae.eval()
    batch = OrderedDict((k, v.float().cuda()) for k, v in O.make_batch(ofl, seed=1, dtype=torch.float64).items())
    return ofl, exp, batch


def _rel(a, ref):
    a, ref = a.detach().double().cpu(), ref.double()
    return float((a - ref).abs().max() / (ref.abs().max() + 1e-300))


def _check_sum(t, cs, tol):
    """checksum triple written by gen_golden.checksum: sum, l2 norm, 8 probe values"""
    t = t.detach().double().cpu().reshape(-1)
    assert t.numel() == cs['numel']
    assert abs(float(t.norm()) - cs['l2']) <= tol * cs['l2'], (float(t.norm()), cs['l2'])
    scale = cs['l2'] / math.sqrt(cs['numel'])
    assert float((t[cs['pos']] - cs['val']).abs().max()) <= tol * max(scale, float(cs['val'].abs().max())) * 4


@pytest.mark.parametrize('cd,tol', [('fp32', 2e-5), ('bf16', 3e-2)])
def test_eval_inference_full_and_partial_batches(golden_dir, cd, tol):
    fx = _fixture(golden_dir)
    ofl, exp, batch = _model(fx, cd)
    vae = exp.mm_vae
    with torch.no_grad():
        lat = vae.inference(OrderedDict(batch))
        assert list(lat['subsets'].keys()) == list(fx['subsets'].keys())
        for k, (mu, lv) in fx['subsets'].items():
            assert _rel(lat['subsets'][k][0], mu) < tol and _rel(lat['subsets'][k][1], lv) < tol, k
        assert _rel(lat['joint'][0], fx['joint'][0]) < tol
        for tag, p in fx['partial'].items():          # plotting.py:74-79,151-159: model.inference(i_batch, num_samples=1)
            lp = vae.inference({m: batch[m][:p['rows']] for m in p['mods']}, num_samples=p['rows'])
            assert list(lp['subsets'].keys()) == list(p['subsets'].keys()), tag
            for k, (mu, lv) in p['subsets'].items():
                assert _rel(lp['subsets'][k][0], mu) < tol and _rel(lp['subsets'][k][1], lv) < tol, (tag, k)
            assert _rel(lp['joint'][0], p['joint'][0]) < tol and _rel(lp['joint'][1], p['joint'][1]) < tol, tag
            for m in ofl.mods:                         # absent modalities: [None, None] (VAEtrimodalMimic.encode:64-93)
                if m not in p['mods']:
                    assert lp['modalities'][m] == [None, None]


def _gaussian_log_pdf(x, mu, logvar):
    log2pi = float(math.log(2.0 * math.pi))
    return torch.sum(-0.5 * log2pi - logvar / 2. - torch.pow(x - mu, 2) / (2. * torch.exp(logvar)), dim=1)


def _log_mean_exp(x, dim=1):
    m = torch.max(x, dim=dim, keepdim=True)[0]
    return m + torch.log(torch.mean(torch.exp(x - m), dim=dim, keepdim=True))


@pytest.mark.parametrize('cd,tol', [('fp32', 2e-5), ('bf16', 3e-2)])
def test_importance_sampled_likelihood_on_BK_rows(golden_dir, cd, tol):
    """evaluation/eval_metrics/likelihood.py:17-93 driven through the product's model API"""
    fx = _fixture(golden_dir)
    ofl, exp, batch = _model(fx, cd)
    vae, K, B = exp.mm_vae, fx['k_imp'], ofl.batch_size
    eps = GE.eval_noise(ofl, torch.float64)['eps_imp']
    for s_key, d in fx['lhood'].items():
        mu, lv = fx['subsets'][s_key]
        z = (eps * torch.exp(0.5 * lv.unsqueeze(0)) + mu.unsqueeze(0)).view(K * B, -1)      # get_latent_samples
        with torch.no_grad():
            gen = vae.generate_sufficient_statistics_from_latents({'content': z.float().cuda(),
                                                                   'style': {m: None for m in ofl.mods}})
            mu_r = mu.unsqueeze(0).repeat(K, 1, 1).view(K * B, -1)
            lv_r = lv.unsqueeze(0).repeat(K, 1, 1).view(K * B, -1)
            log_q = _gaussian_log_pdf(z, mu_r, lv_r)
            log_p = _gaussian_log_pdf(z, torch.zeros_like(z), torch.zeros_like(z))
            rows = []
            for m in ofl.mods:
                x = batch[m]
                xr = x.unsqueeze(0).repeat(K, *([1] * x.dim())).view(K * B, *x.shape[1:])
                lp = gen[m].log_prob(xr)                               # the likelihood object's own log_prob
                assert lp.shape[0] == K * B
                lp = lp.view(K * B, -1).sum(dim=1).double().cpu()
                assert _rel(lp, d['logp_rows'][m]) < tol, (s_key, m)
                _check_sum(gen[m].mean, d['mean'][m], tol)
                rows.append(lp)
                ll_m = float(torch.mean(_log_mean_exp((lp + log_p - log_q).view(B, K), dim=1)))
                assert abs(ll_m - d['ll'][m]) <= tol * abs(d['ll'][m]), (s_key, m, ll_m, d['ll'][m])
            lw = (torch.stack(rows).to(torch.float32).sum(0) + log_p - log_q).view(B, K)
            ll_j = float(torch.mean(_log_mean_exp(lw, dim=1)))
            assert abs(ll_j - d['ll']['joint']) <= tol * abs(d['ll']['joint'])


@pytest.mark.parametrize('cd,tol', [('fp32', 2e-5), ('bf16', 3e-2)])
def test_cond_generation_and_generate(golden_dir, cd, tol, monkeypatch):
    import mopoe_mimic_b200.mmvae as MM
    fx = _fixture(golden_dir)
    ofl, exp, batch = _model(fx, cd)
    vae = exp.mm_vae
    eps = GE.eval_noise(ofl, torch.float64)['eps_cg'].float().cuda()
    monkeypatch.setattr(MM, 'reparameterize', lambda mu, logvar: eps * torch.exp(0.5 * logvar) + mu)
    dists = {k: [fx['subsets'][k][0].float().cuda(), fx['subsets'][k][1].float().cuda()] for k in fx['cond_gen']}
    with torch.no_grad():
        cg = vae.cond_generation(dists, num_samples=ofl.batch_size)
        assert list(cg.keys()) == list(fx['cond_gen'].keys())
        for k, per_mod in fx['cond_gen'].items():
            assert list(cg[k].keys()) == list(per_mod.keys())
            for m, cs in per_mod.items():
                _check_sum(cg[k][m], cs, tol)
        assert _rel(cg['Lateral_PA_text']['PA'], fx['cond_gen_pa_full']) < tol
        torch.manual_seed(fx['generate']['seed'])
        g = vae.generate(fx['generate']['n'])
        for m, cs in fx['generate']['out'].items():
            _check_sum(g[m], cs, tol)
        # default num_samples = flags.batch_size; shapes of the three modalities
        g = vae.generate()
        assert g['PA'].shape == (ofl.batch_size, 1, ofl.img_size, ofl.img_size)
        assert g['text'].shape == (ofl.batch_size, ofl.len_sequence, ofl.num_features)


def test_random_style_dists_and_styles():
    """VAEtrimodalMimic.get_random_styles / get_random_style_dists:95-125"""
    import mopoe_mimic_b200 as P
    fl = P.default_flags(device=torch.device('cuda'), batch_size=4, DIM_img=16, DIM_text=16, class_dim=32,
                         factorized_representation=True, style_pa_dim=8, style_lat_dim=16, style_text_dim=24)
    vae = P.Experiment(fl).mm_vae
    d = vae.get_random_style_dists(5)
    assert [tuple(d[m][i].shape) for m in ('PA', 'Lateral', 'text') for i in (0, 1)] == [(5, 8)] * 2 + [(5, 16)] * 2 + [(5, 24)] * 2
    assert all(float(t.abs().sum()) == 0.0 for v in d.values() for t in v)
    torch.manual_seed(3)
    s = vae.get_random_styles(5)
    torch.manual_seed(3)
    ref = [torch.randn(5, n) for n in (8, 16, 24)]           # CPU generator, modality order, as the reference
    for t, r in zip((s['PA'], s['Lateral'], s['text']), ref):
        assert torch.equal(t.cpu(), r)
    vae.eval()
    with torch.no_grad():                                    # factorized decode from random styles: cat(style, content)
        out = vae.generate(5)
    assert out['Lateral'].shape == (5, 1, 128, 128) and bool(torch.isfinite(out['text']).all())
