"""Golden fixture for the EVALUATION path from the UNMODIFIED reference, eval mode:

  * BaseMMVae.inference on the full batch and on partial-modality batch-1 / batch-2 inputs (utils/plotting.py:74-79,151-159)
  * VAEtrimodalMimic.generate / generate_from_latents / cond_generation (utils/BaseMMVae.py:198-231,
    networks/VAEtrimodalMimic.py:127-152)
  * the importance-sampled likelihood: get_latent_samples -> generate_sufficient_statistics_from_latents on B*K rows
    (text decoded in flags.batch_size chunks, networks/ConvNetworksTextMimic.py:59-64) -> likelihood.log_prob ->
    log_marginal_estimate / log_joint_estimate (evaluation/eval_metrics/likelihood.py:17-93, utils/likelihood.py:82-220)

Build-container only (imports /root/reference through oracle/gen_golden.py).  Writes tests/golden/eval_tri.pt and asserts
that oracle/mopoe_oracle.py reproduces the reference on every quantity (fp64, 1e-9).

    python oracle/gen_golden_eval.py
"""
import os
import sys
from collections import OrderedDict

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import gen_golden as G  # noqa: E402
from oracle import mopoe_oracle as O  # noqa: E402

KW = dict(batch_size=4, DIM_img=16, DIM_text=16, class_dim=32)
K_IMP = 3


def eval_state(fl, dtype):
    """deterministic state with non-trivial BatchNorm running statistics (eval mode reads them)"""
    st = O.make_state(fl, seed=0, dtype=dtype)
    for k in st:
        if k.endswith('running_var'):
            st[k] = st[k] * 1.7
        if k.endswith('running_mean'):
            st[k] = st[k] + 0.05
    return st


def eval_noise(fl, dtype):
    B, D = fl.batch_size, fl.class_dim
    return dict(eps_imp=O.seeded_uniform('eval.eps_imp', 5, (K_IMP, B, D), -1.5, 1.5, dtype),
                eps_cg=O.seeded_uniform('eval.eps_cg', 6, (B, D), -1.5, 1.5, dtype))


def main():
    torch.set_num_threads(os.cpu_count())
    R = G.import_reference()
    from mimic.evaluation.eval_metrics.likelihood import calc_log_likelihood_batch
    dt = torch.float64
    fl = O.default_flags(**KW)
    fl.rec_weights = {m: 0.33 for m in fl.mods}
    state = eval_state(fl, dt)
    batch = O.make_batch(fl, seed=1, dtype=dt)
    noise = eval_noise(fl, dt)
    G.CURRENT.update(calls=0, noise=[(None, None)] * 8, eps_style=None, style_calls=0)
    exp = G.build_reference_model(R, fl, state)
    vae = exp.mm_vae
    vae.eval()
    cur = {'eps': None}
    R.U.reparameterize = lambda mu, logvar: cur['eps'] * torch.exp(0.5 * logvar) + mu
    import mimic.utils.utils as U2            # get_latent_samples / cond_generation read utils.reparameterize at call time
    assert U2 is R.U
    fx = dict(name='eval_tri', flags=dict(KW), k_imp=K_IMP, dtype='float64')
    ck = G.checksum
    with torch.no_grad():
        # (a) full-batch inference
        lat = vae.inference(OrderedDict(batch))
        fx['subsets'] = OrderedDict((k, (v[0].clone(), v[1].clone())) for k, v in lat['subsets'].items())
        fx['joint'] = (lat['joint'][0].clone(), lat['joint'][1].clone())
        # (b) partial-modality inputs at batch 1 / 2 (plotting.py:74-79,151-159)
        fx['partial'] = OrderedDict()
        for tag, rows, mods in (('PA@1', 1, ('PA',)), ('text@1', 1, ('text',)), ('Lateral_text@2', 2, ('Lateral', 'text'))):
            lp = vae.inference({m: batch[m][:rows] for m in mods}, num_samples=rows)
            fx['partial'][tag] = dict(rows=rows, mods=mods,
                                      subsets=OrderedDict((k, (v[0].clone(), v[1].clone())) for k, v in lp['subsets'].items()),
                                      joint=(lp['joint'][0].clone(), lp['joint'][1].clone()))
        # (c) importance-sampled likelihood for two conditioning subsets
        fx['lhood'] = OrderedDict()
        cur['eps'] = noise['eps_imp']
        for s_key in ('PA', 'Lateral_PA_text'):
            ll = calc_log_likelihood_batch(exp, lat, s_key, exp.subsets[s_key], dict(batch), num_imp_samples=K_IMP)
            mu, lv = lat['subsets'][s_key]
            z = (noise['eps_imp'] * torch.exp(0.5 * lv.unsqueeze(0)) + mu.unsqueeze(0)).view(K_IMP * fl.batch_size, -1)
            gen = vae.generate_sufficient_statistics_from_latents({'content': z, 'style': {m: None for m in fl.mods}})
            rows = OrderedDict()
            for m in fl.mods:
                x = batch[m]
                xr = x.unsqueeze(0).repeat(K_IMP, *([1] * x.dim())).view(K_IMP * fl.batch_size, *x.shape[1:])
                rows[m] = gen[m].log_prob(xr).view(K_IMP * fl.batch_size, -1).sum(dim=1).clone()
            fx['lhood'][s_key] = dict(ll=OrderedDict((k, float(v)) for k, v in ll.items()), logp_rows=rows,
                                      mean=OrderedDict((m, ck('lh.%s.%s' % (s_key, m), gen[m].mean)) for m in fl.mods))
        # (d) cond_generation
        cur['eps'] = noise['eps_cg']
        cg = vae.cond_generation({k: lat['subsets'][k] for k in ('PA', 'Lateral_PA_text')}, num_samples=fl.batch_size)
        fx['cond_gen'] = OrderedDict((k, OrderedDict((m, ck('cg.%s.%s' % (k, m), t)) for m, t in v.items())) for k, v in cg.items())
        fx['cond_gen_pa_full'] = cg['Lateral_PA_text']['PA'].clone()            # one complete image batch
        # (e) generate: z ~ torch.randn on the CPU generator, then .to(device) (VAEtrimodalMimic.py:127-135)
        # (the fp64 reference model cannot take generate()'s fp32 randn directly: the same three statements, cast)
        torch.manual_seed(11)
        z_class = torch.randn(3, fl.class_dim).to(dt)
        gen3 = vae.generate_from_latents({'content': z_class, 'style': vae.get_random_styles(3)})
        fx['generate'] = dict(seed=11, n=3, out=OrderedDict((m, ck('gen.' + m, t)) for m, t in gen3.items()))
        # ---- the oracle on the same inputs (pins its eval / decode path) ----
        st = OrderedDict(state)
        ores = O.forward(st, batch, fl, None, None, train=False)
        worst = 0.0
        for k, (mu, lv) in fx['subsets'].items():
            worst = max(worst, float((ores['latents']['subsets'][k][0] - mu).abs().max()),
                        float((ores['latents']['subsets'][k][1] - lv).abs().max()))
        for tag, p in fx['partial'].items():
            b = {m: batch[m][:p['rows']] for m in p['mods']}
            o = O.forward(st, b, fl, None, None, train=False, present=list(p['mods']))
            assert list(o['latents']['subsets'].keys()) == list(p['subsets'].keys()), (tag, list(o['latents']['subsets']))
            for k, (mu, lv) in p['subsets'].items():
                worst = max(worst, float((o['latents']['subsets'][k][0] - mu).abs().max()))
            worst = max(worst, float((o['latents']['joint'][0] - p['joint'][0]).abs().max()))
        for s_key, d in fx['lhood'].items():
            mu, lv = fx['subsets'][s_key]
            z = (noise['eps_imp'] * torch.exp(0.5 * lv.unsqueeze(0)) + mu.unsqueeze(0)).view(K_IMP * fl.batch_size, -1)
            dec = O.decode(st, fl, z)
            for m in fl.mods:
                x = batch[m]
                xr = x.unsqueeze(0).repeat(K_IMP, *([1] * x.dim())).view(K_IMP * fl.batch_size, *x.shape[1:])
                lp = O.log_prob_rows(m, dec[m], xr)
                worst = max(worst, float(((lp - d['logp_rows'][m]) / d['logp_rows'][m].abs().max()).abs().max()))
            ll = O.importance_likelihoods(fl, K_IMP, dec, batch, z, mu, lv)
            for k, v in d['ll'].items():
                worst = max(worst, abs(float(ll[k]) - v) / abs(v))
        z = noise['eps_cg'] * torch.exp(0.5 * fx['subsets']['Lateral_PA_text'][1]) + fx['subsets']['Lateral_PA_text'][0]
        worst = max(worst, float((O.decode(st, fl, z)['PA'] - fx['cond_gen_pa_full']).abs().max()))
    print('eval_tri: oracle vs reference worst abs/rel err %.2e' % worst)
    assert worst < 1e-9, worst
    out = os.path.join(os.path.dirname(HERE), 'tests', 'golden', 'eval_tri.pt')
    torch.save(fx, out)
    print('wrote', out, os.path.getsize(out), 'bytes')


if __name__ == '__main__':
    main()
