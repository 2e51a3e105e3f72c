"""Generate tests/golden/*.pt by running the UNMODIFIED reference (imported from a writable copy of
/root/reference) on the oracle's deterministic synthetic state / inputs / noise.  Build-container
only: /root/reference does not exist on the GPU box, so nothing at test or bench time imports this.

    python oracle/gen_golden.py            # regenerates every fixture (a few minutes of CPU)

Each fixture holds the case's flags + seeds and the reference's outputs in float64: scalars, latent
tensors, and for large tensors (gradients, BN buffers, reconstructions) a checksum triple
(sum, l2, 8 probe values at seeded positions).  The same script asserts that oracle/mopoe_oracle.py
reproduces the reference to ~1e-10 (fp64), which is what pins the oracle.
"""
import argparse
import os
import shutil
import sys
import time
from collections import OrderedDict
from types import SimpleNamespace

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import mopoe_oracle as O  # noqa: E402

REF_SRC = '/root/reference/mimic'
REF_COPY = '/tmp/mopoe_refcopy'


def import_reference():
    """Copy because mimic/logger/logger.py:19-20 mkdirs at import time (read-only tree)."""
    if not os.path.isdir(os.path.join(REF_COPY, 'mimic')):
        os.makedirs(REF_COPY, exist_ok=True)
        shutil.copytree(REF_SRC, os.path.join(REF_COPY, 'mimic'))
    sys.path.insert(0, os.path.join(REF_COPY, 'mimic'))   # losses.py:2-3 uses `from utils import utils`
    sys.path.insert(0, REF_COPY)
    import mimic.utils.utils as U
    from mimic.evaluation import losses
    from mimic.modalities.MimicLateral import MimicLateral
    from mimic.modalities.MimicPA import MimicPA
    from mimic.modalities.MimicText import MimicText
    from mimic.networks.ConvNetworksImgMimic import DecoderImg, EncoderImg
    from mimic.networks.ConvNetworksTextMimic import DecoderText, EncoderText
    from mimic.networks.VAEtrimodalMimic import VAEtrimodalMimic
    from mimic.utils.BaseExperiment import BaseExperiment
    from mimic.utils.BaseMMVae import BaseMMVae
    return SimpleNamespace(**locals())


CURRENT = {'masks': None, 'eps': None, 'calls': 0, 'noise': None, 'eps_style': None, 'style_calls': 0,
           'eps_style_list': None, 'eps_style_pass': None, 'style_mod': None}


def ref_flags(fl):
    m = fl.method
    return SimpleNamespace(
        device=torch.device('cpu'), batch_size=fl.batch_size, class_dim=fl.class_dim, img_size=fl.img_size,
        image_channels=fl.image_channels, DIM_img=fl.DIM_img, DIM_text=fl.DIM_text,
        text_encoding=getattr(fl, 'text_encoding', 'char'), vocab_size=getattr(fl, 'vocab_size', 0),
        len_sequence=fl.len_sequence, num_features=fl.num_features, alphabet='x' * fl.num_features,
        feature_extractor_img='resnet', factorized_representation=O.factorized(fl),
        style_pa_dim=O.style_dim(fl, 'PA'), style_lat_dim=O.style_dim(fl, 'Lateral'),
        style_text_dim=O.style_dim(fl, 'text'), modality_moe=(m == 'moe'), modality_jsd=(m == 'jsd'), modality_poe=(m == 'poe'),
        joint_elbo=(m == 'joint_elbo'), poe_unimodal_elbos=True, alpha_modalities=list(fl.alpha_modalities),
        beta=fl.beta, beta_style=fl.beta_style, beta_content=fl.beta_content, dataset='testing',
        distributed=False, world_size=1, text_gen_lastlayer='softmax')


def build_reference_model(R, fl, state):
    rf = ref_flags(fl)
    mods = OrderedDict()
    for m in fl.mods:            # experiment.py:80-92: dict order PA, Lateral, text
        if m == 'PA':
            mods[m] = R.MimicPA(R.EncoderImg(rf, rf.style_pa_dim), R.DecoderImg(rf, rf.style_pa_dim), rf)
        elif m == 'Lateral':
            mods[m] = R.MimicLateral(R.EncoderImg(rf, rf.style_lat_dim), R.DecoderImg(rf, rf.style_lat_dim), rf)
        else:
            mods[m] = R.MimicText(R.EncoderText(rf, rf.style_text_dim), R.DecoderText(rf, rf.style_text_dim),
                                  rf.len_sequence, None, None, rf)
    exp = SimpleNamespace(flags=rf, modalities=mods)
    exp.subsets = R.BaseExperiment.set_subsets(exp)
    exp.rec_weights = dict(fl.rec_weights)
    exp.style_weights = dict(fl.style_weights)

    class GenericMMVae(R.BaseMMVae):
        """~30-line shim: VAEtrimodalMimic.forward:31-62 / encode:64-93 generalised to any
        modality set and tolerant of absent modalities in the decode loop (needed by calc_poe_loss)."""

        def __init__(self, flags, modalities, subsets):
            super().__init__(flags, modalities, subsets)
            for m, mod in modalities.items():      # registration order of VAEtrimodalMimic.__init__:15-20
                setattr(self, O.ENC_NAME[m], mod.encoder)
            for m, mod in modalities.items():
                setattr(self, O.DEC_NAME[m], mod.decoder)

        def encode(self, input_batch):
            lat = {}
            for m in self.modalities:           # VAEtrimodalMimic.encode:64-93: content first, style after
                if m in input_batch:
                    out = getattr(self, O.ENC_NAME[m])(input_batch[m])
                    lat[m] = list(out[:2])
                    if len(out) == 4:
                        lat[m + '_style'] = list(out[2:])
            return lat

        def forward(self, input_batch):
            noise = CURRENT['noise'][CURRENT['calls']]
            CURRENT['masks'], CURRENT['eps'] = noise
            if CURRENT['eps_style_list'] is not None:
                CURRENT['eps_style_pass'] = CURRENT['eps_style_list'][CURRENT['calls']]
            CURRENT['calls'] += 1
            latents = self.inference(input_batch)
            results = {'latents': latents}
            div = self.calc_joint_divergence(latents['mus'], latents['logvars'], latents['weights'])
            results['group_distr'] = latents['joint']
            z = R.U.reparameterize(latents['joint'][0], latents['joint'][1])
            results.update(div)
            rec = {}
            for m, mod in self.modalities.items():
                if m in input_batch:
                    dec = getattr(self, O.DEC_NAME[m])
                    s_emb = None
                    if self.flags.factorized_representation:       # VAEtrimodalMimic.forward:49-51
                        s_mu, s_logvar = latents['modalities'][m + '_style']
                        CURRENT['style_mod'] = m
                        s_emb = R.U.reparameterize(mu=s_mu, logvar=s_logvar)
                    if m == 'text':
                        rec[m] = mod.likelihood(logits=dec(s_emb, z)[0])
                    else:
                        rec[m] = mod.likelihood(*dec(s_emb, z))
            results['rec'] = rec
            return results

        def get_random_styles(self, n): return {m: None for m in self.modalities}
        def get_random_style_dists(self, n): return {}
        def generate_sufficient_statistics_from_latents(self, latents): raise NotImplementedError
        def save_networks(self): pass

    trimodal_native = tuple(fl.mods) == ('PA', 'Lateral', 'text') and fl.method != 'poe'
    if trimodal_native:          # pin the shipped class wherever it runs
        vae = R.VAEtrimodalMimic(rf, mods, exp.subsets)
        orig_forward = vae.forward

        def fwd(batch):
            CURRENT['masks'], CURRENT['eps'] = CURRENT['noise'][CURRENT['calls']]
            CURRENT['calls'] += 1
            return orig_forward(batch)
        vae.forward = fwd
    else:
        vae = GenericMMVae(rf, mods, exp.subsets)
    ref_sd = vae.state_dict()
    assert list(ref_sd.keys()) == list(state.keys()), 'state_dict key order differs from oracle.param_spec'
    for k, v in ref_sd.items():
        assert tuple(v.shape) == tuple(state[k].shape), (k, v.shape, state[k].shape)
    vae = vae.to(next(iter(state.values())).dtype)
    vae.load_state_dict(state)
    # inject dropout masks / eps: the RNG stream itself is not reproduced
    for name, mod in vae.named_modules():
        if isinstance(mod, (torch.nn.Dropout, torch.nn.Dropout2d)):
            mod.forward = (lambda x, n=name, md=mod: x * CURRENT['masks'][n] * 2.0 if md.training else x)
    def reparameterize(mu, logvar):
        if CURRENT['style_mod'] is not None:        # the shim's style sample of modality `style_mod` in the current pass
            m, CURRENT['style_mod'] = CURRENT['style_mod'], None
            return CURRENT['eps_style_pass'][m] * torch.exp(0.5 * logvar) + mu
        # VAEtrimodalMimic.forward: the content sample first, then one style sample per modality in self.modalities order
        if CURRENT['eps_style'] is not None and mu.shape[1] != fl.class_dim:
            m = list(fl.mods)[CURRENT['style_calls'] % len(fl.mods)]
            CURRENT['style_calls'] += 1
            return CURRENT['eps_style'][m] * torch.exp(0.5 * logvar) + mu
        return CURRENT['eps'] * torch.exp(0.5 * logvar) + mu
    R.U.reparameterize = reparameterize
    exp.mm_vae = vae
    return exp


def reference_step(R, exp, batch):
    """run_epochs.basic_routine_epoch:52-96 restated (run_epochs.py itself needs termcolor)."""
    fl = exp.flags
    results = exp.mm_vae(batch)
    log_probs, weighted = R.losses.calc_log_probs(exp, results, (batch, None))
    klds = R.losses.calc_klds(exp, results)
    if fl.modality_moe or fl.joint_elbo or fl.modality_jsd:
        klds_style = R.losses.calc_klds_style(exp, results) if fl.factorized_representation else None
        total = R.losses.calc_joint_elbo_loss(exp, klds_style, results['joint_divergence'], fl.beta_style,
                                              fl.beta_content, weighted, fl.beta)
    else:
        klds_style = R.losses.calc_klds_style(exp, results) if fl.factorized_representation else None
        total = R.losses.calc_poe_loss(exp, exp.modalities, results['joint_divergence'], klds, klds_style, batch,
                                       exp.mm_vae, log_probs)
    return dict(results=results, log_probs=log_probs, klds=klds, total_loss=total)


def checksum(name, t, nprobe=8):
    t = t.detach().double().reshape(-1)
    n = t.numel()
    pos = (O.seeded_uniform('probe:' + name, 7, (min(nprobe, n),)) * n).long().clamp_(0, n - 1)
    return dict(sum=float(t.sum()), l2=float(t.norm()), pos=pos, val=t[pos].clone(), numel=n)


CASES = OrderedDict([
    # config 1 of BASELINE.json: the reference's CPU-runnable case
    ('cfg1_tri_128_b16_joint', dict(batch_size=16)),
    ('small_tri_joint', dict(batch_size=8, DIM_img=16, DIM_text=16, class_dim=32)),
    ('small_tri_moe', dict(batch_size=8, DIM_img=16, DIM_text=16, class_dim=32, method='moe')),
    ('small_tri_poe', dict(batch_size=6, DIM_img=16, DIM_text=16, class_dim=32, method='poe')),
    ('small_patext_joint', dict(batch_size=8, DIM_img=16, DIM_text=16, class_dim=32, mods=('PA', 'text'))),
    ('small_patext_moe', dict(batch_size=8, DIM_img=16, DIM_text=16, class_dim=32, mods=('PA', 'text'),
                              method='moe')),
    ('small_patext_poe', dict(batch_size=5, DIM_img=16, DIM_text=16, class_dim=32, mods=('PA', 'text'),
                              method='poe')),
    ('small_tri_256_joint', dict(batch_size=4, DIM_img=8, DIM_text=8, class_dim=64, img_size=256)),
    ('small_tri_64_joint', dict(batch_size=4, DIM_img=8, DIM_text=8, class_dim=16, img_size=64)),
    ('small_tri_joint_ragged', dict(batch_size=8, DIM_img=16, DIM_text=16, class_dim=32, actual_batch=5)),
    ('small_tri_style', dict(batch_size=8, DIM_img=16, DIM_text=16, class_dim=32,
                            style_dims={'PA': 8, 'Lateral': 16, 'text': 24})),
    ('small_tri_word', dict(batch_size=8, DIM_img=16, DIM_text=16, class_dim=32, text_encoding='word', vocab_size=48,
                           len_sequence=128)),
    # combinations (pins for the oracle; GPU parity cases for them are next-round work)
    ('small_tri_jsd_style', dict(batch_size=8, DIM_img=16, DIM_text=16, class_dim=32, method='jsd',
                                style_dims={'PA': 8, 'Lateral': 8, 'text': 16})),
    ('small_tri_word_style_moe', dict(batch_size=8, DIM_img=16, DIM_text=16, class_dim=32, method='moe', text_encoding='word',
                                     vocab_size=48, len_sequence=128, style_dims={'PA': 8, 'Lateral': 8, 'text': 8})),
    ('small_tri_poe_style', dict(batch_size=6, DIM_img=16, DIM_text=16, class_dim=32, method='poe',
                                style_dims={'PA': 8, 'Lateral': 8, 'text': 16})),
    ('small_tri_jsd', dict(batch_size=8, DIM_img=16, DIM_text=16, class_dim=32, method='jsd')),
    ('small_patext_jsd', dict(batch_size=9, DIM_img=16, DIM_text=16, class_dim=32, mods=('PA', 'text'), method='jsd')),
])


def run_case(R, name, kw, outdir):
    kw = dict(kw)
    actual = kw.pop('actual_batch', None)
    fl = O.default_flags(**kw)
    if 'rec_weights' not in kw:
        fl.rec_weights = {m: 0.33 for m in fl.mods}
    B = actual or fl.batch_size
    dt = torch.float64
    state = O.make_state(fl, seed=0, dtype=dt)
    batch = O.make_batch(fl, seed=1, dtype=dt, batch=B)
    noise = [O.make_noise(fl, seed=2 + i, dtype=dt, batch=B) for i in range(1 + len(fl.mods))]
    t0 = time.time()
    eps_style = O.make_style_noise(fl, seed=2, dtype=dt, batch=B)
    uni_es = ({m: O.make_style_noise(fl, seed=3 + i, dtype=dt, batch=B) for i, m in enumerate(fl.mods)}
              if eps_style is not None else None)
    # ---- reference
    CURRENT.update(calls=0, noise=noise, eps_style=eps_style, style_calls=0, style_mod=None, eps_style_pass=eps_style,
                   eps_style_list=([eps_style] + [uni_es[m] for m in fl.mods]) if uni_es is not None else None)
    exp = build_reference_model(R, fl, state)
    exp.mm_vae.train()
    out = reference_step(R, exp, OrderedDict(batch))
    exp.mm_vae.zero_grad()
    out['total_loss'].backward()
    ref_grads = OrderedDict((k, p.grad) for k, p in exp.mm_vae.named_parameters())
    ref_sd = exp.mm_vae.state_dict()
    t1 = time.time()
    # ---- oracle on the same data
    st2 = OrderedDict((k, v.clone()) for k, v in state.items())
    uni = {m: noise[1 + i] for i, m in enumerate(fl.mods)}
    orc = O.step_with_grads_full(st2, batch, fl, noise[0][0], noise[0][1], uni_masks=uni, eps_style=eps_style,
                                 uni_eps_style=uni_es)
    t2 = time.time()

    def rel(a, b):
        a, b = a.detach().double(), b.detach().double()
        return float((a - b).abs().max() / (b.abs().max() + 1e-300))
    errs = {'total_loss': rel(orc['total_loss'], out['total_loss'])}
    for k in out['klds']:
        errs['kld.' + k] = rel(orc['klds'][k], out['klds'][k])
    for k in out['log_probs']:
        errs['logp.' + k] = rel(orc['log_probs'][k], out['log_probs'][k])
    # a conv bias feeding a train-mode BN has an analytically ZERO gradient (values ~1e-13 of rounding
    # noise): normalise by the tensor's scale floored at 1e-6 of the largest gradient magnitude
    gscale = max(float(g.abs().max()) for g in ref_grads.values() if g is not None)

    def relg(a, b):
        return float((a.detach() - b).abs().max() / (b.abs().max() + 1e-6 * gscale))
    gerrs = sorted(((relg(orc['grads'][k], g), k) for k, g in ref_grads.items() if g is not None), reverse=True)
    if gerrs[0][0] > 1e-9:
        print(gerrs[:12])
    gmax = gerrs[0][0]
    missing = [k for k, g in ref_grads.items() if g is None]
    errs['grads'] = gmax
    bnmax = max(rel(v, ref_sd[k]) for k, v in orc['results']['bn_updates'].items())
    errs['bn'] = bnmax
    worst = max(errs.values())
    print('%-28s ref %.1fs oracle %.1fs  loss %.6f  worst rel err %.2e (grads %.1e bn %.1e) nograd=%d'
          % (name, t1 - t0, t2 - t1, float(out['total_loss'].detach()), worst, gmax, bnmax, len(missing)))
    assert worst < 1e-9, errs
    # ---- fixture
    res = out['results']
    lat = res['latents']
    fx = dict(
        name=name, flags=kw, actual_batch=B, seeds=dict(state=0, batch=1, noise=2), dtype='float64',
        state_keys=[(k, tuple(v.shape)) for k, v in state.items()],
        subset_keys=list(exp.subsets.keys()),
        total_loss=float(out['total_loss'].detach()), joint_divergence=float(res['joint_divergence'].detach()),
        individual_divs=res['individual_divs'].detach().clone(),
        klds=OrderedDict((k, float(v)) for k, v in out['klds'].items()),
        log_probs=OrderedDict((k, float(v)) for k, v in out['log_probs'].items()),
        enc=OrderedDict((m, (lat['modalities'][m][0].detach().clone(), lat['modalities'][m][1].detach().clone()))
                        for m in fl.mods),
        subsets=OrderedDict((k, (v[0].detach().clone(), v[1].detach().clone())) for k, v in lat['subsets'].items()),
        mus=lat['mus'].detach().clone(), logvars=lat['logvars'].detach().clone(),
        joint=(lat['joint'][0].detach().clone(), lat['joint'][1].detach().clone()),
        rec=OrderedDict(), grads=OrderedDict(), bn=OrderedDict(), no_grad=missing, grad_scale=gscale)
    for m in fl.mods:
        r = res['rec'][m]
        fx['rec'][m] = checksum('rec.' + m, r.loc if m != 'text' else r.logits)
    for k, g in ref_grads.items():
        if g is not None:
            fx['grads'][k] = checksum(k, g)
    for k, v in ref_sd.items():
        if 'running_' in k:
            fx['bn'][k] = checksum(k, v)
    torch.save(fx, os.path.join(outdir, name + '.pt'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default=None)
    ap.add_argument('--out', default=os.path.join(os.path.dirname(HERE), 'tests', 'golden'))
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    R = import_reference()
    for name, kw in CASES.items():
        if a.only and a.only not in name:
            continue
        run_case(R, name, kw, a.out)


if __name__ == '__main__':
    main()
