"""Shared harness for the GPU parity tests: build the product model from the oracle's deterministic
state, inject the oracle's dropout masks / eps (converted to the product's layouts), run one step on
both and compare."""
import os
from collections import OrderedDict

import torch

from oracle import mopoe_oracle as O


def oracle_flags(**kw):
    fl = O.default_flags(**kw)
    if 'rec_weights' not in kw:
        fl.rec_weights = {m: 0.33 for m in fl.mods}
    return fl


def product_flags(ofl, compute_dtype, device='cuda'):
    import mopoe_mimic_b200 as P
    return P.default_flags(device=torch.device(device), batch_size=ofl.batch_size, class_dim=ofl.class_dim,
                           img_size=ofl.img_size, DIM_img=ofl.DIM_img, DIM_text=ofl.DIM_text,
                           len_sequence=ofl.len_sequence, num_features=ofl.num_features, method=ofl.method,
                           mods=tuple(ofl.mods), beta=ofl.beta, compute_dtype=compute_dtype,
                           text_encoding=getattr(ofl, 'text_encoding', 'char'), vocab_size=getattr(ofl, 'vocab_size', 0),
                           factorized_representation=O.factorized(ofl), style_pa_dim=O.style_dim(ofl, 'PA'),
                           style_lat_dim=O.style_dim(ofl, 'Lateral'), style_text_dim=O.style_dim(ofl, 'text'))


def masks_to_product(masks, device):
    """oracle keep-masks (float {0,1}; 2-D [B,C,1,1], 1-D [B,C,L]) -> uint8 [B,C] / [B,L,C] on device"""
    out = {}
    for k, m in masks.items():
        if m.dim() == 4:
            out[k] = m.reshape(m.shape[0], m.shape[1]).to(torch.uint8).contiguous().to(device)
        else:
            out[k] = m.permute(0, 2, 1).to(torch.uint8).contiguous().to(device)
    return out


def make_case(kw, actual_batch=None, dtype=torch.float32, seeds=(0, 1, 2)):
    ofl = oracle_flags(**kw)
    B = actual_batch or ofl.batch_size
    state = O.make_state(ofl, seeds[0], dtype)
    batch = O.make_batch(ofl, seeds[1], dtype, B)
    noise = [O.make_noise(ofl, seeds[2] + i, dtype, B) for i in range(1 + len(ofl.mods))]
    ofl.eps_style = O.make_style_noise(ofl, seeds[2], dtype, B)        # None unless the case is factorized
    # poe: every unimodal pass reparameterises its own style latent
    ofl.eps_style_uni = ({m: O.make_style_noise(ofl, seeds[2] + 1 + i, dtype, B) for i, m in enumerate(ofl.mods)}
                         if ofl.eps_style is not None else None)
    return ofl, state, batch, noise


def run_oracle(ofl, state, batch, noise):
    st = OrderedDict((k, v.clone()) for k, v in state.items())
    uni = {m: noise[1 + i] for i, m in enumerate(ofl.mods)}
    return O.step_with_grads(st, batch, ofl, noise[0][0], noise[0][1], uni_masks=uni, eps_style=getattr(ofl, 'eps_style', None),
                             uni_eps_style=getattr(ofl, 'eps_style_uni', None))


def run_product(ofl, state, batch, noise, compute_dtype):
    import mopoe_mimic_b200 as P
    fl = product_flags(ofl, compute_dtype)
    exp = P.Experiment(fl)
    vae = exp.mm_vae
    vae.load_state_dict({k: v.float() for k, v in state.items()})
    exp.set_optimizer()        # flattens params/grads
    vae.train()
    dev = fl.device
    es = getattr(ofl, 'eps_style', None)
    es = {m: v.float().to(dev) for m, v in es.items()} if es is not None else None
    esu = getattr(ofl, 'eps_style_uni', None)
    per_pass = [es] + [({k: v.float().to(dev) for k, v in esu[m].items()} if esu is not None else es) for m in ofl.mods]
    vae.rt.schedule = [(masks_to_product(m, dev), e.float().to(dev), per_pass[i]) for i, (m, e) in enumerate(noise)]
    b = OrderedDict((k, v.float().to(dev)) for k, v in batch.items())
    out = P.basic_routine_epoch(exp, (b, None))
    exp.optimizer.zero_grad()
    out['total_loss'].backward()
    torch.cuda.synchronize()
    grads = OrderedDict((k, p.grad.detach().cpu()) for k, p in vae.named_parameters())
    return exp, out, grads


def smooth_grads(ofl, state, batch, noise, compute_dtype):
    """Gradients of a SMOOTH surrogate loss (mean-square of every reconstruction + the joint divergence) on both
    sides.  The ELBO's Laplace term is piecewise linear: fp32 rounding flips a handful of sign(x - loc) decisions
    per image, which moves decoder gradients by ~1e-3..1e-2 in ANY fp32 implementation; the surrogate exercises the
    identical conv / BN / dropout / fusion backward kernels without that chaos, so it can be compared tightly."""
    import mopoe_mimic_b200 as P

    def oracle_grads(dt):
        st = OrderedDict((k, (v.clone().to(dt) if v.is_floating_point() else v.clone())) for k, v in state.items())
        params = {k: v for k, v in st.items() if v.is_floating_point() and 'running_' not in k}
        for v in params.values():
            v.requires_grad_(True)
        b = OrderedDict((k, v.to(dt)) for k, v in batch.items())
        masks = {k: v.to(dt) for k, v in noise[0][0].items()}
        es = getattr(ofl, 'eps_style', None)
        res = O.forward(st, b, ofl, masks, noise[0][1].to(dt), True,
                        eps_style={m: v.to(dt) for m, v in es.items()} if es is not None else None)
        loss = sum((r ** 2).mean() for r in res['rec'].values()) * 100.0 + 5.0 * res['joint_divergence']
        loss.backward()
        # (parameters the step never runs — the word encoder's resblock_7/8 at len_sequence <= 500 — have no gradient)
        return loss, OrderedDict((k, v.grad.detach().clone()) for k, v in params.items() if v.grad is not None)
    loss_o, g_o = oracle_grads(torch.float32)
    _, g_o64 = oracle_grads(torch.float64)        # truth: the same fp32-rounded inputs evaluated in fp64
    smooth_grads.truth = g_o64
    fl = product_flags(ofl, compute_dtype)
    exp = P.Experiment(fl)
    vae = exp.mm_vae
    vae.load_state_dict({k: v.detach().float() for k, v in state.items()})
    exp.set_optimizer()
    vae.train()
    dev = fl.device
    es = getattr(ofl, 'eps_style', None)
    es = {m: v.float().to(dev) for m, v in es.items()} if es is not None else None
    vae.rt.schedule = [(masks_to_product(noise[0][0], dev), noise[0][1].float().to(dev), es)]
    out = vae(OrderedDict((k, v.float().to(dev)) for k, v in batch.items()))
    recs = []
    for m, r in out['rec'].items():
        recs.append(torch.log_softmax(r._scores, -1) if m == 'text' else r.loc)
    loss_p = sum((r ** 2).mean() for r in recs) * 100.0 + 5.0 * out['joint_divergence']
    exp.optimizer.zero_grad()
    loss_p.backward()
    torch.cuda.synchronize()
    g_p = OrderedDict((k, p.grad.detach().cpu()) for k, p in vae.named_parameters())
    return float(loss_o), float(loss_p), g_o, g_p


def rel_err(a, b, floor=0.0):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + floor + 1e-300))


def compare_step(orc, out, grads, state_after=None):
    """returns dict of relative errors (max-abs / max-abs) per quantity"""
    errs = OrderedDict()
    errs['total_loss'] = rel_err(out['total_loss'], orc['total_loss'])
    errs['joint_div'] = rel_err(out['results']['joint_divergence'], orc['results']['joint_divergence'])
    for k, v in orc['klds'].items():
        errs['kld.' + k] = rel_err(out['klds'][k], v)
    for k, v in orc['log_probs'].items():
        errs['logp.' + k] = rel_err(out['log_probs'][k], v)
    olat, plat = orc['results']['latents'], out['results']['latents']
    for m, (mu, lv) in olat['modalities'].items():
        errs['enc_mu.' + m] = rel_err(plat['modalities'][m][0], mu)
        errs['enc_lv.' + m] = rel_err(plat['modalities'][m][1], lv)
    for k, (mu, lv) in olat['subsets'].items():
        errs['sub_mu.' + k] = rel_err(plat['subsets'][k][0], mu)
        errs['sub_lv.' + k] = rel_err(plat['subsets'][k][1], lv)
    errs['joint_mu'] = rel_err(plat['joint'][0], olat['joint'][0])
    for m, (smu, slv) in olat.get('styles', {}).items():
        errs['style_mu.' + m] = rel_err(plat['modalities'][m + '_style'][0], smu)
        errs['style_lv.' + m] = rel_err(plat['modalities'][m + '_style'][1], slv)
    errs['mus'] = rel_err(plat['mus'], olat['mus'])
    errs['ind_divs'] = rel_err(out['results']['individual_divs'], orc['results']['individual_divs'])
    if orc['results'].get('dyn_prior') is not None:
        errs['dyn_mu'] = rel_err(out['results']['dyn_prior'][0], orc['results']['dyn_prior'][0])
        errs['dyn_lv'] = rel_err(out['results']['dyn_prior'][1], orc['results']['dyn_prior'][1])
    errs['z'] = rel_err(plat['_z'], orc['results']['z'])
    for m, r in orc['results']['rec'].items():
        pr = out['results']['rec'][m]
        errs['rec.' + m] = rel_err(pr.loc if m != 'text' else pr.logits, r)
    gscale = max(float(g.abs().max()) for g in orc['grads'].values())
    worst = (0.0, None)
    for k, g in orc['grads'].items():
        e = rel_err(grads[k], g, floor=1e-4 * gscale)
        errs['grad.' + k] = e
        if e > worst[0]:
            worst = (e, k)
    errs['_worst_grad'] = worst
    return errs


# ---- per-tensor gradient comparison (bf16 parity) --------------------------------------------------------------------------
ZERO_GRAD_SUFFIXES = ('downsample.0.bias', 'upsample.0.bias')      # conv bias feeding a train-mode BatchNorm: analytically 0


def grad_table(g_test, g_ref):
    """per parameter tensor: relative L2 error, max-abs error / max-abs value, cosine, and the tensor's share of the whole
    gradient's L2 norm (tensors that carry ~nothing of the gradient are rounding noise in ANY implementation)"""
    tot = sum(float(r.double().norm()) ** 2 for r in g_ref.values()) ** 0.5
    rows = []
    for k, r in g_ref.items():
        t, r = g_test[k].double().reshape(-1), r.double().reshape(-1)
        nr, nt = float(r.norm()), float(t.norm())
        rows.append(dict(name=k, numel=r.numel(), share=nr / max(tot, 1e-300),
                         rel_l2=float((t - r).norm()) / max(nr, 1e-300),
                         max_rel=float((t - r).abs().max()) / max(float(r.abs().max()), 1e-300),
                         cos=float(t @ r) / max(nt * nr, 1e-300)))
    return rows


def device_noise(ofl, B, seed, device='cuda'):
    """dropout keep-masks in the PRODUCT's layouts (uint8 [B,C] for Dropout2d, [B,L,C] for Dropout) and eps, generated on
    the device — for comparisons between two product runs at sizes where the CPU mask generator is too slow"""
    g = torch.Generator(device=device).manual_seed(seed)
    masks = {}
    for name, shape in O.dropout_sites(ofl, B).items():
        if len(shape) == 4:
            masks[name] = (torch.rand(shape[0], shape[1], device=device, generator=g) < 0.5).to(torch.uint8)
        else:
            masks[name] = (torch.rand(shape[0], shape[2], shape[1], device=device, generator=g) < 0.5).to(torch.uint8)
    eps = torch.randn(B, ofl.class_dim, device=device, generator=g)
    return masks, eps


def run_product_device_noise(ofl, state, batch_dev, noise_dev, compute_dtype, loss='elbo'):
    """one product step (forward + loss + backward) with device-resident inputs / injected noise -> (out, grads on device)"""
    import mopoe_mimic_b200 as P
    fl = product_flags(ofl, compute_dtype)
    exp = P.Experiment(fl)
    vae = exp.mm_vae
    vae.load_state_dict({k: v.float() for k, v in state.items()})
    exp.set_optimizer()
    vae.train()
    vae.rt.schedule = [(dict(noise_dev[0]), noise_dev[1], None)]
    b = OrderedDict((k, v.clone()) for k, v in batch_dev.items())
    if loss == 'elbo':
        out = P.basic_routine_epoch(exp, (b, None))
        lossv = out['total_loss']
    else:
        res = vae(b)
        recs = [torch.log_softmax(r._scores, -1) if m == 'text' else r.loc for m, r in res['rec'].items()]
        lossv = sum((r ** 2).mean() for r in recs) * 100.0 + 5.0 * res['joint_divergence']
        out = {'total_loss': lossv, 'results': res}
    exp.optimizer.zero_grad()
    lossv.backward()
    vae.join_branches()
    torch.cuda.synchronize()
    grads = OrderedDict((k, p.grad.detach().clone()) for k, p in vae.named_parameters())
    return out, grads
