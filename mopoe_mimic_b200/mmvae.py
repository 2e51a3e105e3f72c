"""Multimodal VAE with the reference's model API (utils/BaseMMVae.py:14-231, networks/VAEtrimodalMimic.py:12-163):
same constructor, same method names, same result-dict nesting — so run_epochs.basic_routine_epoch and the
evaluation callers drive it unchanged.  Inference runs ONE fused CUDA kernel for all
subsets / selection / reparameterisation / KLs instead of ~150 tiny torch ops.
"""
import os
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _lib as L
from .fusion import FusionFn, FusionPlan, JsdDivergenceFn, selection_ends, set_subsets, uniform_weights  # noqa: F401
from .modalities import CategoricalLikelihood, LaplaceLikelihood
from .networks import Runtime

ENC_NAME = {'PA': 'encoder_pa', 'Lateral': 'encoder_lat', 'text': 'encoder_text'}
DEC_NAME = {'PA': 'decoder_pa', 'Lateral': 'decoder_lat', 'text': 'decoder_text'}
LHOOD_NAME = {'PA': 'lhood_pa', 'Lateral': 'lhood_lat', 'text': 'lhood_text'}


def reweight_weights(w):
    return w / w.sum()


def method_of(flags):
    if getattr(flags, 'modality_moe', False):
        return 'moe'
    if getattr(flags, 'modality_jsd', False):
        return 'jsd'
    if getattr(flags, 'modality_poe', False):
        return 'poe'
    if getattr(flags, 'joint_elbo', False):
        return 'joint_elbo'
    raise ValueError('flags select no fusion method (modality_moe / modality_poe / joint_elbo)')


class BaseMMVae(nn.Module):
    def __init__(self, flags, modalities, subsets):
        super().__init__()
        self.num_modalities = len(modalities.keys())
        self.flags = flags
        self.modalities = modalities
        self.subsets = subsets
        object.__setattr__(self, 'rt', Runtime(flags))
        self._plans = {}
        self.set_fusion_functions()

    # ---- method selection (BaseMMVae.py:51-69) ----------------------------------------------------------------
    def set_fusion_functions(self):
        w = torch.tensor(list(self.flags.alpha_modalities), dtype=torch.float32)
        self.weights = reweight_weights(w)
        self.method = method_of(self.flags)
        self.calc_joint_divergence = self.divergence_static_prior
        if self.method in ('moe', 'jsd'):
            self.modality_fusion, self.fusion_condition = self.moe_fusion, self.fusion_condition_moe
            if self.method == 'jsd':
                self.calc_joint_divergence = self.divergence_dynamic_prior
        elif self.method == 'poe':
            self.modality_fusion, self.fusion_condition = self.poe_fusion, self.fusion_condition_poe
        else:
            self.modality_fusion, self.fusion_condition = self.poe_fusion, self.fusion_condition_joint

    def fusion_condition_moe(self, subset, input_batch=None):
        return len(subset) == 1

    def fusion_condition_poe(self, subset, input_batch=None):
        return len(subset) == len(input_batch.keys())

    def fusion_condition_joint(self, subset, input_batch=None):
        return True

    # ---- engine helpers -------------------------------------------------------------------------------------------
    def _eng(self, device):
        return self.rt.eng(device)

    def _plan(self, present, B):
        key = (tuple(present), B)
        if key not in self._plans:
            keys = list(self.subsets.keys())
            members = [[m.name for m in self.subsets[k]] for k in keys]
            self._plans[key] = FusionPlan(list(self.modalities.keys()), present, keys, members, self.method, B,
                                          self.flags.class_dim, self.flags.batch_size)
        return self._plans[key]

    def _eps(self, B, device):
        if self.rt.injected_eps is not None:
            e = self.rt.injected_eps
            assert tuple(e.shape) == (B, self.flags.class_dim)
            return e
        return torch.randn(B, self.flags.class_dim, device=device)

    # ---- standalone fusion API (kept for callers; the hot path uses inference()) --------------------------------------
    def _kernel_rows(self, mus, logvars, fuse_mode, prior):
        """run the fusion kernel on m <= 4 stacked experts as one all-member subset"""
        m, B, D = mus.shape
        names = ['e%d' % i for i in range(m)]
        plan = FusionPlan(names, names, ['all'], [names], 'poe' if prior else 'joint_elbo', B, D, self.flags.batch_size)
        plan.cfg.fuse_mode = fuse_mode
        eng = self._eng(mus.device)
        out = FusionFn.apply(plan, eng, torch.zeros(B, D, device=mus.device), *[mus[i] for i in range(m)],
                             *[logvars[i] for i in range(m)])
        return out

    def poe_fusion(self, mus, logvars, weights=None):
        """Product of experts over stacked [m, B, D] (BaseMMVae.py:113-128 -> mm_div.poe)."""
        if mus.shape[0] > 4:
            raise NotImplementedError('standalone poe_fusion takes at most 4 experts')
        out = self._kernel_rows(mus, logvars, 0, self.method == 'poe')
        return [out[0][0], out[1][0]]

    def moe_fusion(self, mus, logvars, weights=None):
        """Positional mixture selection over stacked [S, B, D] (BaseMMVae.py:101-111 -> utils.py:55-77).
        Pure row-range gathering (no arithmetic): done with device copies."""
        if weights is None:
            weights = self.weights
        w = reweight_weights(torch.as_tensor(weights, dtype=torch.float32).cpu())
        ends = selection_ends(mus.shape[1], w)
        starts = [0] + ends[:-1]
        mu = torch.cat([mus[k, starts[k]:ends[k], :] for k in range(w.shape[0])])
        lv = torch.cat([logvars[k, starts[k]:ends[k], :] for k in range(w.shape[0])])
        return [mu, lv]

    def divergence_static_prior(self, mus, logvars, weights=None):
        """sum_k w_k KL(N(mu_k, var_k) || N(0, I)) / batch_size  (BaseMMVae.py:71-85, mm_div.py:90-110)."""
        if weights is None:
            weights = self.weights
        w = reweight_weights(weights.clone().float()).to(mus.device)
        klds = []
        for k0 in range(0, mus.shape[0], 4):        # identity "mixture" subsets -> per-row KLs from the kernel
            m = min(4, mus.shape[0] - k0)
            B, D = mus.shape[1:]
            names = ['e%d' % i for i in range(m)]
            plan = FusionPlan(names, names, names, [[n] for n in names], 'moe', B, D, self.flags.batch_size)
            out = FusionFn.apply(plan, self._eng(mus.device), torch.zeros(B, D, device=mus.device),
                                 *[mus[k0 + i] for i in range(m)], *[logvars[k0 + i] for i in range(m)])
            klds.append(out[5])
        klds = torch.cat(klds)
        return {'joint_divergence': (w * klds).sum(dim=0), 'individual_divs': klds, 'dyn_prior': None}

    def divergence_dynamic_prior(self, mus, logvars, weights=None):
        """alpha-JSD against the dynamic prior alpha_poe(weights, mus, logvars)  (BaseMMVae.py:87-99, mm_div.py:67-87).
        mus / logvars: [K, B, D] (jsd mode stacks the unimodal posteriors and the N(0,I) prior); the weights are used
        as given (no re-normalisation), exactly as the reference does."""
        if weights is None:
            weights = self.weights
        K = mus.shape[0]
        if K > 5:
            raise NotImplementedError('divergence_dynamic_prior takes at most 5 stacked experts')
        w_host = [float(x) for x in torch.as_tensor(weights, dtype=torch.float32).cpu()] if not getattr(
            weights, 'is_cuda', False) else self._weights_host(weights)
        kl, dyn_mu, dyn_lv = JsdDivergenceFn.apply(self._eng(mus.device), tuple(w_host), float(self.flags.batch_size),
                                                   *[mus[k] for k in range(K)], *[logvars[k] for k in range(K)])
        w_dev = torch.as_tensor(weights, dtype=torch.float32).to(mus.device)
        return {'joint_divergence': (w_dev * kl).sum(dim=0), 'individual_divs': kl, 'dyn_prior': [dyn_mu, dyn_lv]}

    def _weights_host(self, weights):
        """host copy of a (small) device weight vector, cached by identity: keeps the step free of D2H syncs"""
        cache = self.__dict__.setdefault('_w_host_cache', {})
        key = (weights.data_ptr(), weights._version)
        if key not in cache:
            cache[key] = [float(x) for x in weights.detach().float().cpu()]
        return cache[key]

    # ---- inference (BaseMMVae.py:139-196) ---------------------------------------------------------------------------
    def inference(self, input_batch, num_samples=None):
        enc_mods = self.encode(input_batch)
        present = [m for m in self.modalities if m in input_batch]
        first = enc_mods[present[0]][0]
        B = first.shape[0]
        plan = self._plan(present, B)
        eng = self._eng(first.device)
        eps = self._eps(B, first.device)
        mus_in = [enc_mods[m][0] for m in plan.mods]
        lvs_in = [enc_mods[m][1] for m in plan.mods]
        sub_mu, sub_lv, jmu, jlv, z, kl, nan_flag = FusionFn.apply(plan, eng, eps, *mus_in, *lvs_in)
        latents = {'modalities': enc_mods}
        distr_subsets = OrderedDict((k, [sub_mu[i], sub_lv[i]]) for i, k in enumerate(plan.keys))
        st = plan.stacked
        if self.method == 'jsd':             # unimodal posteriors + one zero row for the prior component
            uni = st[:-1]
            zrow = torch.zeros(1, B, self.flags.class_dim, device=first.device)
            latents['mus'] = torch.cat((sub_mu[uni[0]:uni[-1] + 1] if uni == list(range(uni[0], uni[-1] + 1)) else sub_mu[uni], zrow))
            latents['logvars'] = torch.cat((sub_lv[uni[0]:uni[-1] + 1] if uni == list(range(uni[0], uni[-1] + 1)) else sub_lv[uni], zrow))
        elif st == list(range(len(plan.keys))):
            latents['mus'], latents['logvars'] = sub_mu, sub_lv
        else:
            latents['mus'] = sub_mu[st[0]:st[-1] + 1] if st == list(range(st[0], st[-1] + 1)) else sub_mu[st]
            latents['logvars'] = sub_lv[st[0]:st[-1] + 1] if st == list(range(st[0], st[-1] + 1)) else sub_lv[st]
        if getattr(plan, 'weights_dev', None) is None or plan.weights_dev.device != first.device:
            plan.weights_dev = plan.weights.to(first.device)       # once per plan (keeps the step graph-capturable)
        latents['weights'] = plan.weights_dev
        latents['joint'] = [jmu, jlv]
        latents['subsets'] = distr_subsets
        # fused by-products (private): reparameterised sample, per-subset KLs, NaN flag
        latents['_z'] = z
        latents['_klds'] = OrderedDict((k, kl[i]) for i, k in enumerate(plan.keys))
        if self.method != 'jsd':
            latents['_kl_stacked'] = kl[st[0]:st[-1] + 1] if st == list(range(st[0], st[-1] + 1)) else kl[st]
        else:
            plan.weights_host = [float(x) for x in plan.weights]
        latents['_nan_flag'] = nan_flag
        return latents

    # ---- generation API (BaseMMVae.py:198-231) ----------------------------------------------------------------------
    def generate(self, num_samples=None):
        if num_samples is None:
            num_samples = self.flags.batch_size
        dev = next(self.parameters()).device
        # drawn from the CPU generator and moved, as the reference does (VAEtrimodalMimic.py:127-135): the same
        # torch.manual_seed gives the same samples on both implementations
        z_class = torch.randn(num_samples, self.flags.class_dim).to(dev)
        random_latents = {'content': z_class, 'style': self.get_random_styles(num_samples)}
        return self.generate_from_latents(random_latents)

    def generate_from_latents(self, latents):
        suff_stats = self.generate_sufficient_statistics_from_latents(latents)
        return {m_key: suff_stats[m_key].mean for m_key in latents['style'].keys()}

    def cond_generation(self, latent_distributions, num_samples=None):
        if num_samples is None:
            num_samples = self.flags.batch_size
        style_latents = self.get_random_styles(num_samples)
        cond_gen_samples = {}
        for key in latent_distributions.keys():
            mu, logvar = latent_distributions[key]
            content_rep = reparameterize(mu=mu, logvar=logvar)
            cond_gen_samples[key] = self.generate_from_latents({'content': content_rep, 'style': style_latents})
        return cond_gen_samples


def reparameterize(mu, logvar):
    """utils.reparameterize (utils/utils.py:45-48) for callers outside the fused path (cond_generation)."""
    std = logvar.mul(0.5).exp()
    return torch.randn_like(std).mul(std).add(mu)


class _GradReady(torch.autograd.Function):
    """identity on the latent sample fed to every decoder; its backward runs once ALL decoders have produced their input
    gradient — i.e. every decoder parameter gradient of the step is final (each decoder's backward kernels are queued
    on its stream before the kernel that produces dz) — and fires rt.on_decoders_done (the bucketed gradient exchange)"""

    @staticmethod
    def forward(ctx, z, rt):
        ctx.rt = rt
        return z.view_as(z)

    @staticmethod
    def backward(ctx, g):
        cb = getattr(ctx.rt, 'on_decoders_done', None)
        if cb is not None:
            cb()
        return g, None


class MMVaeMimic(BaseMMVae):
    """Any non-empty subset of {PA, Lateral, text}; VAEtrimodalMimic is the 3-modality instance."""

    def __init__(self, flags, modalities, subsets):
        super().__init__(flags, modalities, subsets)
        for m in modalities:                               # registration order of VAEtrimodalMimic.__init__:15-20
            setattr(self, ENC_NAME[m], modalities[m].encoder)
        for m in modalities:
            setattr(self, DEC_NAME[m], modalities[m].decoder)
        for m in modalities:
            setattr(self, LHOOD_NAME[m], modalities[m].likelihood)
        for m in modalities:
            for net, nm in ((modalities[m].encoder, ENC_NAME[m]), (modalities[m].decoder, DEC_NAME[m])):
                object.__setattr__(net, 'rt', self.rt)
                object.__setattr__(net, 'prefix', nm)
        self.to(flags.device)

    # ---- modality branches on their own CUDA streams ---------------------------------------------------------------
    # The encoders (and the decoders) of the modalities are independent until the fusion (the loss).  Their deep stages
    # are chains of tiny launches that leave most of the 148 SMs idle, so each branch runs on its own stream: the big
    # kernels of one branch fill the gaps of the others.  Autograd replays every backward node on its forward stream, so
    # the backward pass overlaps the same way; under CUDA-graph capture the branches become parallel graph branches.
    def _branches(self, names, fn):
        """run fn(name) for every modality name on ITS stream; returns {name: result}; joined on return.

        The stream is a fixed property of the modality (position in self.modalities: 0 -> the ambient stream, i -> side
        stream i-1), not of the call: poe's unimodal passes (losses.calc_poe_loss) therefore run on the stream that also ran
        the modality's branch of the joint pass, and — autograd replays backward nodes on their forward streams — the two
        passes' read-modify-write accumulations into the SAME flat gradient slots are ordered by the stream instead of
        racing as parallel branches of the captured graph."""
        import os
        cur = torch.cuda.current_stream()
        if self.num_modalities <= 1 or os.environ.get('MOPOE_BRANCH_STREAMS', '1') == '0':
            return {n: fn(n) for n in names}
        order = list(self.modalities.keys())
        side = self.__dict__.setdefault('_side_streams', [])
        while len(side) < len(order) - 1:
            side.append(torch.cuda.Stream())
        out = {}
        used = [order.index(n) for n in names]
        self.__dict__.setdefault('_side_used', set()).update(i for i in used if i > 0)
        for n, i in zip(names, used):
            if i == 0:
                continue
            st = side[i - 1]
            st.wait_stream(cur)                       # fork: everything enqueued so far (inputs, packed weights) is visible
            with torch.cuda.stream(st):
                out[n] = fn(n)
        if order[0] in names:
            out[order[0]] = fn(order[0])              # the first modality stays on the ambient stream
        for i in used:
            if i > 0:
                cur.wait_stream(side[i - 1])          # join
        return out

    def join_branches(self):
        """make the ambient stream wait for every branch stream (after backward: kernels of our autograd nodes write
        parameter gradients on the branch streams without an AccumulateGrad node the engine would sync on)"""
        cur = torch.cuda.current_stream()
        # only streams forked since the last join: waiting on an idle stream from a CAPTURING stream invalidates the capture
        side = self.__dict__.get('_side_streams', [])
        for i in sorted(self.__dict__.get('_side_used', ())):
            cur.wait_stream(side[i - 1])
        self.__dict__['_side_used'] = set()
        if self.rt.engine is not None:
            self.rt.engine.join_wgrad_sides()      # the weight-gradient side streams of the branches

    @staticmethod
    def _touch(stream, *tensors):
        """tensors allocated on one stream and consumed on another: tell the caching allocator"""
        for t in tensors:
            if torch.is_tensor(t) and t.is_cuda:
                t.record_stream(stream)

    def encode(self, input_batch):
        latents = {}
        present = [m for m in self.modalities if m in input_batch]
        cur = torch.cuda.current_stream()

        def run(m):
            L.require_cuda(input_batch[m])
            self._touch(torch.cuda.current_stream(), input_batch[m])
            out = getattr(self, ENC_NAME[m])(input_batch[m])
            self._touch(cur, *out)
            return out
        outs = self._branches(present, run)
        for m in self.modalities:
            if m in input_batch:
                out = outs[m]
                if len(out) == 4:                  # VAEtrimodalMimic.encode:64-93: content first, style after
                    latents[m + '_style'] = list(out[2:])
                latents[m] = list(out[:2])
            else:
                latents[m + '_style'] = [None, None]
                latents[m] = [None, None]
        return latents

    def _style_sample(self, m_key, s_mu, s_lv):
        """reparameterised style latent and its KL to N(0,I) / batch_size in ONE launch: the fusion kernel on a single
        expert (identity subset) yields z = eps * exp(logvar / 2) + mu and the KL reduction (utils.py:45-48, kl_div.py:8-16)"""
        B, Ds = s_mu.shape
        key = ('style', Ds, B)
        if key not in self._plans:
            self._plans[key] = FusionPlan(['e0'], ['e0'], ['e0'], [['e0']], 'moe', B, Ds, self.flags.batch_size)
        es = self.rt.injected_eps_style
        eps = es[m_key] if es is not None else torch.randn(B, Ds, device=s_mu.device)
        out = FusionFn.apply(self._plans[key], self._eng(s_mu.device), eps, s_mu, s_lv)
        return out[4], out[5][0]

    def _decode(self, m_key, z, z_style=None):
        eng = self._eng(z.device)
        dec = getattr(self, DEC_NAME[m_key])
        if m_key == 'text':
            return CategoricalLikelihood(scores=dec(z_style, z)[0], eng=eng)
        loc, scale = dec(z_style, z)
        return LaplaceLikelihood(loc, scale, eng=eng, scale_value=0.75)   # ConvNetworksImgMimic.py:54

    def forward(self, input_batch):
        """VAEtrimodalMimic.forward:31-62; absent modalities are skipped in the decode loop (the intended
        behaviour for calc_poe_loss — the shipped reference raises KeyError there)."""
        if self.rt.schedule:
            item = self.rt.schedule.pop(0)
            self.rt.injected_masks, self.rt.injected_eps = item[0], item[1]
            self.rt.injected_eps_style = item[2] if len(item) > 2 else None
        latents = self.inference(input_batch)
        results = {'latents': latents}
        if self.method == 'jsd':
            mus_st, lvs_st = latents['mus'], latents['logvars']
            K = mus_st.shape[0]
            w_host = [1.0 / K] * K       # latents['weights'] = (1 / float(K)) * ones(K), used un-normalised (:187, mm_div.py:86)
            kl, dyn_mu, dyn_lv = JsdDivergenceFn.apply(self._eng(mus_st.device), tuple(w_host), float(self.flags.batch_size),
                                                       *[mus_st[k] for k in range(K)], *[lvs_st[k] for k in range(K)])
            results['joint_divergence'] = (latents['weights'] * kl).sum(dim=0)
            results['individual_divs'] = kl
            results['dyn_prior'] = [dyn_mu, dyn_lv]
        else:
            w = reweight_weights(latents['weights'])
            klds = latents['_kl_stacked']
            results['joint_divergence'] = (w * klds).sum(dim=0)
            results['individual_divs'] = klds
            results['dyn_prior'] = None
        results['group_distr'] = latents['joint']
        class_embeddings = latents['_z']
        if (getattr(self.rt, 'on_decoders_done', None) is not None and torch.is_grad_enabled() and class_embeddings.requires_grad
                and len(input_batch) == self.num_modalities):
            class_embeddings = _GradReady.apply(class_embeddings, self.rt)
        results_rec = {}
        factorized = bool(getattr(self.flags, 'factorized_representation', False))
        if factorized:
            latents['_klds_style'] = {}
        cur = torch.cuda.current_stream()

        def run(m_key):
            s_emb, kl_s = None, None
            self._touch(torch.cuda.current_stream(), class_embeddings)
            if factorized:              # VAEtrimodalMimic.forward:49-51: s_emb = reparameterize(style mu, logvar)
                s_mu, s_lv = latents['modalities'][m_key + '_style']
                self._touch(torch.cuda.current_stream(), s_mu, s_lv)
                s_emb, kl_s = self._style_sample(m_key, s_mu, s_lv)
                self._touch(cur, kl_s)
            rec = self._decode(m_key, class_embeddings, s_emb)
            self._touch(cur, rec._scores if m_key == 'text' else rec.loc)
            return rec, kl_s
        decs = self._branches([m for m in self.modalities if m in input_batch and input_batch[m] is not None], run)
        for m_key in self.modalities:
            if m_key in decs:
                results_rec[m_key], kl_s = decs[m_key]
                if factorized:
                    latents['_klds_style'][m_key + '_style'] = kl_s
        results['rec'] = results_rec
        return results

    def get_random_styles(self, num_samples):
        """VAEtrimodalMimic.get_random_styles:95-110"""
        if not getattr(self.flags, 'factorized_representation', False):
            return {m: None for m in self.modalities}
        dims = {'PA': self.flags.style_pa_dim, 'Lateral': self.flags.style_lat_dim, 'text': self.flags.style_text_dim}
        dev = next(self.parameters()).device
        return {m: torch.randn(num_samples, dims[m]).to(dev) for m in self.modalities}       # CPU generator, as the reference

    def get_random_style_dists(self, num_samples):
        """VAEtrimodalMimic.get_random_style_dists:109-125: N(0, I) style posteriors [zeros(n, style_dim)] * 2 — the
        flags' style dims are used whether or not the representation is factorized, as in the reference"""
        dims = {'PA': self.flags.style_pa_dim, 'Lateral': self.flags.style_lat_dim, 'text': self.flags.style_text_dim}
        dev = next(self.parameters()).device
        return {m: [torch.zeros(num_samples, dims[m], device=dev), torch.zeros(num_samples, dims[m], device=dev)]
                for m in self.modalities}

    def generate_sufficient_statistics_from_latents(self, latents):
        content = latents['content']
        style = latents.get('style') or {}
        return {m: self._decode(m, content, style.get(m)) for m in self.modalities}

    def save_networks(self):
        names = {'PA': ('encoder_save_m1', 'decoder_save_m1'), 'Lateral': ('encoder_save_m2', 'decoder_save_m2'),
                 'text': ('encoder_save_m3', 'decoder_save_m3')}
        for m in self.modalities:
            e, d = names[m]
            torch.save(getattr(self, ENC_NAME[m]).state_dict(), os.path.join(self.flags.dir_checkpoints, getattr(self.flags, e)))
            torch.save(getattr(self, DEC_NAME[m]).state_dict(), os.path.join(self.flags.dir_checkpoints, getattr(self.flags, d)))

    # ---- flat parameter / gradient storage (one Adam launch, one all-reduce buffer) ---------------------------------
    def flatten_(self, alloc=None):
        """Move every parameter (and its .grad) into ONE flat fp32 buffer each.  alloc(total) -> (flat, flat_grads)
        lets the data-parallel exchange supply NVLink-shared (symmetric) buffers instead of plain device memory."""
        params = [p for p in self.parameters()]
        ALIGN = 64                      # elements: every tensor starts 256-B aligned (kernels use 16-B vector loads)
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        dev = params[0].device
        if alloc is not None:
            flat, flat_g = alloc(total)
            assert flat.numel() == total and flat_g.numel() == total and flat.dtype == flat_g.dtype == torch.float32
            flat.zero_()
            flat_g.zero_()
        else:
            flat = torch.zeros(total, dtype=torch.float32, device=dev)
            flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        object.__setattr__(self, 'flat_offsets', {name: off for (name, _), off in zip(self.named_parameters(), offs)})
        for p, off in zip(params, offs):
            n = p.numel()
            flat[off:off + n].copy_(p.data.reshape(-1))
            p.data = flat[off:off + n].view(p.shape)
            p.grad = flat_g[off:off + n].view(p.shape)
        object.__setattr__(self, 'flat_params', flat)
        object.__setattr__(self, 'flat_grads', flat_g)
        object.__setattr__(self, 'flat_numel', total)
        return flat, flat_g


class VAEtrimodalMimic(MMVaeMimic):
    def __init__(self, flags, modalities, subsets):
        if list(modalities.keys()) != ['PA', 'Lateral', 'text']:
            raise ValueError('VAEtrimodalMimic needs modalities PA, Lateral, text (in this order)')
        super().__init__(flags, modalities, subsets)


class VAETextMimic(MMVaeMimic):
    """The text-only VAE of the reference (networks/VAEtrimodalMimic.py:166-256): the same inference / forward / generate
    path with the single modality 'text' (one subset, so every fusion mode degenerates to that posterior)."""

    def __init__(self, flags, modalities, subsets):
        if list(modalities.keys()) != ['text']:
            raise ValueError('VAETextMimic needs exactly the modality text')
        super().__init__(flags, modalities, subsets)
