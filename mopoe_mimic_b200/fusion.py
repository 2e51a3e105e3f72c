"""Host side of the fused MoPoE kernel: subset enumeration, bit-exact mixture-selection boundaries,
and the autograd.Function around mopoe_fusion_fwd / mopoe_fusion_bwd.

Reference: utils/BaseExperiment.py:66-82 (set_subsets), utils/BaseMMVae.py:101-196 (fusion / inference),
utils/utils.py:55-77 (mixture_component_selection).
"""
import ctypes as C
from collections import OrderedDict
from itertools import chain, combinations

import torch

from . import _lib as L


def set_subsets(modalities):
    """BaseExperiment.set_subsets: powerset in itertools order over list(modalities); key =
    '_'.join(sorted(names)); value = modality objects sorted by name.  ASCII sort: Lateral < PA < text."""
    xs = list(modalities)
    subsets = OrderedDict()
    for mod_names in chain.from_iterable(combinations(xs, n) for n in range(len(xs) + 1)):
        subsets['_'.join(sorted(mod_names))] = [modalities[m] for m in sorted(mod_names)]
    return subsets


def selection_ends(num_samples, w):
    """Exclusive batch-row end of every mixture component — the integer logic of
    utils.mixture_component_selection (utils/utils.py:62-75) evaluated with the SAME fp32 tensor ops
    (floor(num_samples * w[k]) on an fp32 weight), so the index ranges are bit-exact.  Host-side integer
    plumbing: it decides which rows each component owns; no activation data passes through here."""
    w = torch.as_tensor(w, dtype=torch.float32, device='cpu')
    K = w.shape[0]
    ends = []
    for k in range(K):
        i_start = 0 if k == 0 else ends[k - 1]
        if k == K - 1:
            i_end = num_samples
        else:
            i_end = i_start + int(torch.floor(num_samples * w[k]))
        ends.append(int(i_end))
    ends[-1] = num_samples
    return ends


def uniform_weights(n):
    """(1 / float(n)) * torch.ones(n) followed by reweight_weights (w / w.sum()) — BaseMMVae.py:167-168, 187, 104."""
    w = (1 / float(n)) * torch.ones(n)
    return w / w.sum()


class FusionPlan:
    """Static description of one fusion call: which subsets exist, which are stacked into the joint mixture."""

    def __init__(self, mod_names, present, subset_keys, subset_members, method, B, D, norm):
        self.mods = [m for m in mod_names if m in present]             # experts fed to the kernel, model order
        idx = {m: i for i, m in enumerate(self.mods)}
        self.keys, members = [], []
        for key, mem in zip(subset_keys, subset_members):
            if key == '' or not all(m in idx for m in mem):
                continue
            self.keys.append(key)
            members.append(mem)                                         # sorted by name, as in the reference
        self.method = method
        if method in ('moe', 'jsd'):
            stacked = [i for i, mem in enumerate(members) if len(mem) == 1]
        elif method == 'poe':
            stacked = [i for i, mem in enumerate(members) if len(mem) == len(self.mods)]
        else:
            stacked = list(range(len(members)))
        if method == 'jsd':
            stacked = stacked + [-1]          # the prior N(0, I) is one more mixture component (BaseMMVae.py:180-186)
        self.stacked = stacked
        self.B, self.D = B, D
        cfg = L.FusionCfg()
        cfg.M, cfg.B, cfg.D = len(self.mods), B, D
        cfg.nsub, cfg.S = len(members), len(stacked)
        cfg.fuse_mode = 1 if method in ('moe', 'jsd') else 0
        cfg.prior_expert = 1 if method == 'poe' else 0
        cfg.kl_chunks = B
        cfg.norm = float(norm)
        for s, mem in enumerate(members):
            bits = 0
            for m in mem:
                bits |= 1 << idx[m]
            cfg.members[s] = bits
            # members in the reference's stacking order (sorted by name); in mixture mode the j-th stacked
            # member owns the batch rows [mem_end[j-1], mem_end[j])
            ends_sorted = selection_ends(B, uniform_weights(len(mem)))
            cfg.mem_cnt[s] = len(mem)
            for j, m in enumerate(mem):
                cfg.mem_idx[s][j] = idx[m]
                cfg.mem_end[s][j] = ends_sorted[j]
        ends = selection_ends(B, uniform_weights(len(stacked)))
        for j, s in enumerate(stacked):
            cfg.stacked[j] = s
            cfg.sel_end[j] = ends[j]
        self.cfg = cfg
        self.sel_end = ends
        self.weights = (1 / float(len(stacked))) * torch.ones(len(stacked))


def _ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return arr


class JsdDivergenceFn(torch.autograd.Function):
    """(mu_1..mu_K, lv_1..lv_K) -> (kl [K], dyn_mu [B,D], dyn_lv [B,D]): alpha-JSD terms against the dynamic prior
    (mopoe_jsd_divergence_fwd / _bwd; mm_div.calc_alphaJSD_modalities).  alpha: K python floats."""

    @staticmethod
    def forward(ctx, eng, alpha, norm, *mulv):
        K = len(mulv) // 2
        mus = [t.contiguous().float() for t in mulv[:K]]
        lvs = [t.contiguous().float() for t in mulv[K:]]
        B, D = mus[0].shape
        dev = mus[0].device
        kl = torch.empty(K, device=dev)
        dyn_mu, dyn_lv = torch.empty(B, D, device=dev), torch.empty(B, D, device=dev)
        al = (C.c_float * K)(*[float(a) for a in alpha])
        L.call('mopoe_jsd_divergence_fwd', K, B, D, _ptr_array(mus), _ptr_array(lvs), al, float(norm), L.ptr(dyn_mu),
               L.ptr(dyn_lv), L.ptr(kl), L.ptr(eng.ws64(K * B)), L.stream_ptr())
        ctx.alpha, ctx.norm, ctx.K = [float(a) for a in alpha], float(norm), K
        ctx.save_for_backward(*mus, *lvs)
        ctx.mark_non_differentiable(dyn_mu, dyn_lv)
        return kl, dyn_mu, dyn_lv

    @staticmethod
    def backward(ctx, d_kl, _dm, _dl):
        K = ctx.K
        mulv = ctx.saved_tensors
        mus, lvs = mulv[:K], mulv[K:]
        B, D = mus[0].shape
        dmu = [torch.empty_like(t) for t in mus]
        dlv = [torch.empty_like(t) for t in lvs]
        al = (C.c_float * K)(*ctx.alpha)
        L.call('mopoe_jsd_divergence_bwd', K, B, D, _ptr_array(mus), _ptr_array(lvs), al, ctx.norm,
               L.ptr(d_kl.contiguous().float()), _ptr_array(dmu), _ptr_array(dlv), L.stream_ptr())
        return (None, None, None, *dmu, *dlv)


class FusionFn(torch.autograd.Function):
    """(mu_1..mu_M, lv_1..lv_M, eps) -> (sub_mu [nsub,B,D], sub_lv, joint_mu, joint_lv, z, kl [nsub], nan_flag)"""

    @staticmethod
    def forward(ctx, plan, eng, eps, *mulv):
        M = plan.cfg.M
        mus = [t.contiguous().float() for t in mulv[:M]]
        lvs = [t.contiguous().float() for t in mulv[M:]]
        B, D, ns = plan.B, plan.D, plan.cfg.nsub
        dev = mus[0].device
        sub_mu = torch.empty(ns, B, D, device=dev)
        sub_lv = torch.empty(ns, B, D, device=dev)
        jmu, jlv, z = (torch.empty(B, D, device=dev) for _ in range(3))
        kl = torch.empty(ns, device=dev)
        nan_flag = torch.zeros(1, dtype=torch.int32, device=dev)
        eps = eps.contiguous().float()
        mu_p, lv_p = _ptr_array(mus), _ptr_array(lvs)
        # algorithmic bytes: read 2 M D + eps D, write subsets 2 nsub D + joint 2 D + z D, fp32
        L.annotate(kind='fusion_fwd', bytes=4 * B * D * (2 * len(mus) + 1 + 2 * ns + 3))
        L.call('mopoe_fusion_fwd', C.byref(plan.cfg), mu_p, lv_p, L.ptr(eps), L.ptr(sub_mu), L.ptr(sub_lv), L.ptr(jmu),
               L.ptr(jlv), L.ptr(z), L.ptr(kl), L.ptr(nan_flag), L.ptr(eng.ws64(ns * B)), L.stream_ptr())
        ctx.plan, ctx.eng = plan, eng
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(eps, sub_mu, sub_lv, *mus, *lvs)
        ctx.mark_non_differentiable(nan_flag)
        return sub_mu, sub_lv, jmu, jlv, z, kl, nan_flag

    @staticmethod
    def backward(ctx, d_smu, d_slv, d_jmu, d_jlv, d_z, d_kl, _d_flag):
        plan = ctx.plan
        eps, sub_mu, sub_lv, *mulv = ctx.saved_tensors
        M = plan.cfg.M
        mus, lvs = mulv[:M], mulv[M:]

        def c(t):
            return t.contiguous().float() if t is not None else None
        d_smu, d_slv, d_jmu, d_jlv, d_z, d_kl = map(c, (d_smu, d_slv, d_jmu, d_jlv, d_z, d_kl))
        dmu = [torch.empty_like(t) for t in mus]
        dlv = [torch.empty_like(t) for t in lvs]
        # read d_z + the experts' 2 M D, write their gradients 2 M D (the saved subset tensors come out of L2 at these sizes)
        L.annotate(kind='fusion_bwd', bytes=4 * plan.B * plan.D * (1 + 4 * M))
        L.call('mopoe_fusion_bwd', C.byref(plan.cfg), _ptr_array(mus), _ptr_array(lvs), L.ptr(eps), L.ptr(sub_mu),
               L.ptr(sub_lv), L.ptr(d_z), L.ptr(d_jmu), L.ptr(d_jlv), L.ptr(d_smu), L.ptr(d_slv), L.ptr(d_kl),
               _ptr_array(dmu), _ptr_array(dlv), L.stream_ptr())
        return (None, None, None, *dmu, *dlv)
