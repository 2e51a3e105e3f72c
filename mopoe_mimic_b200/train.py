"""The training step: run_epochs.basic_routine_epoch (run_epochs.py:52-96) and the step tail of
run_epochs.train (:128-142) — forward, ELBO, backward, (all-reduce), Adam — plus the experiment
object the reference passes around (only the fields the hot path reads).
"""
import ctypes as C
from collections import OrderedDict
from types import SimpleNamespace

import torch
import torch.distributed as dist

from . import _lib as L
from .fusion import set_subsets
from .losses import calc_joint_elbo_loss, calc_klds, calc_klds_style, calc_log_probs, calc_poe_loss
from .mmvae import VAETextMimic  # noqa: F401
from .mmvae import MMVaeMimic, VAEtrimodalMimic
from .modalities import MimicLateral, MimicPA, MimicText
from .networks import DecoderImg, DecoderText, EncoderImg, EncoderText


class NaNInLatent(Exception):
    """utils/exceptions.py:1-2"""


class CudaOutOfMemory(Exception):
    """utils/exceptions.py:4-5"""


def default_flags(**kw):
    """The ~30 flag fields the hot path reads, with the reference's defaults (utils/flags.py, BaseFlags.py)."""
    f = dict(device=torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else torch.device('cpu'),
             batch_size=16, class_dim=128, img_size=128, image_channels=1, DIM_img=128, DIM_text=128,
             text_encoding='char', len_sequence=1024, num_features=71, alphabet='x' * 71, vocab_size=0,
             feature_extractor_img='resnet', factorized_representation=False,
             style_pa_dim=0, style_lat_dim=0, style_text_dim=0,
             modality_moe=False, modality_jsd=False, modality_poe=False, joint_elbo=True, poe_unimodal_elbos=True,
             alpha_modalities=[0.25, 0.25, 0.25, 0.25], beta=5.0, beta_style=1.0, beta_content=1.0,
             rec_weight_m1=0.33, rec_weight_m2=0.33, rec_weight_m3=0.33,
             initial_learning_rate=1e-3, beta_1=0.9, beta_2=0.999,
             dataset='testing', distributed=False, world_size=1, compute_dtype='bf16',
             mods=('PA', 'Lateral', 'text'))
    f.update(kw)
    m = f.pop('method', None)
    if m is not None:
        f['modality_moe'], f['modality_poe'], f['joint_elbo'] = m == 'moe', m == 'poe', m == 'joint_elbo'
        f['modality_jsd'] = m == 'jsd'
    return SimpleNamespace(**f)


class Experiment:
    """Stand-in for MimicExperiment (utils/experiment.py:40-72) restricted to the hot path:
    .flags .modalities .subsets .mm_vae .optimizer .rec_weights .style_weights"""

    def __init__(self, flags):
        self.flags = flags
        self.modalities = self.set_modalities()
        self.subsets = set_subsets(self.modalities)
        self.rec_weights = {'PA': flags.rec_weight_m1, 'Lateral': flags.rec_weight_m2, 'text': flags.rec_weight_m3}
        self.style_weights = {'PA': 1.0, 'Lateral': 1.0, 'text': 1.0}
        self.mm_vae = self.set_model()
        self.optimizer = None

    def set_modalities(self):
        """experiment.py:80-92 — dict order PA, Lateral, text"""
        fl = self.flags
        mods = OrderedDict()
        fz = bool(getattr(fl, 'factorized_representation', False))      # style latents only when factorized
        sd = {'PA': fl.style_pa_dim if fz else 0, 'Lateral': fl.style_lat_dim if fz else 0,
              'text': fl.style_text_dim if fz else 0}
        for m in getattr(fl, 'mods', ('PA', 'Lateral', 'text')):
            if m == 'PA':
                mods[m] = MimicPA(EncoderImg(fl, sd['PA']), DecoderImg(fl, sd['PA']), fl)
            elif m == 'Lateral':
                mods[m] = MimicLateral(EncoderImg(fl, sd['Lateral']), DecoderImg(fl, sd['Lateral']), fl)
            elif m == 'text':
                mods[m] = MimicText(EncoderText(fl, sd['text']), DecoderText(fl, sd['text']), fl.len_sequence, None, None, fl)
            else:
                raise ValueError(m)
        return mods

    def set_model(self):
        if list(self.modalities.keys()) == ['PA', 'Lateral', 'text']:
            return VAEtrimodalMimic(self.flags, self.modalities, self.subsets)
        if list(self.modalities.keys()) == ['text']:
            return VAETextMimic(self.flags, self.modalities, self.subsets)
        return MMVaeMimic(self.flags, self.modalities, self.subsets)

    def set_optimizer(self, exchange=None):
        """experiment.py:171-178: Adam(lr, betas), eps 1e-8, no weight decay — as ONE fused launch over the flat buffers.
        exchange: a dp.PeerExchange — the flat buffers then live in NVLink-shared memory and step() becomes the fused
        reduce-scatter + Adam + all-gather kernel (the DDP all-reduce + optimizer.step of the reference)."""
        self.optimizer = FlatAdam(self.mm_vae, self.flags.initial_learning_rate, (self.flags.beta_1, self.flags.beta_2),
                                  exchange=exchange)
        return self.optimizer


class FlatAdam:
    """torch.optim.Adam semantics over the model's flat parameter / gradient buffers (mopoe_adam_flat)."""

    def __init__(self, model, lr, betas=(0.9, 0.999), eps=1e-8, exchange=None):
        self.model = model
        self.exchange = exchange
        if not hasattr(model, 'flat_params'):
            model.flatten_(exchange.alloc if exchange is not None else None)
        elif exchange is not None:
            raise RuntimeError('the model was flattened before the peer exchange was attached')
        self.p, self.g = model.flat_params, model.flat_grads
        if exchange is not None:
            exchange.connect()
        self.m = torch.zeros_like(self.p)
        self.v = torch.zeros_like(self.p)
        self.lr, self.betas, self.eps = lr, betas, eps
        self.grad_scale = 1.0
        dev = self.p.device
        # step count and bias-correction coefficients live on the device: the step stays CUDA-graph capturable
        self.step_t = torch.zeros(1, dtype=torch.int32, device=dev)
        self.coef = torch.zeros(2, dtype=torch.float32, device=dev)
        self._advanced = False
        self._xs = None
        self._setup_buckets()

    def _setup_buckets(self):
        """Data-parallel overlap (north_star item 4): the decoders' gradients are final long before the encoders' backward
        ends, so their slice of the flat buffers is exchanged (reduce-scatter + Adam + all-gather) on a side stream UNDER
        the encoders' backward; only the encoders' bucket runs after it.  Needs the decoders' parameters to form the
        contiguous tail of the flat buffer (registration order: encoders, then decoders).  poe's unimodal passes revisit
        the decoders, so poe keeps the single exchange.  MOPOE_DP_BUCKETS=0 disables it."""
        import os
        ex, model = self.exchange, self.model
        self.cut = None
        if os.environ.get('MOPOE_DP_BUCKETS', '1') == '0' or getattr(model, 'method', None) == 'poe' or not self.p.is_cuda:
            return
        offs = getattr(model, 'flat_offsets', None)
        if not offs:
            return
        dec = [o for n, o in offs.items() if n.startswith('decoder_')]
        enc = [o for n, o in offs.items() if not n.startswith('decoder_')]
        if not dec or not enc or max(enc) > min(dec):
            return
        cut = min(dec)
        self.cut = cut                                             # single GPU: the same split for the plain Adam kernel
        if ex is not None:
            ex.set_buckets([(cut, self.p.numel()), (0, cut)])      # launch order: decoders first
        self._xs = torch.cuda.Stream()
        self.side_blocks = int(os.environ.get('MOPOE_DPX_SIDE_BLOCKS', '64'))
        model.rt.on_decoders_done = self.begin_exchange

    def _advance(self, eng):
        L.call('mopoe_step_advance', L.ptr(eng.rng_step), L.ptr(self.step_t), L.ptr(self.coef), float(self.lr),
               float(self.betas[0]), float(self.betas[1]), L.stream_ptr())
        self._advanced = True

    def begin_exchange(self):
        """fired from backward when every decoder gradient is final: exchange the decoders' bucket on the side stream"""
        eng = self.model.rt.eng(self.p.device)
        cur = torch.cuda.current_stream()
        self._xs.wait_stream(cur)
        with torch.cuda.stream(self._xs):
            eng.join_wgrad_sides(final=False)       # the decoders' weight gradients were accumulated on side streams
            self._advance(eng)
            if self.exchange is not None:
                self.exchange.adam_step(self.m, self.v, self.coef, self.betas, self.eps, bucket=0, max_blocks=self.side_blocks)
            else:
                self._adam_range(self.cut, self.p.numel())

    def _adam_range(self, lo, hi):
        L.annotate(kind='adam', bytes=28 * (hi - lo))            # p, g, m, v read; p, m, v written (fp32)
        L.call('mopoe_adam_flat_dev', L.ptr(self.p[lo:hi]), L.ptr(self.g[lo:hi]), L.ptr(self.m[lo:hi]), L.ptr(self.v[lo:hi]),
               hi - lo, L.ptr(self.coef), float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.grad_scale),
               L.stream_ptr())

    @property
    def step_count(self):
        return int(self.step_t.item())

    def zero_grad(self, set_to_none=False):
        self.g.zero_()

    def step(self):
        eng = self.model.rt.eng(self.p.device)
        # the backward kernels of the modality branches accumulate straight into the flat gradient views on their own
        # streams and return None to autograd (no AccumulateGrad node orders them): the reference loop
        # `loss.backward(); optimizer.step()` must therefore join the branches before anything reads flat_grads
        if hasattr(self.model, 'join_branches'):
            self.model.join_branches()
        started = self._advanced        # the decoders' bucket was started from inside backward (begin_exchange)
        if started:
            torch.cuda.current_stream().wait_stream(self._xs)
        else:
            self._advance(eng)
        self._advanced = False
        if self.exchange is not None and self.exchange.buckets is not None:
            if not started:             # (a step whose backward did not fire the hook: exchange every bucket here)
                self.exchange.adam_step(self.m, self.v, self.coef, self.betas, self.eps, bucket=0)
            self.exchange.adam_step(self.m, self.v, self.coef, self.betas, self.eps, bucket=1)
        elif self.exchange is not None:
            # reduce-scatter + Adam + all-gather in one kernel over peer memory; moments of a slice live on its owner
            self.exchange.adam_step(self.m, self.v, self.coef, self.betas, self.eps)
        elif started:                   # single GPU: the decoders' Adam already ran under the encoders' backward
            self._adam_range(0, self.cut)
        else:
            self._adam_range(0, self.p.numel())
        eng.invalidate_packs()      # the kernel rewrote the weights in place: packed copies are stale


def basic_routine_epoch(exp, batch):
    """run_epochs.basic_routine_epoch:52-96.  One device->host read (the NaN flag) instead of nine."""
    flags = exp.flags
    mm_vae = exp.mm_vae
    eng = mm_vae.rt.engine
    if eng is not None:
        eng.begin_step()             # same Philox offsets every step; weights changed by the flat Adam are re-packed in one launch
    batch_d = batch[0]
    for m_key in batch_d.keys():
        batch_d[m_key] = batch_d[m_key].to(flags.device, non_blocking=True)
    results = mm_vae(batch_d)
    if flags.dataset != 'testing' and int(results['latents']['_nan_flag'].item()) != 0:
        raise NaNInLatent('The latent representations contain NaNs')
    log_probs, weighted_log_prob = calc_log_probs(exp, results, batch)
    group_divergence = results['joint_divergence']
    klds = calc_klds(exp, results)
    klds_style = calc_klds_style(exp, results) if getattr(flags, 'factorized_representation', False) else None
    if flags.modality_jsd or flags.modality_moe or flags.joint_elbo:
        total_loss = calc_joint_elbo_loss(exp, klds_style, group_divergence, flags.beta_style, flags.beta_content,
                                          weighted_log_prob, flags.beta)
    elif flags.modality_poe:
        total_loss = calc_poe_loss(exp, exp.modalities, group_divergence, klds, klds_style, batch_d, mm_vae, log_probs)
    else:
        raise ValueError('no fusion method selected')
    return {'results': results, 'log_probs': log_probs, 'total_loss': total_loss, 'klds': klds}


def forward_backward(exp, batch):
    """forward + ELBO + zero_grad + backward (everything of the step before the gradient exchange)"""
    out = basic_routine_epoch(exp, batch)
    exp.optimizer.zero_grad()
    out['total_loss'].backward()
    if hasattr(exp.mm_vae, 'join_branches'):
        exp.mm_vae.join_branches()
    return out


def attach_allreduce(exp, allreduce):
    """What wrapping the model in DistributedDataParallel does in the reference (utils/utils.py:179-185): rank 0's
    parameters go to every rank once, and the optimizer divides the summed gradients by the world size (DDP's mean).
    Idempotent; called by train_step / GraphedTrainStep whenever an all-reduce is passed."""
    if allreduce is None or getattr(allreduce, '_attached_to', None) is exp.optimizer:
        return
    from .dp import broadcast_flat
    # the gradients are only final after the all-reduce that runs between backward and step(): no part of the optimizer
    # may start from inside backward (the single-GPU / peer-exchange overlap of FlatAdam)
    rt = getattr(exp.mm_vae, 'rt', None)
    if rt is not None:
        rt.on_decoders_done = None
    exp.optimizer.grad_scale = float(allreduce.grad_scale)
    broadcast_flat(exp.mm_vae.flat_params, 0, getattr(allreduce, 'group', None))
    allreduce._attached_to = exp.optimizer


def train_step(exp, batch, allreduce=None):
    """run_epochs.train:118-131 for one batch: forward + loss, zero_grad, backward, (DP all-reduce), Adam."""
    attach_allreduce(exp, allreduce)
    out = forward_backward(exp, batch)
    if allreduce is not None:
        allreduce(exp.mm_vae.flat_grads)
    exp.optimizer.step()
    return out


class GraphedTrainStep:
    """The training step captured ONCE into CUDA graphs and replayed per batch: ~1000 kernel launches become one
    graph launch, so the host never gates the device.  Single GPU: one graph (forward, ELBO, backward, Adam).
    Data parallel: graph A (forward, ELBO, backward) -> NCCL all-reduce of the flat gradient buffer, launched by
    the host between the graphs -> graph B (Adam).

    Everything step-dependent lives in device memory (dropout step counter, Adam step / bias corrections), the
    batch is copied into static input buffers, and the scalars the reference logs come back as one packed vector.
    The NaN-in-latent guard (utils.check_latents) is read from `nan_flag` after the replay instead of mid-step."""

    def __init__(self, exp, example_batch, allreduce=None, warmup=2, token_indices=False):
        """token_indices: keep the byte indices of the character text in a static device buffer next to the one-hot rows
        and run the text encoder's first layer as a gather over them (meant for the 1-byte-per-token wire
        format — one-hot float batches still work: their indices are recovered with an argmax per call)."""
        flags = exp.flags
        dev = flags.device
        self.exp = exp
        self.allreduce = allreduce
        attach_allreduce(exp, allreduce)
        self._copy_stream = None
        self.static = {k: torch.empty(v.shape, dtype=torch.float32, device=dev) for k, v in example_batch.items()}
        for k, v in example_batch.items():
            self.static[k].copy_(v)
        self.static_idx = None
        if token_indices and 'text' in self.static and self.static['text'].dim() == 3 and self.static['text'].shape[-1] <= 255:
            from .blocks import attach_token_indices
            self.static_idx = self.static['text'].argmax(dim=-1).to(torch.uint8)
            attach_token_indices(self.static['text'], self.static_idx)
        saved_dataset = flags.dataset
        flags.dataset = 'testing'            # no mid-step .item() while capturing; the flag is checked after replay
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):   # >= 1: allocates workspaces / plans / tensor-map entry points eagerly
                train_step(exp, (dict(self.static), None), allreduce)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        self.graph_b = None
        with torch.cuda.graph(self.graph):
            out = forward_backward(exp, (dict(self.static), None))
            self.stats = packed_stats(out)
            self.nan_flag = out['results']['latents']['_nan_flag']
            if allreduce is None:
                exp.optimizer.step()          # (with a PeerExchange this IS the gradient exchange: still one graph)
        if allreduce is not None:
            self.graph_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_b, pool=self.graph.pool()):
                exp.optimizer.step()
        flags.dataset = saved_dataset
        self.keys = (['total_loss', 'joint_divergence'] + ['kld.' + k for k in out['klds']]
                     + ['log_prob.' + k for k in out['log_probs']])

    def _wire(self, batch, k):
        """compact wire format of input k, expanded on the device into the fp32 static input the model reads:
        'text_u8'  char text as one byte per token [B, L]   (instead of fp32 one-hot rows [B, L, 71]: 1 KB vs 291 KB / report)
        'img_u8'   8-bit image [B, 1, px, px]                (instead of fp32 in [0, 1]: ToTensor() runs on the device)"""
        t, st = batch[k], self.static[k]
        if t.dtype != torch.uint8:
            return None
        if k == 'text' and t.dim() == 2 and st.dim() == 3:
            return 'text_u8'
        if k != 'text' and tuple(t.shape) == tuple(st.shape):
            return 'img_u8'
        raise ValueError('uint8 input %r of shape %s does not match the model input %s' % (k, tuple(t.shape), tuple(st.shape)))

    def _expand(self, k, kind, src_dev):
        st = self.static[k]
        if kind == 'text_u8':
            L.call('mopoe_onehot_u8', L.ptr(src_dev), src_dev.numel(), st.shape[-1], L.ptr(st), L.stream_ptr())
            if self.static_idx is not None:
                self.static_idx.copy_(src_dev.view_as(self.static_idx), non_blocking=True)
        else:
            L.call('mopoe_u8_to_unit', L.ptr(src_dev), src_dev.numel(), L.ptr(st), L.stream_ptr())

    def _stage_host_batch(self, batch):
        """Host (pinned) batch -> one of two device staging sets on a COPY stream, so the H2D transfer of step i+1
        runs under step i's graph instead of in front of its own (the loader side of run_epochs.py:61-62)."""
        kinds = {k: self._wire(batch, k) for k in self.static}
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
            self._staging = [{k: torch.empty_like(t) for k, t in self.static.items()} for _ in range(2)]
            self._staged = [torch.cuda.Event(), torch.cuda.Event()]
            self._consumed = [None, None]
            self._slot = 0
        for k, kind in kinds.items():
            if kind is not None and (k, kind) not in self._staging[0]:
                for sset in self._staging:
                    sset[(k, kind)] = torch.empty(batch[k].shape, dtype=torch.uint8, device=self.static[k].device)
        i = self._slot
        self._slot ^= 1
        cs, cur = self._copy_stream, torch.cuda.current_stream()
        if self._consumed[i] is not None:
            cs.wait_event(self._consumed[i])          # the step that last read this staging set has copied it out
        with torch.cuda.stream(cs):
            for k, kind in kinds.items():
                self._staging[i][k if kind is None else (k, kind)].copy_(batch[k], non_blocking=True)
            self._staged[i].record(cs)
        cur.wait_event(self._staged[i])
        for k, t in self.static.items():
            if kinds[k] is not None:
                self._expand(k, kinds[k], self._staging[i][(k, kinds[k])])
            else:
                t.copy_(self._staging[i][k], non_blocking=True)
                if k == 'text' and self.static_idx is not None:
                    self.static_idx.copy_(t.argmax(dim=-1))
        if self._consumed[i] is None:
            self._consumed[i] = torch.cuda.Event()
        self._consumed[i].record(cur)

    def __call__(self, batch):
        if all(not batch[k].is_cuda for k in self.static):
            self._stage_host_batch(batch)
        else:
            for k, t in self.static.items():
                kind = self._wire(batch, k)
                if kind is not None:
                    self._expand(k, kind, batch[k].contiguous())
                else:
                    t.copy_(batch[k], non_blocking=True)
                    if k == 'text' and self.static_idx is not None:
                        self.static_idx.copy_(t.argmax(dim=-1))
        self.graph.replay()
        if self.graph_b is not None:
            self.allreduce(self.exp.mm_vae.flat_grads)
            self.graph_b.replay()
        return self.stats

    def check_latents(self):
        if self.exp.flags.dataset != 'testing' and int(self.nan_flag.item()) != 0:
            raise NaNInLatent('The latent representations contain NaNs')


def packed_stats(out):
    """One device vector with every scalar the reference logs per step (run_epochs.py:133-142): total loss,
    joint divergence, per-subset KLs, per-modality log-probs -> a single D2H copy instead of ~18 .item() syncs."""
    vals = [out['total_loss'].detach().reshape(1), out['results']['joint_divergence'].detach().reshape(1)]
    vals += [v.detach().reshape(1) for v in out['klds'].values()]
    vals += [v.detach().reshape(1) for v in out['log_probs'].values()]
    return torch.cat([v.float() for v in vals])
