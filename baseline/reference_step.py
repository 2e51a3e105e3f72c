"""The reference arm of bench.py: the UNMODIFIED reference (Jimmy2027/MoPoE-MIMIC) timed on the host cores.

`vendor()` copies /root/reference/mimic into baseline/_ref/mimic (git-ignored, shipped to the GPU box by gpurun; the
reference is pure Python and is not pip-installable as a wheel: its setup pins torch~=1.6).  Nothing here is product code
and the product never imports it.  What runs is the reference's own model / loss / optimizer code:

    VAEtrimodalMimic(flags, modalities, subsets)          mimic/networks/VAEtrimodalMimic.py:12-62
    EncoderImg / DecoderImg / EncoderText / DecoderText    mimic/networks/ConvNetworks*Mimic.py
    losses.calc_log_probs / calc_klds / calc_joint_elbo_loss   mimic/evaluation/losses.py:6-89
    torch.optim.Adam(lr, betas)                            mimic/utils/experiment.py:171-178

driven by the ~10 lines of run_epochs.basic_routine_epoch:52-96 + train:128-131 restated below (mimic/run_epochs.py itself
imports termcolor and mimic/utils/experiment.py imports matplotlib — neither is in this image — so `flags` / `exp` are bare
namespaces with the fields the path reads).
"""
import os
import shutil
import sys
import time
from collections import OrderedDict
from types import SimpleNamespace

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = '/root/reference/mimic'
REF_DST = os.path.join(HERE, '_ref')


def vendor():
    """copy the reference tree next to this file (build container only; no-op when it is already there or absent)"""
    dst = os.path.join(REF_DST, 'mimic')
    if os.path.isdir(dst) or not os.path.isdir(REF_SRC):
        return os.path.isdir(dst)
    os.makedirs(REF_DST, exist_ok=True)
    shutil.copytree(REF_SRC, dst, ignore=shutil.ignore_patterns('notebooks', '__pycache__', '*.ipynb', 'logs'))
    return True


def available():
    return os.path.isdir(os.path.join(REF_DST, 'mimic', 'networks'))


def load():
    """import the vendored reference modules (both import roots)"""
    sys.path.insert(0, os.path.join(REF_DST, 'mimic'))      # evaluation/losses.py:2-3 uses `from utils import utils`
    sys.path.insert(0, REF_DST)
    from mimic.evaluation import losses
    from mimic.modalities.MimicLateral import MimicLateral
    from mimic.modalities.MimicPA import MimicPA
    from mimic.modalities.MimicText import MimicText
    from mimic.networks.ConvNetworksImgMimic import DecoderImg, EncoderImg
    from mimic.networks.ConvNetworksTextMimic import DecoderText, EncoderText
    from mimic.networks.VAEtrimodalMimic import VAEtrimodalMimic
    from mimic.utils.BaseExperiment import BaseExperiment
    return SimpleNamespace(**locals())


def build(R, batch_size, lr, img_size=128, class_dim=128):
    """tri-modal joint_elbo experiment with the reference's defaults, weights = its own default init"""
    import torch
    fl = SimpleNamespace(
        device=torch.device('cpu'), batch_size=batch_size, class_dim=class_dim, img_size=img_size, image_channels=1,
        DIM_img=128, DIM_text=128, text_encoding='char', vocab_size=0, len_sequence=1024, num_features=71,
        alphabet='x' * 71, feature_extractor_img='resnet', factorized_representation=False, style_pa_dim=0,
        style_lat_dim=0, style_text_dim=0, modality_moe=False, modality_jsd=False, modality_poe=False, joint_elbo=True,
        poe_unimodal_elbos=True, alpha_modalities=[0.25, 0.25, 0.25, 0.25], beta=5.0, beta_style=1.0, beta_content=1.0,
        dataset='testing', distributed=False, world_size=1, text_gen_lastlayer='softmax')
    mods = OrderedDict()
    mods['PA'] = R.MimicPA(R.EncoderImg(fl, 0), R.DecoderImg(fl, 0), fl)
    mods['Lateral'] = R.MimicLateral(R.EncoderImg(fl, 0), R.DecoderImg(fl, 0), fl)
    mods['text'] = R.MimicText(R.EncoderText(fl, 0), R.DecoderText(fl, 0), fl.len_sequence, None, None, fl)
    exp = SimpleNamespace(flags=fl, modalities=mods)
    exp.subsets = R.BaseExperiment.set_subsets(exp)
    exp.rec_weights = {'PA': 0.33, 'Lateral': 0.33, 'text': 0.33}
    exp.style_weights = {'PA': 1.0, 'Lateral': 1.0, 'text': 1.0}
    exp.mm_vae = R.VAEtrimodalMimic(fl, mods, exp.subsets)
    exp.optimizer = torch.optim.Adam(list(exp.mm_vae.parameters()), lr=lr, betas=(0.9, 0.999))
    exp.mm_vae.train()
    return exp


def train_step(R, exp, batch):
    """run_epochs.basic_routine_epoch:52-96 followed by run_epochs.train:128-131"""
    fl = exp.flags
    results = exp.mm_vae(batch)
    log_probs, weighted_log_prob = R.losses.calc_log_probs(exp, results, (batch, None))
    R.losses.calc_klds(exp, results)
    total_loss = R.losses.calc_joint_elbo_loss(exp, None, results['joint_divergence'], fl.beta_style, fl.beta_content,
                                               weighted_log_prob, fl.beta)
    exp.optimizer.zero_grad()
    total_loss.backward()
    exp.optimizer.step()
    return float(total_loss.detach())


def timed_run(batch_size, lr, steps, warmup, min_seconds=None, img_size=128, class_dim=128):
    """seconds per step of the unmodified reference (fp32, all host threads) on synthetic inputs of the bench's shapes"""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    R = load()
    torch.manual_seed(0)
    exp = build(R, batch_size, lr, img_size, class_dim)
    g = torch.Generator().manual_seed(1)
    batch = OrderedDict([
        ('PA', torch.rand(batch_size, 1, img_size, img_size, generator=g)),
        ('Lateral', torch.rand(batch_size, 1, img_size, img_size, generator=g)),
        ('text', torch.nn.functional.one_hot(torch.randint(0, 71, (batch_size, 1024), generator=g), 71).float())])
    times, loss = [], None
    for it in range(warmup + steps):
        if min_seconds is not None and it >= warmup + 2 and sum(times) >= min_seconds:
            break
        t0 = time.perf_counter()
        loss = train_step(R, exp, OrderedDict(batch))
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return sum(times) / len(times), len(times), cores, loss
