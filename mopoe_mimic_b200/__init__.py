"""mopoe_mimic_b200 — B200-native MoPoE-MIMIC training step (hand-written sm_100a CUDA behind the reference's
Python model API).  See DESIGN.md / INTEGRATION.md."""
from . import _lib
from .train import (CudaOutOfMemory, Experiment, FlatAdam, GraphedTrainStep, NaNInLatent, basic_routine_epoch, default_flags,  # noqa: F401
                    forward_backward, packed_stats, train_step)
from .mmvae import BaseMMVae, MMVaeMimic, VAETextMimic, VAEtrimodalMimic  # noqa: F401
from .networks import DecoderImg, DecoderText, EncoderImg, EncoderText  # noqa: F401
from .modalities import MimicLateral, MimicPA, MimicText  # noqa: F401
from .losses import (calc_joint_elbo_loss, calc_kl_divergence, calc_klds, calc_klds_style, calc_log_probs,  # noqa: F401
                     calc_poe_loss)
from .fusion import set_subsets  # noqa: F401
