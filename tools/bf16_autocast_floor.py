"""How far does ANY bf16 implementation of this model land from fp32, tensor by tensor?  The CPU oracle (the reference's op
sequence in plain torch) run twice on the same weights / inputs / dropout masks / eps: in fp32, and under
torch.autocast(bfloat16) — convolutions and linears in bf16 with fp32 accumulation, bf16 activations between them, the
library's own kernels.  The per-tensor relative-L2 gap of the gradients is the floor set by bf16 storage + ReLU gate
flips, independent of our kernels.  Runs on the CPU (no GPU needed):

    python tools/bf16_autocast_floor.py [case ...]
"""
import os
import sys
from collections import OrderedDict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mopoe_oracle as O  # noqa: E402
from tests import helpers as H  # noqa: E402

CASES = OrderedDict([
    ('smoke64', dict(batch_size=8, DIM_img=64, DIM_text=64, class_dim=64)),
    ('mid_joint', dict(batch_size=16, DIM_img=64, DIM_text=64, class_dim=64)),
])


def main():
    torch.set_num_threads(os.cpu_count())
    for name in (sys.argv[1:] or list(CASES)):
        ofl, state, batch, noise = H.make_case(CASES[name])
        ref = H.run_oracle(ofl, state, batch, noise)
        with torch.autocast('cpu', dtype=torch.bfloat16):
            low = H.run_oracle(ofl, state, batch, noise)
        rows = [r for r in H.grad_table({k: v.float() for k, v in low['grads'].items()}, ref['grads'])
                if not r['name'].endswith(H.ZERO_GRAD_SUFFIXES)]
        rl = sorted(r['rel_l2'] for r in rows)
        big = [r for r in rows if r['share'] > 1e-3]
        print('%-10s ELBO loss fp32 %.4f autocast-bf16 %.4f (rel %.1e) | gradient rel-L2 per tensor: p50 %.3f p90 %.3f | tensors '
              'with >0.1%% of |g| (n=%d): max rel-L2 %.3f, min cosine %.4f'
              % (name, float(ref['total_loss']), float(low['total_loss']),
                 abs(float(ref['total_loss']) - float(low['total_loss'])) / abs(float(ref['total_loss'])),
                 rl[len(rl) // 2], rl[int(0.9 * len(rl))], len(big), max(r['rel_l2'] for r in big), min(r['cos'] for r in big)))
        kinds = {}
        for r in rows:
            if r['share'] > 1e-4:
                kinds.setdefault('.'.join(r['name'].split('.')[-2:]), []).append(r['rel_l2'])
        for k in ('conv1.weight', 'conv2.weight', 'bn1.weight', 'bn1.bias', 'bn2.weight', '0.weight'):
            if k in kinds:
                v = sorted(kinds[k])
                print('      %-14s n=%3d median %.3f max %.3f' % (k, len(v), v[len(v) // 2], v[-1]))


if __name__ == '__main__':
    main()
