"""Per-tensor gradient parity of the bf16 product (the benchmarked dtype): developer report behind the tolerance table in
DESIGN.md §6 and the assertions of tests/test_gpu_bf16.py.

    python tools/bf16_parity.py [case ...]      # on a B200; writes gpurun_out/bf16_parity.json

For every case: the bf16 product against the fp32 oracle (small / cfg1 sizes, CPU) or against the fp32 product (config 2,
B = 256: same weights, inputs, dropout masks and eps), for the ELBO loss and for the smooth surrogate loss (no Laplace
sign() decisions), as relative L2 / max-abs / cosine per parameter tensor.  For the oracle cases the oracle's own
fp32-vs-fp64 gap is printed beside it: that is the floor ANY fp32 implementation has."""
import json
import os
import sys
import time
from collections import OrderedDict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mopoe_oracle as O  # noqa: E402
from tests import helpers as H  # noqa: E402

MID = dict(batch_size=16, DIM_img=64, DIM_text=64, class_dim=64)
CASES = OrderedDict([
    ('smoke64', dict(batch_size=8, DIM_img=64, DIM_text=64, class_dim=64)),
    ('mid_joint', dict(MID)),
    ('mid_moe', dict(MID, method='moe')),
    ('mid_poe', dict(MID, method='poe', batch_size=12)),
    ('mid_patext', dict(MID, mods=('PA', 'text'))),
    ('mid_256px', dict(batch_size=8, DIM_img=32, DIM_text=32, class_dim=64, img_size=256)),
    ('cfg1', dict(batch_size=16)),
])


def summarize(tag, rows, out):
    rows = [r for r in rows if not r['name'].endswith(H.ZERO_GRAD_SUFFIXES)]
    rl = sorted(r['rel_l2'] for r in rows)
    q = lambda f: rl[min(len(rl) - 1, int(f * len(rl)))]
    big = [r for r in rows if r['share'] > 1e-3]
    print('%-34s n=%3d rel_l2 p50 %.2e p90 %.2e p99 %.2e max %.2e | tensors with >0.1%% of |g|: n=%d max rel_l2 %.2e min cos %.5f'
          % (tag, len(rows), q(.5), q(.9), q(.99), rl[-1], len(big), max(r['rel_l2'] for r in big), min(r['cos'] for r in big)))
    for r in sorted(rows, key=lambda r: -r['rel_l2'])[:6]:
        print('      %-72s numel %8d share %.1e rel_l2 %.2e max_rel %.2e cos %.4f'
              % (r['name'][-72:], r['numel'], r['share'], r['rel_l2'], r['max_rel'], r['cos']))
    out[tag] = rows


def oracle_case(name, kw, out):
    t0 = time.time()
    ofl, state, batch, noise = H.make_case(kw)
    orc = H.run_oracle(ofl, state, batch, noise)
    exp, res, grads = H.run_product(ofl, state, batch, noise, 'bf16')
    errs = H.compare_step(orc, res, grads)
    fwd = {k: v for k, v in errs.items() if not k.startswith('grad.') and k != '_worst_grad'}
    print('[%s] forward: loss %.2e, worst %s' % (name, errs['total_loss'], sorted(fwd.items(), key=lambda kv: -kv[1])[:4]))
    summarize(name + ':elbo:bf16-vs-oracle32', H.grad_table(grads, orc['grads']), out)
    del exp
    lo, lp, g_o, g_p = H.smooth_grads(ofl, state, batch, noise, 'bf16')
    truth = H.smooth_grads.truth
    summarize(name + ':smooth:bf16-vs-oracle64', H.grad_table(g_p, truth), out)
    summarize(name + ':smooth:oracle32-vs-oracle64', H.grad_table(g_o, truth), out)
    print('[%s] %.0f s' % (name, time.time() - t0), flush=True)


def cfg2_case(out, B=256):
    ofl = H.oracle_flags(batch_size=B)
    state = O.make_state(ofl, 0, torch.float32)
    batch = OrderedDict((k, v.cuda()) for k, v in O.make_batch(ofl, 1, torch.float32).items())
    noise = H.device_noise(ofl, B, 2)
    for loss in ('elbo', 'smooth'):
        o32, g32 = H.run_product_device_noise(ofl, state, batch, noise, 'fp32', loss)
        o16, g16 = H.run_product_device_noise(ofl, state, batch, noise, 'bf16', loss)
        print('[cfg2 %s] loss fp32 %.6f bf16 %.6f rel %.2e' % (loss, float(o32['total_loss']), float(o16['total_loss']),
                                                              abs(float(o32['total_loss']) - float(o16['total_loss'])) / abs(float(o32['total_loss']))))
        summarize('cfg2_B%d:%s:bf16-vs-fp32-product' % (B, loss), H.grad_table({k: v.cpu() for k, v in g16.items()},
                                                                                {k: v.cpu() for k, v in g32.items()}), out)
        del o32, g32, o16, g16
        torch.cuda.empty_cache()


def main():
    want = sys.argv[1:] or list(CASES) + ['cfg2']
    out = OrderedDict()
    for name in want:
        if name == 'cfg2':
            cfg2_case(out)
        else:
            oracle_case(name, CASES[name], out)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'bf16_parity.json'), 'w') as f:
        json.dump(out, f)


if __name__ == '__main__':
    main()
