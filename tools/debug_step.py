"""Developer probe (not a pytest file): one small training step, product vs oracle, error per quantity.

    python tests/debug_step.py [fp32|bf16] [case-kwargs as k=v ...]
"""
import os
import sys
import time
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import helpers as H  # noqa: E402


def main():
    cd = sys.argv[1] if len(sys.argv) > 1 else 'fp32'
    kw = dict(batch_size=8, DIM_img=16, DIM_text=16, class_dim=32)
    actual = None
    for a in sys.argv[2:]:
        k, v = a.split('=')
        if k == 'actual':
            actual = int(v)
        elif k == 'mods':
            kw[k] = tuple(v.split(','))
        elif k == 'method':
            kw[k] = v
        else:
            kw[k] = int(v)
    ofl, state, batch, noise = H.make_case(kw, actual)
    t0 = time.time()
    orc = H.run_oracle(ofl, state, batch, noise)
    print('oracle %.1fs loss %.6f' % (time.time() - t0, float(orc['total_loss'])), flush=True)
    t0 = time.time()
    exp, out, grads = H.run_product(ofl, state, batch, noise, cd)
    print('product %.1fs loss %.6f' % (time.time() - t0, float(out['total_loss'])), flush=True)
    errs = H.compare_step(orc, out, grads)
    worst = errs.pop('_worst_grad')
    ng = [(k, v) for k, v in errs.items() if not k.startswith('grad.')]
    print('max non-grad err %.3e (%s)' % max((v, k) for k, v in ng))
    if os.environ.get('VERBOSE'):
        for k, v in ng:
            print('%-28s %.3e' % (k, v))
    g = sorted(((v, k) for k, v in errs.items() if k.startswith('grad.')), reverse=True)
    print('worst grads:')
    for v, k in g[:int(os.environ.get('NG', '6'))]:
        print('  %-70s %.3e' % (k, v))
    print('median grad err %.3e' % g[len(g) // 2][0])


if __name__ == '__main__':
    try:
        main()
    except Exception:
        traceback.print_exc()
        sys.exit(1)
