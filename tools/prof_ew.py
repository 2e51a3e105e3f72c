"""Developer probe (not a pytest file): time the BatchNorm forward / backward passes on the model's largest layer shape
with CUDA events (and serve as the short program profiled by `ncu --set full -k regex:bn_bwd_apply`).

    python tests/prof_ew.py [B]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mopoe_mimic_b200 import _lib as L  # noqa: E402
from mopoe_mimic_b200.engine import Act, Engine  # noqa: E402


def timeit(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    eng = Engine('cuda', torch.bfloat16, L.IMPL_TC)
    dt = torch.bfloat16
    H = W = 64
    Cc = 128

    def act(ph=0):
        return Act(torch.randn(B, H + 2 * ph, W + 2 * ph, Cc, device='cuda', dtype=dt), B, H, W, Cc, ph, ph)
    x, a1, dy, dxs, out = act(1), act(), act(), act(), act(1)
    mask = (torch.rand(B * Cc, device='cuda') > 0.5).to(torch.uint8)
    gamma, beta = torch.rand(Cc, device='cuda') + 0.5, torch.randn(Cc, device='cuda')
    dg, db = torch.zeros(Cc, device='cuda'), torch.zeros(Cc, device='cuda')
    stats = eng.bn_stats(x, None, L.MASK_NONE)
    mb = x.t.numel() * 2 / 1e6
    nb = B * H * W * Cc * 2 / 1e6
    ms = timeit(lambda: eng.bn_stats(x, None, L.MASK_NONE))
    print('bn_stats (reduce+finalize)      %.3f ms  %.0f GB/s' % (ms, nb / ms))
    ms = timeit(lambda: eng.bn_apply(x, None, L.MASK_NONE, stats, gamma, beta, True, a1))
    print('bn_apply (1R 1W)                %.3f ms  %.0f GB/s' % (ms, 2 * nb / ms))
    ms = timeit(lambda: eng.bn_bwd(dy, a1, 1.0, x, None, L.MASK_NONE, stats, gamma, dg, db, dxs, out))
    print('bn_bwd bn1 (reduce 3R, apply 4R 1W) %.3f ms  %.0f GB/s' % (ms, 8 * nb / ms))
    ms = timeit(lambda: eng.bn_bwd(dy, a1, 1.0, x, mask, L.MASK_BC, stats, gamma, dg, db, None, out))
    print('bn_bwd bn2 (reduce 3R, apply 3R 1W) %.3f ms  %.0f GB/s' % (ms, 7 * nb / ms))
    ms = timeit(lambda: eng.combine(a1, stats, gamma, beta, dy, mask, L.MASK_BC, 2.0, 0.3, out))
    print('combine (2R 1W)                 %.3f ms  %.0f GB/s' % (ms, 3 * nb / ms))
    dr, dc = act(1), act(1)
    t1, t2, t3 = a1.t, dy.t, dxs.t
    ms = timeit(lambda: t1.copy_(t2))
    print('torch copy_ (1R 1W)             %.3f ms  %.0f GB/s' % (ms, 2 * nb / ms))
    ms = timeit(lambda: torch.add(t1, t2, out=t3))
    print('torch add (2R 1W)               %.3f ms  %.0f GB/s' % (ms, 3 * nb / ms))
    ms = timeit(lambda: torch.addcmul(t1, t2, t3, out=out.t[:, 1:-1, 1:-1].reshape(-1)[:t1.numel()].view_as(t1)) if False else torch.addcmul(t1, t2, t3, out=t1))
    print('torch addcmul (3R 1W)           %.3f ms  %.0f GB/s' % (ms, 4 * nb / ms))
    ms = timeit(lambda: eng.combine_bwd(dy, 2.0, a1, stats, gamma, dg, db, mask, L.MASK_BC, 0.3, dr, dc))
    print('combine_bwd (reduce 2R, apply 2R 2W) %.3f ms  %.0f GB/s' % (ms, 6 * nb / ms))


if __name__ == '__main__':
    main()
