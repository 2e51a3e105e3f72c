"""Developer probe (not a pytest file): the HBM-bound likelihood / fusion kernels at config-2 sizes (B = 256), timed with
CUDA events with an L2 flush between launches, and the short program profiled by
`ncu --set full -k regex:"categorical|laplace|fusion"`.

    python tools/prof_nll.py [B]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mopoe_mimic_b200.blocks import CategoricalLogProbSumFn, LaplaceLogProbSumFn  # noqa: E402
from mopoe_mimic_b200.engine import Engine  # noqa: E402
from mopoe_mimic_b200.fusion import FusionFn, FusionPlan, set_subsets  # noqa: E402

PEAK = 6551.7


def timeit(fn, flush, iters=5):
    fn()
    ts = []
    for _ in range(iters):
        flush.zero_()                      # > L2: every timed launch starts cold, as inside the training step
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    eng = Engine('cuda', torch.float32)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    g = torch.Generator(device='cuda').manual_seed(0)
    L_, V = 1024, 71
    # categorical: one-hot fp32 target (the reference's wire format) and index target
    y = (torch.randn(B, L_, V, device='cuda', generator=g) * 3).requires_grad_(True)
    idx = torch.randint(0, V, (B, L_), device='cuda', generator=g)
    tgt = torch.nn.functional.one_hot(idx, V).float()
    gout = torch.tensor(-0.33 / B, device='cuda')
    for name, t in (('one-hot fp32 target', tgt), ('index target', idx.float())):
        out = CategoricalLogProbSumFn.apply(y, t, eng)
        f = timeit(lambda: CategoricalLogProbSumFn.apply(y.detach(), t, eng), flush)
        bw = timeit(lambda: torch.autograd.grad(out, y, gout, retain_graph=True), flush)
        nb = 4 * B * L_ * V * (3 if t.dim() == 3 else 2)
        print('categorical %-20s fwd %.1f us  bwd %.1f us  -> %.0f MB in %.1f us = %.0f GB/s = %.2f of %.0f'
              % (name, f * 1e3, bw * 1e3, nb / 1e6, (f + bw) * 1e3, nb / 1e9 / ((f + bw) * 1e-3), nb / 1e9 / ((f + bw) * 1e-3) / PEAK, PEAK))
    # Laplace, one 128-px image modality
    loc = torch.randn(B, 1, 128, 128, device='cuda', generator=g).requires_grad_(True)
    x = torch.rand(B, 1, 128, 128, device='cuda', generator=g)
    out = LaplaceLogProbSumFn.apply(loc, x, 0.75, eng)
    f = timeit(lambda: LaplaceLogProbSumFn.apply(loc.detach(), x, 0.75, eng), flush)
    bw = timeit(lambda: torch.autograd.grad(out, loc, gout, retain_graph=True), flush)
    nb = 12 * loc.numel()
    print('laplace 128px                    fwd %.1f us  bwd %.1f us  -> %.0f MB in %.1f us = %.0f GB/s = %.2f of %.0f'
          % (f * 1e3, bw * 1e3, nb / 1e6, (f + bw) * 1e3, nb / 1e9 / ((f + bw) * 1e-3), nb / 1e9 / ((f + bw) * 1e-3) / PEAK, PEAK))
    # fused MoPoE: 3 modalities, 7 subsets, D = 128
    mods = ['PA', 'Lateral', 'text']
    D = 128

    class _M:
        def __init__(self, n):
            self.name = n
    sub = set_subsets({m: _M(m) for m in mods})
    plan = FusionPlan(mods, mods, list(sub.keys()), [[m.name for m in v] for v in sub.values()], 'joint_elbo', B, D, B)
    mus = [torch.randn(B, D, device='cuda', generator=g).requires_grad_(True) for _ in mods]
    lvs = [(torch.randn(B, D, device='cuda', generator=g) * 0.5).requires_grad_(True) for _ in mods]
    eps = torch.randn(B, D, device='cuda', generator=g)
    outs = FusionFn.apply(plan, eng, eps, *mus, *lvs)
    f = timeit(lambda: FusionFn.apply(plan, eng, eps, *[m.detach() for m in mus], *[l.detach() for l in lvs]), flush)
    loss = outs[4].sum() + outs[5].sum()
    bw = timeit(lambda: torch.autograd.grad(loss, mus + lvs, retain_graph=True), flush)
    nf, nbk = 4 * B * D * (2 * 3 + 1 + 2 * 7 + 3), 4 * B * D * 13
    print('fusion M=3 S=7 D=128 B=%d: fwd %.1f us (%.2f MB, %.0f GB/s)  bwd (+2 torch sum kernels) %.1f us (%.2f MB)'
          % (B, f * 1e3, nf / 1e6, nf / 1e9 / (f * 1e-3), bw * 1e3, nbk / 1e6))


if __name__ == '__main__':
    main()
