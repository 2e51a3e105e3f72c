"""Developer probe: fp32-mode accuracy of the BatchNorm passes against fp64 (run once per library mode and compare)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mopoe_mimic_b200 import _lib as L  # noqa: E402
from mopoe_mimic_b200.engine import Act, Engine  # noqa: E402


def main():
    eng = Engine('cuda', torch.float32, L.IMPL_AUTO)
    for (B, H, W, Cc) in [(16, 64, 64, 16), (16, 32, 32, 32), (16, 1, 512, 16), (16, 8, 8, 64), (16, 64, 64, 128)]:
        g = torch.Generator(device='cuda').manual_seed(1)
        x64 = (torch.randn(B, H, W, Cc, generator=g, device='cuda') * 1.3 + 0.4).double()
        dy64 = torch.randn(B, H, W, Cc, generator=g, device='cuda').double()
        x = Act(x64.float().contiguous(), B, H, W, Cc, 0, 0)
        dy = Act(dy64.float().contiguous(), B, H, W, Cc, 0, 0)
        gamma = torch.rand(Cc, device='cuda') + 0.5
        beta = torch.randn(Cc, device='cuda') * 0.3
        st = eng.bn_stats(x, None, L.MASK_NONE)
        mean, var = x64.mean(dim=(0, 1, 2)), x64.var(dim=(0, 1, 2), unbiased=False)
        is64 = 1 / torch.sqrt(var + 1e-5)
        e_mean = float(((st[0].double() - mean).abs() / (mean.abs() + 1e-3)).max())
        e_is = float(((st[1].double() - is64).abs() / is64).max())
        a = Act(torch.empty(B, H, W, Cc, device='cuda'), B, H, W, Cc, 0, 0)
        eng.bn_apply(x, None, L.MASK_NONE, st, gamma, beta, True, a)
        ref = torch.relu((x64 - mean) * is64 * gamma.double() + beta.double())
        e_app = float((a.t.double() - ref).abs().max())
        dg, db = torch.zeros(Cc, device='cuda'), torch.zeros(Cc, device='cuda')
        dx = Act(torch.empty(B, H, W, Cc, device='cuda'), B, H, W, Cc, 0, 0)
        eng.bn_bwd(dy, a, 1.0, x, None, L.MASK_NONE, st, gamma, dg, db, None, dx)
        gate = (ref > 0).double()
        gg = dy64 * gate
        xh = (x64 - mean) * is64
        n = B * H * W
        sg, sgx = gg.sum(dim=(0, 1, 2)), (gg * xh).sum(dim=(0, 1, 2))
        rdx = gamma.double() * is64 * (gg - sg / n - xh * sgx / n)
        e_db = float(((db.double() - sg).abs() / (gg.abs().sum(dim=(0, 1, 2)))).max())
        e_dg = float(((dg.double() - sgx).abs() / ((gg * xh).abs().sum(dim=(0, 1, 2)))).max())
        flips = int(((a.t > 0).double() != gate).sum())
        e_dx = float((dx.t.double() - rdx).abs().max() / rdx.abs().max())
        # combine + next-block statistics in one pass
        c64 = torch.randn(B, H, W, Cc, generator=g, device='cuda').double()
        cc = Act(c64.float().contiguous(), B, H, W, Cc, 0, 0)
        y = Act(torch.empty(B, H + 2, W + 2, Cc, device='cuda'), B, H, W, Cc, 1, 1)
        rm, rv = torch.zeros(Cc, device='cuda'), torch.ones(Cc, device='cuda')
        _, stn = eng.combine(x, st, gamma, beta, cc, None, L.MASK_NONE, 2.0, 0.3, y, bn=(rm, rv))
        yv = y.interior().double()
        ym, yvv = yv.mean(dim=(0, 1, 2)), yv.var(dim=(0, 1, 2), unbiased=False)
        e_cm = float(((stn[0].double() - ym).abs() / (ym.abs() + 1e-3)).max())
        e_cis = float(((stn[1].double() - 1 / torch.sqrt(yvv + 1e-5)).abs() * torch.sqrt(yvv + 1e-5)).max())
        st2 = eng.bn_stats(y, None, L.MASK_NONE)
        e_sm = float(((st2[0].double() - ym).abs() / (ym.abs() + 1e-3)).max())
        e_sis = float(((st2[1].double() - 1 / torch.sqrt(yvv + 1e-5)).abs() * torch.sqrt(yvv + 1e-5)).max())
        print('   combine_bn: mean %.2e invstd %.2e   | separate bn_stats: mean %.2e invstd %.2e' % (e_cm, e_cis, e_sm, e_sis))
        print('%-16s mean %.2e invstd %.2e apply %.2e dbeta %.2e dgamma %.2e dx %.2e gate flips %d' % (
            '%dx%dx%dx%d' % (B, H, W, Cc), e_mean, e_is, e_app, e_db, e_dg, e_dx, flips))


if __name__ == '__main__':
    main()
