"""Developer probe (not a pytest file): the two fused GEMM epilogues of a residual block against the launches they replace,
on the model's largest block (image encoder block 1 at B = 256: 128 -> 256 channels, 64 -> 32 px).  Times with CUDA events;
also the short program profiled by `ncu --set full -k regex:persist_kernel`.

    python tools/prof_fused_epi.py [B]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mopoe_mimic_b200 import _lib as L  # noqa: E402
from mopoe_mimic_b200.engine import Act, Engine  # noqa: E402


def timed(fn, iters=5):
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
    iters = int(os.environ.get('PROF_ITERS', '5'))
    eng = Engine('cuda', torch.bfloat16, L.IMPL_TC)
    dt = torch.bfloat16
    ci, co, sp = 128, 256, 64
    OH = sp // 2
    g = torch.Generator(device='cuda').manual_seed(1)
    rnd = lambda *s: torch.randn(*s, device='cuda', dtype=dt, generator=g)
    a2 = Act(rnd(B, sp + 2, sp + 2, ci), B, sp, sp, ci, 1, 1)
    wc = rnd(co, 16 * ci) * 0.02
    r = Act(rnd(B, OH, OH, co), B, OH, OH, co, 0, 0)
    st3 = torch.stack((torch.zeros(co, device='cuda'), torch.ones(co, device='cuda')))
    gamma, beta = torch.ones(co, device='cuda'), torch.zeros(co, device='cuda')
    m2 = (torch.rand(B * co, device='cuda') < 0.5).to(torch.uint8)
    rm, rv = torch.zeros(co, device='cuda'), torch.ones(co, device='cuda')
    y = Act.empty(B, OH, OH, co, 1, 1, dt, 'cuda')
    fl = 2.0 * B * OH * OH * co * 16 * ci

    def unfused_fwd():
        c = eng.gemm_down(a2, wc, None, 4, 2, 1, co)
        eng.combine(r, st3, gamma, beta, c, m2, L.MASK_BC, 2.0, 0.3, y, bn=(rm, rv))

    def fused_fwd():
        res = dict(r=r, stats=st3, gamma=gamma, beta=beta, a=2.0, b=0.3, mask=m2, mode=L.MASK_BC, next_bn=(rm, rv))
        assert eng.gemm_down(a2, wc, None, 4, 2, 1, co, out=y, res=res) is not None

    t_plain = timed(lambda: eng.gemm_down(a2, wc, None, 4, 2, 1, co), iters)
    t_unf = timed(unfused_fwd, iters)
    t_fus = timed(fused_fwd, iters)
    print('conv2 128->256 @64 (M=%d N=%d K=%d): plain GEMM %.3f ms (%.0f TF/s) | GEMM + combine_bn %.3f ms | fused residual epilogue '
          '%.3f ms (%.0f TF/s)' % (B * OH * OH, co, 16 * ci, t_plain, fl / t_plain / 1e9, t_unf, t_fus, fl / t_fus / 1e9), flush=True)

    # input gradient of that conv2 (4 sub-pixel phases) feeding the bn2 backward
    dc = Act(rnd(B, OH + 2, OH + 2, co), B, OH, OH, co, 1, 1)
    wph = [rnd(ci, 4 * co) * 0.02 for _ in range(4)]
    hh = Act(rnd(B, sp, sp, ci), B, sp, sp, ci, 0, 0)
    m1 = (torch.rand(B * ci, device='cuda') < 0.5).to(torch.uint8)
    st2 = eng.bn_stats(hh, m1, L.MASK_BC)
    g2, b2 = torch.ones(ci, device='cuda'), torch.zeros(ci, device='cuda')
    a2o = eng.bn_apply(hh, m1, L.MASK_BC, st2, g2, b2, True, Act.empty(B, sp, sp, ci, 0, 0, dt, 'cuda'))
    dg, db = torch.zeros(ci, device='cuda'), torch.zeros(ci, device='cuda')
    dh = Act.empty(B, sp, sp, ci, 0, 0, dt, 'cuda')

    def unfused_bwd():
        da2 = eng.gemm_up(dc, wph, None, ci)
        eng.bn_bwd(da2, a2o, 1.0, hh, m1, L.MASK_BC, st2, g2, dg, db, None, dh, accumulate=True, beta=b2)

    def fused_bwd():
        da2, sums = eng.gemm_up(dc, wph, None, ci, bnb=dict(x=hh, mask=m1, mode=L.MASK_BC, stats=st2, gamma=g2, beta=b2,
                                                             dgamma=dg, dbeta=db, accumulate=True))
        assert sums is not None
        eng.bn_bwd(da2, a2o, 1.0, hh, m1, L.MASK_BC, st2, g2, dg, db, None, dh, accumulate=True, beta=b2, sums=sums)

    t_plain = timed(lambda: eng.gemm_up(dc, wph, None, ci), iters)
    t_unf = timed(unfused_bwd, iters)
    t_fus = timed(fused_bwd, iters)
    print('its input gradient (4 phases, M=4x%d N=%d K=%d): plain GEMM %.3f ms (%.0f TF/s) | GEMM + bn_bwd (reduce + apply) %.3f ms | '
          'sums in the epilogue + apply %.3f ms' % (B * OH * OH, ci, 4 * co, t_plain, fl / t_plain / 1e9, t_unf, t_fus), flush=True)


if __name__ == '__main__':
    main()
