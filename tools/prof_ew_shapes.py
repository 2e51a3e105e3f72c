"""Developer probe: the BatchNorm / residual passes on every activation geometry of the benchmarked model (B = 256,
128 px, char text), timed with CUDA events.  Run once per library mode and diff:
    python tools/prof_ew_shapes.py > a.txt;  MOPOE_EW_STAGED=0 python tools/prof_ew_shapes.py > b.txt
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mopoe_mimic_b200 import _lib as L  # noqa: E402
from mopoe_mimic_b200.engine import Act, Engine  # noqa: E402

SHAPES = [  # (H, W, C, nd)
    (64, 64, 128, 2), (32, 32, 256, 2), (16, 16, 384, 2), (8, 8, 512, 2), (4, 4, 640, 2), (1, 1, 640, 2),
    (1, 512, 128, 1), (1, 256, 256, 1), (1, 128, 384, 1), (1, 64, 512, 1), (1, 32, 512, 1), (1, 16, 512, 1), (1, 8, 640, 1),
    (1, 4, 640, 1), (1, 1, 640, 1),
]


def timeit(fn, iters=10):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    eng = Engine('cuda', torch.bfloat16, L.IMPL_TC)
    dt = torch.bfloat16
    print('%-18s %8s %8s %8s %8s %8s %8s %8s   (us; MB = one tensor)' % ('H x W x C', 'MB', 'stats', 'apply', 'bwd_red', 'bwd_app',
                                                                       'combine', 'comb_bwd'))
    for H, W, Cc, nd in SHAPES:
        def act(p=0):
            ph = p if nd == 2 else 0
            return Act(torch.randn(B, H + 2 * ph, W + 2 * p, Cc, device='cuda', dtype=dt), B, H, W, Cc, ph, p)
        x, a1, dy, dxs, out, dr, dc = act(1), act(), act(), act(), act(1), act(1), act(1)
        if nd == 2:
            mask, mode = (torch.rand(B * Cc, device='cuda') > 0.5).to(torch.uint8), L.MASK_BC
        else:
            mask, mode = (torch.rand(B * W * Cc, device='cuda') > 0.5).to(torch.uint8), L.MASK_ELEM
        gamma, beta = torch.rand(Cc, device='cuda') + 0.5, torch.randn(Cc, device='cuda')
        dg, db = torch.zeros(Cc, device='cuda'), torch.zeros(Cc, device='cuda')
        stats = eng.bn_stats(x, None, L.MASK_NONE)
        sums = eng.f32(2, Cc)
        nc = eng.nchunk(B * H * W, Cc)
        ws = eng.ws64(2 * nc * Cc)
        import ctypes as C
        t = []
        t.append(timeit(lambda: eng.bn_stats(x, None, L.MASK_NONE)))
        t.append(timeit(lambda: eng.bn_apply(x, mask, mode, stats, gamma, beta, True, a1)))
        t.append(timeit(lambda: L.call('mopoe_bn_bwd_reduce', C.byref(dy.view()), C.byref(a1.view()), 1.0, C.byref(x.view()), L.ptr(mask),
                                       mode, L.ptr(stats[0]), L.ptr(stats[1]), L.ptr(ws), nc, L.ptr(dg), L.ptr(db), 1, L.ptr(sums),
                                       None, None, None, L.stream_ptr())))
        t.append(timeit(lambda: L.call('mopoe_bn_bwd_apply', C.byref(dy.view()), C.byref(a1.view()), 1.0, C.byref(x.view()), L.ptr(mask),
                                       mode, L.ptr(stats[0]), L.ptr(stats[1]), L.ptr(gamma), L.ptr(sums), C.byref(dxs.view()),
                                       C.byref(out.view()), None, L.stream_ptr())))
        t.append(timeit(lambda: eng.combine(a1, stats, gamma, beta, dy, mask, mode, 2.0, 0.3, out)))
        t.append(timeit(lambda: eng.combine_bwd(dy, 2.0, a1, stats, gamma, dg, db, mask, mode, 0.3, dr, dc)))
        print('%-18s %8.1f %8.1f %8.1f %8.1f %8.1f %8.1f %8.1f' % ('%dx%dx%d' % (H, W, Cc), B * H * W * Cc * 2 / 1e6, *t))


if __name__ == '__main__':
    main()
