"""Developer probe (not a pytest file): the single-channel image layers (first conv / last deconv) at B=256, 128 px.

    python tests/prof_direct.py [B]
"""
import ctypes as C
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
    dt, Cc, H = torch.bfloat16, 128, 128
    x = torch.rand(B, 1, H, H, device='cuda')
    w = torch.randn(Cc, 1, 3, 3, device='cuda') * 0.1
    y = Act.empty(B, H // 2, H // 2, Cc, 1, 1, dt, 'cuda')
    mb = B * (H // 2) ** 2 * Cc * 2 / 1e6
    ms = timeit(lambda: L.call('mopoe_conv3x3s2_c1_fwd', L.ptr(x), L.ptr(w), B, H, H, C.byref(y.view()), L.stream_ptr()))
    print('conv3x3s2_c1_fwd   %.3f ms  (writes %.0f MB -> %.0f GB/s)' % (ms, mb, mb / ms))
    dy = Act(torch.randn(B, H // 2 + 2, H // 2 + 2, Cc, device='cuda', dtype=dt), B, H // 2, H // 2, Cc, 1, 1)
    nc = eng.nchunk(B * (H // 2) ** 2, Cc)
    ws = eng.ws64(nc * 9 * Cc + nc + 64)
    dw = torch.zeros(Cc, 1, 3, 3, device='cuda')
    ms = timeit(lambda: L.call('mopoe_conv3x3s2_c1_wgrad', L.ptr(x), C.byref(dy.view()), B, H, H, L.ptr(dw), 0, L.ptr(ws), nc,
                               L.stream_ptr()))
    print('conv3x3s2_c1_wgrad %.3f ms  (reads %.0f MB -> %.0f GB/s)' % (ms, mb, mb / ms))
    xa = Act(torch.randn(B, H // 2, H // 2, Cc, device='cuda', dtype=dt), B, H // 2, H // 2, Cc, 0, 0)
    bias = torch.zeros(1, device='cuda')
    out = torch.empty(B, 1, H, H, device='cuda')
    nb = L.load().mopoe_deconv3x3s2_c1_fwd_ws(C.byref(xa.view()))
    wsd = eng.wsf(nb)
    ms = timeit(lambda: L.call('mopoe_deconv3x3s2_c1_fwd', C.byref(xa.view()), L.ptr(w), L.ptr(bias), L.ptr(out), L.ptr(wsd), nb,
                               L.stream_ptr()))
    print('deconv3x3s2_c1_fwd %.3f ms  (reads %.0f MB -> %.0f GB/s)' % (ms, mb, mb / ms))
    dx = Act.empty(B, H // 2, H // 2, Cc, 0, 0, dt, 'cuda')
    db = torch.zeros(1, device='cuda')
    ms = timeit(lambda: L.call('mopoe_deconv3x3s2_c1_bwd', C.byref(xa.view()), L.ptr(w), L.ptr(out), C.byref(dx.view()), L.ptr(dw),
                               L.ptr(db), 0, L.ptr(ws), nc, L.stream_ptr()))
    print('deconv3x3s2_c1_bwd %.3f ms  (dx + dw + dbias; reads+writes %.0f MB -> %.0f GB/s)' % (ms, 2 * mb, 2 * mb / ms))


if __name__ == '__main__':
    main()
