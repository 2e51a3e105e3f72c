"""Developer probe (not a pytest file): time the tcgen05 implicit-GEMM kernels on the model's real layer shapes
with CUDA events (and serve as the short program profiled by `ncu --set full`).

    python tests/prof_gemm.py [B]          # prints ms / TFLOP/s per shape, fprop+dgrad+wgrad
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mopoe_mimic_b200 import _lib as L  # noqa: E402
from mopoe_mimic_b200.engine import Act, Engine  # noqa: E402


def bench(fn, flops, iters=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    return ms, flops / ms / 1e9


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    only = sys.argv[2] if len(sys.argv) > 2 else None
    eng = Engine('cuda', torch.bfloat16, L.IMPL_TC)
    dt = torch.bfloat16
    # (name, nd, cin, cout, spatial-in) stride-2 k4 p1 conv layers of the image encoder + 1x1 + text
    shapes = [('enc_img.b1 conv 128->256 @64', 2, 128, 256, 64), ('enc_img.b2 conv 256->384 @32', 2, 256, 384, 32),
              ('enc_img.b3 conv 384->512 @16', 2, 384, 512, 16), ('enc_img.b4 conv 512->640 @8', 2, 512, 640, 8),
              ('enc_txt.b1 conv1d 128->256 @512', 1, 128, 256, 512)]
    tot_ms = tot_fl = 0.0
    for name, nd, ci, co, sp in shapes:
        if only and only not in name:
            continue
        H = 1 if nd == 1 else sp
        ph = 0 if nd == 1 else 1
        x = Act(torch.randn(B, H + 2 * ph, sp + 2, ci, device='cuda', dtype=dt), B, H, sp, ci, ph, 1)
        OH, OW = (1 if nd == 1 else sp // 2), sp // 2
        taps = 4 if nd == 1 else 16
        wc = torch.randn(co, taps * ci, device='cuda', dtype=dt) * 0.02
        g = Act(torch.randn(B, OH + 2 * ph, OW + 2, co, device='cuda', dtype=dt), B, OH, OW, co, ph, 1)
        nph = 2 if nd == 1 else 4
        wph = [torch.randn(ci, (taps // nph) * co, device='cuda', dtype=dt) * 0.02 for _ in range(nph)]
        fl = 2.0 * B * OH * OW * co * taps * ci
        r = []
        r.append(bench(lambda: eng.gemm_down(x, wc, None, 4, 2, 1, co), fl))
        r.append(bench(lambda: eng.gemm_up(g, wph, None, ci), fl))
        r.append(bench(lambda: eng.wgrad_down(x, 4, 2, 1, g), fl))
        print('%-34s fprop %.3f ms %6.0f TF | dgrad(4 phases) %.3f ms %6.0f TF | wgrad %.3f ms %6.0f TF'
              % (name, r[0][0], r[0][1], r[1][0], r[1][1], r[2][0], r[2][1]), flush=True)
        tot_ms += sum(v[0] for v in r)
        tot_fl += 3 * fl
    # 1x1 conv 128 @64x64
    if not only or '1x1' in only:
        a = Act(torch.randn(B, 64, 64, 128, device='cuda', dtype=dt), B, 64, 64, 128, 0, 0)
        w1 = torch.randn(128, 128, device='cuda', dtype=dt) * 0.05
        fl = 2.0 * B * 64 * 64 * 128 * 128
        ms, tf = bench(lambda: eng.gemm_rows(a, w1, None, 128), fl)
        print('%-34s fprop %.3f ms %6.0f TF (HBM-bound: %.0f GB/s)' % ('1x1 128->128 @64', ms, tf, 2 * a.t.numel() * 2 / ms / 1e6))
    if tot_ms:
        print('total %.2f ms  %.0f TFLOP/s' % (tot_ms, tot_fl / tot_ms / 1e9))


if __name__ == '__main__':
    main()
