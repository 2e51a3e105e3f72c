"""Algorithmic HBM traffic of the BatchNorm / residual passes of ONE training step (config 2: B = 256, 128 px, tri-modal,
bf16 activations) per kernel class, divided by the in-graph device times of profiles/r1_in_graph_kernel_times.txt.

    python profiles/hbm_rooflines.py > profiles/r1_hbm_rooflines.txt

Per residual block with E_in = B*H*W*C_in and E_out = B*OH*OW*C_out activation elements (2 bytes each):
  stats reduce  (reduce_rows<0>)        reads  2*E_in + E_out              (bn1, bn2, shortcut BN)
  bn_apply                              reads  2*E_in, writes 2*E_in
  combine                               reads  2*E_out, writes E_out
  bwd reduce    (reduce_rows<1>)        reads  2*E_out + 3*E_in + 3*E_in    (combine, bn2, bn1)
  bwd apply bn1 (<gate, addend>)        reads  4*E_in, writes E_in
  bwd apply bn2 (<gate>)                reads  3*E_in, writes E_in
  bwd apply combine (<second output>)   reads  2*E_out, writes 2*E_out
"""
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mopoe_oracle as O  # noqa: E402  (block tables only: cites the reference's constructors)

PEAK = 6551.7     # GB/s, MEASURED_PEAKS.json hbm copy bandwidth


def blocks(flags, B):
    out = []
    for _ in range(2):      # PA, Lateral
        h = flags.img_size // 2
        for ci, co, k, s, p in O.img_encoder_blocks(flags):
            oh = (h + 2 * p - k) // s + 1
            out.append((B * h * h * ci, B * oh * oh * co))
            h = oh
        h = 1
        for ci, co, k, s, p in O.img_decoder_blocks(flags):
            oh = (h - 1) * s - 2 * p + k
            out.append((B * h * h * ci, B * oh * oh * co))
            h = oh
    L_ = flags.len_sequence // 2
    for ci, co, k, s, p in O.text_encoder_blocks(flags):
        ol = (L_ + 2 * p - k) // s + 1
        out.append((B * L_ * ci, B * ol * co))
        L_ = ol
    L_ = 1
    for ci, co, k, s, p in O.text_decoder_blocks(flags):
        ol = (L_ - 1) * s - 2 * p + k
        out.append((B * L_ * ci, B * ol * co))
        L_ = ol
    return out


def main():
    fl = O.default_flags(batch_size=256)
    bl = blocks(fl, 256)
    ein, eout = sum(b[0] for b in bl), sum(b[1] for b in bl)
    classes = {
        'reduce_rows_kernel<__nv_bfloat16, 0>': (2 * ein + eout) * 2,
        'bn_apply_kernel': 4 * ein * 2,
        'combine_kernel': 3 * eout * 2,
        'reduce_rows_kernel<__nv_bfloat16, 1>': (2 * eout + 6 * ein) * 2,
        'bn_bwd_apply_oneshot_kernel<__nv_bfloat16, true, true, false>': 5 * ein * 2,
        'bn_bwd_apply_oneshot_kernel<__nv_bfloat16, true, false, false>': 4 * ein * 2,
        'bn_bwd_apply_oneshot_kernel<__nv_bfloat16, false, false, true>': 4 * eout * 2,
        'adam_kernel': 7 * 4 * 153067136,
    }
    times = {}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'r1_in_graph_kernel_times.txt')
    for line in open(path):
        m = re.match(r'^(.*?)\s+(\d+)\s+([\d.]+) us\s+[\d.]+%\s*$', line.rstrip())
        if m:
            times[m.group(1).strip()] = float(m.group(3))
    print('# %d residual blocks, sum E_in = %.3e, sum E_out = %.3e elements (B = 256)' % (len(bl), ein, eout))
    print('%-66s %10s %10s %9s %6s' % ('kernel class', 'alg. MB', 'time us', 'GB/s', 'frac'))
    tot_b = tot_t = 0.0
    for name, nbytes in classes.items():
        cand = [k for k in times if name.split('<')[0] in k]
        if name.startswith('reduce_rows'):
            cand = [k for k in cand if (', %s>' % name[-2]) in k]
        if name.startswith('bn_bwd_apply_oneshot'):
            tag = name[name.index('<'):].replace('<__nv_bfloat16, ', '').rstrip('>')
            cand = [k for k in cand if tag in k]
        if not cand:
            continue
        t = times[cand[0]]
        gbs = nbytes / t / 1e3
        tot_b += nbytes
        tot_t += t
        print('%-66s %10.1f %10.1f %9.0f %6.2f' % (name, nbytes / 1e6, t, gbs, gbs / PEAK))
    print('%-66s %10.1f %10.1f %9.0f %6.2f' % ('TOTAL (HBM-bound passes above)', tot_b / 1e6, tot_t, tot_b / tot_t / 1e3,
                                               tot_b / tot_t / 1e3 / PEAK))


if __name__ == '__main__':
    main()
