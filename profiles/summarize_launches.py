"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name (device time, share)."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in csv.DictReader(lines):
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', re.sub(r'<.*', '', r['Kernel Name']))
        v = float(r['Metric Value'].replace(',', ''))
        u = r['Metric Unit']
        v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print('%-52s %7s %12s %7s' % ('kernel', 'launches', 'time (us)', 'share'))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-52s %7d %12.1f %6.1f%%' % (k[:52], v[0], v[1], 100 * v[1] / tot))
    print('%-52s %7d %12.1f' % ('TOTAL', sum(v[0] for v in agg.values()), tot))


if __name__ == '__main__':
    main(sys.argv[1])
