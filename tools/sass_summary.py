"""Count the SASS mnemonics that prove the tcgen05 / TMEM / TMA path per kernel of libmopoe_b200.so
(B200_PROFILING.md: UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTMALDG / UTMASTG = TMA load / store,
SYNCS = mbarrier).      python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'mopoe_mimic_b200', 'libmopoe_b200.so')
OPS = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTCBAR', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'ELECT', 'SYNCS', 'HMMA', 'FFMA', 'LDGSTS', 'RED', 'MULTIMEM']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    kernels = OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = kernels.setdefault(m.group(1), {k: 0 for k in OPS})
            cur['_n'] = 0
            continue
        if cur is None:
            continue
        m = re.match(r'\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m:
            cur['_n'] += 1
            op = m.group(1)
            for k in OPS:
                if op.startswith(k):
                    cur[k] += 1
    demangle = subprocess.run(['c++filt'], input='\n'.join(kernels), capture_output=True, text=True).stdout.splitlines()
    print('# SASS mnemonic counts per kernel of mopoe_mimic_b200/libmopoe_b200.so (cuobjdump -sass, sm_100a)')
    print('%-78s %6s ' % ('kernel', 'instr') + ' '.join('%7s' % k for k in OPS))
    for (name, c), dn in zip(kernels.items(), demangle):
        short = re.sub(r'\(.*', '', dn)[:78]
        print('%-78s %6d ' % (short, c['_n']) + ' '.join('%7d' % c[k] for k in OPS))


if __name__ == '__main__':
    main()
