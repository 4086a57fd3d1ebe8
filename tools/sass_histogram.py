#!/usr/bin/env python
"""Per-kernel SASS opcode evidence for libd2p.so (runs on the CPU box: cuobjdump only).

  python tools/sass_histogram.py > profiles/r02_sass_opcodes.txt

For every kernel: instruction count and the counts of the Blackwell-specific mnemonics
(UTCHMMA/UTCQMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, UBLKCP =
cp.async.bulk, UTMALDG/UTMASTG = TMA tensor copies, SYNCS = mbarrier, UCGABAR = cluster barrier)
plus HMMA/FFMA/MUFU for context."""
import collections
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(HERE, 'demo2program_b200', 'libd2p.so')
KEYS = ['UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'LDTM', 'STTM', 'UTCBAR', 'UTCCP', 'UBLKCP', 'UBLKRED', 'UTMALDG',
        'UTMASTG', 'SYNCS', 'UCGABAR', 'HMMA', 'FFMA', 'MUFU', 'LDGSTS', 'REDUX', 'ATOM', 'RED']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    kern, hist = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            kern = m.group(1)
            hist[kern] = collections.Counter()
            continue
        m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
        if m and kern:
            hist[kern][m.group(1)] += 1
            hist[kern]['_n'] += 1
    dem = subprocess.run(['c++filt'], input='\n'.join(hist), capture_output=True, text=True).stdout.splitlines()
    print('# SASS opcode histogram of %s (sm_100a), one line per kernel' % os.path.relpath(LIB, HERE))
    print('# columns: instructions | ' + ' '.join(KEYS))
    tot = collections.Counter()
    for (k, c), name in zip(hist.items(), dem):
        short = re.sub(r'\(.*', '', name.replace('(anonymous namespace)::', '').replace('void ', ''))
        cols = ' '.join('%s=%d' % (key, c[key]) for key in KEYS if c[key])
        print('%-58s %6d | %s' % (short[:58], c['_n'], cols))
        tot.update(c)
    print('# total: %d kernels, %d instructions | %s' % (
        len(hist), tot['_n'], ' '.join('%s=%d' % (key, tot[key]) for key in KEYS if tot[key])))


if __name__ == '__main__':
    main()
