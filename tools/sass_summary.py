"""Count the SASS mnemonics that identify how each kernel of libvtaco_b200.so runs
(UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, SYNCS = mbarrier,
FFMA2 = packed fp32 FMA, REDG...F32x4 = 16-byte vector atomics, ...).  CPU-only (cuobjdump).

usage: python tools/sass_summary.py > profiles/r01_sass_mnemonics.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'vtaco_b200', 'lib', 'libvtaco_b200.so')
KEYS = ['UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'SYNCS', 'ELECT', 'UBLKCP', 'FFMA2', 'FADD2', 'FFMA', 'REDG', 'ATOMG', 'LDS', 'LDG', 'STG']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    ins = re.compile(r'^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[T0-9]+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Za-z0-9_]+)*)')
    for ln in sass.split('\n'):
        m = re.match(r'\s*Function : (\S+)', ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = ins.match(ln)
        if m and cur:
            counts[cur][m.group(1)] += 1
            if m.group(1) == 'REDG' and 'F32x4' in m.group(2):
                counts[cur]['REDG.F32x4'] += 1
    names = subprocess.run(['c++filt'], input='\n'.join(counts.keys()), capture_output=True, text=True).stdout.split('\n')
    print('# SASS mnemonics per kernel (cuobjdump -sass vtaco_b200/lib/libvtaco_b200.so)\n')
    print('| kernel | instructions | ' + ' | '.join(KEYS) + ' | REDG.F32x4 |')
    print('|---|---|' + '---|' * (len(KEYS) + 1))
    for (f, c), name in zip(counts.items(), names):
        tot = sum(v for k, v in c.items() if k != 'REDG.F32x4')
        if tot == 0:
            continue
        name = re.sub(r'\(.*', '', name).replace('void ', '').replace('vtaco::', '')
        print('| `%s` | %d | ' % (name, tot) + ' | '.join(str(c.get(k, 0)) for k in KEYS) + ' | %d |' % c.get('REDG.F32x4', 0))


if __name__ == '__main__':
    sys.exit(main())
