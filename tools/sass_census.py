"""SASS opcode census of libhfl_b200.so: per kernel, the counts of the Blackwell-native instructions
(UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA, UTCBAR = tcgen05.commit,
SYNCS = mbarrier) next to the legacy tensor path (HMMA = mma.sync) and the packed / mixed fp32 math.

    python tools/sass_census.py > profiles/r02_sass_census.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'hotformerloc_b200', 'libhfl_b200.so')
WANT = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTCBAR', 'SYNCS', 'HMMA', 'LDGSTS', 'FFMA2', 'FHFMA',
        'FHADD', 'MUFU', 'LDSM']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    name, counts, total = None, {}, {}
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r'\(.*', '', name).replace('void ', '')
            counts[name] = collections.Counter()
            total[name] = 0
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m and name:
            op = m.group(1)
            total[name] += 1
            for w in WANT:
                if op.startswith(w):
                    counts[name][w] += 1
    print('# SASS census of hotformerloc_b200/libhfl_b200.so (cuobjdump -sass, sm_100a); columns = opcode prefixes')
    print('%-52s %7s ' % ('kernel', 'instrs') + ' '.join('%7s' % w for w in WANT))
    for k in sorted(counts):
        print('%-52s %7d ' % (k[:52], total[k]) + ' '.join('%7d' % counts[k][w] for w in WANT))


if __name__ == '__main__':
    sys.exit(main())
