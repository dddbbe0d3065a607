"""Counts the Blackwell-specific SASS mnemonics per kernel of libxview_b200.so (no GPU needed):
UTC*MMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA loads /
stores, UTCBAR = tcgen05.commit, and the legacy HMMA (must be absent)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'modular_semantic_segmentation_b200', 'libxview_b200.so')
WANT = ['UTCHMMA', 'UTCHMMA.2CTA', 'LDTM', 'UTMALDG', 'UTMALDG.2CTA', 'UTMASTG', 'UTCBAR', 'HMMA']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    name = None
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            name = subprocess.run(['c++filt', m.group(1)], capture_output=True,
                                  text=True).stdout.strip()
            name = name.replace('xv::(anonymous namespace)::', '').replace('void ', '')
            name = re.sub(r'\((xv::)?[A-Za-z_ ].*', '', name)
            counts[name] = collections.Counter()
            continue
        m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m and name:
            op = m.group(1)
            base = op.split('.')[0]
            if base in ('UTCHMMA', 'UTMALDG') and '.2CTA' in op:
                counts[name][base + '.2CTA'] += 1
            elif base in WANT:
                counts[name][base] += 1
    cols = WANT
    print('| kernel | ' + ' | '.join(cols) + ' |')
    print('|---|' + '---|' * len(cols))
    for k, c in counts.items():
        if not (c['UTCHMMA'] or c['UTCHMMA.2CTA'] or c['UTMALDG'] or c['UTMALDG.2CTA'] or
                c['UTMASTG'] or c['HMMA']):
            continue
        print('| `%s` | ' % k[:70] + ' | '.join(str(c[x]) for x in cols) + ' |')
    legacy = sum(c['HMMA'] for c in counts.values())
    print('\nkernels in the library: %d; legacy HMMA instructions: %d' % (len(counts), legacy))


if __name__ == '__main__':
    sys.exit(main())
