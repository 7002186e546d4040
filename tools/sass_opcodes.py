#!/usr/bin/env python
"""Per-kernel counts of the SASS opcodes that prove the Blackwell path (cuobjdump -sass of the built library):

    python tools/sass_opcodes.py [path/to/libbiom3_b200.so] > profiles/r02_sass_opcodes.md

UTCHMMA = tcgen05.mma (kind::f16), UTMALDG / UTMASTG = TMA tensor loads / stores (cp.async.bulk.tensor), LDTM / STTM =
tcgen05.ld / tcgen05.st (tensor memory), UTCBAR = tcgen05.commit, HMMA = legacy mma.sync, MUFU = special-function unit.
Static instruction counts of the shipped binary, not execution counts."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'biom3_b200', 'libbiom3_b200.so')
COLS = ['UTCHMMA', 'UTMALDG', 'UTMASTG', 'LDTM', 'STTM', 'UTCBAR', 'SYNCS', 'HMMA', 'MUFU', 'FFMA', 'LDG', 'STG', 'total']

sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(['cu++filt', n], capture_output=True, text=True).stdout.strip() or n
counts, order, cur = {}, [], None
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and cur:
        op = m.group(1)
        counts[cur]['total'] += 1
        for c in COLS[:-1]:
            if op.startswith(c):
                counts[cur][c] += 1


def short(name: str) -> str:
    d = demangle(name)
    d = re.sub(r'\((int|bool)\)', '', d)
    d = re.sub(r'\(.*$', '', d)
    d = d.replace('void ', '').replace('(anonymous namespace)::', '')
    return d


print(f'# SASS opcode counts per kernel, `{os.path.relpath(lib, ROOT)}` ({os.path.getsize(lib) / 1e6:.1f} MB), sm_100a')
print()
print('Static counts from `cuobjdump -sass` (tools/sass_opcodes.py).  UTCHMMA = `tcgen05.mma`, UTMALDG / UTMASTG = TMA tensor '
      'load / store, LDTM / STTM = `tcgen05.ld` / `tcgen05.st`, UTCBAR = `tcgen05.commit`, SYNCS = mbarrier ops, HMMA = legacy '
      '`mma.sync`.')
print()
print('| kernel | ' + ' | '.join(COLS) + ' |')
print('|---|' + '---|' * len(COLS))
for k in sorted(order, key=lambda k_: -counts[k_]['UTCHMMA'] * 100000 - counts[k_]['total']):
    print(f'| `{short(k)}` | ' + ' | '.join(str(counts[k][c]) for c in COLS) + ' |')
