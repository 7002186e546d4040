#!/usr/bin/env python
"""Warp-stall samples of one kernel launch per CUDA source line (needs -lineinfo + --import-source on):
    python tools/ncu_lines.py rep kernel-regex launch-skip [n]"""
import csv
import subprocess
import sys

rep, regex, skip = sys.argv[1], sys.argv[2], sys.argv[3]
n = int(sys.argv[4]) if len(sys.argv) > 4 else 30
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', f'regex:{regex}',
                      '--launch-skip', skip, '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
tot_all = 0
out = []
fname = ''
hdr = None
for r in rows:
    if r and r[0] == 'File Path':
        fname = r[1].split('/')[-1]
    elif r and r[0] == 'Function Name':
        func = r[1][:100]
    elif r and r[0] == 'Line No':
        hdr = r
        si = hdr.index('Warp Stall Sampling (All Samples)')
        ie = hdr.index('Instructions Executed')
        stall = [i for i, h in enumerate(hdr) if h.startswith('stall_')]
    elif hdr and r and r[0].strip().isdigit():
        try:
            s = int(r[si])
        except ValueError:
            continue
        top = sorted([(int(r[i]) if r[i].isdigit() else 0, hdr[i][6:]) for i in stall], reverse=True)[:2]
        out.append((s, int(r[ie]) if r[ie].isdigit() else 0, f'{fname}:{r[0]}', r[1].strip(), top))
        tot_all += s
print(func)
print('total samples', tot_all)
for s, inst, loc, line, top in sorted(out, key=lambda x: -x[0])[:n]:
    print(f'{100 * s / max(tot_all, 1):5.1f}%  inst={inst:9d}  {loc:24s} {line[:90]:90s} {top}')
