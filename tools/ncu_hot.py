#!/usr/bin/env python
"""Top stall lines of one kernel launch in an .ncu-rep:  python tools/ncu_hot.py rep regex skip [n]"""
import csv, subprocess, sys, collections
rep, regex, skip = sys.argv[1], sys.argv[2], sys.argv[3]
n = int(sys.argv[4]) if len(sys.argv) > 4 else 25
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{regex}', '--launch-skip', skip,
                      '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
print(rows[0][1][:100])
hdr = rows[1]
si, src, ie = hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Source'), hdr.index('Instructions Executed')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data, reasons = [], collections.Counter()
for r in rows[2:]:
    try:
        data.append((int(r[si]), r[src].strip(), int(r[ie]), r))
    except ValueError:
        continue
    for i in stall:
        if r[i].isdigit():
            reasons[hdr[i]] += int(r[i])
tot = sum(d[0] for d in data)
print('total samples', tot, ' by reason:', [(k, round(100 * v / tot, 1)) for k, v in reasons.most_common(8)])
for d in sorted(data, key=lambda x: -x[0])[:n]:
    r = d[3]
    top = sorted([(int(r[i]) if r[i].isdigit() else 0, hdr[i][6:]) for i in stall], reverse=True)[:2]
    print(f'{100 * d[0] / tot:5.1f}%  exec={d[2]:8d}  {d[1][:64]:64s} {top}')
