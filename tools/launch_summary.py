#!/usr/bin/env python
"""Share of the step per kernel from an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file x.csv ...`):
    python tools/launch_summary.py gpurun_out/launches_bench.csv > profiles/r02_launches_bench_summary.md"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[h]
ki, mi, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[h + 1:]:
    if len(r) <= vi or r[mi] != 'gpu__time_duration.sum':
        continue
    v = float(r[vi].replace(',', ''))
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[ui], 1.0)
    name = r[ki].split('(')[0].replace('void ', '')
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print('# ncu launch list summary, round 2 final build: `ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv python '
      'bench.py --steps 1 --warmup 1 --no-cpu-baseline` (500 consecutive launches of the first warm-up decode; cold-cache, serialised, '
      'clocks uncapped: compare SHARES, not absolutes)')
print('| share | total us | launches | kernel |\n|---|---|---|---|')
for name, v in tot.most_common():
    print(f'| {100 * v / total:5.1f}% | {v:9.1f} | {cnt[name]} | `{name}` |')
print(f'| 100% | {total:9.1f} | {sum(cnt.values())} | total |')
