#!/usr/bin/env python
"""Compact per-kernel summary of an .ncu-rep (run where ncu is installed, no GPU needed):
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [extra-metric-substring ...]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
extra = sys.argv[2:]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
WANT = [
    'gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__throughput.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu.sum',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max', 'launch__grid_size',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
]
ki = hdr.index('Kernel Name')
for r in rows[2:]:
    print('==', r[ki][:110])
    for i, h in enumerate(hdr):
        if h in WANT or any(e in h for e in extra):
            print(f'   {h} = {r[i]} {units[i]}')
