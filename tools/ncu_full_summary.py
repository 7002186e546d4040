#!/usr/bin/env python
"""Markdown table + per-kernel DRAM traffic from an `ncu --set full` capture of one decode layer (run where ncu is
installed; no GPU needed):
    python tools/ncu_full_summary.py gpurun_out/r02_step_full.ncu-rep profiles/r02_ncu_full_summary.md profiles/ncu_traffic.json"""
import csv
import json
import subprocess
import sys

rep, out_md, out_json = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name, scale=1.0):
    try:
        v = float(r[col[name]].replace(',', ''))
    except (KeyError, ValueError):
        return float('nan')
    u = units[col[name]]
    if u == 'Mbyte':
        v *= 1e6
    elif u == 'Gbyte':
        v *= 1e9
    elif u == 'Kbyte':
        v *= 1e3
    elif u == 'ms':
        v *= 1e3          # -> us
    elif u == 'ns':
        v *= 1e-3
    return v * scale


def label(name):
    if 'local_attention' in name:
        return 'local_attention'
    if 'linear_attention' in name:
        return 'linear_attention'
    if 'gemm_bf16_tcgen05' in name:
        epi = name.split('<')[1].split(',')[2].strip()
        return {'1': 'gemm_qkv', '2': 'gemm_ff1', '5': 'gemm_resid', '6': 'gemm_resid'}.get(epi, 'gemm_epi' + epi)
    return name.split('(')[0].split('::')[-1]


lines, traffic, n_resid = [], {}, 0
for r in rows[2:]:
    name = r[col['Kernel Name']]
    lab = label(name)
    if lab == 'gemm_resid':                      # launch order inside a layer: out-proj (epilogue 6 or 5) then FF2 (epilogue 5)
        lab = 'gemm_out' if n_resid % 2 == 0 else 'gemm_ff2'
        n_resid += 1
    rd, wr = val(r, 'dram__bytes_read.sum'), val(r, 'dram__bytes_write.sum')
    traffic.setdefault(lab, int(rd + wr))
    lines.append('| {} | {:.1f} | {:.2f} | {:.1f} | {:.1f} | {:.1f} | {:.1f} | {:.1f} | {:.1f} | {:.0f} | {:.0f} |'.format(
        lab, val(r, 'gpu__time_duration.sum'), val(r, 'gpc__cycles_elapsed.avg.per_second'),
        val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
        val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
        val(r, 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'),
        rd / 1e6, wr / 1e6, val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
        val(r, 'lts__t_sector_hit_rate.pct'), val(r, 'launch__registers_per_thread')))
with open(out_md, 'w') as f:
    f.write('# ncu --set full --clock-control none --import-source on, one launch of each kernel of a B=64 decode step '
            '(`tools/ncu_step.py`), round 2 final build\n')
    f.write('| kernel | time us | SM GHz | tensor pipe % | issue active % | XU (MUFU) % | DRAM read MB | DRAM write MB | DRAM % | L2 hit % | regs |\n')
    f.write('|---|---|---|---|---|---|---|---|---|---|---|\n')
    f.write('\n'.join(lines) + '\n\n')
    f.write('Clocks are not capped under ncu (single kernels, no sustained power cap): compare shares and utilisation, not absolute '
            'times; in the timed step the SM clock sits at ~1.45 GHz under `sw_power_cap`.\n')
with open(out_json, 'w') as f:
    json.dump(traffic, f, indent=1)
print(open(out_md).read())
