#!/usr/bin/env python
"""BASELINE.json configs[4]: per-step forward microbench, sequence length 1024, batch sweep, bf16 path or the
fp32-class path (bf16x3 split GEMMs + fp32 attention).  Prints one line per batch size: ms per forward,
algorithmic TFLOP/s (109.552 GFLOP per sequence per step) and the fraction of the measured sustained bf16 peak.
Run on the GPU box:  python tools/forward_microbench.py [max_batch] [bf16|fp32]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from biom3_b200 import synthetic  # noqa: E402
from biom3_b200.engine import Engine  # noqa: E402

max_b = int(sys.argv[1]) if len(sys.argv) > 1 else 256
prec = sys.argv[2] if len(sys.argv) > 2 else 'bf16'
peak = 1387.7
p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
if os.path.exists(p):
    peak = json.load(open(p))['bf16_tflops_sustained']
args = synthetic.stage3_args()
eng = Engine(args, synthetic.random_state_dict(args, seed=0), torch.device('cuda'), max_b, precision=prec)
g = torch.Generator().manual_seed(0)
B = 1
while B <= max_b:
    x = torch.randint(0, 29, (B, 1024), generator=g).cuda()
    t = torch.randint(0, 1024, (B,), generator=g).cuda()
    z = synthetic.synthetic_z_c(B, 512, seed=1).cuda()
    for _ in range(3):
        eng.forward(x, t, z)
    torch.cuda.synchronize()
    n = 10 if B <= 64 else 4
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        eng.forward(x, t, z)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    tf = 109.552e9 * B / (ms * 1e-3) / 1e12
    print(json.dumps({'batch': B, 'ms_per_forward': round(ms, 4), 'tflops': round(tf, 1), 'frac_of_sustained_peak': round(tf / peak, 3),
                      'dtype': prec, 'seq_len': 1024}), flush=True)
    B *= 2
