#!/usr/bin/env python
"""Achieved HBM bandwidth of the bandwidth-bound kernels of the path: the all-position categorical
draw (biom3_sample_all: softmax + renormalise + argmax(p / q) at every position) and the unmask scatter, at batch
sizes whose working set exceeds the 126 MB L2.  Algorithmic bytes per sequence: logits 29*1024*4 + noise 29*1024*4 +
tokens 1024*8 (int64 out).  Prints one JSON line per case.  Run on the GPU box: python tools/sampler_bw.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from biom3_b200 import engine  # noqa: E402

peak = 6551.0
p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
if os.path.exists(p):
    peak = json.load(open(p))['hbm_gbs']
L, C = 1024, 29
for B in (64, 512, 2048):
    logits = torch.randn(B, C, L, device='cuda')
    noise = torch.empty(B * L, C, device='cuda').exponential_(1.0)
    for _ in range(3):
        tok = engine.sample_all(logits, noise)
    torch.cuda.synchronize()
    n = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        tok = engine.sample_all(logits, noise)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    nbytes = B * L * (C * 4 * 2 + 8)
    print(json.dumps({'kernel': 'sample_all', 'batch': B, 'us': round(ms * 1e3, 2), 'algorithmic_MB': round(nbytes / 1e6, 1),
                      'GBps': round(nbytes / ms / 1e6, 1), 'frac_of_measured_hbm_peak': round(nbytes / ms / 1e6 / peak, 3),
                      'l2_resident': nbytes < 126e6}), flush=True)
    state = torch.zeros(B, L, dtype=torch.int64, device='cuda')
    path = torch.stack([torch.randperm(L) for _ in range(B)]).cuda()
    G = 64
    for _ in range(3):
        engine.unmask_(state, tok, path, 5, group=G)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        engine.unmask_(state, tok, path, 5, group=G)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(json.dumps({'kernel': 'unmask (incl. inverse-path build, B*L*12 bytes)', 'batch': B, 'group': G, 'us': round(ms * 1e3, 2),
                      'GBps': round(B * L * 12 / ms / 1e6, 1)}), flush=True)
