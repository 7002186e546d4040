#!/usr/bin/env python
"""Device-side evidence that the step loop has no host round trip: %globaltimer stamps written by the first and the last
kernel of every denoising step of one full B = 64 decode (1024 CUDA-graph replays).  Prints one JSON line (run on a B200):
    python tools/step_gaps.py > profiles/r02_step_gaps.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
from biom3_b200 import synthetic  # noqa: E402
from biom3_b200.engine import Engine  # noqa: E402

B, L = 64, 1024
args = synthetic.stage3_args()
eng = Engine(args, synthetic.random_state_dict(args, seed=0), torch.device('cuda'), B)
z = synthetic.synthetic_z_c(1, 512, seed=1).repeat(B, 1).cuda()
path = synthetic.synthetic_paths(B, L, seed=2).cuda()
eng.decode(z, path, num_steps=64, seed=1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eng.decode(z, path, seed=2)
e1.record()
torch.cuda.synchronize()
st = eng.debug_buffer('stamps', (L, 2), torch.int64).numpy().astype(np.int64)
dur = (st[:, 1] - st[:, 0]) / 1e3                      # us, first kernel start -> last kernel end of a step
gap = (st[1:, 0] - st[:-1, 1]) / 1e3                   # us, end of step s -> start of step s + 1
print(json.dumps({
    'what': 'per-step %globaltimer stamps of one full 64-sequence decode: 1024 replays of the 101-kernel step graph, device-resident step counter',
    'decode_ms_cuda_events': round(e0.elapsed_time(e1), 2),
    'first_to_last_stamp_ms': round((st[-1, 1] - st[0, 0]) / 1e6, 2),
    'step_us': {'min': round(float(dur.min()), 1), 'median': round(float(np.median(dur)), 1), 'max': round(float(dur.max()), 1)},
    'gap_between_steps_us': {'min': round(float(gap.min()), 2), 'median': round(float(np.median(gap)), 2),
                             'p99': round(float(np.percentile(gap, 99)), 2), 'max': round(float(gap.max()), 2),
                             'sum_ms': round(float(gap.sum()) / 1e3, 3)},
    'note': 'a host round trip per step (the reference does 4 D2H copies + >= 2 syncs per step) would show as gaps of tens of microseconds; '
            'negative gaps are programmatic dependent launch: the next step\'s first kernel starts under the tail of the last one',
}))
