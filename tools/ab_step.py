#!/usr/bin/env python
"""Quick A/B of the decode step at the benchmark shape (B=64, stage3 config): graph-replayed ms per denoising step
over a partial decode + the un-graphed per-kernel profile.  Environment switches (BIOM3_*) are read at engine creation:
    BIOM3_EPI_PIPE=0 python tools/ab_step.py [steps]"""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from biom3_b200 import synthetic  # noqa: E402
from biom3_b200.engine import Engine  # noqa: E402

B = int(os.environ.get('AB_B', '64'))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 384
args = synthetic.stage3_args()
eng = Engine(args, synthetic.random_state_dict(args, seed=0), torch.device('cuda'), B)
z = synthetic.synthetic_z_c(1, 512, seed=1).repeat(B, 1).cuda()
path = synthetic.synthetic_paths(B, 1024, seed=2).cuda()
eng.decode(z, path, num_steps=128, seed=1)
torch.cuda.synchronize()
times = []
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tokens, _ = eng.decode(z, path, num_steps=steps, seed=1)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1) / steps)
profs = [eng.profile_step(B, B) for _ in range(3)]
prof = {k: round(statistics.median(p[k] for p in profs), 4) for k in profs[0]}
env = {k: v for k, v in os.environ.items() if k.startswith('BIOM3_')}
print(json.dumps({'env': env, 'ms_per_step': [round(t, 4) for t in times], 'checksum': int(tokens.sum()), 'profile': prof}))
