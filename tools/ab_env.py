#!/usr/bin/env python
"""A/B of environment switches (read at biom3_create) in ONE process on one GPU: an engine per setting at the benchmark
shape (B=64, stage3 config), graph-replayed ms per denoising step interleaved over the settings + the un-graphed
per-kernel profile of each.
    python tools/ab_env.py 256 BIOM3_QSOFT_EPI=0,BIOM3_KEXP_EPI=0 BIOM3_QSOFT_EPI=1,BIOM3_KEXP_EPI=0 BIOM3_QSOFT_EPI=1,BIOM3_KEXP_EPI=1"""
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
steps = int(sys.argv[1])
settings = [dict(kv.split('=') for kv in s.split(',') if kv) for s in sys.argv[2:]]
args = synthetic.stage3_args()
sd = synthetic.random_state_dict(args, seed=0)
z = synthetic.synthetic_z_c(1, 512, seed=1).repeat(B, 1).cuda()
path = synthetic.synthetic_paths(B, 1024, seed=2).cuda()
engines = []
for env in settings:
    os.environ.update(env)
    eng = Engine(args, sd, torch.device('cuda'), B)
    eng.decode(z, path, num_steps=64, seed=1)
    engines.append(eng)
torch.cuda.synchronize()
times = [[] for _ in engines]
sums = [0] * len(engines)
for rep in range(2):
    for i, eng in enumerate(engines):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tokens, _ = eng.decode(z, path, num_steps=steps, seed=1)
        e1.record()
        torch.cuda.synchronize()
        times[i].append(round(e0.elapsed_time(e1) / steps, 4))
        sums[i] = int(tokens.sum())
for i, eng in enumerate(engines):
    profs = [eng.profile_step(B, B) for _ in range(3)]
    prof = {k: round(statistics.median(p[k] for p in profs), 4) for k in profs[0]}
    eng.check_inputs()
    print(json.dumps({'env': settings[i], 'ms_per_step': times[i], 'checksum': sums[i], 'profile': prof}))
