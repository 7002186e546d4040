#!/usr/bin/env python
"""Smallest program that runs the real decode step at the benchmark shape (B=64, stage3 config):
a 3-step decode.  Used as the command under ncu (launch list / --set full captures)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from biom3_b200 import synthetic  # noqa: E402
from biom3_b200.engine import Engine  # noqa: E402

B = int(os.environ.get('NCU_B', '64'))
steps = int(os.environ.get('NCU_STEPS', '3'))
args = synthetic.stage3_args()
eng = Engine(args, synthetic.random_state_dict(args, seed=0), torch.device('cuda'), B)
z = synthetic.synthetic_z_c(1, 512, seed=1).repeat(B, 1).cuda()
path = synthetic.synthetic_paths(B, 1024, seed=2).cuda()
tokens, _ = eng.decode(z, path, num_steps=steps, seed=1)
torch.cuda.synchronize()
print('ok', int(tokens.sum()))
