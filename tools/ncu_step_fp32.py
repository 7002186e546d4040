import os, sys
sys.path.insert(0, os.getcwd())
import torch
from biom3_b200 import synthetic
from biom3_b200.engine import Engine
B = 64
args = synthetic.stage3_args()
eng = Engine(args, synthetic.random_state_dict(args, seed=0), torch.device('cuda'), B, precision='fp32')
z = synthetic.synthetic_z_c(1, 512, seed=1).repeat(B, 1).cuda()
path = synthetic.synthetic_paths(B, 1024, seed=2).cuda()
tokens, _ = eng.decode(z, path, num_steps=2, seed=1)
torch.cuda.synchronize()
print('ok', int(tokens.sum()))
