"""fp32-class mode check on a B200: forward vs the committed fp32 fixtures, decode vs fixtures, timing.
Usage: python tools/fp32_check.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from biom3_b200 import synthetic  # noqa: E402
from biom3_b200.engine import Engine  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


def rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max()).item()


def main():
    z = np.load(os.path.join(GOLDEN, 'full_forward_b2.npz'))
    args = synthetic.stage3_args()
    sd = synthetic.random_state_dict(args, seed=int(z['weight_seed']))
    x = torch.from_numpy(z['x'].astype(np.int64)).cuda()
    t = torch.from_numpy(z['t'].astype(np.int64)).cuda()
    zc = torch.from_numpy(z['z_c']).cuda()
    ref = torch.from_numpy(z['logits'])
    for prec in ('bf16', 'fp32'):
        eng = Engine(args, sd, torch.device('cuda'), 2, precision=prec)
        got = eng.forward(x, t, zc).cpu()
        print(f'{prec}: full-config forward rel_err {rel_err(got, ref):.3e}  max softmax diff '
              f'{(torch.softmax(got, 1) - torch.softmax(ref, 1)).abs().max().item():.3e}', flush=True)
        eng.close()
    B = int(os.environ.get('DIAG_B', '64'))
    eng = Engine(args, sd, torch.device('cuda'), B, precision='fp32')
    zc = synthetic.synthetic_z_c(1, 512, seed=1).repeat(B, 1).cuda()
    path = synthetic.synthetic_paths(B, 1024, seed=2).cuda()
    eng.decode(zc, path, num_steps=2, seed=1)
    torch.cuda.synchronize()
    t0 = time.time()
    eng.decode(zc, path, num_steps=8, seed=1)
    torch.cuda.synchronize()
    print(f'fp32 decode B={B}: {(time.time() - t0) / 8 * 1e3:.2f} ms/step')
    print('profile_step:', {k: round(v, 3) for k, v in eng.profile_step(B, B).items()})


if __name__ == '__main__':
    main()
