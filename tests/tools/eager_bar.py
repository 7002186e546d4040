#!/usr/bin/env python
"""The "library kernel" bar of SURVEY.md section 8d: the reference algorithm as plain PyTorch eager ON the B200 (the CPU
oracle's functional forward with its state dict moved to cuda: cuBLAS / cuDNN / ATen kernels, fp32 as the reference runs at
inference, and the same under bf16 autocast), full forward over all positions + draw at every position + the cross-sample
unmask write + the per-step host copies, B = 64.  A bounded number of denoising steps, extrapolated to 1024 (per-step cost is
step independent).  Measurement tool only: nothing here is on the product path.
    python tests/tools/eager_bar.py [steps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from biom3_b200 import synthetic  # noqa: E402
from oracle.model import OracleModel  # noqa: E402
from oracle import sampler as osamp  # noqa: E402

B, L, C = 64, 1024, 29
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
dev = torch.device('cuda')
args = synthetic.stage3_args()
orc = OracleModel(args, synthetic.random_state_dict(args, seed=0))
orc.sd = {k: v.to(dev) for k, v in orc.sd.items()}
z = synthetic.synthetic_z_c(1, 512, seed=1).repeat(B, 1).to(dev)
path = synthetic.synthetic_paths(B, L, seed=2).to(dev)
g = torch.Generator(device='cpu').manual_seed(3)
noise = torch.empty(2 + steps, B * L, C).exponential_(1, generator=g).to(dev)
out = {}
for name, ctx in (('fp32', torch.autocast('cuda', enabled=False)), ('bf16_autocast', torch.autocast('cuda', dtype=torch.bfloat16))):
    stamps = []

    def hook(i, logits):
        torch.cuda.synchronize()
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        stamps.append(e)

    with ctx, torch.no_grad():
        osamp.decode(lambda x, t, y: orc(x, t, y).float(), torch.zeros(B, L, device=dev), torch.zeros(B, device=dev).long(), z, path,
                     noise, L, max_iters=2 + steps, logits_hook=hook)
    torch.cuda.synchronize()
    ms = stamps[2].elapsed_time(stamps[-1]) / (len(stamps) - 3)          # forward-to-forward = one whole reference iteration
    out[name] = {'ms_per_denoising_step': round(ms, 3), 'sequences_per_s_extrapolated': round(B / (ms * 1e-3 * L), 4), 'steps_timed': len(stamps) - 3}
print(json.dumps({'what': 'reference algorithm, PyTorch eager on one B200, B=64 (oracle forward on cuda)', **out,
                  'torch': torch.__version__, 'gpu': torch.cuda.get_device_name(0)}))
