#!/usr/bin/env python
"""Timeline (clock64) of CTA 0 of the local-attention kernel, per 64-key block: where the issuer thread and one softmax
warp of each stream spend their cycles."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from biom3_b200 import engine, _lib
B, H, L = int(os.environ.get('TRACE_B', '64')), 8, 1024
G0, G1 = (0, 6) if B <= 2 else (8, 44)
qkv = torch.randn(3, B, H, L, 32, device='cuda').bfloat16()
for _ in range(2):
    engine.attention_test(qkv, H, int(os.environ.get("TRACE_VARIANT", "1")))
    torch.cuda.synchronize()
torch.cuda.synchronize()
tr = np.zeros((2, 2, 128, 6), dtype=np.int64)
_lib.check(_lib.load().biom3_debug_trace(0, C.c_void_p(tr.ctypes.data), tr.nbytes))
t0 = tr[tr > 0].min()
for s in range(2):
    print(f'---- stream {s}: block, issuer [before S, S issued, before PV, p_ready seen, PV issued], softmax [top, s_full, loaded, exp done, P stored, after epilogue]')
    for g in range(G0, G1):
        i, x = tr[s, 0, g] - t0, tr[s, 1, g] - t0
        print(f'g={g:3d}  issuer {i[0]:7d} {i[1]-i[0]:5d} | {i[2]:7d} wait {i[3]-i[2]:5d} issue {i[4]-i[3]:5d}   softmax top {x[0]:7d} wait {x[1]-x[0]:5d} ld {x[2]-x[1]:5d} math {x[3]-x[2]:5d} st {x[4]-x[3]:5d} epi {x[5]-x[4]:5d}  | period {tr[s,1,g,0]-tr[s,1,g-1,0]:5d}')
per = np.diff(tr[:, 1, 8:120, 0], axis=1)
print('softmax block period: mean', per.mean(), 'median', np.median(per))
for name, a, b in (('wait s_full', 0, 1), ('ld', 1, 2), ('math', 2, 3), ('store+arrive', 3, 4), ('epilogue', 4, 5)):
    d = tr[:, 1, 8:120, b] - tr[:, 1, 8:120, a]
    print(f'softmax {name:14s} mean {d.mean():7.1f}')
gap = tr[:, 1, 9:120, 0] - tr[:, 1, 8:119, 5]
print(f'softmax loop overhead (after epilogue -> next top) mean {gap.mean():7.1f}')
for name, a, b in (('S: waits+issue', 0, 1), ('PV wait p_ready', 2, 3), ('PV issue', 3, 4)):
    d = tr[:, 0, 8:120, b] - tr[:, 0, 8:120, a]
    print(f'issuer {name:16s} mean {d.mean():7.1f}')
