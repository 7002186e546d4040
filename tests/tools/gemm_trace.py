#!/usr/bin/env python
"""clock64 timeline of CTA 0 of the pair-tiled tcgen05 GEMM (MMA issuer, epilogue warp 0, TMA producer), per tile, for the
four GEMM shapes of a decode layer at B = 64 (run on the GPU box)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ['BIOM3_GEMM_TRACE'] = '1'
import numpy as np
import torch
from biom3_b200 import engine, _lib
M = int(os.environ.get('TRACE_M', '65536'))
for (N, K, epi, name) in [(2048, 512, 2, 'ff1 bias+gelu'), (1536, 512, 0, 'qkv-like store'), (512, 512, 5, 'out-proj split resid'), (512, 512, 6, 'out-proj split resid, TMA ring'),
                          (512, 2048, 5, 'ff2 split resid'), (512, 2048, 6, 'ff2 split resid, TMA ring')]:
    if os.environ.get('TRACE_ONLY') and os.environ['TRACE_ONLY'] not in name:
        continue
    A = (torch.randn(M, K, device='cuda') * 0.5).bfloat16()
    W = (torch.randn(N, K, device='cuda') * 0.1).bfloat16()
    bias = torch.randn(N, device='cuda')
    out = torch.zeros(2, M, N, device='cuda', dtype=torch.bfloat16) if epi in (5, 6) else torch.zeros(M, N, device='cuda', dtype=torch.bfloat16)
    for _ in range(3):
        engine.gemm_test(A, W, bias, epi, 256, out=out, pair=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    engine.gemm_test(A, W, bias, epi, 256, out=out, pair=True)
    e1.record()
    torch.cuda.synchronize()
    tr = np.zeros((3, 64, 4), dtype=np.int64)
    _lib.check(_lib.load().biom3_debug_trace(1, C.c_void_p(tr.ctypes.data), tr.nbytes))
    tiles = (M // 256) * (N // 256) / 74
    reps = int(os.environ.get('TRACE_REPS', '1'))
    print(f'==== {name}: N={N} K={K}: {e0.elapsed_time(e1) * 1e3:.1f} us, {tiles:.1f} tiles per CTA pair, {2.0 * M * N * K / e0.elapsed_time(e1) / 1e9:.0f} TFLOP/s')
    t0 = tr[0, 0, 0]
    nt = int(min(tiles, 24))
    print('tile | MMA: start  wait_acc_empty  issue_all | EPI: wait_acc_full  until_release  rest_of_tile | PROD: first->last req | MMA period, EPI period')
    for i in range(2, nt):
        m, e, pr = tr[0, i], tr[1, i], tr[2, i]
        print(f'{i:4d} | {m[0]-t0:8d} {m[1]-m[0]:6d} {m[2]-m[1]:6d} | {e[0]-t0:8d} wait {e[1]-e[0]:6d} rel {e[2]-e[1]:6d} rest {e[3]-e[2]:6d} | {pr[0]-t0:8d} {pr[1]-pr[0]:6d} | {m[0]-tr[0,i-1,0]:6d} {e[0]-tr[1,i-1,0]:6d}')
