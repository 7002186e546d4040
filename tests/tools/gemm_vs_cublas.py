#!/usr/bin/env python
"""The four GEMM shapes of a decode layer at B = 64: this library's tcgen05 kernel (plain bf16 store, and the fused epilogue
the step uses) against torch.matmul (cuBLAS) on the same operands, each run back to back for ~0.7 s so that both sit in the
same power-capped regime as the decode step.  cuBLAS has no epilogue work here: it is the bar for the mainloop only."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from biom3_b200 import engine

def sustained(fn, secs=0.7):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    while time.time() - t0 < secs:
        for _ in range(20): fn()
        n += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

M = 65536
for name, N, K, epi in (('qkv', 1536, 512, 0), ('out', 512, 512, 6), ('ff1', 2048, 512, 2), ('ff2', 512, 2048, 5)):
    A = (torch.randn(M, K, device='cuda') * 0.5).bfloat16()
    W = (torch.randn(N, K, device='cuda') * 0.1).bfloat16()
    bias = torch.randn(N, device='cuda')
    Wt = W.t()
    out_c = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
    t_cublas = sustained(lambda: torch.matmul(A, Wt, out=out_c))
    out0 = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
    t_plain = sustained(lambda: engine.gemm_test(A, W, None, 0, 256, out=out0, pair=True))
    oute = torch.zeros(2, M, N, device='cuda', dtype=torch.bfloat16) if epi in (5, 6) else torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
    t_epi = sustained(lambda: engine.gemm_test(A, W, bias, epi, 256, out=oute, pair=True))
    fl = 2.0 * M * N * K
    print(f'{name}: N={N} K={K}: cuBLAS {t_cublas:7.1f} us {fl / t_cublas / 1e6:6.0f} TF | ours plain store {t_plain:7.1f} us {fl / t_plain / 1e6:6.0f} TF | ours epi {epi} {t_epi:7.1f} us {fl / t_epi / 1e6:6.0f} TF', flush=True)
