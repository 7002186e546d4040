#!/usr/bin/env python
"""Does the epilogue overlap the mainloop?  Time per tile of the pair-tiled GEMM as a function of K with a fixed epilogue:
overlapped -> max(T_epi, a*K); serialised -> T_epi + a*K.  BIOM3_EPI_SKIP=1 gives the mainloop-only line.
    python tools/gemm_ksweep.py            (run on the GPU box)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run():
    import torch
    from biom3_b200 import engine
    M = 65536
    skip = os.environ.get('BIOM3_EPI_SKIP', '0')
    ares = os.environ.get('KSWEEP_ARES', '0') == '1'          # A-resident pair tiling (bf16 epilogues, K <= 512)
    shapes = [(2048, 2, 'ff1 bias+gelu bf16'), (1536, 0, 'store bf16'), (512, 5, 'split resid')]
    if ares:
        shapes = [(N, e, 'ARES ' + n) for (N, e, n) in shapes if e != 5]
    for (N, epi, name) in shapes:
        for K in ((128, 256, 512) if ares else (128, 256, 512, 1024, 2048)):
            A = (torch.randn(M, K, device='cuda') * 0.5).bfloat16()
            W = (torch.randn(N, K, device='cuda') * 0.1).bfloat16()
            bias = torch.randn(N, device='cuda')
            out = torch.zeros(2, M, N, device='cuda', dtype=torch.bfloat16) if epi == 5 else torch.zeros(M, N, device='cuda', dtype=torch.bfloat16)
            for _ in range(3):
                engine.gemm_test(A, W, bias, epi, 256, out=out, pair=True, ares=ares)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                engine.gemm_test(A, W, bias, epi, 256, out=out, pair=True, ares=ares)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 10 * 1e3
            tiles_per_pair = (M / 256) * (N / 256) / 74
            print(f'skip={skip} {name:20s} K={K:5d}: {us:8.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s  {us / tiles_per_pair:6.3f} us/tile', flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 1:
        run()
    else:
        for skip in ['0', '1']:
            subprocess.run([sys.executable, os.path.abspath(__file__), 'run'], env=dict(os.environ, BIOM3_EPI_SKIP=skip))
