#!/usr/bin/env python
"""GEMM mainloop / epilogue ceilings at the benchmark shapes (run on the GPU box).
BIOM3_EPI_SKIP=1|2 turns the epilogue into a no-op (test hook) to expose the TMA+MMA pipeline rate."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def run():
    import torch
    from biom3_b200 import engine
    M = 65536
    for (N, K, epi, name) in [(1536, 512, 0, 'qkv-like bf16 store'), (2048, 512, 2, 'ff1 bias+gelu'), (512, 2048, 3, 'ff2 resid'), (512, 512, 3, 'out resid')]:
        A = (torch.randn(M, K, device='cuda') * 0.5).bfloat16()
        W = (torch.randn(N, K, device='cuda') * 0.1).bfloat16()
        bias = torch.randn(N, device='cuda')
        out = torch.zeros(M, N, device='cuda', dtype=torch.float32 if epi == 3 else torch.bfloat16)
        for pair in (False, True):
            for _ in range(3):
                engine.gemm_test(A, W, bias, epi, 256, out=out, pair=pair)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                engine.gemm_test(A, W, bias, epi, 256, out=out, pair=pair)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(f'  {name:22s} pair={int(pair)}: {ms*1e3:7.1f} us  {2.0*M*N*K/ms/1e9:7.1f} TFLOP/s', flush=True)

if __name__ == '__main__':
    if len(sys.argv) > 1:
        run()
    else:
        for skip in ['0', '1', '2', '3']:
            print(f'BIOM3_EPI_SKIP={skip}', flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__), 'run'], env=dict(os.environ, BIOM3_EPI_SKIP=skip))
