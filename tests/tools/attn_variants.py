#!/usr/bin/env python
"""Times the local-attention kernel (NL = H = 8) and the linear-attention kernel (NL = 0, H = 8) separately at the decode
step's per-layer size (B = 64, L = 1024).  (Round 2 used this script with many more kernel variants; their logs are under
profiles/r02_attn_*.log.)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from biom3_b200 import engine


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


B, H, L = 64, 8, 1024
qkv = (torch.randn(3, B, H, L, 32, device='cuda')).bfloat16()
print(f'local attention: {timeit(lambda: engine.attention_test(qkv, H, 0)):.1f} us (includes a 33 MB memset)', flush=True)
print(f'linear attention: {timeit(lambda: engine.attention_test(qkv, 0, 0)):.1f} us (includes a 33 MB memset)', flush=True)
out = torch.zeros(B * L, H * 32, device='cuda', dtype=torch.bfloat16)
print(f'memset alone: {timeit(lambda: out.zero_()):.1f} us')
