#!/usr/bin/env python
"""Times the local-attention variants (NL = H = 8) and the linear-attention cluster sizes (NL = 0, H = 8) separately at
the decode step's per-layer size (B = 64, L = 1024), and checks the cluster kernel against the fp32 reference."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from biom3_b200 import engine
from oracle.upstream_blocks import linear_attention


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


g = torch.Generator().manual_seed(0)
for (B, H, L) in [(1, 2, 128), (2, 3, 256), (2, 4, 512), (3, 8, 1024), (2, 2, 2048)]:
    qkv = (torch.randn(3, B, H, L, 32, generator=g) * 1.5).bfloat16()
    q, k, v = (t.float() for t in qkv)
    ref = linear_attention(q, k, v).transpose(1, 2).reshape(B * L, H * 32)
    for cl in (0, 1, 2, 4, 8):
        out = engine.attention_test(qkv.cuda(), 0, 3 + 100 * cl).float().cpu()
        err = ((out - ref).abs().max() / ref.abs().max()).item()
        print(f'linear B={B} H={H} L={L} cluster={cl}: rel_err={err:.3e}', flush=True)

B, H, L = 64, 8, 1024
qkv = (torch.randn(3, B, H, L, 32, device='cuda')).bfloat16()
for variant in [int(v) for v in os.environ.get('LOCAL_VARIANTS', '3,4,11,12,13').split(',')]:
    print(f'local variant {variant}: {timeit(lambda: engine.attention_test(qkv, H, variant)):.1f} us (includes a 33 MB memset)', flush=True)
for cl in (0, 1, 2, 4, 8):
    print(f'linear cluster {cl}: {timeit(lambda: engine.attention_test(qkv, 0, 3 + 100 * cl)):.1f} us (includes a 33 MB memset)', flush=True)
out = torch.zeros(B * L, H * 32, device='cuda', dtype=torch.bfloat16)
print(f'memset alone: {timeit(lambda: out.zero_()):.1f} us')
