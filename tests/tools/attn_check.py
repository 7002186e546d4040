#!/usr/bin/env python
"""Attention kernels vs a torch fp32 reference on random head-major qkv (run on the GPU box)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from biom3_b200 import engine
from oracle.upstream_blocks import LocalAttention, linear_attention

def ref(qkv, NL):
    q, k, v = (t.float() for t in qkv)          # [B,H,L,32]
    B, H, L, _ = q.shape
    lo = LocalAttention(128)(q[:, :NL], k[:, :NL], v[:, :NL])
    go = linear_attention(q[:, NL:], k[:, NL:], v[:, NL:])
    return torch.cat([lo, go], 1).transpose(1, 2).reshape(B * L, H * 32)

VARIANTS = [int(v) for v in os.environ.get('ATTN_VARIANTS', '0').split(',')]
g = torch.Generator().manual_seed(0)
for (B, H, L, NL, amp) in [(1, 2, 128, 1, 1.5), (2, 4, 256, 2, 1.5), (2, 16, 1024, 8, 1.5), (2, 4, 512, 2, 4.0), (3, 4, 1024, 3, 6.0)]:
    qkv = (torch.randn(3, B, H, L, 32, generator=g) * amp).bfloat16()      # amp >= 4: peaked rows, large block-to-block maxima
    r = ref(qkv, NL)
    for variant in VARIANTS:
        out = engine.attention_test(qkv.cuda(), NL, variant).float().cpu()
        torch.cuda.synchronize()
        e_loc = ((out[:, :NL * 32] - r[:, :NL * 32]).abs().max() / r[:, :NL * 32].abs().max()).item()
        e_lin = ((out[:, NL * 32:] - r[:, NL * 32:]).abs().max() / r[:, NL * 32:].abs().max()).item()
        print(f'B={B} H={H} L={L} NL={NL} amp={amp} variant={variant}: local rel_err={e_loc:.3e} linear rel_err={e_lin:.3e}', flush=True)
        if e_loc > 2e-2:
            d = (out[:, :32] - r[:, :32]).abs()
            print('   per-window max err (head 0):', [round(d[i * 128:(i + 1) * 128].max().item(), 4) for i in range(L // 128)][:8])
            print('   out[0,:4]', out[0, :4].tolist(), 'ref', r[0, :4].tolist(), ' out[200,:4]' if L > 200 else '', out[min(200, L - 1), :4].tolist(), r[min(200, L - 1), :4].tolist())
B, H, L, NL = 64, 16, 1024, 8
qkv = (torch.randn(3, B, H, L, 32, device='cuda') * 1.0).bfloat16()
for variant in VARIANTS:
    for _ in range(3):
        engine.attention_test(qkv, NL, variant)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        engine.attention_test(qkv, NL, variant)
    e1.record()
    torch.cuda.synchronize()
    print(f'variant {variant}: local+linear attention {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per layer-call at B=64', flush=True)
