#!/usr/bin/env python
"""Bring-up diagnostics on a B200: runs each stage in its own subprocess (a hang or fault in one
does not take the others down) and prints numeric detail, not just pass/fail.

    python tests/tools/gpu_diag.py [stage ...]      # stages: gemm sampler stages forward decode perf
"""
from __future__ import annotations

import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

SMALL = dict(diffusion_steps=256, transformer_dim=256, transformer_heads=8, transformer_depth=2,
             transformer_local_heads=4, transformer_local_size=128, text_emb_dim=64)


def rel_err(a, b):
    import torch
    a = a.double()
    b = b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def stage_gemm():
    import torch
    from biom3_b200 import engine
    dev = 'cuda'
    g = torch.Generator().manual_seed(0)
    for (M, N, K, bn) in [(128, 256, 64, 256), (128, 128, 64, 128), (256, 512, 128, 256), (1024, 1536, 512, 256),
                          (4096, 512, 2048, 256), (4096, 512, 2048, 128), (65536, 512, 512, 256)]:
        A = (torch.randn(M, K, generator=g) * 0.5).to(dev).bfloat16()
        W = (torch.randn(N, K, generator=g) * 0.1).to(dev).bfloat16()
        ref = A.float() @ W.float().t()
        out = engine.gemm_test(A, W, None, 4, bn)
        torch.cuda.synchronize()
        e = rel_err(out, ref)
        print(f'gemm f32-out M={M} N={N} K={K} bn={bn}: rel_err={e:.3e}', flush=True)
        if e > 1e-2:
            bad = ((out - ref).abs() > 1e-2 * ref.abs().max())
            rows = bad.any(1).nonzero().flatten()
            cols = bad.any(0).nonzero().flatten()
            print('   bad rows', rows[:16].tolist(), '... n=', rows.numel(), ' bad cols', cols[:16].tolist(), 'n=', cols.numel())
            print('   out[0,:8]', out[0, :8].tolist(), '\n   ref[0,:8]', ref[0, :8].tolist())
            print('   out nonzero frac', (out != 0).float().mean().item())
    M, N, K = 1024, 512, 512
    A = (torch.randn(M, K, generator=g) * 0.5).to(dev).bfloat16()
    W = (torch.randn(N, K, generator=g) * 0.1).to(dev).bfloat16()
    bias = torch.randn(N, generator=g).to(dev)
    ref = A.float() @ W.float().t()
    out = engine.gemm_test(A, W, None, 0, 256)
    print(f'gemm bf16-out: rel_err={rel_err(out.float(), ref):.3e}')
    out = engine.gemm_test(A, W, bias, 2, 256)
    print(f'gemm bias+gelu: rel_err={rel_err(out.float(), torch.nn.functional.gelu(ref + bias)):.3e}')
    resid = torch.randn(M, N, generator=g).to(dev)
    out = engine.gemm_test(A, W, bias, 3, 256, out=resid.clone())
    print(f'gemm bias+resid: rel_err={rel_err(out, resid + ref + bias):.3e}')
    # timing of the big shapes
    for (M, N, K, epi, bn) in [(65536, 1536, 512, 0, 256), (65536, 2048, 512, 2, 256), (65536, 512, 2048, 3, 256),
                               (65536, 512, 512, 3, 256), (65536, 512, 2048, 3, 128), (65536, 512, 512, 3, 128)]:
        A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
        W = (torch.randn(N, K, device=dev) * 0.1).bfloat16()
        bias = torch.randn(N, device=dev)
        out = torch.zeros(M, N, device=dev, dtype=torch.float32 if epi == 3 else torch.bfloat16)
        for _ in range(3):
            engine.gemm_test(A, W, bias, epi, bn, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            engine.gemm_test(A, W, bias, epi, bn, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f'gemm time M={M} N={N} K={K} epi={epi} bn={bn}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s', flush=True)


def stage_sampler():
    import torch
    from biom3_b200 import engine, synthetic
    from oracle import sampler as osamp
    B, L, C = 5, 512, 29
    g = torch.Generator().manual_seed(5)
    logits = torch.randn(B, C, L, generator=g) * 2
    noise = torch.empty(B * L, C).exponential_(1, generator=g)
    ref = osamp.sample_tokens(logits, noise)
    got = engine.sample_all(logits.cuda(), noise.cuda()).cpu()
    print('sample_all mismatches:', (ref != got).sum().item(), 'of', ref.numel())
    path = synthetic.synthetic_paths(B, L, seed=9)
    state = torch.zeros(B, 1, L, dtype=torch.long)
    sg = torch.zeros(B, L, dtype=torch.long, device='cuda')
    for t in range(0, 40):
        tok = torch.randint(1, C, (B, L), generator=g)
        osamp.unmask(state, tok, path, torch.full((B, 1), t))
        engine.unmask_(sg, tok.cuda(), path.cuda(), t)
    print('unmask mismatches:', (state[:, 0] != sg.cpu()).sum().item())


def _build(args_over, B, seed=11, perturb=True):
    import torch
    from biom3_b200 import synthetic
    from biom3_b200.engine import Engine
    from oracle.model import OracleModel
    args = synthetic.stage3_args(**args_over)
    sd = synthetic.random_state_dict(args, seed=seed, perturb_norm=perturb)
    eng = Engine(args, sd, torch.device('cuda'), B)
    return args, sd, eng, OracleModel(args, sd)


def stage_stages():
    """depth-1 model: compare every intermediate buffer with the oracle."""
    import torch
    import torch.nn.functional as F
    from biom3_b200 import synthetic
    over = dict(SMALL, transformer_depth=1)
    B = 2
    args, sd, eng, orc = _build(over, B)
    L, D, H = args.diffusion_steps, args.transformer_dim, args.transformer_heads
    g = torch.Generator().manual_seed(3)
    x = torch.randint(0, 29, (B, L), generator=g)
    t = torch.tensor([5, 200])
    z = synthetic.synthetic_z_c(B, args.text_emb_dim, seed=4)
    logits = eng.forward(x.cuda(), t.cuda(), z.cuda()).cpu()
    # oracle pieces
    T, Y = orc.cond_vectors(t, z)
    Ttab = eng.debug_buffer('Ttab', (L, 1, D), torch.float32)
    print('Ttab rel_err (rows t):', rel_err(Ttab[t, 0], T[..., 0]))
    Yd = eng.debug_buffer('Y', (B, 1, D), torch.float32)
    print('Y rel_err:', rel_err(Yd[:, 0], Y[..., 0]))
    cv = eng.debug_buffer('cvec', (B, 1, D), torch.float32)
    print('cvec rel_err:', rel_err(cv[:, 0], T[..., 0] + Y[..., 0]))
    u0 = orc.embed(x) + T[:, None, :, 0] + Y[:, None, :, 0]
    p = 'transformer.transformer_blocks.0.0.layers.layers.0.'
    a1 = F.layer_norm(u0, (D,), sd[p + '0.norm.weight'], sd[p + '0.norm.bias'], 1e-5)
    q = F.linear(a1, sd[p + '0.fn.to_q.weight']); k = F.linear(a1, sd[p + '0.fn.to_k.weight']); v = F.linear(a1, sd[p + '0.fn.to_v.weight'])
    qkv_ref = torch.stack([z_.reshape(B, L, H, 32).transpose(1, 2) for z_ in (q, k, v)])   # [3,B,H,L,32]
    qkv = eng.debug_buffer('qkv', (3, B, H, L, 32), torch.bfloat16).float()
    for i, n in enumerate('qkv'):
        print(f'{n} rel_err:', rel_err(qkv[i], qkv_ref[i]))
    from oracle.upstream_blocks import linear_attention
    NL = args.transformer_local_heads
    lo = orc.local(qkv_ref[0][:, :NL], qkv_ref[1][:, :NL], qkv_ref[2][:, :NL])
    go = linear_attention(qkv_ref[0][:, NL:], qkv_ref[1][:, NL:], qkv_ref[2][:, NL:])
    att_ref = torch.cat([lo, go], 1).transpose(1, 2).reshape(B, L, D)
    att = eng.debug_buffer('att', (B, L, D), torch.bfloat16).float()
    print('att local rel_err:', rel_err(att[..., :NL * 32], att_ref[..., :NL * 32]))
    print('att linear rel_err:', rel_err(att[..., NL * 32:], att_ref[..., NL * 32:]))
    for w in range(L // 128):
        print(f'   local window {w} rel_err:', rel_err(att[:, w * 128:(w + 1) * 128, :NL * 32], att_ref[:, w * 128:(w + 1) * 128, :NL * 32]))
    u1 = u0 + F.linear(att_ref, sd[p + '0.fn.to_out.weight'], sd[p + '0.fn.to_out.bias'])
    a2 = F.layer_norm(u1, (D,), sd[p + '1.norm.weight'], sd[p + '1.norm.bias'], 1e-5)
    hid_ref = F.gelu(F.linear(a2, sd[p + '1.fn.fn.w1.weight'], sd[p + '1.fn.fn.w1.bias']))
    hid = eng.debug_buffer('hid', (B, L, 4 * D), torch.bfloat16).float()
    print('hid rel_err:', rel_err(hid, hid_ref))
    u2 = u1 + F.linear(hid_ref, sd[p + '1.fn.fn.w2.weight'], sd[p + '1.fn.fn.w2.bias'])
    ab = eng.debug_buffer('a', (B, L, D), torch.bfloat16).float()
    if os.environ.get('BIOM3_SPLIT_RESID', '1') != '0' and os.environ.get('BIOM3_LO8', '1') != '0':
        ud = ab                                                  # hi plane only (the 8-bit tiled lo plane is not decoded here)
    elif os.environ.get('BIOM3_SPLIT_RESID', '1') != '0':        # residual stream stored as bf16 hi ('a') + lo
        ud = ab + eng.debug_buffer('u_lo', (B, L, D), torch.bfloat16).float()
    else:
        ud = eng.debug_buffer('u', (B, L, D), torch.float32)
    print('u final rel_err:', rel_err(ud, u2))
    print('bf16 copy of u rel_err:', rel_err(ab, u2))
    ref = orc(x, t, z)
    print('logits rel_err:', rel_err(logits, ref), ' max|ref|', ref.abs().max().item())


def stage_forward():
    import numpy as np
    import torch
    from biom3_b200 import synthetic
    B = 3
    args, sd, eng, orc = _build(SMALL, B)
    L = args.diffusion_steps
    g = torch.Generator().manual_seed(3)
    x = torch.randint(0, 29, (B, L), generator=g)
    t = torch.tensor([0, 100, 255])
    z = synthetic.synthetic_z_c(B, args.text_emb_dim, seed=4)
    got = eng.forward(x.cuda(), t.cuda(), z.cuda()).cpu()
    ref = orc(x, t, z)
    print('small forward logits rel_err:', rel_err(got, ref))
    eng.close()
    # full config vs the committed fixture from the real reference
    zf = np.load(os.path.join(ROOT, 'tests', 'golden', 'full_forward_b2.npz'))
    args = synthetic.stage3_args()
    sd = synthetic.random_state_dict(args, seed=int(zf['weight_seed']))
    from biom3_b200.engine import Engine
    t0 = time.time()
    eng = Engine(args, sd, torch.device('cuda'), 2)
    print(f'full engine build {time.time() - t0:.1f}s')
    got = eng.forward(torch.from_numpy(zf['x'].astype(np.int64)).cuda(), torch.from_numpy(zf['t'].astype(np.int64)).cuda(),
                      torch.from_numpy(zf['z_c']).cuda()).cpu()
    ref = torch.from_numpy(zf['logits'])
    print('FULL forward logits rel_err (max-norm):', rel_err(got, ref), ' max|ref|', ref.abs().max().item(),
          ' mean abs err', (got - ref).abs().mean().item())
    pg, pr = torch.softmax(got, 1), torch.softmax(ref, 1)
    print('FULL forward probs max abs err:', (pg - pr).abs().max().item())


def stage_decode():
    import torch
    from biom3_b200 import synthetic
    from oracle import sampler as osamp
    B = 3
    args, sd, eng, orc = _build(SMALL, B)
    L, C = args.diffusion_steps, 29
    z = synthetic.synthetic_z_c(1, args.text_emb_dim, seed=4).repeat(B, 1)
    path = synthetic.synthetic_paths(B, L, seed=6)
    noise = synthetic.synthetic_noise(L, B, L, C, seed=7)
    margins = []

    def hook(i, logits):
        p = torch.softmax(logits, 1).permute(0, 2, 1).reshape(B * L, C) / noise[i]
        top2 = p.topk(2, -1).values
        margins.append(((top2[:, 0] - top2[:, 1]) / top2[:, 0]).reshape(B, L))

    states, times = osamp.decode(orc, torch.zeros(B, L), torch.zeros(B).long(), z, path, noise, L, logits_hook=hook)
    tokens, traj = eng.decode(z.cuda(), path.cuda(), noise=noise.cuda(), want_traj=True)
    torch.cuda.synchronize()
    traj = traj.cpu().long()
    ref = torch.from_numpy(__import__('numpy').stack(states))[:, :, 0]
    neq = (traj != ref)
    print('decode: traj mismatching entries', neq.sum().item(), 'of', neq.numel(), ' final-token mismatches',
          (tokens.cpu() != ref[-1]).sum().item(), 'of', B * L)
    if neq.any():
        s = neq.flatten(1).any(1).nonzero()[0].item()
        bl = neq[s].nonzero()[0].tolist()
        print(f'   first divergence at step {s}, (b,l)={bl}: got {traj[s][bl[0], bl[1]].item()} ref {ref[s][bl[0], bl[1]].item()}'
              f' oracle top-2 relative margin there {margins[s][bl[0], bl[1]].item():.3e}')


def stage_perf():
    import torch
    from biom3_b200 import synthetic
    from biom3_b200.engine import Engine
    args = synthetic.stage3_args()
    sd = synthetic.random_state_dict(args, seed=0)
    B = int(os.environ.get('DIAG_B', '64'))
    eng = Engine(args, sd, torch.device('cuda'), B)
    z = synthetic.synthetic_z_c(1, 512, seed=1).repeat(B, 1).cuda()
    path = synthetic.synthetic_paths(B, 1024, seed=2).cuda()
    for steps in (8, 64):
        eng.decode(z, path, num_steps=steps, seed=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.decode(z, path, num_steps=steps, seed=1)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f'decode B={B} steps={steps}: {ms:.1f} ms total, {ms / steps:.3f} ms/step -> {B / (ms / steps * 1024 / 1e3):.3f} seq/s', flush=True)
    prof = eng.profile_step(B, B)
    print('profile_step:', {k: round(v, 3) for k, v in prof.items()})


STAGES = dict(gemm=stage_gemm, sampler=stage_sampler, stages=stage_stages, forward=stage_forward,
              decode=stage_decode, perf=stage_perf)

if __name__ == '__main__':
    if len(sys.argv) >= 3 and sys.argv[1] == '--run':
        STAGES[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or list(STAGES)
    rc = 0
    for n in names:
        print(f'===== {n} =====', flush=True)
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), '--run', n], timeout=int(os.environ.get('DIAG_TIMEOUT', '420')))
            print(f'===== {n}: exit {r.returncode} in {time.time() - t0:.1f}s', flush=True)
            rc |= r.returncode != 0
        except subprocess.TimeoutExpired:
            print(f'===== {n}: TIMEOUT', flush=True)
            rc |= 1
    sys.exit(rc)
