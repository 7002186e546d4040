import sys, torch
sys.path.insert(0, '/root/repo')
from biom3_b200 import engine
g = torch.Generator().manual_seed(0)
for (M, N, K) in [(256, 256, 64), (256, 256, 512), (512, 512, 128), (1024, 1536, 512), (65536, 512, 2048)]:
    A = (torch.randn(M, K, generator=g) * 0.5).cuda().bfloat16()
    W = (torch.randn(N, K, generator=g) * 0.1).cuda().bfloat16()
    ref = A.float() @ W.float().t()
    out = engine.gemm_test(A, W, None, 4, 256, pair=True)
    torch.cuda.synchronize()
    err = ((out - ref).abs().max() / ref.abs().max()).item()
    print(f'pair gemm M={M} N={N} K={K}: rel_err={err:.3e}', flush=True)
    if err > 1e-3:
        bad = (out - ref).abs() > 1e-2 * ref.abs().max()
        rows = bad.any(1).nonzero().flatten(); cols = bad.any(0).nonzero().flatten()
        print('  bad rows n=', rows.numel(), rows[:8].tolist(), rows[-8:].tolist(), ' bad cols n=', cols.numel(), cols[:8].tolist(), cols[-8:].tolist())
        print('  out[0,:4]', out[0,:4].tolist(), 'ref', ref[0,:4].tolist(), ' out[128,:4]', out[128,:4].tolist(), 'ref', ref[128,:4].tolist())
        print('  out[0,128:132]', out[0,128:132].tolist(), 'ref', ref[0,128:132].tolist())
