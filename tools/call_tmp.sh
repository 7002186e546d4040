cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests -q -m gpu -x -k "attention or small or decode" 2>&1 | tail -3
python - <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from biom3_b200 import engine
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
qkv = (torch.randn(3, 64, 8, 1024, 32, device='cuda')).bfloat16()
for v in (0, 2, 0, 2):
    print(f'variant {v}: {timeit(lambda: engine.attention_test(qkv, 8, v)):.1f} us (variant 2 includes the norm helper kernel)', flush=True)
P
for v in 0 1 0 1; do BIOM3_ATTN_BOUND=$v timeout 300 python tools/ab_step.py 256 >> gpurun_out/r02b_ab_attn_bound.jsonl 2>gpurun_out/ab_err.log; done
cat gpurun_out/r02b_ab_attn_bound.jsonl
