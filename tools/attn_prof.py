import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from biom3_b200 import engine
qkv = (torch.randn(3, 64, 16, 1024, 32, device='cuda')).bfloat16()
for v in (0, 3):
    for _ in range(2):
        engine.attention_test(qkv, 8, v)
torch.cuda.synchronize()
print('ok')
