import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hotformerloc_b200 import ops
M, C = 1_050_000, 256
y = torch.randn(M, C, device='cuda').to(torch.bfloat16)
W1 = (torch.randn(4 * C, C, device='cuda') / 16).to(torch.bfloat16)
W2 = (torch.randn(C, 4 * C, device='cuda') / 32).to(torch.bfloat16)
b1, b2 = torch.randn(4 * C, device='cuda'), torch.randn(C, device='cuda')
x = torch.randn(M, C, device='cuda'); xb = torch.empty(M, C, device='cuda', dtype=torch.bfloat16)
for _ in range(3):
    ops.mlp_fused(y, W1, b1, W2, b2, res=x, out_f32=x, out_bf16=xb)
torch.cuda.synchronize()
